// snk_hostxfer.cpp -- host half of the packed observation transport of snk_step_host / snk_reset_host.
//
// An observation byte is 0 or 1 (snake_env.py:484-492), so the NHWC block a host caller asks for carries
// one bit of information per byte.  The PCIe link (~56 GB/s) is two orders of magnitude slower than HBM,
// so the host entry points ship one CHANNEL-BIT byte per (cell, frame) -- bit c = channel c -- and widen
// it to the reference's eight 0/1 bytes here, on the host cores, chunk by chunk while later chunks are
// still in flight.  No game rule and no part of the observation encode runs here: this is bit -> byte
// widening of what the GPU produced, nothing else.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "snk_hostxfer.h"

namespace snk {
namespace {

struct WidenLut {
  uint64_t v[256];
  WidenLut() {
    for (int b = 0; b < 256; ++b) {
      uint64_t x = 0;
      for (int c = 0; c < 8; ++c) x |= (uint64_t)((b >> c) & 1) << (8 * c);   // little-endian: byte c = bit c
      v[b] = x;
    }
  }
};
const WidenLut g_lut;

void widen_scalar(const uint8_t* src, uint8_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) memcpy(dst + 8 * i, &g_lut.v[src[i]], 8);
}

#if defined(__x86_64__)
// 8 channel-bit bytes = one 64-bit lane mask -> 64 output bytes, written around the cache.
__attribute__((target("avx512f,avx512bw"))) void widen_avx512(const uint8_t* src, uint8_t* dst, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + 8 * i) & 63)) { memcpy(dst + 8 * i, &g_lut.v[src[i]], 8); ++i; }
  const __m512i one = _mm512_set1_epi8(1);
  for (; i + 8 <= n; i += 8) {
    uint64_t k;
    memcpy(&k, src + i, 8);
    _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + 8 * i), _mm512_maskz_mov_epi8((__mmask64)k, one));
  }
  _mm_sfence();
  for (; i < n; ++i) memcpy(dst + 8 * i, &g_lut.v[src[i]], 8);
}
bool cpu_has_avx512bw() {
  static const bool has = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
  const char* off = getenv("SNK_NO_AVX512");          // tests exercise the scalar path on AVX-512 hosts with it
  return has && !(off && *off == '1');
}
#endif

}  // namespace

void widen_bits(const uint8_t* src, uint8_t* dst, size_t n) {
#if defined(__x86_64__)
  if (cpu_has_avx512bw() && ((uintptr_t)dst & 7) == 0) { widen_avx512(src, dst, n); return; }
#endif
  widen_scalar(src, dst, n);
}

// ---- a small persistent pool: submit() slices a range over the workers, wait() drains it ----------
struct WidenPool::Impl {
  struct Task { const uint8_t* src; uint8_t* dst; size_t n; };
  std::vector<std::thread> workers;
  std::deque<Task> queue;
  std::mutex mu;
  std::condition_variable cv_work, cv_idle;
  size_t in_flight = 0;
  bool stop = false;

  void run() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || !queue.empty(); });
        if (queue.empty()) return;
        t = queue.front();
        queue.pop_front();
      }
      widen_bits(t.src, t.dst, t.n);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--in_flight == 0) cv_idle.notify_all();
      }
    }
  }
};

WidenPool::WidenPool(int threads) : impl_(new Impl), threads_(threads < 1 ? 1 : threads) {
  for (int t = 0; t < threads_; ++t) impl_->workers.emplace_back([this] { impl_->run(); });
}

WidenPool::~WidenPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv_work.notify_all();
  for (auto& w : impl_->workers) w.join();
  delete impl_;
}

void WidenPool::submit(const uint8_t* src, uint8_t* dst, size_t n) {
  if (n == 0) return;
  // slices are multiples of 64 source bytes so every slice but the first starts on a 64-byte output line
  size_t per = (n + (size_t)threads_ - 1) / (size_t)threads_;
  per = (per + 63) / 64 * 64;
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    for (size_t lo = 0; lo < n; lo += per) {
      const size_t len = lo + per < n ? per : n - lo;
      impl_->queue.push_back({src + lo, dst + 8 * lo, len});
      ++impl_->in_flight;
    }
  }
  impl_->cv_work.notify_all();
}

void WidenPool::wait() {
  std::unique_lock<std::mutex> lk(impl_->mu);
  impl_->cv_idle.wait(lk, [&] { return impl_->in_flight == 0; });
}

int default_host_threads() {
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  if (pthread_getaffinity_np(pthread_self(), sizeof set, &set) == 0) n = CPU_COUNT(&set);
  if (n < 1) n = (int)std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  return n > 32 ? 32 : n;
}

}  // namespace snk
