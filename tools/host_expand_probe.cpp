// host_expand_probe.cpp -- how fast can this box's host cores expand channel-bit bytes (1 byte per
// observation cell) into the reference's uint8 0/1 NHWC layout (8 bytes per cell)?  Sizes the
// packed host transport of snk_step_host.
//   g++ -O3 -std=c++17 -pthread -o host_expand_probe host_expand_probe.cpp && ./host_expand_probe
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

static uint64_t g_lut[256];

static void expand_lut(const uint8_t* src, uint8_t* dst, size_t n) {
  uint64_t* o = reinterpret_cast<uint64_t*>(dst);
  for (size_t i = 0; i < n; ++i) o[i] = g_lut[src[i]];
}

__attribute__((target("avx512f,avx512bw"))) static void expand_avx512(const uint8_t* src, uint8_t* dst, size_t n) {
  const __m512i one = _mm512_set1_epi8(1);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t k;
    memcpy(&k, src + i, 8);
    _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i * 8), _mm512_maskz_mov_epi8((__mmask64)k, one));
  }
  for (; i < n; ++i) reinterpret_cast<uint64_t*>(dst)[i] = g_lut[src[i]];
  _mm_sfence();
}

int main() {
  for (int b = 0; b < 256; ++b) {
    uint64_t v = 0;
    for (int c = 0; c < 8; ++c) v |= (uint64_t)((b >> c) & 1) << (8 * c);
    g_lut[b] = v;
  }
  const size_t n = (size_t)256 << 20;           // packed bytes -> 2 GiB expanded
  uint8_t* src = (uint8_t*)aligned_alloc(64, n);
  uint8_t* dst = (uint8_t*)aligned_alloc(64, n * 8);
  for (size_t i = 0; i < n; ++i) src[i] = (uint8_t)(i * 2654435761u >> 13);
  memset(dst, 0, n * 8);
  const bool has512 = __builtin_cpu_supports("avx512bw");
  printf("hardware_concurrency %u, avx512bw %d\n", std::thread::hardware_concurrency(), (int)has512);
  for (int mode = 0; mode < (has512 ? 2 : 1); ++mode)
    for (int T : {1, 2, 4, 8, 16, 32}) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        const size_t per = (n / T + 63) / 64 * 64;
        for (int t = 0; t < T; ++t) {
          const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
          if (lo >= hi) break;
          th.emplace_back([=] { (mode ? expand_avx512 : expand_lut)(src + lo, dst + lo * 8, hi - lo); });
        }
        for (auto& x : th) x.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best) best = s;
      }
      printf("%s threads %2d: %.1f ms, %.1f GB/s written\n", mode ? "avx512-nt" : "lut      ", T, best * 1e3,
             n * 8 / best / 1e9);
    }
  uint64_t chk = 0;
  for (size_t i = 0; i < n * 8; i += 4097) chk += dst[i];
  printf("check %llu\n", (unsigned long long)chk);
  return 0;
}
