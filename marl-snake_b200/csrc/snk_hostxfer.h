// snk_hostxfer.h -- host half of the packed observation transport (see snk_hostxfer.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace snk {

// dst[8*i + c] = (src[i] >> c) & 1 for i < n, c < 8 (single thread).
void widen_bits(const uint8_t* src, uint8_t* dst, size_t n);

class WidenPool {
 public:
  explicit WidenPool(int threads);
  ~WidenPool();
  WidenPool(const WidenPool&) = delete;
  WidenPool& operator=(const WidenPool&) = delete;
  void submit(const uint8_t* src, uint8_t* dst, size_t n);   // returns at once
  void wait();                                               // until everything submitted is written
  int threads() const { return threads_; }

 private:
  struct Impl;
  Impl* impl_;
  int threads_;
};

int default_host_threads();   // cores this process may run on, at most 32

}  // namespace snk
