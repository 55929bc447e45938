// write_probe.cu -- what is this B200's ceiling for a WRITE-dominated stream?  (roofline context:
// the step kernel writes 4.8 GB and reads 0.7 GB per launch; MEASURED_PEAKS.json is a 50/50 copy.)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_probe write_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void wr(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n, size_t nread) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(i, 1, 2, 3);
  for (; i < n; i += stride) {
    if (MODE == 3) { if (i < nread) v = __ldcs(src + i); }
    if (MODE == 0) dst[i] = v;
    else if (MODE == 1 || MODE == 3) __stcs(dst + i, v);
    else __stwt(dst + i, v);
  }
}

int main() {
  const size_t bytes = 4800ull << 20, n = bytes / 16;
  uint4 *d, *s;
  cudaMalloc(&d, bytes); cudaMalloc(&s, bytes);
  cudaMemset(s, 1, bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const char* names[4] = {"st.global", "st.global.cs", "st.global.wt", "st.cs + 14% ld.cs"};
  for (int mode = 0; mode < 4; ++mode)
    for (int blocks : {148 * 4, 148 * 16, 148 * 64, 148 * 256}) {
      auto run = [&]() {
        const size_t nread = n / 7;
        if (mode == 0) wr<0><<<blocks, 256>>>(d, s, n, nread);
        if (mode == 1) wr<1><<<blocks, 256>>>(d, s, n, nread);
        if (mode == 2) wr<2><<<blocks, 256>>>(d, s, n, nread);
        if (mode == 3) wr<3><<<blocks, 256>>>(d, s, n, nread);
      };
      run(); run();
      cudaEventRecord(a);
      for (int k = 0; k < 10; ++k) run();
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double moved = (double)bytes * (mode == 3 ? 1.0 + 1.0 / 7 : 1.0);
      printf("%-20s blocks %6d  %.1f GB/s\n", names[mode], blocks, moved * 10 / (ms * 1e-3) / 1e9);
    }
  return 0;
}
