// rw_phase_probe.cu -- does separating the step kernel's DRAM reads from its writes in time pay?
// The step kernel moves, per tile of 8 environments, 5248 B of records in, 5248 B back and 30976 B of
// observations out, all interleaved; a pure write stream is faster than that mix.  This probe times
//   A  pure streaming write of the same bytes
//   B  the interleaved per-tile pattern (one warp per tile, like the kernel)
//   C  the same pattern when the records of a chunk were pulled into L2 first (separate prefetch launch,
//      chunk = `chunk_tiles` tiles) so that the tile warps only WRITE to DRAM
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rw_phase_probe rw_phase_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#ifndef REC_ENV_BYTES
#define REC_ENV_BYTES 544                            // 656 with the separate direction plane
#endif
constexpr int REC_TILE = 8 * REC_ENV_BYTES, OBS_TILE = 30976;     // bytes per tile (8 envs x record, 8 x 3872 B)

__global__ void pure_write(uint4* __restrict__ dst, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(i, 1, 2, 3);
  for (; i < n; i += stride) __stcs(dst + i, v);
}

// one warp per tile: records in (to registers), records back, observations out
template <bool L2ONLY>
__global__ void tile_pattern(uint8_t* __restrict__ recs, uint8_t* __restrict__ obs, int tile0, int ntiles) {
  const int tile = tile0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= tile0 + ntiles) return;
  const int lane = threadIdx.x & 31;
  uint4* r = reinterpret_cast<uint4*>(recs + (size_t)tile * REC_TILE);
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint4 keep[11];   // up to 11 x 512 B per warp
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int i = lane + 32 * k;
    keep[k] = make_uint4(0, 0, 0, 0);
    if (i < REC_TILE / 16) { keep[k] = L2ONLY ? __ldcg(r + i) : __ldcs(r + i); acc.x ^= keep[k].x; }
  }
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int i = lane + 32 * k;
    if (i < REC_TILE / 16) { keep[k].y += 1; r[i] = keep[k]; }
  }
  uint4* o = reinterpret_cast<uint4*>(obs + (size_t)tile * OBS_TILE);
  for (int i = lane; i < OBS_TILE / 16; i += 32) __stcs(o + i, acc);
}

// variant E: like tile_pattern, but only a pseudo-random `keep_of_41` of every 41 record sectors (32 B) are
// written back -- what a dirty-sector write-back of the records would cost
__global__ void tile_pattern_sparse(uint8_t* __restrict__ recs, uint8_t* __restrict__ obs, int ntiles, int keep_of_41) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= ntiles) return;
  const int lane = threadIdx.x & 31;
  uint4* r = reinterpret_cast<uint4*>(recs + (size_t)tile * REC_TILE);
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint4 keep[11];   // up to 11 x 512 B per warp
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int i = lane + 32 * k;
    keep[k] = make_uint4(0, 0, 0, 0);
    if (i < REC_TILE / 16) { keep[k] = __ldcs(r + i); acc.x ^= keep[k].x; }
  }
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int i = lane + 32 * k;                       // 16-byte unit; sector = i / 2
    const unsigned sector = (unsigned)(i >> 1) + (unsigned)tile * 164u;
    if (i < REC_TILE / 16 && (int)((sector * 2654435761u >> 16) % 41u) < keep_of_41) { keep[k].y += 1; r[i] = keep[k]; }
  }
  uint4* o = reinterpret_cast<uint4*>(obs + (size_t)tile * OBS_TILE);
  for (int i = lane; i < OBS_TILE / 16; i += 32) __stcs(o + i, acc);
}

__global__ void prefetch_l2(const uint8_t* __restrict__ recs, size_t bytes) {
  // one 128-byte line per thread: prefetch.global.L2 brings the line into L2 without a register destination
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 128;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 128;
  for (; i < bytes; i += stride) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(recs + i));
}

int main(int argc, char** argv) {
  const int ntiles = 131072;                               // 1,048,576 envs
  const size_t rec_bytes = (size_t)ntiles * REC_TILE, obs_bytes = (size_t)ntiles * OBS_TILE;
  uint8_t *recs, *obs;
  cudaMalloc(&recs, rec_bytes); cudaMalloc(&obs, obs_bytes);
  cudaMemset(recs, 1, rec_bytes); cudaMemset(obs, 0, obs_bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time_it = [&](const char* name, double bytes, auto&& fn) {
    fn(); fn();
    cudaEventRecord(a);
    for (int k = 0; k < 10; ++k) fn();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-56s %.3f ms  %.0f GB/s   (%s)\n", name, ms / 10, bytes * 10 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  const double total = 2.0 * rec_bytes + obs_bytes;
  time_it("A pure write of obs+record bytes", (double)(rec_bytes + obs_bytes),
          [&] { pure_write<<<148 * 16, 256>>>(reinterpret_cast<uint4*>(obs), obs_bytes / 16);
                pure_write<<<148 * 16, 256>>>(reinterpret_cast<uint4*>(recs), rec_bytes / 16); });
  for (int wpb : {1, 4})
    time_it(wpb == 1 ? "B interleaved tiles, 1 warp/CTA" : "B interleaved tiles, 4 warps/CTA", total,
            [&] { tile_pattern<false><<<(ntiles + wpb - 1) / wpb, 32 * wpb>>>(recs, obs, 0, ntiles); });
  for (int keep : {41, 24, 16, 12, 8}) {
    char name[128];
    snprintf(name, sizeof name, "E interleaved tiles, %d of 41 record sectors written back", keep);
    time_it(name, (double)rec_bytes * (1.0 + keep / 41.0) + obs_bytes,
            [&] { tile_pattern_sparse<<<ntiles, 32>>>(recs, obs, ntiles, keep); });
  }
  for (int chunk : {16384})
    for (int wpb : {1, 4}) {
      char name[128];
      snprintf(name, sizeof name, "C prefetch launch + tiles, chunk %d tiles (%.0f MB rec), %d w/CTA", chunk, chunk * REC_TILE / 1e6, wpb);
      time_it(name, total, [&] {
        for (int t0 = 0; t0 < ntiles; t0 += chunk) {
          const int n = ntiles - t0 < chunk ? ntiles - t0 : chunk;
          prefetch_l2<<<148 * 4, 256>>>(recs + (size_t)t0 * REC_TILE, (size_t)n * REC_TILE);
          tile_pattern<true><<<(n + wpb - 1) / wpb, 32 * wpb>>>(recs, obs, t0, n);
        }
      });
    }
  // D: prefetch of chunk c+1 overlapped with the tiles of chunk c on a second stream
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t ev[64]; for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  for (int chunk : {8192, 16384})
    time_it(chunk == 8192 ? "D prefetch(c+1) on 2nd stream || tiles(c), chunk 8192" : "D prefetch(c+1) on 2nd stream || tiles(c), chunk 16384", total, [&] {
      int c = 0;
      prefetch_l2<<<148 * 4, 256, 0, s2>>>(recs, (size_t)chunk * REC_TILE);
      cudaEventRecord(ev[0], s2);
      for (int t0 = 0; t0 < ntiles; t0 += chunk, ++c) {
        const int n = ntiles - t0 < chunk ? ntiles - t0 : chunk;
        cudaStreamWaitEvent(s1, ev[c], 0);
        if (t0 + chunk < ntiles) {
          const int n2 = ntiles - t0 - chunk < chunk ? ntiles - t0 - chunk : chunk;
          prefetch_l2<<<148 * 4, 256, 0, s2>>>(recs + (size_t)(t0 + chunk) * REC_TILE, (size_t)n2 * REC_TILE);
          cudaEventRecord(ev[c + 1], s2);
        }
        tile_pattern<true><<<n, 32, 0, s1>>>(recs, obs, t0, n);
      }
      cudaStreamSynchronize(s1);
    });
  return 0;
}
