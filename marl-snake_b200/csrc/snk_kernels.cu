// snk_kernels.cu -- sm_100a kernels of the batched Snake-v1 step path.
//
// One CTA owns a tile of E consecutive environments.  Their records (grid, body-direction plane,
// snake table, counters, episode statistics) are contiguous in HBM, so the tile is staged into
// shared memory with 128-bit coalesced loads, stepped there, and written back the same way.
//
//   phase L  one thread per environment runs the step rules (env_step_logic, snk_core.cuh)
//   phase R  rare events, one WARP per environment: fruit respawn = k-th empty cell by
//            ballot/popc rank-select over the grid; auto-reset = spawn sampling + overlap vote
//   phase A  one warp per (environment, viewer): egocentric crop -> 1 byte of channel bits per
//            cell, written to a staging area in output order (and to the frame-stack history)
//   phase B  whole CTA: staging bytes -> NHWC uint8 observation, 128-bit coalesced stores
//
// Reference: SnakeEnv.step / reset / _encode (envs/snake_env.py:301-414, 131-159, 474-519) and the
// vector worker's auto-reset (wrappers.py:138-146).
#include <cuda_runtime.h>
#include <stdint.h>

#include "snk_core.cuh"
#include "snk_kernels.h"

namespace snk {

enum : uint8_t { F_RESET = 1, F_INIT = 2, F_SKIP = 4 };

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- phase R helpers (warp-cooperative) -----------------------------------------------------------

// Place k fruits on the k drawn ranks of the row-major empty-cell list (all ranks refer to the
// grid as it is before any of them is placed; duplicates collapse)    core/grid_util.py:126-133
__device__ void place_fruits_warp(const KParams& p, Rec& r, uint32_t env_local, int k, int purpose) {
  const Dims& d = p.d;
  const uint32_t lane = lane_id();
  const int HW = d.HW;
  int n_empty = 0;
  for (int base = 0; base < HW; base += 32) {
    const int c = base + (int)lane;
    const bool emp = c < HW && r.grid[c] == EMPTY;
    n_empty += __popc(__ballot_sync(0xffffffffu, emp));
  }
  if (n_empty == 0 || k <= 0) return;              // reference draws nothing when no cell is empty
  int rank = -1;
  if ((int)lane < k) {
    if (d.rng_mode == RNG_PHILOX) {
      rank = (int)draw_below(d, env_local, r.hdr->event, purpose, lane, (uint32_t)n_empty);
    } else {
      const int64_t lo = p.replay_off[env_local], hi = p.replay_off[env_local + 1];
      const int64_t at = lo + r.hdr->cursor + lane;
      if (at >= hi) { atomicOr(p.err, ERR_REPLAY_UNDERRUN); rank = 0; }
      else {
        rank = p.replay[at];
        if (rank < 0 || rank >= n_empty) { atomicOr(p.err, ERR_REPLAY_RANGE); rank = 0; }
      }
    }
  }
  __syncwarp();
  if (lane == 0 && d.rng_mode == RNG_REPLAY) r.hdr->cursor += (uint32_t)k;
  int run = 0, mycell = -1;
  for (int base = 0; base < HW; base += 32) {
    const int c = base + (int)lane;
    const bool emp = c < HW && r.grid[c] == EMPTY;
    const uint32_t b = __ballot_sync(0xffffffffu, emp);
    const int pc = __popc(b);
    if (rank >= run && rank < run + pc) mycell = base + (int)__fns(b, 0, rank - run + 1);
    run += pc;
  }
  __syncwarp();
  if (mycell >= 0) r.grid[mycell] = (uint8_t)FRUIT;
  __syncwarp();
}

// SnakeEnv.reset for one environment record in shared memory      envs/snake_env.py:131-159, 576-596
__device__ void reset_env_warp(const KParams& p, Rec& r, uint32_t env_local) {
  const Dims& d = p.d;
  const uint32_t lane = lane_id();
  const int ns = d.ns, K = d.K, W = d.W;
  for (int c = (int)lane; c < d.HW; c += 32) r.grid[c] = wall_or_empty(c, d.H, d.W);
  __syncwarp();

  uint64_t entry = 0;
  const bool me = (int)lane < ns;
  if (d.rng_mode == RNG_REPLAY) {
    if (me) {
      const int64_t lo = p.replay_off[env_local], hi = p.replay_off[env_local + 1];
      const int64_t at = lo + r.hdr->cursor + lane;
      int pick = 0;
      if (at >= hi) atomicOr(p.err, ERR_REPLAY_UNDERRUN);
      else {
        pick = p.replay[at];
        if (pick < 0 || (uint32_t)pick >= d.n_cand) { atomicOr(p.err, ERR_REPLAY_RANGE); pick = 0; }
      }
      entry = p.spawn[pick];
    }
    __syncwarp();
    if (lane == 0) r.hdr->cursor += (uint32_t)ns;
  }
  // sample poses until no two snakes share a cell (np.random.permutation(n)[:ns] retry loop, :579-586;
  // ns independent uniform picks conditioned on "no shared cell" have the same law)
  for (uint32_t attempt = 0;; ++attempt) {
    if (d.rng_mode == RNG_PHILOX && me)
      entry = p.spawn[draw_below(d, env_local, r.hdr->event, DRAW_SPAWN, attempt * (uint32_t)ns + lane, d.n_cand)];
    if (me) {
      int c = spawn_head(entry);
      r.grid[c] = (uint8_t)(BODY + 10 * lane);
      for (int j = 1; j < K; ++j) { c += dir_delta(spawn_link(entry, j), W); r.grid[c] = (uint8_t)(BODY + 10 * lane); }
    }
    __syncwarp();
    bool clash = false;
    if (me) {
      int c = spawn_head(entry);
      clash = r.grid[c] != BODY + 10 * lane;
      for (int j = 1; j < K; ++j) { c += dir_delta(spawn_link(entry, j), W); clash |= r.grid[c] != BODY + 10 * lane; }
    }
    const uint32_t any_clash = __ballot_sync(0xffffffffu, clash);
    if (!any_clash) break;
    if (d.rng_mode == RNG_REPLAY) { if (lane == 0) atomicOr(p.err, ERR_REPLAY_RANGE); break; }
    if (attempt + 1 >= SPAWN_ATTEMPT_CAP) { if (lane == 0) atomicOr(p.err, ERR_SPAWN_GIVEUP); break; }
    __syncwarp();
    if (me) {
      int c = spawn_head(entry);
      r.grid[c] = (uint8_t)EMPTY;
      for (int j = 1; j < K; ++j) { c += dir_delta(spawn_link(entry, j), W); r.grid[c] = (uint8_t)EMPTY; }
    }
    __syncwarp();
  }
  __syncwarp();
  // snake table + body-direction plane; serialised per snake because plane bytes hold 4 cells
  for (int s = 0; s < ns; ++s) {
    if ((int)lane == s) {
      int c = spawn_head(entry);
      r.head[s] = (uint16_t)c;
      r.grid[c] = (uint8_t)(HEAD + 10 * s);
      for (int j = 1; j < K; ++j) {
        const int l = spawn_link(entry, j);
        c += dir_delta(l, W);
        dirp_set(r.dirp, c, (l + 2) & 3);            // toward the head
      }
      r.tail[s] = (uint16_t)c;
      r.grid[c] = (uint8_t)(TAIL + 10 * s);
      r.len[s] = (uint16_t)K;
      r.dir[s] = (uint8_t)((spawn_link(entry, 1) + 2) & 3);   // coords[0] - coords[1]  core/snake.py:58-61
      r.alive[s] = 1;
      r.score[s] = 0.0; r.steps[s] = 0; r.fruits[s] = 0; r.kills[s] = 0;
    }
    __syncwarp();
  }
  place_fruits_warp(p, r, env_local, d.nfruits, DRAW_RESET_FRUIT);
  if (lane == 0) { r.hdr->alive_counter = ns; r.hdr->episode_length = 0; }
  __syncwarp();
}

// ---- the fused step / reset / encode kernel -------------------------------------------------------
//
// Template parameters are compile-time copies of configuration values (0 = read the runtime value):
// the BASELINE shapes get specialised instances so index arithmetic folds to shifts and multiplies.
template <int kNS, int kW, int kOH, int kOW, int kFS>
struct Shape {
  const Dims& d;
  __device__ __forceinline__ explicit Shape(const Dims& dd) : d(dd) {}
  __device__ __forceinline__ int ns() const { return kNS ? kNS : d.ns; }
  __device__ __forceinline__ int W() const { return kW ? kW : d.W; }
  __device__ __forceinline__ int oh() const { return kOH ? kOH : d.oh; }
  __device__ __forceinline__ int ow() const { return kOW ? kOW : d.ow; }
  __device__ __forceinline__ int ohw() const { return oh() * ow(); }
  __device__ __forceinline__ int fs() const { return kFS ? kFS : d.fs; }
  __device__ __forceinline__ int lut_stride() const { return 10 * ns() + 6; }   // cell codes 0 .. 10*(ns-1)+5
};

__device__ __forceinline__ void st_cs_128(void* p, uint2 a, uint2 b) { __stcs(reinterpret_cast<uint4*>(p), make_uint4(a.x, a.y, b.x, b.y)); }
__device__ __forceinline__ void st_cs_64(void* p, uint2 a) { __stcs(reinterpret_cast<uint2*>(p), a); }

template <int kNS, int kW, int kOH, int kOW, int kFS>
__global__ void __launch_bounds__(SNK_MAX_THREADS)
snk_tile_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const Shape<kNS, kW, kOH, kOW, kFS> sh(d);
  const int E = p.E;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, nwarps = nt >> 5;
  const uint32_t lane = lane_id();
  const int e0 = blockIdx.x * E;
  const int ne = min(E, d.N - e0);
  const int ns = sh.ns();
  const int fs = sh.fs(), ohw = sh.ohw(), ow = sh.ow(), W = sh.W();
  const int LS = sh.lut_stride();

  // shared memory carve-up (offsets computed on the host side the same way, see tile_smem_bytes)
  uint8_t* s_rec = smem;
  uint8_t* s_lut = s_rec + (size_t)E * d.rec_bytes;                    // fs==1: uint2[ns*LS]; else uint8[ns*LS]
  uint8_t* s_stage = s_lut + round_up(ns * LS * (fs == 1 ? 8 : 1), 16);   // fs>1 only
  uint8_t* s_scr = s_stage + (fs == 1 ? 0 : round_up(E * d.stage_env_bytes, 16));
  uint8_t* s_fruit = s_scr + (size_t)E * d.scr_bytes;
  uint8_t* s_flag = s_fruit + E;

  // ---- stage the tile's records: HBM -> shared, 128-bit coalesced
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.recs + (size_t)e0 * d.rec_bytes);
    uint4* dst = reinterpret_cast<uint4*>(s_rec);
    const int n16 = ne * (d.rec_bytes >> 4);
    for (int i = tid; i < n16; i += nt) dst[i] = __ldcs(src + i);
  }
  // ---- per-viewer lookup table: cell code -> channel bits (expanded to 8 bytes when fs == 1)
  for (int idx = tid; idx < ns * LS; idx += nt) {
    const int v = idx / LS, code = idx - v * LS;
    const uint32_t bits = cell_bits((uint32_t)code, (uint32_t)v);
    if (fs == 1) reinterpret_cast<uint2*>(s_lut)[idx] = make_uint2(spread4(bits & 15u), spread4(bits >> 4));
    else s_lut[idx] = (uint8_t)bits;
  }
  __syncthreads();

  // ---- phase L: one thread per environment
  int deaths = 0, ended = 0;
  double ret_sum = 0.0; uint32_t len_sum = 0, fruit_sum = 0, kill_sum = 0;
  if (tid < ne) {
    const int e = e0 + tid;
    uint8_t* base = s_rec + (size_t)tid * d.rec_bytes;
    Rec r = rec_view(base, d);
    uint8_t flag = 0, fruit = 0;
    if (p.mode == MODE_STEP) {
      r.hdr->event += 1;
      uint32_t err = 0;
      const StepResult res = env_step_logic(d, base, s_scr + (size_t)tid * d.scr_bytes,
                                            p.actions + (size_t)e * ns, p.rew + (size_t)e * ns,
                                            p.done + (size_t)e * ns, &err);
      if (err) atomicOr(p.err, err);
      fruit = res.fruit_taken;
      deaths = res.deaths;
      if (p.fin) p.fin[e] = res.finished;
      if (res.finished) {
        ended = 1;
        len_sum = r.hdr->episode_length;
        for (int i = 0; i < ns; ++i) {
          const size_t o = (size_t)e * ns + i;
          if (p.rank) p.rank[o] = competition_rank(r.score, ns, i);
          if (p.ep_scores) p.ep_scores[o] = r.score[i];
          if (p.ep_steps) p.ep_steps[o] = (int32_t)r.steps[i];
          if (p.ep_fruits) p.ep_fruits[o] = (int32_t)r.fruits[i];
          if (p.ep_kills) p.ep_kills[o] = (int32_t)r.kills[i];
          ret_sum += r.score[i]; fruit_sum += r.fruits[i]; kill_sum += r.kills[i];
        }
        for (int i = 0; i < ns; ++i) { r.score[i] = 0.0; r.steps[i] = 0; r.fruits[i] = 0; r.kills[i] = 0; }   // :412
        if (d.auto_reset) flag |= F_RESET;                                        // wrappers.py:141-143
      }
    } else {
      const bool sel = p.mask == nullptr || p.mask[e] != 0;
      if (!sel) flag = F_SKIP;
      else if (p.mode == MODE_RESET) { r.hdr->event += 1; flag = F_RESET; }
      else flag = F_INIT;
    }
    s_fruit[tid] = fruit;
    s_flag[tid] = flag;
  }
  if (p.mode == MODE_STEP && warp * 32 < ne) {        // warp-aggregated rollout statistics
    const int dsum = __reduce_add_sync(0xffffffffu, deaths);
    const uint32_t any_end = __ballot_sync(0xffffffffu, ended);
    if (lane == 0 && dsum) atomicAdd(p.stats + STAT_DEATHS, (double)dsum);
    if (any_end) {
      const uint32_t ls = __reduce_add_sync(0xffffffffu, len_sum);
      const uint32_t fsum = __reduce_add_sync(0xffffffffu, fruit_sum);
      const uint32_t ks = __reduce_add_sync(0xffffffffu, kill_sum);
      if (ended) atomicAdd(p.stats + STAT_RETURN, ret_sum);
      if (lane == 0) {
        atomicAdd(p.stats + STAT_EPISODES, (double)__popc(any_end));
        atomicAdd(p.stats + STAT_EP_STEPS, (double)ls);
        atomicAdd(p.stats + STAT_FRUITS, (double)fsum);
        atomicAdd(p.stats + STAT_KILLS, (double)ks);
      }
    }
  }
  __syncthreads();

  // ---- phase R: rare events, one warp per environment
  for (int el = warp; el < ne; el += nwarps) {
    const uint8_t fruit = s_fruit[el], flag = s_flag[el];
    if (!fruit && !(flag & F_RESET)) continue;
    Rec r = rec_view(s_rec + (size_t)el * d.rec_bytes, d);
    if (fruit) place_fruits_warp(p, r, (uint32_t)(e0 + el), fruit, DRAW_STEP_FRUIT);
    if (flag & F_RESET) reset_env_warp(p, r, (uint32_t)(e0 + el));
  }
  __syncthreads();

  const bool want_obs = p.obs != nullptr;

  if (fs == 1) {
    // ---- write the records back: shared -> HBM (nothing below modifies them)
    {
      uint4* dst = reinterpret_cast<uint4*>(p.recs + (size_t)e0 * d.rec_bytes);
      const uint4* src = reinterpret_cast<const uint4*>(s_rec);
      const int n16 = ne * (d.rec_bytes >> 4);
      for (int i = tid; i < n16; i += nt) dst[i] = src[i];
    }
    // ---- fused encode: one warp per (environment, viewer); each lane produces 16 output bytes
    //      (two cells x 8 channels) per iteration straight from the staged grid through the LUT.
    //      snake_env.py:474-519.  A viewer's block of ohw*8 bytes is only 8-byte aligned when ohw is
    //      odd, so its 16-byte units are laid out from the address parity and the two end units may
    //      be half units.
    if (want_obs) {
      const uint2* lut_all = reinterpret_cast<const uint2*>(s_lut);
      const int H = d.H, V = d.V;
      const int pairs = ne * ns;
      for (int pv = warp; pv < pairs; pv += nwarps) {
        const int el = pv / ns, v = pv - el * ns;
        if (s_flag[el] & F_SKIP) continue;
        const uint8_t* base = s_rec + (size_t)el * d.rec_bytes;
        const uint8_t* grid = base;
        const uint8_t alive = base[d.off_snk + 7 * ns + v];
        // crop centre: own head, or cell (0,0) when the viewer has no head in the grid   :500-502
        const int hc = alive ? (int)((const uint16_t*)(base + d.off_snk))[v] : 0;
        int r0 = 0, c0 = 0;
        if (V > 0) { const int hr = hc / W; r0 = hr - V; c0 = hc - hr * W - V; }
        uint8_t* outv = p.obs + ((size_t)(e0 + el) * ns + v) * (size_t)ohw * 8;
        const int shift = (int)((reinterpret_cast<uintptr_t>(outv) >> 3) & 1);
        const uint2* lut = lut_all + v * LS;
        const int units = (ohw + shift + 1) >> 1;
        for (int u = (int)lane; u < units; u += 32) {
          const int ca = 2 * u - shift, cb = ca + 1;
          const bool va = ca >= 0, vb = cb < ohw;
          uint2 qa = make_uint2(0, 0), qb = make_uint2(0, 0);
          {
            const int c = va ? ca : 0;
            const int i = c / ow, j = c - i * ow;
            const int rr = r0 + i, cc = c0 + j;
            uint32_t code = 0;
            if ((unsigned)rr < (unsigned)H && (unsigned)cc < (unsigned)W) code = grid[rr * W + cc];
            qa = lut[code];
          }
          {
            const int c = vb ? cb : 0;
            const int i = c / ow, j = c - i * ow;
            const int rr = r0 + i, cc = c0 + j;
            uint32_t code = 0;
            if ((unsigned)rr < (unsigned)H && (unsigned)cc < (unsigned)W) code = grid[rr * W + cc];
            qb = lut[code];
          }
          uint8_t* dst = outv + (ptrdiff_t)ca * 8;
          if (va && vb) st_cs_128(dst, qa, qb);
          else if (va) st_cs_64(dst, qa);
          else st_cs_64(dst + 8, qb);
        }
      }
    }
    return;
  }

  // ================= frame_stack > 1: channel-bit frames staged in output order ======================
  // ---- older frames of the stack: history rows -> staging (oldest-first)
  if (want_obs) {
    const int rows = ne * ns * fs;
    for (int row = warp; row < rows; row += nwarps) {
      const int el = row / (ns * fs);
      const int rem = row - el * ns * fs;
      const int v = rem / fs, slot = rem - v * fs;
      if (s_flag[el] & (F_RESET | F_INIT | F_SKIP)) continue;
      const int hpos = (int)((const EnvHdr*)(s_rec + (size_t)el * d.rec_bytes + d.off_hdr))->hpos;
      if (slot == hpos) continue;                  // about to be overwritten by the new frame
      int f = slot - hpos - 1; if (f < 0) f += fs;
      const uint8_t* src = p.hist + (size_t)(e0 + el) * d.hist_env_bytes + (size_t)(v * fs + slot) * d.ohw_p;
      uint8_t* dst = s_stage + (size_t)el * d.stage_env_bytes + (size_t)v * ohw * fs + f;
      for (int c4 = (int)lane * 4; c4 < ohw; c4 += 128) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + c4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c4 + k < ohw) dst[(size_t)(c4 + k) * fs] = (uint8_t)(w >> (8 * k));
      }
    }
  }
  // ---- phase A: encode the new frame, one warp per (environment, viewer)     snake_env.py:474-519
  {
    const int H = d.H, V = d.V;
    const int pairs = ne * ns;
    for (int pv = warp; pv < pairs; pv += nwarps) {
      const int el = pv / ns, v = pv - el * ns;
      const uint8_t flag = s_flag[el];
      if (flag & F_SKIP) continue;
      const uint8_t* base = s_rec + (size_t)el * d.rec_bytes;
      const uint8_t* grid = base;
      const uint8_t alive = base[d.off_snk + 7 * ns + v];
      const int hc = alive ? (int)((const uint16_t*)(base + d.off_snk))[v] : 0;
      int r0 = 0, c0 = 0;
      if (V > 0) { const int hr = hc / W; r0 = hr - V; c0 = hc - hr * W - V; }
      const bool init = (flag & (F_RESET | F_INIT)) != 0;
      const int hpos = (int)((const EnvHdr*)(base + d.off_hdr))->hpos;
      const uint8_t* lut = s_lut + v * LS;
      uint8_t* stg = s_stage + (size_t)el * d.stage_env_bytes + (size_t)v * ohw * fs;
      uint8_t* hrow = p.hist + (size_t)(e0 + el) * d.hist_env_bytes + (size_t)(v * fs) * d.ohw_p;
      for (int cell = (int)lane; cell < ohw; cell += 32) {
        const int i = cell / ow, j = cell - i * ow;
        const int rr = r0 + i, cc = c0 + j;
        uint32_t code = 0;
        if ((unsigned)rr < (unsigned)H && (unsigned)cc < (unsigned)W) code = grid[rr * W + cc];
        const uint8_t bits = lut[code];
        if (!init) {
          stg[(size_t)cell * fs + (fs - 1)] = bits;
          hrow[(size_t)hpos * d.ohw_p + cell] = bits;
        } else {                                   // reset: every slot holds the first frame (:452-457)
          for (int f = 0; f < fs; ++f) {
            stg[(size_t)cell * fs + f] = bits;
            hrow[(size_t)f * d.ohw_p + cell] = bits;
          }
        }
      }
    }
  }
  __syncthreads();
  if (tid < ne && !(s_flag[tid] & F_SKIP)) {
    EnvHdr* h = (EnvHdr*)(s_rec + (size_t)tid * d.rec_bytes + d.off_hdr);
    h->hpos = (s_flag[tid] & (F_RESET | F_INIT)) ? 0u : (h->hpos + 1u) % (uint32_t)fs;
  }
  __syncthreads();

  // ---- write the records back: shared -> HBM
  {
    uint4* dst = reinterpret_cast<uint4*>(p.recs + (size_t)e0 * d.rec_bytes);
    const uint4* src = reinterpret_cast<const uint4*>(s_rec);
    const int n16 = ne * (d.rec_bytes >> 4);
    for (int i = tid; i < n16; i += nt) dst[i] = src[i];
  }

  // ---- phase B: channel bits -> NHWC uint8, coalesced
  if (want_obs) {
    uint8_t* out = p.obs + (size_t)e0 * d.obs_env_bytes;
    if (p.mode == MODE_STEP && p.vec16) {
      const int total = ne * d.stage_env_bytes;             // staging bytes == 8-byte output units
      const int n16 = total >> 1;
      const uint16_t* s2 = reinterpret_cast<const uint16_t*>(s_stage);
      uint4* o4 = reinterpret_cast<uint4*>(out);
      for (int u = tid; u < n16; u += nt) {
        const uint32_t two = s2[u];
        uint4 q;
        q.x = spread4(two & 15u); q.y = spread4((two >> 4) & 15u);
        q.z = spread4((two >> 8) & 15u); q.w = spread4(two >> 12);
        __stcs(o4 + u, q);
      }
      if ((total & 1) && tid == 0) {
        const uint32_t b = s_stage[total - 1];
        __stcs(reinterpret_cast<uint2*>(out) + (total - 1), make_uint2(spread4(b & 15u), spread4(b >> 4)));
      }
    } else {
      for (int el = 0; el < ne; ++el) {
        if (s_flag[el] & F_SKIP) continue;
        const uint8_t* stg = s_stage + (size_t)el * d.stage_env_bytes;
        uint2* o2 = reinterpret_cast<uint2*>(out + (size_t)el * d.obs_env_bytes);
        for (int u = tid; u < d.stage_env_bytes; u += nt) {
          const uint32_t b = stg[u];
          __stcs(o2 + u, make_uint2(spread4(b & 15u), spread4(b >> 4)));
        }
      }
    }
  }
}

// ---- state export / import (parity + checkpoint interface; thread per environment) ----------------
__global__ void snk_get_state_kernel(const Dims d, const uint8_t* __restrict__ recs, StateView sv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  Rec r = rec_view(const_cast<uint8_t*>(recs) + (size_t)e * d.rec_bytes, d);
  if (sv.grid) for (int c = 0; c < d.HW; ++c) sv.grid[(size_t)e * d.HW + c] = r.grid[c];
  if (sv.alive_counter) sv.alive_counter[e] = r.hdr->alive_counter;
  if (sv.episode_length) sv.episode_length[e] = (int32_t)r.hdr->episode_length;
  for (int i = 0; i < d.ns; ++i) {
    const size_t o = (size_t)e * d.ns + i;
    const bool al = r.alive[i] != 0;
    if (sv.head) sv.head[o] = al ? r.head[i] : -1;
    if (sv.tail) sv.tail[o] = al ? r.tail[i] : -1;
    if (sv.length) sv.length[o] = al ? r.len[i] : 0;
    if (sv.dir) sv.dir[o] = r.dir[i];
    if (sv.alive) sv.alive[o] = r.alive[i];
    if (sv.cells) {
      int32_t* cl = sv.cells + o * sv.max_cells;
      for (int k = 0; k < sv.max_cells; ++k) cl[k] = -1;
      if (al) {
        int c = r.tail[i];
        for (int k = (int)r.len[i] - 1; k >= 0; --k) {
          if (k < sv.max_cells) cl[k] = c;
          c += dir_delta(dirp_get(r.dirp, c), d.W);
        }
      }
    }
  }
}

__global__ void snk_set_state_kernel(const Dims d, uint8_t* __restrict__ recs, StateView sv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  Rec r = rec_view(recs + (size_t)e * d.rec_bytes, d);
  for (int c = 0; c < d.HW; ++c) r.grid[c] = sv.grid[(size_t)e * d.HW + c];
  for (int i = 0; i < d.ns; ++i) {
    const size_t o = (size_t)e * d.ns + i;
    const int32_t* cl = sv.cells + o * sv.max_cells;
    const int len = sv.length[o];
    r.alive[i] = sv.alive[o];
    r.dir[i] = sv.dir[o];
    r.len[i] = (uint16_t)(sv.alive[o] ? len : 0);
    r.head[i] = 0; r.tail[i] = 0;
    if (sv.alive[o] && len > 0) {
      r.head[i] = (uint16_t)cl[0];
      r.tail[i] = (uint16_t)cl[len - 1];
      for (int k = 1; k < len; ++k) {
        const int diff = cl[k - 1] - cl[k];
        const int dd = diff == -d.W ? 0 : diff == 1 ? 1 : diff == d.W ? 2 : 3;
        dirp_set(r.dirp, cl[k], dd);
      }
    }
    r.score[i] = 0.0; r.steps[i] = 0; r.fruits[i] = 0; r.kills[i] = 0;
  }
  r.hdr->alive_counter = sv.alive_counter[e];
  r.hdr->episode_length = (uint32_t)sv.episode_length[e];
  r.hdr->hpos = 0;
}

__global__ void snk_init_records_kernel(const Dims d, uint8_t* __restrict__ recs) {
  const size_t n = (size_t)d.N * d.rec_bytes / 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(recs)[i] = make_uint4(0, 0, 0, 0);
}

// ---- launch wrappers -----------------------------------------------------------------------------
size_t tile_smem_bytes(const Dims& d, int E) {
  const int LS = 10 * d.ns + 6;
  size_t b = (size_t)E * d.rec_bytes + (size_t)round_up(d.ns * LS * (d.fs == 1 ? 8 : 1), 16);
  if (d.fs > 1) b += (size_t)round_up(E * d.stage_env_bytes, 16);
  b += (size_t)E * d.scr_bytes + 2 * (size_t)E + 16;
  return b;
}

template <int kNS, int kW, int kOH, int kOW, int kFS>
static cudaError_t launch_instance(const KParams& p, int threads, size_t smem_bytes, cudaStream_t stream) {
  static size_t configured = 0;
  auto kern = snk_tile_kernel<kNS, kW, kOH, kOW, kFS>;
  if (smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    configured = smem_bytes;
  }
  const int grid = (p.d.N + p.E - 1) / p.E;
  kern<<<grid, threads, smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

// Specialised instances for the BASELINE shapes; everything else runs the generic instance.
cudaError_t launch_tile_kernel(const KParams& p, int threads, size_t smem_bytes, cudaStream_t stream) {
  const Dims& d = p.d;
  const bool generic = p.force_generic != 0;
#define SNK_TRY(NS, W_, OH, OW, FS)                                                              \
  if (!generic && d.ns == NS && d.W == W_ && d.oh == OH && d.ow == OW && d.fs == FS)             \
    return launch_instance<NS, W_, OH, OW, FS>(p, threads, smem_bytes, stream);
  SNK_TRY(4, 20, 11, 11, 1)      // cfg1 / cfg5: 20x20, 4 snakes, vision 5
  SNK_TRY(4, 20, 20, 20, 1)      // cfg2: full-grid observation
  SNK_TRY(4, 20, 11, 11, 4)      // cfg3: frame_stack 4
  SNK_TRY(16, 64, 15, 15, 1)     // cfg4: 64x64, 16 snakes, vision 7
#undef SNK_TRY
  return launch_instance<0, 0, 0, 0, 0>(p, threads, smem_bytes, stream);
}

cudaError_t launch_get_state(const Dims& d, const uint8_t* recs, const StateView& sv, cudaStream_t s) {
  snk_get_state_kernel<<<(d.N + 127) / 128, 128, 0, s>>>(d, recs, sv);
  return cudaGetLastError();
}
cudaError_t launch_set_state(const Dims& d, uint8_t* recs, const StateView& sv, cudaStream_t s) {
  snk_set_state_kernel<<<(d.N + 127) / 128, 128, 0, s>>>(d, recs, sv);
  return cudaGetLastError();
}
cudaError_t launch_init_records(const Dims& d, uint8_t* recs, cudaStream_t s) {
  snk_init_records_kernel<<<592, 256, 0, s>>>(d, recs);
  return cudaGetLastError();
}

}  // namespace snk
