// snk_hostsim.cpp -- TEST INFRASTRUCTURE: sequential host driver around the step rules of
// marl-snake_b200/csrc/snk_core.cuh (the same source the CUDA kernels compile), so the rule
// implementation, record layout, Philox stream and frame-stack ring can be checked against the
// golden vectors on a box without a GPU.  Never linked into libsnk.so, never used by the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/snk.h"
#include "../../marl-snake_b200/csrc/snk_core.cuh"

namespace snk { int64_t spawn_enumerate(int H, int W, int K, uint64_t* packed_out, int32_t* cells_out, int64_t cap, int64_t limit = 0, const uint8_t* walls = nullptr); }
using namespace snk;

struct HostSim {
  Dims d;
  std::vector<uint8_t> recs, hist, scr;
  std::vector<uint64_t> spawn;
  std::vector<int32_t> replay;
  std::vector<int64_t> replay_off;
  std::vector<uint8_t> walls;     // custom wall layout (cell codes EMPTY / WALL), empty = walled box
  uint32_t err = 0;
};

static void place_fruits(HostSim& h, Rec& r, uint32_t env, int k, int purpose) {
  const Dims& d = h.d;
  const int n_empty = count_empty_seq(r.grid, d.HW);
  if (n_empty == 0 || k <= 0) return;
  std::vector<int> cells;
  for (int j = 0; j < k; ++j) {
    int rank;
    if (d.rng_mode == RNG_PHILOX) rank = (int)draw_below(d, env, r.hdr->event, purpose, (uint32_t)j, (uint32_t)n_empty);
    else {
      const int64_t at = h.replay_off[env] + r.hdr->cursor + j;
      if (at >= h.replay_off[env + 1]) { h.err |= ERR_REPLAY_UNDERRUN; rank = 0; }
      else { rank = h.replay[at]; if (rank < 0 || rank >= n_empty) { h.err |= ERR_REPLAY_RANGE; rank = 0; } }
    }
    cells.push_back(nth_empty_seq(r.grid, d.HW, rank));
  }
  if (d.rng_mode == RNG_REPLAY) r.hdr->cursor += (uint32_t)k;
  for (int c : cells) r.grid[c] = (uint8_t)FRUIT;
}

static void reset_env(HostSim& h, Rec& r, uint32_t env) {
  const Dims& d = h.d;
  const int ns = d.ns, K = d.K, W = d.W;
  for (int c = 0; c < d.HW; ++c) r.grid[c] = h.walls.empty() ? wall_or_empty(c, d.H, d.W) : h.walls[c];
  std::vector<uint64_t> entry(ns);
  if (d.rng_mode == RNG_REPLAY) {
    for (int i = 0; i < ns; ++i) {
      const int64_t at = h.replay_off[env] + r.hdr->cursor + i;
      int pick = 0;
      if (at >= h.replay_off[env + 1]) h.err |= ERR_REPLAY_UNDERRUN;
      else { pick = h.replay[at]; if (pick < 0 || (uint32_t)pick >= d.n_cand) { h.err |= ERR_REPLAY_RANGE; pick = 0; } }
      entry[i] = h.spawn[pick];
    }
    r.hdr->cursor += (uint32_t)ns;
  }
  for (uint32_t attempt = 0;; ++attempt) {
    if (d.rng_mode == RNG_PHILOX)
      for (int i = 0; i < ns; ++i)
        entry[i] = h.spawn[draw_below(d, env, r.hdr->event, DRAW_SPAWN, attempt * (uint32_t)ns + i, d.n_cand)];
    bool clash = false;
    for (int i = 0; i < ns; ++i) {
      int c = spawn_head(entry[i]);
      for (int j = 0; j < K; ++j) {
        if (j) c += dir_delta(spawn_link(entry[i], j), W);
        if (r.grid[c] != EMPTY) clash = true;
        r.grid[c] = (uint8_t)(BODY + 10 * i);
      }
    }
    if (!clash) break;
    if (d.rng_mode == RNG_REPLAY) { h.err |= ERR_REPLAY_RANGE; break; }
    if (attempt + 1 >= SPAWN_ATTEMPT_CAP) { h.err |= ERR_SPAWN_GIVEUP; break; }
    for (int i = 0; i < ns; ++i) {
      int c = spawn_head(entry[i]);
      for (int j = 0; j < K; ++j) { if (j) c += dir_delta(spawn_link(entry[i], j), W); r.grid[c] = (uint8_t)EMPTY; }
    }
  }
  for (int s = 0; s < ns; ++s) {
    int c = spawn_head(entry[s]);
    r.head[s] = (uint16_t)c;
    r.grid[c] = (uint8_t)(HEAD + 10 * s);
    for (int j = 1; j < K; ++j) {
      const int l = spawn_link(entry[s], j);
      c += dir_delta(l, W);
      r.grid[c] = (uint8_t)((j == K - 1 ? TAIL : BODY) + 10 * s);
      set_body_dir(d, r, c, (l + 2) & 3);
    }
    r.tail[s] = (uint16_t)c;
    r.len[s] = (uint16_t)K;
    r.dir[s] = (uint8_t)((spawn_link(entry[s], 1) + 2) & 3);
    r.alive[s] = 1;
    stats_zero(d, r, s);
  }
  place_fruits(h, r, env, d.nfruits, DRAW_RESET_FRUIT);
  r.hdr->alive_counter = (int16_t)ns;
  r.hdr->episode_length = 0;
}

// frame encode + stack, mirroring phases A/B of the kernel (history ring in `hist`)
static void encode_env(HostSim& h, Rec& r, uint32_t env, bool init, uint8_t* obs) {
  const Dims& d = h.d;
  const int fs = d.fs, ohw = d.ohw;
  std::vector<uint8_t> stage((size_t)d.stage_env_bytes, 0);
  const int hpos = (int)r.hdr->hpos;
  for (int v = 0; v < d.ns; ++v) {
    uint8_t* hrow = fs > 1 ? h.hist.data() + (size_t)env * d.hist_env_bytes + (size_t)(v * fs) * d.ohw_p : nullptr;
    uint8_t* stg = stage.data() + (size_t)v * ohw * fs;
    if (fs > 1 && !init)
      for (int slot = 0; slot < fs; ++slot) {
        if (slot == hpos) continue;
        int f = slot - hpos - 1; if (f < 0) f += fs;
        for (int c = 0; c < ohw; ++c) stg[(size_t)c * fs + f] = hrow[(size_t)slot * d.ohw_p + c];
      }
    const int hc = r.alive[v] ? r.head[v] : 0;
    int r0 = 0, c0 = 0;
    if (d.V > 0) { r0 = hc / d.W - d.V; c0 = hc % d.W - d.V; }
    for (int cell = 0; cell < ohw; ++cell) {
      const int rr = r0 + cell / d.ow, cc = c0 + cell % d.ow;
      uint32_t bits = 0;
      if (rr >= 0 && rr < d.H && cc >= 0 && cc < d.W) bits = cell_bits(cell_code(d, r.grid[rr * d.W + cc]), (uint32_t)v);
      if (fs == 1) stg[cell] = (uint8_t)bits;
      else if (!init) { stg[(size_t)cell * fs + fs - 1] = (uint8_t)bits; hrow[(size_t)hpos * d.ohw_p + cell] = (uint8_t)bits; }
      else for (int f = 0; f < fs; ++f) { stg[(size_t)cell * fs + f] = (uint8_t)bits; hrow[(size_t)f * d.ohw_p + cell] = (uint8_t)bits; }
    }
  }
  if (fs > 1) r.hdr->hpos = (uint8_t)(init ? 0u : ((uint32_t)r.hdr->hpos + 1u) % (uint32_t)fs);
  if (obs)
    for (int u = 0; u < d.stage_env_bytes; ++u) {
      const uint32_t b = stage[u];
      const uint32_t lo = spread4(b & 15u), hi = spread4(b >> 4);
      memcpy(obs + (size_t)u * 8, &lo, 4);
      memcpy(obs + (size_t)u * 8 + 4, &hi, 4);
    }
}

extern "C" {

void* hs_create_map(const snk_config* c, const uint8_t* walls);
void* hs_create(const snk_config* c) { return hs_create_map(c, nullptr); }
void* hs_create_map(const snk_config* c, const uint8_t* walls) {
  HostSim* h = new HostSim();
  if (walls) { h->walls.resize((size_t)c->height * c->width); for (size_t i = 0; i < h->walls.size(); ++i) h->walls[i] = walls[i] ? WALL : EMPTY; }
  const uint8_t* wm = walls ? h->walls.data() : nullptr;
  Dims& d = h->d;
  memset(&d, 0, sizeof d);
  d.N = c->num_envs; d.H = c->height; d.W = c->width; d.ns = c->num_snakes; d.K = c->snake_length;
  d.V = c->vision_range; d.fs = c->frame_stack;
  d.nfruits = c->num_fruits < 0 ? (int)(c->num_snakes * 0.8 + 0.5) : c->num_fruits;
  d.auto_reset = c->auto_reset; d.done_mode = c->done_mode; d.rng_mode = c->rng_mode; d.observer = c->observer;
  d.dig = default_dig(d.ns);
  d.seed_lo = (uint32_t)c->seed; d.seed_hi = (uint32_t)(c->seed >> 32);
  d.env_off_lo = (uint32_t)c->env_id_offset; d.env_off_hi = (uint32_t)(c->env_id_offset >> 32);
  d.r_fruit = c->reward_fruit; d.r_kill = c->reward_kill; d.r_lose = c->reward_lose;
  d.r_win = c->reward_win; d.r_time = c->reward_time; d.max_steps = c->max_episode_steps;
  finalize_layout(d);
  const int64_t n = spawn_enumerate(d.H, d.W, d.K, nullptr, nullptr, 0, 0, wm);
  h->spawn.resize((size_t)n);
  spawn_enumerate(d.H, d.W, d.K, h->spawn.data(), nullptr, n, 0, wm);
  d.n_cand = (uint32_t)n;
  h->recs.assign((size_t)d.N * d.rec_bytes, 0);
  h->hist.assign((size_t)d.N * d.hist_env_bytes + 16, 0);
  h->scr.assign((size_t)d.scr_bytes + 16, 0);
  h->replay_off.assign((size_t)d.N + 1, 0);
  return h;
}
void hs_destroy(void* p) { delete (HostSim*)p; }
int hs_rec_bytes(void* p) { return ((HostSim*)p)->d.rec_bytes; }
uint32_t hs_errors(void* p) { return ((HostSim*)p)->err; }

void hs_set_replay(void* p, const int32_t* draws, const int64_t* off) {
  HostSim* h = (HostSim*)p;
  h->replay_off.assign(off, off + h->d.N + 1);
  h->replay.assign(draws, draws + off[h->d.N]);
  for (int e = 0; e < h->d.N; ++e) rec_view(h->recs.data() + (size_t)e * h->d.rec_bytes, h->d).hdr->cursor = 0;
}

void hs_reset(void* p, uint8_t* obs) {
  HostSim* h = (HostSim*)p;
  const Dims& d = h->d;
  for (int e = 0; e < d.N; ++e) {
    Rec r = rec_view(h->recs.data() + (size_t)e * d.rec_bytes, d);
    r.hdr->event += 1;
    reset_env(*h, r, (uint32_t)e);
    encode_env(*h, r, (uint32_t)e, true, obs ? obs + (size_t)e * d.obs_env_bytes : nullptr);
  }
}

void hs_step(void* p, const uint8_t* actions, uint8_t* obs, double* rew, uint8_t* done, uint8_t* fin,
             int32_t* rank, double* ep_scores, int32_t* ep_steps, int32_t* ep_fruits, int32_t* ep_kills) {
  HostSim* h = (HostSim*)p;
  const Dims& d = h->d;
  const int ns = d.ns;
  for (int e = 0; e < d.N; ++e) {
    uint8_t* base = h->recs.data() + (size_t)e * d.rec_bytes;
    Rec r = rec_view(base, d);
    r.hdr->event += 1;
    const StepResult res = env_step_logic(d, base, h->scr.data(), actions + (size_t)e * ns, rew + (size_t)e * ns,
                                          done + (size_t)e * ns, &h->err);
    if (fin) fin[e] = res.finished;
    bool do_reset = false;
    if (res.finished) {
      for (int i = 0; i < ns; ++i) {
        const size_t o = (size_t)e * ns + i;
        if (rank) rank[o] = competition_rank(r.score, ns, i);
        if (ep_scores) ep_scores[o] = r.score[i];
        if (ep_steps) ep_steps[o] = (int32_t)cnt_get(d, r, CNT_STEPS, i);
        if (ep_fruits) ep_fruits[o] = (int32_t)cnt_get(d, r, CNT_FRUITS, i);
        if (ep_kills) ep_kills[o] = (int32_t)cnt_get(d, r, CNT_KILLS, i);
      }
      for (int i = 0; i < ns; ++i) stats_zero(d, r, i);
      do_reset = d.auto_reset != 0;
    }
    if (res.fruit_taken) place_fruits(*h, r, (uint32_t)e, res.fruit_taken, DRAW_STEP_FRUIT);
    if (do_reset) reset_env(*h, r, (uint32_t)e);
    encode_env(*h, r, (uint32_t)e, do_reset, obs ? obs + (size_t)e * d.obs_env_bytes : nullptr);
  }
}

void hs_get_grid(void* p, uint8_t* grid, int32_t* counter, int32_t* cursor) {
  HostSim* h = (HostSim*)p;
  const Dims& d = h->d;
  for (int e = 0; e < d.N; ++e) {
    Rec r = rec_view(h->recs.data() + (size_t)e * d.rec_bytes, d);
    for (int c = 0; c < d.HW; ++c) grid[(size_t)e * d.HW + c] = (uint8_t)cell_code(d, r.grid[c]);
    if (counter) counter[e] = r.hdr->alive_counter;
    if (cursor) cursor[e] = (int32_t)r.hdr->cursor;
  }
}

// scenario import (same contract as snk_set_state)
void hs_set_state(void* p, const uint8_t* grid, const uint8_t* alive, const uint8_t* dir, const int32_t* length,
                  const int32_t* cells, int max_cells, const int32_t* counter, const int32_t* ep_len, uint8_t* obs) {
  HostSim* h = (HostSim*)p;
  const Dims& d = h->d;
  for (int e = 0; e < d.N; ++e) {
    Rec r = rec_view(h->recs.data() + (size_t)e * d.rec_bytes, d);
    memcpy(r.grid, grid + (size_t)e * d.HW, (size_t)d.HW);
    for (int i = 0; i < d.ns; ++i) {
      const size_t o = (size_t)e * d.ns + i;
      const int32_t* cl = cells + o * max_cells;
      const int len = length[o];
      r.alive[i] = alive[o]; r.dir[i] = dir[o]; r.len[i] = (uint16_t)(alive[o] ? len : 0);
      r.head[i] = 0; r.tail[i] = 0;
      if (alive[o] && len > 0) {
        r.head[i] = (uint16_t)cl[0]; r.tail[i] = (uint16_t)cl[len - 1];
        for (int k = 1; k < len; ++k) {
          const int diff = cl[k - 1] - cl[k];
          set_body_dir(d, r, cl[k], diff == -d.W ? 0 : diff == 1 ? 1 : diff == d.W ? 2 : 3);
        }
      }
      stats_zero(d, r, i);
    }
    r.hdr->alive_counter = (int16_t)counter[e];
    r.hdr->episode_length = (uint32_t)ep_len[e];
    r.hdr->hpos = 0;
    encode_env(*h, r, (uint32_t)e, true, obs ? obs + (size_t)e * d.obs_env_bytes : nullptr);
  }
}

}  // extern "C"
