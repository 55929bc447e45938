// snk_spawn.cpp -- host-side enumeration of the spawn poses.
//
// The reference recomputes, on every reset, every self-avoiding path of `snake_length` empty cells
// (dfs_sweep_empty, core/grid_util.py:73-99) and samples from that list.  The list depends only on
// (height, width, snake_length), so it is built once per handle here, in the reference's order:
// start cells row-major; neighbours tried as (dr,dc) = (0,1),(1,0),(0,-1),(-1,0) (grid_util.py:7-11);
// an extension is refused when it would leave the FIRST cell (the head) with all four neighbours
// wall / path / the cell being appended (_head_blocked, :102-110).
#include <stdint.h>

#include <vector>

#include "snk_kernels.h"

namespace snk {
namespace {

struct Walker {
  int H, W, K;
  std::vector<int> path;       // flat cells
  std::vector<int> links;      // direction codes (0 UP 1 RIGHT 2 DOWN 3 LEFT) between consecutive cells
  uint64_t* packed; int32_t* cells; int64_t cap; int64_t n = 0;
  int64_t limit = 0;           // > 0: stop enumerating once this many poses were seen (the count explodes with length)
  const uint8_t* walls = nullptr;   // [H*W], nonzero = not empty; null = the walled box

  bool free_cell(int r, int c) const {
    if (walls) return r >= 0 && r < H && c >= 0 && c < W && walls[r * W + c] == 0;
    return r > 0 && r < H - 1 && c > 0 && c < W - 1;
  }
  bool on_path(int cell) const {
    for (int p : path) if (p == cell) return true;
    return false;
  }
  bool head_boxed(int extra) const {
    static const int DR[4] = {0, 1, 0, -1}, DC[4] = {1, 0, -1, 0};
    const int r0 = path[0] / W, c0 = path[0] % W;
    for (int s = 0; s < 4; ++s) {
      const int r = r0 + DR[s], c = c0 + DC[s], cell = r * W + c;
      if (free_cell(r, c) && !on_path(cell) && cell != extra) return false;
    }
    return true;
  }
  void emit() {
    if (n < cap) {
      if (packed) {
        uint64_t e = (uint64_t)path[0];
        for (int j = 1; j < K; ++j) e |= (uint64_t)links[j - 1] << (16 + 2 * (j - 1));
        packed[n] = e;
      }
      if (cells) for (int j = 0; j < K; ++j) cells[n * K + j] = path[j];
    }
    ++n;
  }
  void grow() {
    if (limit > 0 && n > limit) return;
    if ((int)path.size() == K) { emit(); return; }
    // DFS shift order (0,1),(1,0),(0,-1),(-1,0) expressed as direction codes RIGHT, DOWN, LEFT, UP
    static const int DR[4] = {0, 1, 0, -1}, DC[4] = {1, 0, -1, 0}, CODE[4] = {1, 2, 3, 0};
    const int r = path.back() / W, c = path.back() % W;
    for (int s = 0; s < 4; ++s) {
      const int nr = r + DR[s], nc = c + DC[s], cell = nr * W + nc;
      if (!free_cell(nr, nc) || on_path(cell)) continue;
      if (head_boxed(cell)) continue;
      path.push_back(cell); links.push_back(CODE[s]);
      grow();
      path.pop_back(); links.pop_back();
    }
  }
};

}  // namespace

// Returns the number of poses; fills up to `cap` entries of whichever outputs are non-null.  With limit > 0
// the walk stops soon after `limit` poses (the return value is then > limit, not the true count).
int64_t spawn_enumerate(int H, int W, int K, uint64_t* packed_out, int32_t* cells_out, int64_t cap, int64_t limit,
                        const uint8_t* walls) {
  Walker w{H, W, K, {}, {}, packed_out, cells_out, cap};
  w.limit = limit;
  w.walls = walls;
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c)
      if (w.free_cell(r, c)) { w.path.assign(1, r * W + c); w.links.clear(); w.grow(); }
  return w.n;
}

}  // namespace snk
