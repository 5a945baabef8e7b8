// Host-only check of the tile enumeration of the tcgen05 trailing update (ba_solve_tc.cuh):
// decode_tile / count_tiles must enumerate, exactly once each, the 128 x 64 tiles that touch the
// lower triangle of a trailing matrix of n_nb 64-wide column blocks.  Built and run by
// tests/test_tc_host_logic.py (no GPU involved).
#include <cuda_runtime.h>
#include <cstdio>
#include <set>
#include <utility>
#include "../../pysfm_b200/csrc/ba_solve_tc.cuh"

int main() {
  for (int n_nb = 1; n_nb <= 600; ++n_nb) {
    const int ntiles = ba::tc::count_tiles(n_nb);
    std::set<std::pair<int, int>> seen;
    for (int t = 0; t < ntiles; ++t) {
      int mbl, nbl;
      ba::tc::decode_tile(t, mbl, nbl);
      // the tile covers rows 128 mbl .. +127 and columns 64 nbl .. +63 of the trailing matrix
      if (mbl < 0 || nbl < 0 || nbl >= n_nb || 128 * mbl >= 64 * n_nb) { printf("n_nb %d: tile %d out of range (%d, %d)\n", n_nb, t, mbl, nbl); return 1; }
      if (64 * nbl > 128 * mbl + 127) { printf("n_nb %d: tile %d (%d, %d) lies above the diagonal\n", n_nb, t, mbl, nbl); return 1; }
      if (!seen.insert(std::make_pair(mbl, nbl)).second) { printf("n_nb %d: tile (%d, %d) twice\n", n_nb, mbl, nbl); return 1; }
    }
    // every 64 x 64 block (i >= j) of the lower triangle is covered by exactly one tile
    for (int i = 0; i < n_nb; ++i)
      for (int j = 0; j <= i; ++j)
        if (!seen.count(std::make_pair(i / 2, j))) { printf("n_nb %d: block (%d, %d) not covered\n", n_nb, i, j); return 1; }
  }
  // instruction descriptor: kind::i8, D = S32, A and B signed 8-bit, K-major, M = 128, N = 64 .. 256
  if (ba::tc::instr_desc(64) != ((2u << 4) | (1u << 7) | (1u << 10) | (8u << 17) | (8u << 24))) { printf("instr_desc(64)\n"); return 1; }
  if (ba::tc::instr_desc(256) != ((2u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24))) { printf("instr_desc(256)\n"); return 1; }
  // pipeline depth that fits the 227 KB of shared memory: 3 stages of 72 KB at 6 slices, 64-byte K steps
  if (ba::tc::syrk_stages<6, 64>() != 3 || ba::tc::syrk_stages<7, 64>() != 2 || ba::tc::syrk_stages<4, 64>() != 4 || ba::tc::syrk_stages<6, 128>() != 1) {
    printf("stages %d %d %d %d\n", ba::tc::syrk_stages<6, 64>(), ba::tc::syrk_stages<7, 64>(), ba::tc::syrk_stages<4, 64>(), ba::tc::syrk_stages<6, 128>());
    return 1;
  }
  printf("ok\n");
  return 0;
}
