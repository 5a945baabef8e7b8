// In-situ rate of the solver's tile loops on ONE SM: tile_dmma (k-loop product), the progressive
// panel loop and the fold, with operands in shared memory exactly as in chol_dataflow_kernel.
//   tile_bench        prints clocks per call against the FP64 tensor-pipe bound (64 FMA/clk/SM)
#include "../../pysfm_b200/csrc/ba_solve.cu"
#include <cstdio>
using namespace ba;
__device__ __forceinline__ int row_block_of_warp(int w) { return w < 4 ? w : 11 - w; }   // pairs (0,7)(1,6)(2,5)(3,4) per scheduler

// row-block update with the row block as a COMPILE-TIME constant: no predicate, no branch in the loop
template <int R>
__device__ __forceinline__ void diag_rows_dmma_t(RowTiles& W, const double* __restrict__ P, int lane, int m_lo = 0, int m_hi = NB) {
  const int g = lane >> 2, t4 = lane & 3;
  const double* p = P + t4 * LDT + g;
#pragma unroll 2
  for (int m0 = m_lo; m0 < m_hi; m0 += 4) {
    const double* pm = p + m0 * LDT;
    double b[R + 1];
#pragma unroll
    for (int c = 0; c <= R; ++c) b[c] = pm[8 * c];
    const double a = -b[R];
#pragma unroll
    for (int c = 0; c <= R; ++c) dmma884(W.t[c][0], W.t[c][1], a, b[c]);
  }
}
__device__ __forceinline__ void diag_rows_dmma_sw(RowTiles& W, const double* __restrict__ P, int r, int lane) {
  switch (r) {
    case 0: diag_rows_dmma_t<0>(W, P, lane); break;
    case 1: diag_rows_dmma_t<1>(W, P, lane); break;
    case 2: diag_rows_dmma_t<2>(W, P, lane); break;
    case 3: diag_rows_dmma_t<3>(W, P, lane); break;
    case 4: diag_rows_dmma_t<4>(W, P, lane); break;
    case 5: diag_rows_dmma_t<5>(W, P, lane); break;
    case 6: diag_rows_dmma_t<6>(W, P, lane); break;
    default: diag_rows_dmma_t<7>(W, P, lane); break;
  }
}

__global__ void __launch_bounds__(256, 1) k_tile(double* out, long long* clk, int mode, int reps) {
  extern __shared__ __align__(16) double sm[];
  double* P = sm; double* Q = sm + kTileDoubles; double* Ls = sm + 2 * kTileDoubles;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, gq = lane >> 2, t4 = lane & 3;
  for (int e = tid; e < 3 * kTileDoubles; e += 256) sm[e] = 1e-3 * ((e * 7) % 13);
  __syncthreads();
  const int R0 = 32 * (wid >> 2), C0 = 16 * (wid & 3);
  Frag acc; acc.zero();
  RowTiles W;
  for (int c = 0; c < 8; ++c) W.t[c][0] = W.t[c][1] = 0.0;
  double e0s = 0, e1s = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    if (mode == 0) tile_dmma<true>(acc, P, Q, R0, C0, lane);
    if (mode == 1) diag_rows_dmma(W, P, row_block_of_warp(wid), lane);
    if (mode == 2) { tile_dmma<true>(acc, P, Q, R0, C0, lane); diag_rows_dmma(W, P, row_block_of_warp(wid), lane); }
    if (mode == 3) {   // progressive panel loop, all 8 blocks
      for (int q = 0; q < 8; ++q) {
        const double* pa = P + t4 * LDT + 8 * wid + gq;
        const double* pb = Q + t4 * LDT + 8 * q + gq;
        double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
        double a0 = pa[0], a1 = pa[4 * LDT], b0 = pb[0], b1 = pb[4 * LDT];
        const int kend = 8 * q + 8;
#pragma unroll 1
        for (int m0 = 8; m0 < kend; m0 += 8) {
          const double na0 = pa[m0 * LDT], na1 = pa[(m0 + 4) * LDT];
          const double nb0 = pb[m0 * LDT], nb1 = pb[(m0 + 4) * LDT];
          dmma884(e0, e1, a0, b0);
          dmma884(f0, f1, a1, b1);
          a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
        }
        dmma884(e0, e1, a0, b0);
        dmma884(f0, f1, a1, b1);
        e0 += f0; e1 += f1;
        const int col = 8 * q + 2 * t4, row = 8 * wid + gq;
        Ls[col * LDT + row] = e0;
        Ls[(col + 1) * LDT + row] = e1;
        e0s += e0; e1s += e1;
      }
    }
    if (mode == 5) {   // progressive panel loop, tiles in pairs (q, q+1): four accumulator chains
      for (int q = 0; q < 8; q += 2) {
        const double* pa = P + t4 * LDT + 8 * wid + gq;
        const double* pb = Q + t4 * LDT + 8 * q + gq;
        double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
        const int kend = 8 * q + 8;
#pragma unroll 2
        for (int m0 = 0; m0 < kend; m0 += 8) {
          const double a0 = pa[m0 * LDT], a1 = pa[(m0 + 4) * LDT];
          const double b0 = pb[m0 * LDT], b1 = pb[(m0 + 4) * LDT];
          const double c0 = pb[m0 * LDT + 8], c1 = pb[(m0 + 4) * LDT + 8];
          dmma884(e0, e1, a0, b0);
          dmma884(g0, g1, a0, c0);
          dmma884(f0, f1, a1, b1);
          dmma884(h0, h1, a1, c1);
        }
        {
          const double a0 = pa[kend * LDT], a1 = pa[(kend + 4) * LDT];
          const double c0 = pb[kend * LDT + 8], c1 = pb[(kend + 4) * LDT + 8];
          dmma884(g0, g1, a0, c0);
          dmma884(h0, h1, a1, c1);
        }
        e0 += f0; e1 += f1; g0 += h0; g1 += h1;
        const int col = 8 * q + 2 * t4, row = 8 * wid + gq;
        Ls[col * LDT + row] = e0;
        Ls[(col + 1) * LDT + row] = e1;
        Ls[(col + 8) * LDT + row] = g0;
        Ls[(col + 9) * LDT + row] = g1;
        e0s += e0 + g0; e1s += e1 + g1;
      }
    }
    if (mode == 6) diag_rows_dmma_sw(W, P, row_block_of_warp(wid), lane);
    if (mode == 7) diag_rows_dmma_sw(W, P, wid, lane);
    if (mode == 4) {   // fold of all 8 column blocks
      const double* q0 = Ls + t4 * LDT + gq;
      const int r = row_block_of_warp(wid);
#pragma unroll 2
      for (int m0 = 0; m0 < 64; m0 += 4) {
        const double* q = q0 + m0 * LDT;
        const double av = -q[8 * r];
        double bv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) bv[c] = q[8 * c];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= r) dmma884(W.t[c][0], W.t[c][1], av, bv[c]);
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) clk[0] = t1 - t0;
  double s = e0s + e1s;
  for (int c = 0; c < 8; ++c) s += W.t[c][0] + W.t[c][1];
  for (int mi = 0; mi < 4; ++mi) for (int ni = 0; ni < 2; ++ni) s += acc.v[mi][ni][0] + acc.v[mi][ni][1];
  out[blockIdx.x * 256 + tid] = s;
}

// ---- what limits ONE warp: accumulators x operand source -------------------------------------
// NACC independent accumulators per k-step, operands either fixed registers or loaded from shared
// memory every step (as in the tile loops); launched with 1, 2 or 8 warps.
template <int NACC, bool LDS>
__global__ void __launch_bounds__(256, 1) k_issue(double* out, long long* clk, int reps) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, gq = lane >> 2, t4 = lane & 3;
  for (int e = tid; e < 2 * kTileDoubles; e += blockDim.x) sm[e] = 1e-3 * ((e * 7) % 13);
  __syncthreads();
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c0[i] = c1[i] = 0.0;
  const double* p = sm + t4 * LDT + gq + 8 * (wid & 7);
  double a = p[0], b[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) b[i] = p[8 * (i & 7) + 1];
  const long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
#pragma unroll 4
    for (int m0 = 0; m0 < 64; m0 += 4) {
      if (LDS) {
        a = p[m0 * LDT];
#pragma unroll
        for (int i = 0; i < NACC; ++i) b[i] = p[m0 * LDT + 8 * (i & 7) + kTileDoubles * (i >> 3)];
      }
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b[i]);
    }
  }
  const long long t1 = clock64();
  if (tid == 0) clk[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[tid] = s;
}
template <int NACC, bool LDS>
void run_issue(double* out, long long* clk, int warps) {
  const int reps = 50;
  cudaFuncSetAttribute(k_issue<NACC, LDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTileDoubles * 8);
  k_issue<NACC, LDS><<<1, 32 * warps, 2 * kTileDoubles * 8>>>(out, clk, reps);
  k_issue<NACC, LDS><<<1, 32 * warps, 2 * kTileDoubles * 8>>>(out, clk, reps);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("issue: %d warp(s), %2d accumulators, operands from %s: %6.1f clk per DMMA per warp  (%s)\n", warps, NACC, LDS ? "smem" : "regs",
         (double)h / reps / 16 / NACC, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&clk, 8);
  cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kTileDoubles * 8);
  const char* names[] = {"tile_dmma 64^3 (ideal 4096)", "diag_rows_dmma lower half (ideal 2304)", "both (ideal 6400)", "progressive panel, 8 blocks (ideal 2304)", "fold, 8 blocks (ideal 2304)", "progressive panel in tile pairs (ideal 2304)", "diag rows, row block compile-time, pairs (0,7)..(3,4) (ideal 2304)", "diag rows, row block compile-time, r = warp (ideal 3072)"};
  for (int mode = 0; mode < 8; ++mode) {
    const int reps = 50;
    k_tile<<<1, 256, 3 * kTileDoubles * 8>>>(out, clk, mode, reps);
    k_tile<<<1, 256, 3 * kTileDoubles * 8>>>(out, clk, mode, reps);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("mode %d  %-62s %8.0f clk per call  (%s)\n", mode, names[mode], (double)h / reps, cudaGetErrorString(cudaGetLastError()));
  }
  for (int warps : {1, 2, 4, 8}) {
    run_issue<1, false>(out, clk, warps); run_issue<2, false>(out, clk, warps); run_issue<4, false>(out, clk, warps); run_issue<8, false>(out, clk, warps);
    run_issue<1, true>(out, clk, warps); run_issue<2, true>(out, clk, warps); run_issue<4, true>(out, clk, warps); run_issue<8, true>(out, clk, warps);
    
  }
  return 0;
}
