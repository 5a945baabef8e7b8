// Reduced camera system:  solve_motion_normal_eqns  (bundle_adjuster.py:281-312).
//
// The reference flattens S (nc',nc',6,6) to a dense (6nc' x 6nc') matrix, drops masked
// rows/columns and calls numpy.linalg.solve (LU).  S is symmetric positive definite whenever
// the damped normal equations are, so this path factors it with a tiled FP64 Cholesky
// instead; a non-positive pivot is the "ill-conditioned" signal (the reference's
// LinAlgError -> NormalEquationsIllconditioned, :302-305).
//
// Storage: the elimination kernel accumulated S[row*ld + col] for row <= col, which read
// column-major is the LOWER triangle of the same symmetric matrix: A(i,j) = sys[j*ld + i],
// i >= j.  ld is a multiple of the 64-wide tile; finalize_system_kernel puts ones on the
// padded diagonal (and on masked parameters) so the factorisation needs no edge cases.
// The right-hand side rides along as one extra row of every panel, so the forward
// substitution costs no extra pass; the backward substitution is one dataflow kernel.
#include <cuda_runtime.h>

#include "ba_context.h"

namespace ba {

constexpr int NB = kSolveTile;  // 64

// ------------------------------------------------------------------------------------------
// Mask / padding: a frozen or padded parameter p gets row/column p = e_p and rhs[p] = 0, which
// leaves the solution of the free parameters untouched and yields dC[p] = 0 (:296-309).
__global__ void finalize_system_kernel(double* __restrict__ sys, int ld, int n_sys,
                                       const unsigned char* __restrict__ mask, bool have_mask) {
  const int j = blockIdx.x;  // column
  double* col = sys + (size_t)j * ld;
  const bool free_j = j < n_sys && (!have_mask || mask[j]);
  if (!free_j) {
    for (int i = j + threadIdx.x; i < ld; i += blockDim.x) col[i] = (i == j) ? 1.0 : 0.0;
    if (threadIdx.x == 0) sys[(size_t)ld * ld + j] = 0.0;
  } else if (have_mask) {
    for (int i = j + threadIdx.x; i < n_sys; i += blockDim.x)
      if (!mask[i]) col[i] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// Diagonal tile: in-place Cholesky of A_kk (64x64, lower) in shared memory.
// Unscaled right-looking sweep (one barrier per column): after step j the columns still hold
// L[r][j]*L[j][j]; the scaling by 1/L[j][j] happens once at the end.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ sys, int ld, int k,
                                                        double* __restrict__ status) {
  __shared__ double T[NB][NB + 1];
  __shared__ int s_bad;
  double* Akk = sys + (size_t)(k * NB) * ld + k * NB;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < NB * NB; e += 256) {
    const int c = e / NB, r = e % NB;
    T[r][c] = (r >= c) ? Akk[(size_t)c * ld + r] : 0.0;
  }
  const int tr = tid & 15, tc = tid >> 4;
  for (int j = 0; j < NB; ++j) {
    __syncthreads();
    const double d = T[j][j];
    if (!(d > 0.0)) {  // also catches NaN
      if (tid == 0) s_bad = 1;
      break;           // uniform: every thread reads the same d
    }
    const double id = 1.0 / d;
    for (int c = j + 1 + tc; c < NB; c += 16) {
      const double lc = T[c][j] * id;
      for (int r = c + tr; r < NB; r += 16) T[r][c] -= T[r][j] * lc;
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) *status = 1.0;
    // leave an identity factor behind so later kernels stay finite
    for (int e = tid; e < NB * NB; e += 256) {
      const int c = e / NB, r = e % NB;
      if (r >= c) Akk[(size_t)c * ld + r] = (r == c) ? 1.0 : 0.0;
    }
    return;
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int c = e / NB, r = e % NB;
    if (r >= c) {
      const double s = sqrt(T[c][c]);
      Akk[(size_t)c * ld + r] = (r == c) ? s : T[r][c] / s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Panel: X = B L_kk^{-T} for every row tile below the diagonal (blockIdx.x < ntiles) and for
// the right-hand-side row (blockIdx.x == ntiles, a 1 x 64 "tile": y_k^T = b_k^T L_kk^{-T}).
// Column sweep: x[:,c] = b[:,c]/L[c][c]; b[:,c+1:] -= x[:,c] (x) L[c+1:,c].
__global__ void __launch_bounds__(256) chol_trsm_kernel(double* __restrict__ sys, int ld, int k,
                                                        int ntiles) {
  extern __shared__ double sm_trsm[];
  double (*L)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_trsm);                    // L[r][c]
  double (*B)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_trsm + NB * (NB + 1));    // B[c][r]
  const int tid = threadIdx.x;
  const double* Lkk = sys + (size_t)(k * NB) * ld + k * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int c = e / NB, r = e % NB;
    L[r][c] = (r >= c) ? Lkk[(size_t)c * ld + r] : 0.0;
  }
  const bool is_rhs = (int)blockIdx.x == ntiles;
  double* rhs = sys + (size_t)ld * ld + k * NB;
  double* Bik = sys + (size_t)(k * NB) * ld + (size_t)(k + 1 + blockIdx.x) * NB;
  const int rows = is_rhs ? 1 : NB;
  if (is_rhs) {
    if (tid < NB) B[tid][0] = rhs[tid];
  } else {
    for (int e = tid; e < NB * NB; e += 256) {
      const int c = e / NB, r = e % NB;
      B[c][r] = Bik[(size_t)c * ld + r];
    }
  }
  // thread layout for the rank-1 updates: r = tid % 64, column group = tid / 64 (4 groups)
  const int r = tid & 63, cg = tid >> 6;
  for (int c = 0; c < NB; ++c) {
    __syncthreads();
    if (r < rows) {
      const double x = B[c][r] / L[c][c];
      for (int c2 = c + 1 + cg; c2 < NB; c2 += 4) B[c2][r] -= x * L[c2][c];
    }
  }
  __syncthreads();
  // final scaling: column c of the result is B[c][:] / L[c][c] (each thread above only used
  // the quotient locally so that no barrier separates read and write of B[c][r])
  if (is_rhs) {
    if (tid < NB) rhs[tid] = B[tid][0] / L[tid][tid];
  } else {
    for (int e = tid; e < NB * NB; e += 256) {
      const int c = e / NB, rr = e % NB;
      Bik[(size_t)c * ld + rr] = B[c][rr] / L[c][c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Trailing update after panel k:  A_ij -= L_ik L_jk^T for k < j <= i  (tile list flattened in
// blockIdx.x), plus the rhs row  b_j -= L_jk y_k  (blockIdx.x >= ntri).
__global__ void __launch_bounds__(256) chol_update_kernel(double* __restrict__ sys, int ld, int k,
                                                          int nrem, int ntri) {
  extern __shared__ double sm[];
  double* Li = sm;              // [m][r]  64 x 64
  double* Lj = sm + NB * NB;    // [m][c]
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= ntri) {
    // rhs tile: b_j -= L_jk y_k
    const int j = k + 1 + ((int)blockIdx.x - ntri);
    const double* Ljk = sys + (size_t)(k * NB) * ld + (size_t)j * NB;
    double* rhs = sys + (size_t)ld * ld;
    const double* yk = rhs + k * NB;
    // 4 threads per row, each a quarter of the dot product
    const int r = tid >> 2, q = tid & 3;
    double acc = 0.0;
    for (int m = q * 16; m < q * 16 + 16; ++m) acc += Ljk[(size_t)m * ld + r] * yk[m];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0) rhs[j * NB + r] -= acc;
    return;
  }
  // decode (i, j) from the flattened lower-triangular tile index
  int t = blockIdx.x, ti = 0;
  while (t >= ti + 1) { t -= ti + 1; ++ti; }  // ti = row within trailing block, t = col
  const int i = k + 1 + ti, j = k + 1 + t;
  (void)nrem;
  const double* Lik = sys + (size_t)(k * NB) * ld + (size_t)i * NB;
  const double* Ljk = sys + (size_t)(k * NB) * ld + (size_t)j * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int m = e / NB, r = e % NB;
    Li[m * NB + r] = Lik[(size_t)m * ld + r];
    Lj[m * NB + r] = Ljk[(size_t)m * ld + r];
  }
  __syncthreads();
  const int tr = tid & 15, tc = tid >> 4;  // rows 4*tr.., cols 4*tc..
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
  for (int m = 0; m < NB; ++m) {
    const double2 a01 = *reinterpret_cast<const double2*>(Li + m * NB + 4 * tr);
    const double2 a23 = *reinterpret_cast<const double2*>(Li + m * NB + 4 * tr + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(Lj + m * NB + 4 * tc);
    const double2 b23 = *reinterpret_cast<const double2*>(Lj + m * NB + 4 * tc + 2);
    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bv[b];
  }
  double* Aij = sys + (size_t)(j * NB) * ld + (size_t)i * NB;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    double* colp = Aij + (size_t)(4 * tc + b) * ld + 4 * tr;
    double2 v0 = *reinterpret_cast<double2*>(colp);
    double2 v1 = *reinterpret_cast<double2*>(colp + 2);
    v0.x -= acc[0][b]; v0.y -= acc[1][b]; v1.x -= acc[2][b]; v1.y -= acc[3][b];
    *reinterpret_cast<double2*>(colp) = v0;
    *reinterpret_cast<double2*>(colp + 2) = v1;
  }
}

// ------------------------------------------------------------------------------------------
// Backward substitution L^T x = y as one dataflow kernel: CTA b owns tile k = T-1-b, folds in
// every finished x_i (i > k) as soon as its flag is up, then solves its 64x64 triangle.
// CTAs only ever wait on lower block indices, which the hardware dispatches first.
__global__ void __launch_bounds__(256) chol_backsolve_kernel(const double* __restrict__ sys, int ld,
                                                             int T, double* __restrict__ x,
                                                             unsigned int* __restrict__ flags,
                                                             unsigned int epoch) {
  __shared__ double acc[NB];
  __shared__ double part[4][NB];
  __shared__ double Lkk[NB][NB + 1];
  const int tid = threadIdx.x;
  const int k = T - 1 - (int)blockIdx.x;
  const double* y = sys + (size_t)ld * ld;
  if (tid < NB) acc[tid] = y[k * NB + tid];
  const double* Lk = sys + (size_t)(k * NB) * ld + (size_t)k * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int c = e / NB, r = e % NB;
    Lkk[r][c] = (r >= c) ? Lk[(size_t)c * ld + r] : 0.0;
  }
  __syncthreads();
  // acc[c] -= sum_r L_ik[r][c] x_i[r]:  column c of tile (i,k) is contiguous in r
  const int c = tid & 63, q = tid >> 6;
  for (int i = T - 1; i > k; --i) {
    if (tid == 0) {
      while (atomicAdd(&flags[i], 0u) != epoch) { __nanosleep(64); }
      __threadfence();
    }
    __syncthreads();
    const double* Lik = sys + (size_t)(k * NB + c) * ld + (size_t)i * NB + q * 16;
    const double* xi = x + i * NB + q * 16;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += Lik[r] * __ldcg(xi + r);
    part[q][c] = s;
    __syncthreads();
    if (tid < NB) acc[tid] -= part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
    __syncthreads();
  }
  // triangular solve L_kk^T x_k = acc, column sweep from the bottom
  for (int r = NB - 1; r >= 0; --r) {
    __syncthreads();
    if (tid == 0) acc[r] = acc[r] / Lkk[r][r];
    __syncthreads();
    if (tid < r) acc[tid] -= Lkk[r][tid] * acc[r];
  }
  __syncthreads();
  if (tid < NB) x[k * NB + tid] = acc[tid];
  __threadfence();
  __syncthreads();
  if (tid == 0) atomicExch(&flags[k], epoch);
}

// ------------------------------------------------------------------------------------------
static unsigned int g_epoch_seed = 1;

cudaError_t launch_solve(Context& c, bool have_mask, cudaStream_t st) {
  const int ld = c.ld, T = ld / NB;
  cudaError_t e;
  if ((e = cudaMemsetAsync(&c.scalars->status, 0, sizeof(double), st)) != cudaSuccess) return e;
  finalize_system_kernel<<<ld, 128, 0, st>>>(c.sys, ld, c.n_sys, c.cam_mask, have_mask);
  c.launches += 1;
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  2 * NB * NB * (int)sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  2 * NB * (NB + 1) * (int)sizeof(double))) != cudaSuccess) return e;
    attr_set = true;
  }
  for (int k = 0; k < T; ++k) {
    chol_diag_kernel<<<1, 256, 0, st>>>(c.sys, ld, k, &c.scalars->status);
    const int nrem = T - k - 1;
    chol_trsm_kernel<<<nrem + 1, 256, 2 * NB * (NB + 1) * sizeof(double), st>>>(c.sys, ld, k, nrem);
    c.launches += 2;
    if (nrem > 0) {
      const int ntri = nrem * (nrem + 1) / 2;
      chol_update_kernel<<<ntri + nrem, 256, 2 * NB * NB * sizeof(double), st>>>(c.sys, ld, k, nrem, ntri);
      c.launches += 1;
    }
  }
  // flags live after the two tickets in c.counters; a fresh epoch per call avoids a memset
  const unsigned int epoch = ++g_epoch_seed;
  chol_backsolve_kernel<<<T, 256, 0, st>>>(c.sys, ld, T, c.dC, c.counters + 8, epoch);
  c.launches += 1;
  return cudaGetLastError();
}

}  // namespace ba
