// Reduced camera system:  solve_motion_normal_eqns  (bundle_adjuster.py:281-312).
//
// The reference flattens S (nc',nc',6,6) to a dense (6nc' x 6nc') matrix, drops masked
// rows/columns and calls numpy.linalg.solve (LU).  S is symmetric positive definite whenever
// the damped normal equations are, so this path factors it with an FP64 tile Cholesky; a
// non-positive pivot is the "ill-conditioned" signal (the reference's LinAlgError ->
// NormalEquationsIllconditioned, :302-305).
//
// Two kernels per solve:
//
//   expand_system_kernel   packed upper 6x6 blocks (the buffer the elimination kernel
//                          accumulates and ranks all-reduce) -> dense lower-triangular
//                          column-major A with leading dimension ld (multiple of the 64-wide
//                          tile).  Frozen parameters (param_mask, :296-309) and the padding
//                          get unit rows/columns and a zero right-hand side, which leaves
//                          the free parameters' solution untouched and yields dC = 0 there.
//
//   chol_dataflow_kernel   ONE persistent launch for factorisation, forward and backward
//                          substitution.  Work items are 64x64 tiles of the lower triangle in
//                          column-major order, handed out through an atomic ticket; a tile
//                          task is LEFT-LOOKING: it accumulates  A_ij - sum_k L_ik L_jk^T  in
//                          registers, consuming the L tiles of earlier columns as soon as
//                          their ready-flags go up (acquire/release through L2), and is
//                          written exactly once.  Diagonal tasks factor their tile in shared
//                          memory with a Gauss-Jordan sweep that yields L_jj and L_jj^{-1}
//                          together, so every panel tile below is a plain GEMM with the
//                          inverse and the forward substitution y_j = L_jj^{-1}(b_j - ...)
//                          rides along.  When the tile tickets run out the CTAs take the
//                          backward-substitution tasks x_k = L_kk^{-T}(y_k - sum_i L_ik^T x_i)
//                          from a second ticket.  Tickets are handed out in dependency order
//                          and the grid never exceeds the number of co-resident CTAs, so a
//                          waiting CTA always waits on a task that is already running.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_context.h"

namespace ba {

constexpr int NB = kSolveTile;     // 64
constexpr int NBP = NB + 1;        // padded row of the diagonal-tile work arrays
constexpr int kSolveThreads = 256;

// ------------------------------------------------------------------------------------------
// packed block index of (a, b), a <= b < nc:  rows of the upper block triangle back to back
__host__ __device__ __forceinline__ size_t packed_block(int a, int b, int nc) {
  return (size_t)a * nc - (size_t)a * (a - 1) / 2 + (b - a);
}

// A(p, q), p >= q  (column-major lower, A[q*ld + p])  <-  packed block (q/6, p/6)[q%6][p%6]
__global__ void __launch_bounds__(256)
expand_system_kernel(const double* __restrict__ packed, int nc, int n_sys, int ld,
                     const unsigned char* __restrict__ mask, bool have_mask,
                     double* __restrict__ A, double* __restrict__ rhs) {
  const int q = blockIdx.x;  // column
  const bool free_q = q < n_sys && (!have_mask || mask[q]);
  double* col = A + (size_t)q * ld;
  if (!free_q) {
    for (int p = q + threadIdx.x; p < ld; p += blockDim.x) col[p] = (p == q) ? 1.0 : 0.0;
    if (threadIdx.x == 0) rhs[q] = 0.0;
    return;
  }
  const int a = q / 6, rr = q - 6 * a;
  const double* row = packed + packed_block(a, a, nc) * 36 + rr * 6;  // block (a, a), row rr
  for (int p = q + threadIdx.x; p < ld; p += blockDim.x) {
    double v = 0.0;
    if (p < n_sys && (!have_mask || mask[p])) {
      const int b = p / 6, cc = p - 6 * b;
      v = row[(size_t)(b - a) * 36 + cc];
    }
    col[p] = v;
  }
  if (threadIdx.x == 0) rhs[q] = packed[packed_block(nc - 1, nc - 1, nc) * 36 + 36 + q];
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct CholArgs {
  double* __restrict__ A;         // [ld*ld] dense lower, column-major; overwritten by L
  double* __restrict__ rhs;       // [ld] b -> y (forward substitution)
  double* __restrict__ x;         // [ld] solution
  double* __restrict__ LinvT;     // [T][NB*NB]  LinvT[m*NB + c] = (L_jj^{-1})[c][m]
  unsigned int* __restrict__ flags;   // [T*T] tile (i,j) ready == epoch ; [T*T + k] x_k ready
  unsigned int* __restrict__ tickets; // [0] tile tasks, [1] back-substitution tasks
  double* __restrict__ status;    // set to 1 on a non-positive pivot
  int ld, T;
  unsigned int epoch;
};

// Block-wide wait until *f == epoch (thread 0 spins with acquire loads).
__device__ __forceinline__ void wait_flag(const unsigned int* f, unsigned int epoch) {
  if (threadIdx.x == 0) {
    while (ld_acquire(f) != epoch) __nanosleep(20);
  }
}

// Stage the 64x64 tile whose (r, m) element is at src[m*ld + r] into smem dst[m*NB + r].
__device__ __forceinline__ void stage_tile(double* dst, const double* src, int ld) {
  // 64 columns x 512 B; 16 B per cp.async; 2048 chunks / 256 threads = 8 each
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int chunk = it * kSolveThreads + threadIdx.x;
    const int m = chunk >> 5, r2 = (chunk & 31) * 2;
    cp_async16(dst + m * NB + r2, src + (size_t)m * ld + r2);
  }
}

// acc[a][b] += sum_m P[m][4tr+a] * Q[m][4tc+b]
__device__ __forceinline__ void tile_mma(double acc[4][4], const double* __restrict__ P,
                                         const double* __restrict__ Q, int tr, int tc) {
#pragma unroll 8
  for (int m = 0; m < NB; ++m) {
    const double2 a01 = *reinterpret_cast<const double2*>(P + m * NB + 4 * tr);
    const double2 a23 = *reinterpret_cast<const double2*>(P + m * NB + 4 * tr + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(Q + m * NB + 4 * tc);
    const double2 b23 = *reinterpret_cast<const double2*>(Q + m * NB + 4 * tc + 2);
    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bv[b];
  }
}

// Shared memory map (doubles):
//   buf[2][2][NB*NB]   double-buffered operand tiles (P, Q) of the k loop            128 KB
//   aliases used after the k loop of a task:
//     W  [NB][NBP]  = buf            diagonal tile being factored (full symmetric)
//     M  [NB][NBP]  = buf + NB*NBP   running L~^{-1} (strict lower part)
//     Cs [NB*NB]    = buf            panel task: C transposed to [m][r]
//     Bs [NB*NB]    = buf + NB*NB    panel task: LinvT tile [m][c]
//   vec[3][NB]      y_k staging, b_j accumulator, scratch
constexpr int kSolveSmemDoubles = 4 * NB * NB + 4 * NB;
constexpr size_t kSolveSmemBytes = kSolveSmemDoubles * sizeof(double);

__global__ void __launch_bounds__(kSolveThreads, 1) chol_dataflow_kernel(const CholArgs g) {
  extern __shared__ __align__(16) double sm[];
  double* const buf = sm;
  double* const vec = sm + 4 * NB * NB;
  __shared__ int s_task;
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int tr = tid & 15, tc = tid >> 4;
  const int T = g.T, ld = g.ld;
  const int ntasks = T * (T + 1) / 2;
  const unsigned int epoch = g.epoch;

  // ======================================= factorisation ===================================
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = (int)atomicAdd(&g.tickets[0], 1u);
    __syncthreads();
    const int t = s_task;
    if (t >= ntasks) break;
    // column-major enumeration of the lower triangle: column j holds T - j tasks
    int j = 0, rem = t;
    {
      // solve rem < T - j incrementally from a closed-form guess
      const double Tf = (double)T + 0.5;
      j = (int)(Tf - sqrt(Tf * Tf - 2.0 * (double)t));
      if (j < 0) j = 0;
      if (j > T - 1) j = T - 1;
      while (j > 0 && (size_t)j * T - (size_t)j * (j - 1) / 2 > (size_t)t) --j;
      while ((size_t)(j + 1) * T - (size_t)(j + 1) * j / 2 <= (size_t)t) ++j;
      rem = t - (int)((size_t)j * T - (size_t)j * (j - 1) / 2);
    }
    const int i = j + rem;
    const bool diag = (i == j);

    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    double bacc = 0.0;  // diag task, tid < NB: sum_k (L_jk y_k)[tid]

    // ---- k loop: acc += L_ik L_jk^T, operands double-buffered through cp.async ------------
    // stage s of step k: P = buf + (2*(k&1))*NB*NB, Q = P + NB*NB (Q unused when diag)
    auto issue = [&](int k) {
      double* P = buf + (size_t)(2 * (k & 1)) * NB * NB;
      stage_tile(P, g.A + (size_t)(k * NB) * ld + (size_t)i * NB, ld);
      if (!diag) stage_tile(P + NB * NB, g.A + (size_t)(k * NB) * ld + (size_t)j * NB, ld);
      cp_async_commit();
    };
    if (j > 0) {
      wait_flag(&g.flags[(size_t)i * T + 0], epoch);
      if (!diag) wait_flag(&g.flags[(size_t)j * T + 0], epoch);
      __syncthreads();
      issue(0);
    }
    for (int k = 0; k < j; ++k) {
      if (k + 1 < j) {
        wait_flag(&g.flags[(size_t)i * T + k + 1], epoch);
        if (!diag) wait_flag(&g.flags[(size_t)j * T + k + 1], epoch);
        __syncthreads();
        issue(k + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      if (diag && tid < NB) vec[tid] = __ldcg(g.rhs + k * NB + tid);  // y_k
      __syncthreads();
      const double* P = buf + (size_t)(2 * (k & 1)) * NB * NB;
      const double* Q = diag ? P : P + NB * NB;
      tile_mma(acc, P, Q, tr, tc);
      if (diag && tid < NB) {
        double s = 0.0;
#pragma unroll 8
        for (int m = 0; m < NB; ++m) s += P[m * NB + tid] * vec[m];
        bacc += s;
      }
      __syncthreads();
    }

    // ---- C = A_ij - acc ---------------------------------------------------------------------
    const double* Aij = g.A + (size_t)(j * NB) * ld + (size_t)i * NB;
    if (diag) {
      // Only the lower triangle of A_jj is valid in memory: build the full symmetric tile.
      double* W = buf;
      double* M = buf + NB * NBP;
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int r = 4 * tr + a, c = 4 * tc + b;
          if (r >= c) {
            const double v = __ldcg(Aij + (size_t)c * ld + r) - acc[a][b];
            W[r * NBP + c] = v;
            W[c * NBP + r] = v;
          }
          M[r * NBP + c] = 0.0;
        }
      if (tid == 0) s_bad = 0;
      // Gauss-Jordan sweep without pivoting: after step p row r > p of W holds the Schur
      // complement in columns > p, and M (unit lower) accumulates L~^{-1}:  M A = U = D L~^T.
      const int c = tid & 63, rg = tid >> 6;
      for (int p = 0; p < NB; ++p) {
        __syncthreads();
        double d = W[p * NBP + p];
        if (!(d > 0.0)) {          // also catches NaN; uniform across the block
          if (tid == 0) s_bad = 1;
          d = 1.0;
        }
        const double id = 1.0 / d;
        if (c == p) {
          for (int r = p + 1 + rg; r < NB; r += 4) M[r * NBP + p] = -W[r * NBP + p] * id;
        } else if (c > p) {
          const double u = W[p * NBP + c];
          for (int r = p + 1 + rg; r < NB; r += 4) W[r * NBP + c] -= (W[r * NBP + p] * id) * u;
        } else {
          const double u = M[p * NBP + c];
          for (int r = p + 1 + rg; r < NB; r += 4) M[r * NBP + c] -= (W[r * NBP + p] * id) * u;
        }
      }
      __syncthreads();
      if (s_bad) {
        if (tid == 0) *g.status = 1.0;
        // leave an identity factor behind so that dependants stay finite
        for (int e = tid; e < NB * NB; e += kSolveThreads) {
          const int r = e >> 6, cc = e & 63;
          W[r * NBP + cc] = (r == cc) ? 1.0 : 0.0;
          M[r * NBP + cc] = 0.0;
        }
        __syncthreads();
      }
      // scale: L[r][c] = U[c][r] / sqrt(U[c][c]) (r >= c);  Linv[r][c] = M[r][c] / sqrt(U[r][r])
      double* isd = vec + NB;      // 1/sqrt(U[r][r])
      if (tid < NB) {
        isd[tid] = 1.0 / sqrt(W[tid * NBP + tid]);
      }
      __syncthreads();
      double* Ljj = g.A + (size_t)(j * NB) * ld + (size_t)j * NB;
      double* LT = g.LinvT + (size_t)j * NB * NB;
      for (int e = tid; e < NB * NB; e += kSolveThreads) {
        const int cc = e >> 6, r = e & 63;      // r fastest: coalesced column-major store
        if (r >= cc) Ljj[(size_t)cc * ld + r] = W[cc * NBP + r] * isd[cc];
        // LinvT[m = cc][c' = r] = Linv[r][cc]
        double v = 0.0;
        if (r > cc) v = M[r * NBP + cc] * isd[r];
        else if (r == cc) v = isd[r];
        LT[cc * NB + r] = v;
      }
      // forward substitution: y_j = Linv (b_j - bacc);  thread r: sum_m LinvT[m][r] t[m]
      double* tvec = vec + 2 * NB;
      if (tid < NB) tvec[tid] = __ldcg(g.rhs + j * NB + tid) - bacc;
      __syncthreads();
      if (tid < NB) {
        double s = 0.0;
        for (int m = 0; m <= tid; ++m) {
          const double li = (m == tid) ? isd[tid] : M[tid * NBP + m] * isd[tid];
          s += li * tvec[m];
        }
        g.rhs[j * NB + tid] = s;
      }
    } else {
      // ---- panel tile: L_ij = C Linv_jj^T ----------------------------------------------------
      double* Cs = buf;            // [m][r] = C[r][m]
      double* Bs = buf + NB * NB;  // [m][c] = Linv[c][m]
      wait_flag(&g.flags[(size_t)j * T + j], epoch);
      __syncthreads();
      {
        const double* LT = g.LinvT + (size_t)j * NB * NB;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int chunk = it * kSolveThreads + tid;
          cp_async16(Bs + chunk * 2, LT + chunk * 2);
        }
        cp_async_commit();
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double* colp = Aij + (size_t)(4 * tc + b) * ld + 4 * tr;
        const double2 v0 = __ldcg(reinterpret_cast<const double2*>(colp));
        const double2 v1 = __ldcg(reinterpret_cast<const double2*>(colp + 2));
        double* d = Cs + (4 * tc + b) * NB + 4 * tr;
        *reinterpret_cast<double2*>(d) = make_double2(v0.x - acc[0][b], v0.y - acc[1][b]);
        *reinterpret_cast<double2*>(d + 2) = make_double2(v1.x - acc[2][b], v1.y - acc[3][b]);
      }
      cp_async_wait<0>();
      __syncthreads();
      double out[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[a][b] = 0.0;
      tile_mma(out, Cs, Bs, tr, tc);
      double* Lij = g.A + (size_t)(j * NB) * ld + (size_t)i * NB;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double* colp = Lij + (size_t)(4 * tc + b) * ld + 4 * tr;
        *reinterpret_cast<double2*>(colp) = make_double2(out[0][b], out[1][b]);
        *reinterpret_cast<double2*>(colp + 2) = make_double2(out[2][b], out[3][b]);
      }
    }
    // ---- publish ------------------------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(&g.flags[(size_t)i * T + j], epoch);
  }

  // ==================================== backward substitution ==============================
  // x_k = L_kk^{-T} (y_k - sum_{i>k} L_ik^T x_i),  k = T-1 .. 0
  double* const part = buf;            // [8 warps][NB] partial sums
  double* const accv = vec;            // [NB]
  double* const LTs = buf + 8 * NB;    // [NB][NBP] padded copy of LinvT
  const int lane = tid & 31, wid = tid >> 5;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = (int)atomicAdd(&g.tickets[1], 1u);
    __syncthreads();
    const int bt = s_task;
    if (bt >= T) break;
    const int k = T - 1 - bt;
    wait_flag(&g.flags[(size_t)k * T + k], epoch);   // L_kk^{-1} and y_k
    __syncthreads();
    {
      const double* LT = g.LinvT + (size_t)k * NB * NB;
      for (int e = tid; e < NB * NB; e += kSolveThreads) LTs[(e >> 6) * NBP + (e & 63)] = __ldcg(LT + e);
    }
    // warp w owns columns 8w .. 8w+7 of every tile; lanes span rows
    double cs[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cs[q] = 0.0;
    for (int i = T - 1; i > k; --i) {
      wait_flag(&g.flags[(size_t)i * T + k], epoch);       // tile (i, k) of L
      wait_flag(&g.flags[(size_t)T * T + i], epoch);       // x_i
      __syncthreads();
      const double x0 = __ldcg(g.x + i * NB + lane), x1 = __ldcg(g.x + i * NB + 32 + lane);
      const double* Lik = g.A + (size_t)(k * NB + 8 * wid) * ld + (size_t)i * NB;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        cs[q] += __ldcg(Lik + (size_t)q * ld + lane) * x0 + __ldcg(Lik + (size_t)q * ld + 32 + lane) * x1;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double s = warp_sum(cs[q]);
      if (lane == 0) accv[8 * wid + q] = s;
    }
    __syncthreads();
    if (tid < NB) accv[tid] = __ldcg(g.rhs + k * NB + tid) - accv[tid];
    __syncthreads();
    // x_k[c] = sum_m Linv[m][c] acc[m] = sum_m LinvT[c][m] acc[m];  4 partial sums per c
    {
      const int c = tid & 63, q = tid >> 6;
      double s = 0.0;
#pragma unroll 4
      for (int m = q * 16; m < q * 16 + 16; ++m) s += LTs[c * NBP + m] * accv[m];
      part[q * NB + c] = s;
    }
    __syncthreads();
    if (tid < NB) g.x[k * NB + tid] = part[tid] + part[NB + tid] + part[2 * NB + tid] + part[3 * NB + tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(&g.flags[(size_t)T * T + k], epoch);
  }
}

// ------------------------------------------------------------------------------------------
cudaError_t launch_solve(Context& c, bool have_mask, cudaStream_t st) {
  const int ld = c.ld, T = ld / NB;
  cudaError_t e;
  if ((e = cudaMemsetAsync(&c.scalars->status, 0, sizeof(double), st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(c.solve_tickets, 0, 2 * sizeof(unsigned int), st)) != cudaSuccess) return e;
  expand_system_kernel<<<ld, 256, 0, st>>>(c.sys, c.n_opt_cam, c.n_sys, ld, c.cam_mask, have_mask,
                                           c.Adense, c.Adense + (size_t)ld * ld);
  c.launches += 1;
  if (!c.solve_attr_set) {
    if ((e = cudaFuncSetAttribute(chol_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSolveSmemBytes)) != cudaSuccess) return e;
    c.solve_attr_set = true;
  }
  CholArgs g;
  g.A = c.Adense;
  g.rhs = c.Adense + (size_t)ld * ld;
  g.x = c.dC;
  g.LinvT = c.LinvT;
  g.flags = c.solve_flags;
  g.tickets = c.solve_tickets;
  g.status = &c.scalars->status;
  g.ld = ld; g.T = T;
  g.epoch = ++c.solve_epoch;   // a fresh epoch per call: flags never need clearing
  const int ntasks = T * (T + 1) / 2;
  int grid = ntasks < c.num_sms ? ntasks : c.num_sms;   // 1 CTA / SM (128 KB smem): all co-resident
  chol_dataflow_kernel<<<grid, kSolveThreads, kSolveSmemBytes, st>>>(g);
  c.launches += 1;
  return cudaGetLastError();
}

}  // namespace ba
