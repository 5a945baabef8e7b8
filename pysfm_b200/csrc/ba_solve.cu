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

// Optional per-task timeline (tools/microbench/solve_bench.cu defines BA_SOLVE_TRACE).
#ifdef BA_SOLVE_TRACE
#define BA_TRACE_DECL unsigned long long* trace;
#define BA_TRACE(rec, slot)                                                        \
  do {                                                                             \
    if (g.trace && threadIdx.x == 0) {                                             \
      unsigned long long gt__;                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__));                     \
      g.trace[(size_t)(rec) * 8 + (slot)] = gt__;                                  \
    }                                                                              \
  } while (0)
#define BA_TRACE_SET(rec, slot, v)                                                 \
  do {                                                                             \
    if (g.trace && threadIdx.x == 0) g.trace[(size_t)(rec) * 8 + (slot)] = (v);    \
  } while (0)
#else
#define BA_TRACE_DECL
#define BA_TRACE(rec, slot) do { } while (0)
#define BA_TRACE_SET(rec, slot, v) do { } while (0)
#endif

struct CholArgs {
  BA_TRACE_DECL
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

// Shared-memory operand tiles are stored k-major with a padded row of LDT doubles:
// element (row r, contraction index m) of an operand lives at tile[m*LDT + r].  LDT = 68 makes
// the DMMA fragment loads (lanes = 4 consecutive m x 8 consecutive r) bank-conflict free.
constexpr int LDT = NB + 4;
constexpr int kTileDoubles = NB * LDT;

// Stage the 64x64 tile whose (r, m) element is at src[m*ld + r] into smem dst[m*LDT + r].
__device__ __forceinline__ void stage_tile(double* dst, const double* src, size_t ld) {
  // 64 columns x 512 B; 16 B per cp.async; 2048 chunks / 256 threads = 8 each
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int chunk = it * kSolveThreads + threadIdx.x;
    const int m = chunk >> 5, r2 = (chunk & 31) * 2;
    cp_async16(dst + m * LDT + r2, src + (size_t)m * ld + r2);
  }
}

// Register tile of one thread in the warp-level DMMA layout.  Warp w owns rows
// R0 = 32 (w & 1) .. +31 and columns C0 = 16 (w >> 1) .. +15 of the 64x64 tile; inside it
//   v[mi][ni][e]  <->  row R0 + 8 mi + (lane >> 2),  column C0 + 8 ni + 2 (lane & 3) + e.
struct Frag {
  double v[4][2][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) v[mi][ni][0] = v[mi][ni][1] = 0.0;
  }
};

// acc += P^T Q over the 64-long contraction:  acc[r][c] += sum_m P[m*LDT + r] * Q[m*LDT + c]
// (FP64 tensor-core path: mma.sync.m8n8k4.f64, SASS DMMA.8x8x4)
__device__ __forceinline__ void tile_dmma(Frag& acc, const double* __restrict__ P,
                                          const double* __restrict__ Q, int R0, int C0, int lane) {
  const int g = lane >> 2, t4 = lane & 3;
  const double* pa = P + t4 * LDT + R0 + g;
  const double* pb = Q + t4 * LDT + C0 + g;
#pragma unroll 4
  for (int m0 = 0; m0 < NB; m0 += 4) {
    double a[4], b[2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = pa[m0 * LDT + 8 * mi];
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) b[ni] = pb[m0 * LDT + 8 * ni];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc.v[mi][ni][0]), "+d"(acc.v[mi][ni][1])
                     : "d"(a[mi]), "d"(b[ni]));
  }
}

// Shared memory map (doubles):
//   buf[4][NB*LDT]     double-buffered operand tiles (P, Q) of the k loop           136 KB
//   aliases used after the k loop of a task:
//     LTs = buf                diagonal task: scaled L_jj^{-1}, transposed, [m][c]
//     Cs  = buf                panel task: C transposed to [m][r]
//     Bs  = buf + NB*LDT       panel task: LinvT tile [m][c]
//   vec[8][NB]  u[2], um[2] (pivot rows of W and M, ping-pong), piv, isd, tvec, yk
constexpr int kSolveSmemDoubles = 4 * kTileDoubles + 8 * NB;
constexpr size_t kSolveSmemBytes = kSolveSmemDoubles * sizeof(double);

__global__ void __launch_bounds__(kSolveThreads, 1) chol_dataflow_kernel(const CholArgs g) {
  extern __shared__ __align__(16) double sm[];
  double* const buf = sm;
  double* const vec = sm + 4 * kTileDoubles;
  double* const u_buf = vec;             // [2][NB]
  double* const um_buf = vec + 2 * NB;   // [2][NB]
  double* const piv = vec + 4 * NB;
  double* const isd = vec + 5 * NB;
  double* const tvec = vec + 6 * NB;
  double* const yk = vec + 7 * NB;
  __shared__ int s_task;
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const int gq = lane >> 2, t4 = lane & 3;
  const int R0 = 32 * (wid & 1), C0 = 16 * (wid >> 1);
  const int T = g.T;
  const size_t ld = (size_t)g.ld;
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
    BA_TRACE_SET(t, 0, ((unsigned long long)i << 32) | (unsigned)j);
    BA_TRACE_SET(t, 1, (unsigned long long)blockIdx.x);
    BA_TRACE(t, 2);   // task grabbed

    Frag acc;
    acc.zero();
    double bacc = 0.0;  // diag task, tid < NB: sum_k (L_jk y_k)[tid]

    // ---- k loop: acc += L_ik L_jk^T, operands double-buffered through cp.async ------------
    auto issue = [&](int k) {
      double* P = buf + (size_t)(2 * (k & 1)) * kTileDoubles;
      stage_tile(P, g.A + (size_t)(k * NB) * ld + (size_t)i * NB, ld);
      if (!diag) stage_tile(P + kTileDoubles, g.A + (size_t)(k * NB) * ld + (size_t)j * NB, ld);
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
      if (diag && tid < NB) yk[tid] = __ldcg(g.rhs + k * NB + tid);  // y_k
      __syncthreads();
      const double* P = buf + (size_t)(2 * (k & 1)) * kTileDoubles;
      const double* Q = diag ? P : P + kTileDoubles;
      tile_dmma(acc, P, Q, R0, C0, lane);
      if (diag && tid < NB) {
        double s = 0.0;
#pragma unroll 8
        for (int m = 0; m < NB; ++m) s += P[m * LDT + tid] * yk[m];
        bacc += s;
      }
      __syncthreads();
    }
    BA_TRACE(t, 3);   // k loop done

    const double* Aij = g.A + (size_t)(j * NB) * ld + (size_t)i * NB;
    if (diag) {
      // ---- W = A_jj - acc (full symmetric tile: only the lower triangle is valid in memory),
      //      M = I;  both live in registers in the DMMA layout -------------------------------
      Frag W, M;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = R0 + 8 * mi + gq, c = C0 + 8 * ni + 2 * t4 + e;
            const int hi = r > c ? r : c, lo = r > c ? c : r;
            W.v[mi][ni][e] = __ldcg(Aij + (size_t)lo * ld + hi) - acc.v[mi][ni][e];
            M.v[mi][ni][e] = (r == c) ? 1.0 : 0.0;
          }
      if (tid == 0) s_bad = 0;
      // Gauss-Jordan sweep without pivoting.  After step p the rows r > p of W hold the Schur
      // complement (kept symmetric, so column p equals row p) and M accumulates the unit
      // lower-triangular L~^{-1}:  M A = U = D L~^T.  One barrier per step: the owners of row
      // p+1 publish it (ping-pong buffers) right after their update of step p.
      auto publish_row = [&](int p) {
        double* u = u_buf + (p & 1) * NB;
        double* um = um_buf + (p & 1) * NB;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
          if (R0 + 8 * mi + gq == p) {
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
              const int c = C0 + 8 * ni + 2 * t4;
              *reinterpret_cast<double2*>(u + c) = make_double2(W.v[mi][ni][0], W.v[mi][ni][1]);
              *reinterpret_cast<double2*>(um + c) = make_double2(M.v[mi][ni][0], M.v[mi][ni][1]);
            }
          }
      };
      publish_row(0);
      for (int p = 0; p < NB; ++p) {
        __syncthreads();
        const double* u = u_buf + (p & 1) * NB;
        const double* um = um_buf + (p & 1) * NB;
        double d = u[p];
        if (!(d > 0.0)) {          // also catches NaN; uniform across the block
          if (tid == 0) s_bad = 1;
          d = 1.0;
        }
        if (tid == 0) piv[p] = d;
        const double id = 1.0 / d;
        double f[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const int r = R0 + 8 * mi + gq;
          f[mi] = (r > p) ? u[r] * id : 0.0;
        }
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const int c = C0 + 8 * ni + 2 * t4;
          const double2 uc = *reinterpret_cast<const double2*>(u + c);
          const double2 mc = *reinterpret_cast<const double2*>(um + c);
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            W.v[mi][ni][0] -= f[mi] * uc.x;
            W.v[mi][ni][1] -= f[mi] * uc.y;
            M.v[mi][ni][0] -= f[mi] * mc.x;
            M.v[mi][ni][1] -= f[mi] * mc.y;
          }
        }
        if (p + 1 < NB) publish_row(p + 1);
      }
      __syncthreads();
      BA_TRACE(t, 4);   // sweep done
      const bool bad = s_bad != 0;
      if (bad && tid == 0) *g.status = 1.0;
      if (tid < NB) isd[tid] = bad ? 1.0 : 1.0 / sqrt(piv[tid]);
      __syncthreads();
      // scaled inverse, transposed, into smem:  LTs[m*LDT + c'] = Linv[c'][m] = M[c'][m] isd[c']
      // (on a failed pivot leave an identity behind so that dependants stay finite)
      double* LTs = buf;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = R0 + 8 * mi + gq, c = C0 + 8 * ni + 2 * t4 + e;
            double v = (c <= r) ? M.v[mi][ni][e] * isd[r] : 0.0;
            if (bad) v = (r == c) ? 1.0 : 0.0;
            LTs[c * LDT + r] = v;
          }
      if (tid < NB) tvec[tid] = __ldcg(g.rhs + j * NB + tid) - bacc;
      __syncthreads();
      double* LT = g.LinvT + (size_t)j * NB * NB;
      for (int e = tid; e < NB * NB; e += kSolveThreads) LT[e] = LTs[(e >> 6) * LDT + (e & 63)];
      // forward substitution: y_j = Linv (b_j - sum_k L_jk y_k);  thread r: sum_m LTs[m][r] t[m]
      if (tid < NB) {
        double s = 0.0;
#pragma unroll 8
        for (int m = 0; m < NB; ++m) s += LTs[m * LDT + tid] * tvec[m];
        g.rhs[j * NB + tid] = s;
      }
    } else {
      // ---- panel tile: L_ij = (A_ij - acc) Linv_jj^T -----------------------------------------
      double* Cs = buf;                  // [m][r] = C[r][m]
      double* Bs = buf + kTileDoubles;   // [m][c] = Linv[c][m]
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = R0 + 8 * mi + gq, c = C0 + 8 * ni + 2 * t4 + e;
            Cs[c * LDT + r] = __ldcg(Aij + (size_t)c * ld + r) - acc.v[mi][ni][e];
          }
      wait_flag(&g.flags[(size_t)j * T + j], epoch);
      __syncthreads();
      BA_TRACE(t, 4);   // diagonal inverse available
      stage_tile(Bs, g.LinvT + (size_t)j * NB * NB, NB);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      Frag out;
      out.zero();
      tile_dmma(out, Cs, Bs, R0, C0, lane);
      double* Lij = g.A + (size_t)(j * NB) * ld + (size_t)i * NB;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = R0 + 8 * mi + gq, c = C0 + 8 * ni + 2 * t4 + e;
            Lij[(size_t)c * ld + r] = out.v[mi][ni][e];
          }
    }
    // ---- publish ------------------------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(&g.flags[(size_t)i * T + j], epoch);
    BA_TRACE(t, 5);   // published
  }

  // ==================================== backward substitution ==============================
  // x_k = L_kk^{-T} (y_k - sum_{i>k} L_ik^T x_i),  k = T-1 .. 0
  double* const part = buf;            // [4][NB] partial sums
  double* const accv = vec;            // [NB]
  double* const LTs = buf + 8 * NB;    // [NB][NBP] padded copy of LinvT
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = (int)atomicAdd(&g.tickets[1], 1u);
    __syncthreads();
    const int bt = s_task;
    if (bt >= T) break;
    const int k = T - 1 - bt;
    BA_TRACE_SET(ntasks + bt, 0, (unsigned long long)k);
    BA_TRACE_SET(ntasks + bt, 1, (unsigned long long)blockIdx.x);
    BA_TRACE(ntasks + bt, 2);
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
    BA_TRACE(ntasks + bt, 3);   // all x_i folded
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
    BA_TRACE(ntasks + bt, 5);
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
#ifdef BA_SOLVE_TRACE
  g.trace = c.solve_trace;
#endif
  const int ntasks = T * (T + 1) / 2;
  int grid = ntasks < c.num_sms ? ntasks : c.num_sms;   // 1 CTA / SM (128 KB smem): all co-resident
  chol_dataflow_kernel<<<grid, kSolveThreads, kSolveSmemBytes, st>>>(g);
  c.launches += 1;
  return cudaGetLastError();
}

}  // namespace ba
