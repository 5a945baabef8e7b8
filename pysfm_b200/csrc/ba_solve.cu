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
//                          registers on the FP64 tensor pipe, consuming the L tiles of earlier
//                          columns as soon as their ready-flags (epoch valued, in L2) go up,
//                          and is written exactly once.  A chain task owns a diagonal tile and
//                          the panel tile left of it: it sweeps the diagonal tile by blocked
//                          Gauss-Jordan in registers / shared memory into L_jj^{-1} (L_jj is
//                          never formed), published row block by row block, so that every
//                          panel tile below is a product with the inverse that tracks the
//                          sweep, and the forward substitution y_j = L_jj^{-1}(b_j - ...) rides
//                          along.  Tiles are published column block by column block (bulk
//                          stores, relaxed flags); the only fences left are the releases of
//                          the sweep and of the last column group of a tile.  When the tile
//                          tickets run out the CTAs take the backward-substitution tasks
//                          x_k = L_kk^{-T}(y_k - sum_i L_ik^T x_i) from a second ticket; x_k is
//                          its own ready-flag (the solution vector starts out as a NaN pattern
//                          and is polled).  Tickets are handed out in dependency order and the
//                          grid never exceeds the number of co-resident CTAs, so a waiting CTA
//                          always waits on a task that is already running.  DESIGN.md 4.2 has
//                          the measured timeline and what bounds it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "ba_context.h"

namespace ba {

constexpr int NB = kSolveTile;     // 64
constexpr int NBP = NB + 1;        // padded row of the diagonal-tile work arrays
constexpr int kSolveThreads = 256;
// A quiet NaN with a payload no arithmetic produces: marks entries of x that are not computed yet.
constexpr long long kNotYet = 0x7ff8dead0000beefLL;
constexpr int kPhaseAll = 0, kPhaseWindow = 1, kPhaseBackward = 2;
constexpr int kDiagTask = 0x40000000;   // task code of the distributed solve: diagonal-update task D_j (low 16 bits = j)

// ------------------------------------------------------------------------------------------
// packed block index of (a, b), a <= b < nc:  rows of the upper block triangle back to back
__host__ __device__ __forceinline__ size_t packed_block(int a, int b, int nc) {
  return (size_t)a * nc - (size_t)a * (a - 1) / 2 + (b - a);
}

// A(p, q), p >= q  (column-major lower, A[q*ld + p])  <-  packed block (q/6, p/6)[q%6][p%6]
__global__ void __launch_bounds__(256)
expand_system_kernel(const double* __restrict__ packed, int nc, int n_sys, int ld,
                     const unsigned char* __restrict__ mask, bool have_mask,
                     double* __restrict__ A, double* __restrict__ rhs,
                     unsigned int* __restrict__ tickets, double* __restrict__ status, double* __restrict__ x,
                     unsigned int* __restrict__ abort_word, double* __restrict__ dist_status, double unit_diag) {
  const int q = blockIdx.x;  // column
  // the solution vector starts out as "not there yet": the backward substitution polls the data itself
  if (threadIdx.x == 0) x[q] = __longlong_as_double(kNotYet);
  if (q == 0 && threadIdx.x == 0) {   // reset the solver's task tickets and status word (saves two memset nodes)
    tickets[0] = 0u;
    tickets[1] = 0u;
    *status = 0.0;
    *abort_word = 0u;
    if (dist_status) *dist_status = 0.0;
  }
  const bool free_q = q < n_sys && (!have_mask || mask[q]);
  double* col = A + (size_t)q * ld;
  if (!free_q) {
    // unit_diag: 1 on a single GPU; in the distributed solve the ranks' contributions are summed,
    // so only rank 0 contributes the unit pivot of a frozen / padding parameter
    for (int p = q + threadIdx.x; p < ld; p += blockDim.x) col[p] = (p == q) ? unit_diag : 0.0;
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
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Release at gpu scope = fence + store (SASS: MEMBAR.ALL.GPU; ST.STRONG.GPU).  It is cumulative:
// stores of OTHER threads of the CTA that a barrier (bar.sync, __syncwarp) ordered before it are
// covered, so no publication below issues a separate __threadfence() first (that second fence
// cost ~0.4 us on every hop of the chain).
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// System-scope variants: in the distributed solve the flags of a rank are written by its peers
// over NVLink, so both sides of every flag use .sys (a .gpu load is not morally strong against
// a store of another GPU).
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {   // peer memory: never cached on this side
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
template <bool SYS>
__device__ __forceinline__ unsigned int ldf_acquire(const unsigned int* p) { return SYS ? ld_acquire_sys(p) : ld_acquire(p); }
template <bool SYS>
__device__ __forceinline__ unsigned int ldf_relaxed(const unsigned int* p) { return SYS ? ld_relaxed_sys(p) : ld_relaxed(p); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Column m of a k-major shared tile (64 contiguous doubles at tile + m*LDT) to the global tile
// (column m at dst + m*ld) as ONE bulk copy on the TMA path, own bulk group; bulk_store_wait()
// waits for the COMPLETION of the caller's copies (not .read: the bytes are in L2 when it
// returns).  Every writer of the tile ran fence.proxy.async before the barrier that precedes the
// call.  Unlike a release (MEMBAR.GPU, which held up the memory traffic of the whole SM for
// ~0.6 us per publication), the wait is for exactly these bytes and stalls nobody else.
__device__ __forceinline__ void bulk_store_column(double* dst, int ld, const double* tile, int m) {
  const unsigned int src = (unsigned int)__cvta_generic_to_shared(tile + m * (NB + 4));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(dst + (size_t)m * ld), "r"(src) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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
__device__ long long g_sweep_clk[64];
__device__ unsigned long long g_dbg_time[256];
// trace builds: the producer chain task whose row-block flag times are recorded, and the first of
// the two consecutive tickets whose group timeline is recorded (set by solve_bench)
__device__ int g_dbg_producer = 36, g_dbg_consumer = 52;
#define BA_GT(idx, cond)                                                           \
  do {                                                                             \
    if (cond) {                                                                    \
      unsigned long long gt__;                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__));                     \
      g_dbg_time[(idx)] = gt__;                                                    \
    }                                                                              \
  } while (0)
#define BA_CLK(idx)                                                                \
  do {                                                                             \
    if (t == 0 && tid == 224) g_sweep_clk[(idx)] = clock64();                      \
  } while (0)
#define BA_CLK0(idx)                                                               \
  do {                                                                             \
    if (t == 0 && tid == 0) g_sweep_clk[(idx)] = clock64();                        \
  } while (0)
// clock read that waits for the value `dep` (a load result or accumulator) to exist
#define BA_CLK_DEP(idx, dep)                                                       \
  do {                                                                             \
    if (t == 0 && tid == 224) {                                                    \
      long long c__;                                                               \
      asm volatile("{ .reg .f64 z; add.f64 z, %1, 0d0000000000000000; mov.u64 %0, %%clock64; }" : "=l"(c__) : "d"(dep)); \
      g_sweep_clk[(idx)] = c__;                                                    \
    }                                                                              \
  } while (0)
#else
#define BA_GT(idx, cond) do { } while (0)
#define BA_CLK(idx) do { } while (0)
#define BA_CLK0(idx) do { } while (0)
#define BA_CLK_DEP(idx, dep) do { } while (0)
#define BA_TRACE_DECL
#define BA_TRACE(rec, slot) do { } while (0)
#define BA_TRACE_SET(rec, slot, v) do { } while (0)
#endif

// Wait-time profile (ba_solver_profile, diagnostics): nanoseconds the polling thread of each wait
// site spent waiting and the time spent inside tasks, summed over the CTAs of the launches since
// the last reset.  Off (g.prof == nullptr) unless BA_OPT_SOLVER_PROFILE is set: two uniform
// branches per wait.
#define BA_PROF_T0() const unsigned long long pt0__ = g.prof ? global_ns() : 0ull
#define BA_PROF_ADD(site) do { if (g.prof) atomicAdd(&g.prof[(site)], global_ns() - pt0__); } while (0)
enum { kProfTask = 0, kProfWaitK = 1, kProfLast = 2, kProfPanel = 3, kProfDiagFlag = 4, kProfPush = 5, kProfContrib = 6,
       kProfYflag = 7, kProfBackward = 8, kProfBarrier = 9, kProfKernel = 10, kProfChainTask = 11, kProfDiagTask = 12 };

struct CholArgs {
  BA_TRACE_DECL
  double* __restrict__ A;         // [ld*ld] dense lower, column-major; overwritten by L
  double* __restrict__ rhs;       // [ld] b -> y (forward substitution)
  double* __restrict__ x;         // [ld] solution
  double* __restrict__ LinvT;     // [T][NB*NB]  LinvT[m*NB + c] = (L_jj^{-1})[c][m]
  double* __restrict__ Wpart;     // [T][NB*NB + NB]  D_j -> C_j: A_jj - sum_{k<j-2} L_jk L_jk^T ([col][row], lower) | sum_{k<j-2} L_jk y_k
  unsigned int* __restrict__ flags;   // [T*T] tile (i,j) ready == epoch ; [T*T + j] D_j's partial diagonal tile ready ; [T*T + T + k] y_k ready ; [T*T + 2T + 8k + b] rows 8b.. of Linv_kk ready ; [T*T + 10T + 8(iT+j) + b] columns 8b.. of L_ij ready
  unsigned int* __restrict__ tickets; // [0] tile tasks, [1] back-substitution tasks
  double* __restrict__ status;    // set to 1 on a non-positive pivot, 2 when a spin-wait ran past the deadline
  int ld, T;
  unsigned int epoch;
  // ---- robustness ----
  unsigned long long* __restrict__ prof;   // [16] wait-time profile (nullptr = off)
  unsigned int* __restrict__ abort;   // [1] != 0: some spin-wait (here or on a peer) gave up; every wait falls through
  unsigned long long spin_limit_ns;   // budget of the whole launch for waiting
  int strict;                         // release/acquire publication of the column-block flags (PTX-model clean, slower)
  int split;                          // 1: diagonal-update tasks D_j carry the W / y part of the chain tasks' k loops (large T);
                                      // 0: the chain task does it all and D_j is empty (small T: the hand-over costs more than it hides)
  // ---- distributed mode (DIST): tiles are owned by ranks, see the header of the kernel ----
  int world, rank;
  const int* __restrict__ tasks;      // this rank's tasks in global ticket order: (i << 16) | j, chain task C_j as (j, j)
  int ntasks;
  // ---- windowed use by the blocked tcgen05 solver (ba_solve_tc.cuh; single GPU only) ----
  int phase;                          // kPhaseAll: factor + substitute; kPhaseWindow: the first `window_tasks` tile tickets only
                                      // (the leading tile columns of the matrix at A), no backward substitution;
                                      // kPhaseBackward: backward substitution only, over a factor finished by earlier launches
  int window_tasks;
  double* base[kMaxPeers];            // DistLayout section of every rank (own rank included), peer-mapped
};

__device__ __forceinline__ void raise_abort(const CholArgs& g) {
  atomicExch(g.abort, 1u);
  *g.status = 2.0;
  if (g.world > 1) {
    const DistLayout d = dist_layout(g.ld);
    for (int p = 0; p < g.world; ++p)
      if (p != g.rank) st_relaxed_sys(reinterpret_cast<unsigned int*>(g.base[p] + d.abort), 1u);
  }
}
// Called on the idle path of every polling loop.  true = stop waiting: this launch (or a peer's)
// has spent its waiting budget -- a peer died, a kernel faulted, or a flag was lost.  The task
// then runs on with whatever is there (finite garbage), the launch ends, status 2 is reported.
__device__ __forceinline__ bool spin_expired(const CholArgs& g, unsigned int& spins, const unsigned long long& t0) {
  if ((++spins & 63u) != 0u) return false;
  if (ld_relaxed_sys(g.abort) != 0u) {
    *g.status = 2.0;
    return true;
  }
  if (global_ns() - t0 < g.spin_limit_ns) return false;
  raise_abort(g);
  return true;
}

// Block-wide wait until *f == epoch (thread 0 spins with acquire loads).
template <bool SYS>
__device__ __forceinline__ void wait_flag(const CholArgs& g, const unsigned int* f, unsigned int epoch, const unsigned long long& t0,
                                          int site = kProfWaitK) {
  if (threadIdx.x == 0) {
    unsigned int spins = 0;
    if (SYS) {
      // Distributed solve: an acquire load at system scope is a load plus MEMBAR.SYS, and that fence
      // waits for every store this thread's SM has on the links (the tile it just pushed to seven
      // peers): a loaded NVLink round trip PER POLL.  The polls are relaxed; what the flag guards is
      // read from L2 (cp.async.cg / ld.cg) after the flag load returned, L2 being the point of
      // coherence for this GPU's memory whoever wrote it; g.strict adds the fence.
      if (ld_relaxed_sys(f) != epoch) {
        BA_PROF_T0();
        while (ld_relaxed_sys(f) != epoch) {
          if (spin_expired(g, spins, t0)) break;
          __nanosleep(20);
        }
        BA_PROF_ADD(site);
      }
      if (g.strict) __threadfence_system();
      return;
    }
    if (ldf_acquire<SYS>(f) == epoch) return;
    BA_PROF_T0();
    while (ldf_acquire<SYS>(f) != epoch) {
      if (spin_expired(g, spins, t0)) break;
      __nanosleep(20);
    }
    BA_PROF_ADD(site);
    (void)site;
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

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// acc (+/-)= P^T Q over the 64-long contraction:  acc[r][c] += sum_m P[m*LDT + r] * Q[m*LDT + c]
// (FP64 tensor-core path: mma.sync.m8n8k4.f64, SASS DMMA.8x8x4)
template <bool NEG>
__device__ __forceinline__ void tile_dmma(Frag& acc, const double* __restrict__ P,
                                          const double* __restrict__ Q, int R0, int C0, int lane,
                                          int m_lo = 0, int m_hi = NB) {
  const int g = lane >> 2, t4 = lane & 3;
  const double* pa = P + t4 * LDT + R0 + g;
  const double* pb = Q + t4 * LDT + C0 + g;
#pragma unroll 4
  for (int m0 = m_lo; m0 < m_hi; m0 += 4) {
    double a[4], b[2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = NEG ? -pa[m0 * LDT + 8 * mi] : pa[m0 * LDT + 8 * mi];
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) b[ni] = pb[m0 * LDT + 8 * ni];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) dmma884(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi], b[ni]);
  }
}

// Diagonal tasks keep the lower triangle of their tile as 8x8 DMMA accumulator tiles owned by
// ROW BLOCK: warp r holds tiles (r, c), c <= r;  t[c][e] <-> row 8r + (lane >> 2), column
// 8c + 2 (lane & 3) + e.  The blocked sweep below updates them in place.
struct RowTiles {
  double t[8][2];
};

// W(R, c) -= sum_m P[m*LDT + 8R + .] * P[m*LDT + 8c + .],  c <= R, over m in [m_lo, m_hi).
// The row block R is a COMPILE-TIME constant, dispatched once per call (rows_dispatch): with a
// run-time r the tile loop reads `if (c <= r) dmma`, the compiler predicates it, and a
// predicated-off DMMA still holds the FP64 tensor pipe for its whole issue slot -- the lower
// triangle then costs as much as the full square (tools/microbench/tile_bench.cu: 4696 clk per
// 64-deep update against 4566 for a full 64^3 product; 3199 with this form and the pairing below;
// a switch INSIDE the loop is worse, 5271: an indirect branch per k step).
template <int R>
__device__ __forceinline__ void diag_rows_dmma_t(RowTiles& W, const double* __restrict__ P, int lane, int m_lo, int m_hi) {
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
__device__ __forceinline__ void diag_rows_dmma(RowTiles& W, const double* __restrict__ P, int r, int lane,
                                               int m_lo = 0, int m_hi = NB) {
  switch (r) {
    case 0: diag_rows_dmma_t<0>(W, P, lane, m_lo, m_hi); break;
    case 1: diag_rows_dmma_t<1>(W, P, lane, m_lo, m_hi); break;
    case 2: diag_rows_dmma_t<2>(W, P, lane, m_lo, m_hi); break;
    case 3: diag_rows_dmma_t<3>(W, P, lane, m_lo, m_hi); break;
    case 4: diag_rows_dmma_t<4>(W, P, lane, m_lo, m_hi); break;
    case 5: diag_rows_dmma_t<5>(W, P, lane, m_lo, m_hi); break;
    case 6: diag_rows_dmma_t<6>(W, P, lane, m_lo, m_hi); break;
    default: diag_rows_dmma_t<7>(W, P, lane, m_lo, m_hi); break;
  }
}

// Row block of the diagonal tile owned by warp w.  Warps w and w + 4 share a scheduler (and its
// share of the FP64 tensor pipe); the lower triangle gives row block r r + 1 tiles, so pairing
// (0,7) (1,6) (2,5) (3,4) puts 9 tiles on every scheduler instead of 6 / 8 / 10 / 12.
__device__ __forceinline__ int row_block_of_warp(int w) { return w < 4 ? w : 11 - w; }

// 1/d for a pivot d > 0: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~20 bits) and one third-order
// Newton step (3 dependent DFMA) -> ~2^-60 relative error, without the special-case branches
// and the two extra DFMAs of the IEEE division that sat on the pivot chain.
__device__ __forceinline__ double pivot_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  e = fma(e, e, e);
  return fma(x, e, x);
}

// v[t4] of a 4-vector held in registers, without dynamic register indexing
__device__ __forceinline__ double sel4(double v0, double v1, double v2, double v3, int t4) {
  const double lo = (t4 & 1) ? v1 : v0, hi = (t4 & 1) ? v3 : v2;
  return (t4 & 2) ? hi : lo;
}

// 8x8 scratch tiles of the blocked sweep: row stride 12 doubles (conflict-free fragment loads)
constexpr int TS = 12;
constexpr int TD = 8 * TS;

// Shared memory map (doubles):
//   buf[4][NB*LDT]     double-buffered operand tiles (P, Q) of the k loop           136 KB
//   aliases used after the k loop of a task:
//     LTs  = buf               diagonal task: L_jj^{-1}, transposed, [m][c]
//     Wcol, Mrow, Lp, Mp = buf + NB*LDT ...   diagonal task: 8x8 scratch tiles of the sweep
//     Cs  = buf                panel task: C transposed to [m][r]
//     Bs  = buf + NB*LDT       panel task: LinvT tile [m][c]
//   vec[8][NB]  accv (backward phase), tvec, yk
constexpr int kSolveSmemDoubles = 4 * kTileDoubles + 8 * NB;
constexpr size_t kSolveSmemBytes = kSolveSmemDoubles * sizeof(double);

// DIST (points sharded over the GPUs of one node, large reduced systems): ONE launch per rank
// does the reduce-scatter of the ranks' contributions, the factorisation and the all-gather of
// the factor, tile by tile, over NVLink peer memory:
//   * every tile task has ONE owner rank (g.tasks: the rank's tasks in global ticket order).  The
//     chain tasks and the band next to the diagonal stay on rank 0, so the critical path never
//     crosses a link; the rest of row i belongs to one rank, so a row's tile-to-tile dependency
//     (L_ij is the last operand of task (i, j+1)) stays local and progressive.
//   * the owner sums A_ij over the ranks' dense contributions (peer loads in rank order, one peer
//     per step of the k loop, so the link latency hides behind the tile products),
//   * and pushes the finished L_ij (bulk copies shared -> peer global), the rows of Linv_jj and
//     y_j to EVERY rank, followed by the tile's flags (system-scope release): every rank ends up
//     with the whole factor and runs the backward substitution on its own copy -- all ranks get
//     the same bits for dC without a broadcast.
template <bool DIST>
__global__ void __launch_bounds__(kSolveThreads, 1) chol_dataflow_kernel(const CholArgs g) {
  extern __shared__ __align__(16) double sm[];
  double* const buf = sm;
  double* const vec = sm + 4 * kTileDoubles;
  double* const tvec = vec + 6 * NB;
  double* const yk = vec + 7 * NB;
  __shared__ int s_task;
  __shared__ __align__(16) unsigned int s_snap[8];   // asynchronous snapshot of the 8 row-block flags a task polls
  __shared__ int s_bad;
  __shared__ unsigned long long s_t0;   // start of the launch: the waiting budget counts from here
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const int gq = lane >> 2, t4 = lane & 3;
  const int R0 = 32 * (wid & 1), C0 = 16 * (wid >> 1);
  const int T = g.T;
  const size_t ld = (size_t)g.ld;
  // C_0, then per column: diagonal-update task, chain task, the panels below
  const int ntasks = DIST ? g.ntasks : g.phase == kPhaseWindow ? g.window_tasks : g.phase == kPhaseBackward ? 0 : 1 + (T - 1) * (T + 2) / 2;
  const unsigned int epoch = g.epoch;
  const DistLayout dl = dist_layout(g.ld);
  if (tid == 0) s_t0 = global_ns();
  __syncthreads();
  if (DIST) {
    // start barrier: every rank's contribution is expanded (its expand kernel precedes this launch
    // in its stream) and nobody is still inside the previous solve; nothing is pushed before it
    if (blockIdx.x == 0 && tid < g.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned int*>(g.base[tid] + dl.bar) + g.rank, epoch);
    }
    if (tid < g.world) {
      const unsigned int* mine = reinterpret_cast<const unsigned int*>(g.base[g.rank] + dl.bar) + tid;
      unsigned int spins = 0;
      BA_PROF_T0();
      while ((int)(ld_acquire_sys(mine) - epoch) < 0) {   // (a peer may already be one solve ahead)
        if (spin_expired(g, spins, s_t0)) break;
        __nanosleep(100);
      }
      if (tid == 0) BA_PROF_ADD(kProfBarrier);
    }
    __syncthreads();
  }

  // ======================================= factorisation ===================================
  // Ticket order:  C_0;  then per column j = 0 .. T-2:  C_{j+1}, (j+2, j), ..., (T-1, j).
  //   panel task (i, j), i >= j + 2:  L_ij = (A_ij - sum_{k<j} L_ik L_jk^T) Linv_jj^T
  //   chain task C_j:  the panel tile (j, j-1) AND the diagonal tile (j, j) in one task, so that
  //                    L_{j,j-1} goes from the tensor pipe straight into the last update of the
  //                    diagonal tile without a global-memory round trip and a flag hop.
  // Both kinds consume Linv_jj ROW BLOCK BY ROW BLOCK while the chain task that owns column j is
  // still sweeping: rows 8cb .. 8cb+7 of Linv give columns 8cb .. 8cb+7 of L_ij, and (chain
  // tasks) each finished column block is folded into the diagonal tile, lazily: when no further
  // rows are waiting, or after the last one.  When the sweep of column j ends, the next chain
  // task is a group of blocks (~2-3 us) away from starting its own sweep.
  unsigned int* const yflag = g.flags + (size_t)T * T + T;
  unsigned int* const rowflag = g.flags + solve_rowflag_base(T);          // 8 per diagonal tile, 32-byte aligned
  unsigned int* const colflag = g.flags + solve_rowflag_base(T) + 8 * T;   // [(i*T + j)*8 + cb]: columns 8cb.. of L_ij are out
  const int Tm = T - 1;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = (int)atomicAdd(&g.tickets[0], 1u);
    __syncthreads();
    const int t = s_task;
    if (t >= ntasks) break;
    bool chain = true;
    bool diag = false;  // diagonal-update task D_j
    int j = 0;          // chain / diag: diagonal tile index;  panel: column
    int pi = 0, pj = 0; // the panel tile (pi, pj) of this task (chain and diag: (j, j - 1))
    if (DIST) {
      const int code = g.tasks[t];
      diag = (code & kDiagTask) != 0;
      pi = (code >> 16) & 0x3fff;
      pj = code & 0xffff;
      chain = !diag && (pi == pj);
      j = pj;
      if (diag) pi = j;
      if ((chain || diag) && j > 0) pj = j - 1;
    } else if (t > 0) {
      // column-major enumeration over columns 0 .. T-2, column jc holding T - jc tasks:
      // D_{jc+1}, C_{jc+1}, (jc+2, jc), ..., (T-1, jc)
      const int tt = t - 1;
      const double Tf = (double)T + 0.5;
      int jc = (int)(Tf - sqrt(Tf * Tf - 2.0 * (double)tt));
      if (jc < 0) jc = 0;
      if (jc > Tm - 1) jc = Tm - 1;
      while (jc > 0 && (size_t)jc * T - (size_t)jc * (jc - 1) / 2 > (size_t)tt) --jc;
      while (jc < Tm - 1 && (size_t)(jc + 1) * T - (size_t)(jc + 1) * jc / 2 <= (size_t)tt) ++jc;
      const int rem = tt - (int)((size_t)jc * T - (size_t)jc * (jc - 1) / 2);
      diag = (rem == 0);
      chain = (rem == 1);
      j = (diag || chain) ? jc + 1 : jc;
      pi = (diag || chain) ? jc + 1 : jc + rem;
      pj = jc;
    }
    const bool has_panel = !diag && !(chain && j == 0);
    BA_TRACE_SET(t, 0, ((unsigned long long)((chain || diag) ? j : pi) << 32) | (unsigned)(chain ? j : diag ? (0x8000 | j) : pj));
    BA_TRACE_SET(t, 1, (unsigned long long)blockIdx.x);
    BA_TRACE(t, 2);   // task grabbed
    BA_PROF_T0();

    // ---- original tiles first: their latency hides behind the k loop --------------------------
    const int r = row_block_of_warp(wid);   // row block owned in the diagonal tile
    Frag acc;            // panel tile  A_{pi,pj} - sum_k L_{pi,k} L_{pj,k}^T   (warp tile R0, C0)
    RowTiles W;          // diagonal tile A_jj - sum_k L_jk L_jk^T, lower triangle, row-block owned
    acc.zero();
    // DIST: the original tile is the sum of the ranks' contributions, fetched from the peers' dense
    // copies right here, two peers (round trips) in flight at a time, added in rank order.  (Riding
    // on the first steps of the k loop instead hid the latency but kept 32 more registers live
    // across the tile products: spills in the loop every task spends its life in.)
    const size_t tile_off = (size_t)(pj * NB) * ld + (size_t)pi * NB;
    auto push_tile_to_peers = [&]() {   // all warps; Ls = buf + 2 tiles holds L_{pi,pj} as [m][row]
      const double* Lsrc = buf + 2 * kTileDoubles;
      for (int p = 0; p < g.world; ++p) {
        if (p == g.rank) continue;
        double* dst = g.base[p] + dl.L + tile_off;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int m = 8 * wid + q;
          const double2 v = *reinterpret_cast<const double2*>(Lsrc + m * LDT + 2 * lane);
          *reinterpret_cast<double2*>(dst + (size_t)m * ld + 2 * lane) = v;
        }
      }
    };
    if (DIST && has_panel) {
      BA_PROF_T0();
      // The four operand buffers are still free: four peers' tiles are staged into them at once with
      // cp.async (one loaded NVLink round trip, ~9 us, per FOUR peers; register loads managed one
      // round trip per quarter of the fragment) and summed in rank order.  (.cg: peer memory is
      // never held in this GPU's L2, and L1 is bypassed, so every solve sees the fresh contribution.)
      for (int p0 = 0; p0 < g.world; p0 += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (p0 + q < g.world) stage_tile(buf + (size_t)q * kTileDoubles, g.base[p0 + q] + dl.contrib + tile_off, ld);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (p0 + q < g.world) {
            const double* Cq = buf + (size_t)q * kTileDoubles;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                  acc.v[mi][ni][e] += Cq[(C0 + 8 * ni + 2 * t4 + e) * LDT + R0 + 8 * mi + gq];
          }
        }
        __syncthreads();   // the buffers are free again (next round / the k loop)
      }
      if (tid == 0) BA_PROF_ADD(kProfContrib);
    }
    if (!DIST && has_panel) {
      const double* Ap = g.A + (size_t)(pj * NB) * ld + (size_t)pi * NB;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int rr = R0 + 8 * mi + gq, cc = C0 + 8 * ni + 2 * t4 + e;
            acc.v[mi][ni][e] = __ldcg(Ap + (size_t)cc * ld + rr);
          }
    }
    double bacc = 0.0, rhs_j = 0.0;   // tid < NB: sum_k (L_jk y_k)[tid], b_j
#pragma unroll
    for (int c = 0; c < 8; ++c) W.t[c][0] = W.t[c][1] = 0.0;
    // The diagonal tile A_jj belongs to the diagonal-update task D_j (C_0 has none and loads its own);
    // the chain task starts from W = 0, takes the updates of the last two columns itself and adds
    // D_j's partial tile before its sweep.
    const bool split = g.split != 0;
    if (diag && !split) {   // nothing to hand over: the chain task keeps its whole k loop
      BA_TRACE(t, 3);
      BA_TRACE(t, 5);
      continue;
    }
    const bool wk = diag || (chain && !split);      // this task's k loop updates W and the forward-substitution sum
    const bool loadW = diag || (chain && (j == 0 || !split));
    if (!DIST && loadW) {
      const double* Ajj = g.A + (size_t)(j * NB) * ld + (size_t)j * NB;
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int row = 8 * r + gq, col = 8 * c + 2 * t4 + e;
          const int hi = row > col ? row : col, lo = row > col ? col : row;
          if (c <= r) W.t[c][e] = __ldcg(Ajj + (size_t)lo * ld + hi);
        }
    }
    if (!DIST && chain && tid < NB) rhs_j = __ldcg(g.rhs + j * NB + tid);
    if (DIST && (loadW || chain)) {
      // diagonal tile (D_j / C_0) and right-hand side (chain): summed here, in rank order.  Both
      // kinds of task are grabbed well before their last operand exists, so these round trips are
      // off the critical path.
      for (int p = 0; p < g.world; ++p) {
        const double* Ajj = g.base[p] + dl.contrib + (size_t)(j * NB) * ld + (size_t)j * NB;
        double w[8][2];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int row = 8 * r + gq, col = 8 * c + 2 * t4 + e;
            const int hi = row > col ? row : col, lo = row > col ? col : row;
            w[c][e] = (loadW && c <= r) ? ld_peer_f64(Ajj + (size_t)lo * ld + hi) : 0.0;
          }
        const double rb = (chain && tid < NB) ? ld_peer_f64(g.base[p] + dl.contrib + ld * ld + j * NB + tid) : 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) { W.t[c][0] += w[c][0]; W.t[c][1] += w[c][1]; }
        rhs_j += rb;
      }
    }

    // ---- k loop over the finished columns k < pj:  P = L_{pi,k}, Q = L_{pj,k} ------------------
    // Steps k < kfull take complete tiles.  A panel or chain task forms  acc -= L_{pi,k} L_{pj,k}^T;
    // the diagonal-update task D_j  W -= L_jk L_jk^T  and the forward-substitution sum  L_jk y_k
    // (one operand tile per step).  The chain task C_j used to do all three: its k loop then took
    // 1.7x a panel task's and every panel task of column j idled that much behind it (the more
    // ranks share a column, the more columns the chain fell behind).
    auto issue = [&](int k) {
      double* P = buf + (size_t)(2 * (k & 1)) * kTileDoubles;
      stage_tile(P, g.A + (size_t)(k * NB) * ld + (size_t)pi * NB, ld);
      if (!diag) stage_tile(P + kTileDoubles, g.A + (size_t)(k * NB) * ld + (size_t)pj * NB, ld);
      cp_async_commit();
    };
    auto wait_k = [&](int k) {
      wait_flag<DIST>(g, &g.flags[(size_t)pi * T + k], epoch, s_t0);
      if (!diag) wait_flag<DIST>(g, &g.flags[(size_t)pj * T + k], epoch, s_t0);
      if (wk) wait_flag<DIST>(g, &yflag[k], epoch, s_t0);
    };
    // The tiles of step k+1 are prefetched while step k computes ONLY if they are already
    // published; otherwise step k runs first and the wait comes after it.  (Blocking on the flags
    // of step k+1 before computing step k put a whole extra tile product behind every late
    // operand -- the last operand of a chain task always is.)
    // The LAST operand column (k = pj - 1) is consumed column block by column block while its
    // two tiles are still being produced (their owners publish 8 flags per tile): the fat last
    // step of a chain task then finishes about one block after the tiles themselves, instead of
    // a whole two-tile product later.
    const int kfull = pj > 0 ? pj - 1 : 0;   // steps 0 .. kfull-1 take complete tiles
    int issued = 0;
    int ready_upto = -1;   // steps <= ready_upto are known to be published (CTA-uniform)
    for (int k = 0; k < kfull; ++k) {
      if (issued == k) {
        wait_k(k);
        __syncthreads();
        issue(k);
        issued = k + 1;
      }
      if (k + 1 < kfull) {
        if (k + 1 > ready_upto) {
          // (re)scan: warp 0 looks at the next eight steps at once (one lane per step), so that deep
          // tasks, whose operands are long finished, pay for a scan and a barrier once per eight
          // steps instead of every step
          if (wid == 0) {
            const int kk = k + 1 + lane;
            bool ok = false;
            if (lane < 8 && kk < kfull) {
              const unsigned int f0 = ldf_relaxed<DIST>(&g.flags[(size_t)pi * T + kk]);
              const unsigned int f1 = diag ? epoch : ldf_relaxed<DIST>(&g.flags[(size_t)pj * T + kk]);
              const unsigned int f2 = wk ? ldf_relaxed<DIST>(&yflag[kk]) : epoch;
              ok = f0 == epoch && f1 == epoch && f2 == epoch;
            }
            const unsigned int mask = __ballot_sync(0xffffffffu, ok) & 0xffu;
            if (lane == 0) {
              const unsigned int m = ~mask & 0xffu;
              const int n = m ? __ffs(m) - 1 : 8;      // consecutive ready steps from k + 1
              if (n > 0) { if (DIST && g.strict) __threadfence_system(); else __threadfence(); }   // acquire for what the relaxed loads saw
              s_task = k + n;
            }
          }
          __syncthreads();
          ready_upto = s_task;
        }
        if (k + 1 <= ready_upto) {
          issue(k + 1);
          issued = k + 2;
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
      } else {
        cp_async_wait<0>();
      }
      if (wk && tid < NB) yk[tid] = __ldcg(g.rhs + k * NB + tid);  // y_k
      __syncthreads();
      const double* P = buf + (size_t)(2 * (k & 1)) * kTileDoubles;
      if (!diag) tile_dmma<true>(acc, P, P + kTileDoubles, R0, C0, lane);
      if (wk) {
        diag_rows_dmma(W, P, r, lane);
        if (tid < NB) {
          double s = 0.0;
#pragma unroll 8
          for (int m = 0; m < NB; ++m) s += P[m * LDT + tid] * yk[m];
          bacc += s;
        }
      }
      __syncthreads();
    }
    if (diag) {
      // D_j is done: partial diagonal tile (lower triangle, [col][row]) and partial sum to the chain task
      double* Wp = g.Wpart + (size_t)j * (NB * NB + NB);
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (c <= r) Wp[(8 * c + 2 * t4 + e) * NB + 8 * r + gq] = W.t[c][e];
      if (tid < NB) Wp[NB * NB + tid] = bacc;
      __syncthreads();
      if (tid == 0) st_release(&g.flags[(size_t)T * T + j], epoch);
      BA_TRACE(t, 3);
      BA_TRACE(t, 5);
      if (tid == 0) BA_PROF_ADD(kProfDiagTask);
      continue;
    }
    if (pj > 0) {
      const int k = pj - 1;
      double* P = buf + (size_t)(2 * (k & 1)) * kTileDoubles;
      double* Q = P + kTileDoubles;
      const double* gP = g.A + (size_t)(k * NB) * ld + (size_t)pi * NB;
      const double* gQ = g.A + (size_t)(k * NB) * ld + (size_t)pj * NB;
      const unsigned int* fP = colflag + ((size_t)pi * T + k) * 8;
      const unsigned int* fQ = colflag + ((size_t)pj * T + k) * 8;
      // Three cursors (column blocks): P fetched up to pf, acc done up to ca (needs P and Q), and,
      // in a chain task, W done up to cw (needs P only).  Q of a chain task is the panel tile of the
      // PREVIOUS chain task, the last thing to arrive; P (a plain panel tile) and y_k are there a
      // little earlier, so the W half of the step and the y GEMV run while the task would
      // otherwise wait for Q, and only the acc half is left behind the last block of Q.
      int ca = 0, pf = 0, cw = chain ? 0 : 8;
      bool gemv_done = !chain;
#pragma unroll 1
      while (ca < 8 || cw < 8) {
        if (tid == 0) {
          unsigned int spins = 0;
          BA_PROF_T0();
          for (;;) {
            unsigned int f[8], h[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              f[q] = (q >= pf) ? ldf_relaxed<DIST>(fP + q) : epoch;
              h[q] = (q >= ca) ? ldf_relaxed<DIST>(fQ + q) : epoch;
            }
            const unsigned int fy = gemv_done ? 0u : ldf_relaxed<DIST>(&yflag[k]);
            int eP = 0, eQ = 0;
            bool runP = true, runQ = true;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              runP = runP && (f[q] == epoch);
              runQ = runQ && (h[q] == epoch);
              if (runP) eP = q + 1;
              if (runQ) eQ = q + 1;
            }
            const int eA = eP < eQ ? eP : eQ;
            const bool yr = !gemv_done && fy == epoch && eP == 8;
            if (eA > ca || eP > cw || yr) {
              // acquire for what the relaxed loads saw: only on request (g.strict).  The operands are
              // fetched from L2 with cp.async.cg after the flag load returned, and L2 is the point of
              // coherence for this GPU's memory whether the producer is an SM or a peer over NVLink
              // (DESIGN.md 4.2); a system-scope fence here cost ~20 % of the k loop.
              if (g.strict) { if (DIST) __threadfence_system(); else __threadfence(); }
              s_task = eP | (eQ << 4) | (yr ? 256 : 0);
              break;
            }
            if (spin_expired(g, spins, s_t0)) {
              s_task = 8 | (8 << 4) | (gemv_done ? 0 : 256);
              break;
            }
            __nanosleep(20);
          }
          BA_PROF_ADD(kProfLast);
        }
        __syncthreads();
        const int st = s_task;
        const int eP = st & 15, eQ = (st >> 4) & 15;
        const bool do_gemv = (st & 256) != 0;
        const int eA = eP < eQ ? eP : eQ;
        // columns of the tiles: 64 contiguous doubles each, 32 chunks of 16 B per column
        for (int ch = tid; ch < 32 * 8 * (eP - pf); ch += kSolveThreads) {
          const int m = 8 * pf + (ch >> 5), r2 = (ch & 31) * 2;
          cp_async16(P + m * LDT + r2, gP + (size_t)m * ld + r2);
        }
        for (int ch = tid; ch < 32 * 8 * (eA - ca); ch += kSolveThreads) {
          const int m = 8 * ca + (ch >> 5), r2 = (ch & 31) * 2;
          cp_async16(Q + m * LDT + r2, gQ + (size_t)m * ld + r2);
        }
        cp_async_commit();
        if (do_gemv && tid < NB) yk[tid] = __ldcg(g.rhs + k * NB + tid);
        cp_async_wait<0>();
        __syncthreads();
        if (eA > ca) tile_dmma<true>(acc, P, Q, R0, C0, lane, 8 * ca, 8 * eA);
        if (eP > cw) diag_rows_dmma(W, P, r, lane, 8 * cw, 8 * eP);
        if (do_gemv) {
          if (tid < NB) {
            double s = 0.0;
#pragma unroll 8
            for (int m = 0; m < NB; ++m) s += P[m * LDT + tid] * yk[m];
            bacc += s;
          }
          gemv_done = true;
        }
        pf = eP;
        if (cw < eP) cw = eP;
        if (ca < eA) ca = eA;
      }
      if (!gemv_done) {
        wait_flag<DIST>(g, &yflag[k], epoch, s_t0, kProfYflag);
        __syncthreads();
        if (tid < NB) {
          yk[tid] = __ldcg(g.rhs + k * NB + tid);
        }
        __syncthreads();
        if (tid < NB) {
          double s = 0.0;
#pragma unroll 8
          for (int m = 0; m < NB; ++m) s += P[m * LDT + tid] * yk[m];
          bacc += s;
        }
      }
      __syncthreads();
    }
    BA_TRACE(t, 3);   // k loop done
    double* const Ws = buf + 3 * kTileDoubles;   // chain: D_j's partial diagonal tile, [col*LDT + row]
    double vpart = 0.0;
    if (chain && j > 0 && split) {
      // D_j was grabbed right before this task and has less to do: its partial tile is (all but
      // always) there by now.  It is staged into the one operand buffer the panel phase leaves
      // alone and added to W after the fold, so the L2 latency hides behind the panel phase.
      wait_flag<DIST>(g, &g.flags[(size_t)T * T + j], epoch, s_t0, kProfDiagFlag);
      __syncthreads();
      const double* Wp = g.Wpart + (size_t)j * (NB * NB + NB);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int chunk = it * kSolveThreads + tid;
        const int m = chunk >> 5, r2 = (chunk & 31) * 2;
        cp_async16(Ws + m * LDT + r2, Wp + (size_t)m * NB + r2);
      }
      cp_async_commit();
      if (tid < NB) vpart = __ldcg(Wp + NB * NB + tid);
    }

    double* const Cs = buf;                       // [m][row] = C[row][m]      (A operand)
    double* const Bs = buf + kTileDoubles;        // [m][c]   = Linv_pj[c][m]  (B operand), filled block by block
    double* const Ls = buf + 2 * kTileDoubles;    // [m][row] = L_{pi,pj}[row][m]
    int last_cb = 0;   // first column block of the last group of the panel tile
    if (has_panel) {
      // ---- progressive panel:  L[:, 8cb..8cb+7] = C[:, 0..8cb+7] Linv[8cb..8cb+7, 0..8cb+7]^T ----
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int rr = R0 + 8 * mi + gq, cc = C0 + 8 * ni + 2 * t4 + e;
            Cs[cc * LDT + rr] = acc.v[mi][ni][e];
          }
      const double* LTp = g.LinvT + (size_t)pj * NB * NB;
      double* Lout = g.A + (size_t)(pj * NB) * ld + (size_t)pi * NB;
      const unsigned int* rf = rowflag + (size_t)pj * 8;
      // Row blocks are taken in GROUPS [cb, ce): whatever the owner of column pj has already
      // published is processed in one go (one fetch, one pipelined batch of DMMAs), so a task
      // that arrives late catches up at tensor-pipe speed instead of one block per round trip.
      // Chain tasks fold each finished column block of L_{j,j-1} into the diagonal tile LAZILY:
      // the next chain task waits for the panel tile, not for W, so while more rows of the
      // inverse are already there the panel goes first and the update of W (column blocks
      // [ub, cb)) is caught up on when the task would otherwise spin, or after the loop.
      auto fold = [&](int b0, int b1) { diag_rows_dmma(W, Ls, r, lane, 8 * b0, 8 * b1); };
      int cb = 0, ub = 0;
      bool have_snap = false;   // s_snap holds (or is about to hold) a snapshot newer than the last scan
      int pub_lo = -1, pub_hi = 0;   // column blocks whose bulk copies are in flight, flags not yet set (CTA-uniform)
      unsigned int pspins = 0;       // idle polls of the row-block flags (warp 1, lane 0)
      // column blocks [b0, b1) of Ls on their way out: at most one column per thread (issuing a bulk
      // copy costs ~25 ns and serialises inside a warp, so the issue is spread over the CTA)
      auto pub_issue = [&](int b0, int b1) {
        const int i = (lane << 3) | wid;
        if (i < 8 * (b1 - b0)) bulk_store_column(Lout, ld, Ls, 8 * b0 + i);
      };
      auto pub_flags = [&]() {   // after every thread's bulk_store_wait() and a barrier
        if (pub_lo >= 0) {
          // relaxed by default: the bulk copies have COMPLETED (bulk_store_wait, not .read), i.e. the
          // bytes are in L2, and consumers fetch from L2 after seeing the flag.  g.strict publishes
          // with a release instead (clean under the PTX memory model; ~0.6 us of stalled memory
          // traffic per group, tests/test_gpu_parity.py holds the two variants to the same bits)
          if (wid == 0 && lane >= pub_lo && lane < pub_hi) {
            unsigned int* f = colflag + ((size_t)pi * T + pj) * 8 + lane;
            if (g.strict) st_release(f, epoch); else st_relaxed(f, epoch);
          }
          pub_lo = -1;
        }
      };
#pragma unroll 1
      while (cb < 8) {
        // Warp 1 scans and fetches (warp 0 may still be inside the release that published the
        // previous group: the two round trips overlap).  The scan is one round trip: all flags at
        // once, relaxed, independent loads.  No acquire follows: the rows are read with
        // cp.async.cg, i.e. from L2, the point of coherence, after the flag load returned; the
        // owner's release made them visible there before the flag.  (An acquire load costs a
        // second round trip per group.)
        if (wid == 1) {
          int e = cb;
          if (have_snap) cp_async_wait<0>();
          if (lane == 0) {
            // first the snapshot that cp.async took while the previous group's DMMAs ran (no
            // round trip); a fresh scan only if it shows nothing new
            unsigned int f[8];
            bool fresh = !have_snap;
            if (have_snap) {
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = (q >= cb) ? s_snap[q] : epoch;
              fresh = s_snap[cb] != epoch;
            }
            if (fresh) {
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = (q >= cb) ? ldf_relaxed<DIST>(rf + q) : epoch;
            }
            bool run = true;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q >= cb) {
                run = run && (f[q] == epoch);
                if (run) e = q + 1;
              }
            // only when nothing is ready yet, spin on the next block
            if (e == cb) e = spin_expired(g, pspins, s_t0) ? 8 : -1;   // nothing new: flush the pending publication, fold pending column blocks into W
            else if (g.strict) { if (DIST) __threadfence_system(); else __threadfence(); }   // acquire for the rows the scan found
            s_task = e;
          }
          e = __shfl_sync(0xffffffffu, e, 0);
          // rows 8q .. 8q+7 of Linv are columns 8q .. of Bs, 8 (q + 1) entries deep: 16-byte chunks
          for (int q = cb; q < e; ++q)
            for (int ch = lane; ch < 4 * (8 * q + 8); ch += 32)
              cp_async16(Bs + (ch >> 2) * LDT + 8 * q + 2 * (ch & 3), LTp + (size_t)(ch >> 2) * NB + 8 * q + 2 * (ch & 3));
          cp_async_commit();
        }
        __syncthreads();
        const int ce = s_task;
        if (ce < 0) {   // CTA-uniform
          if (g.prof && tid == 32) atomicAdd(&g.prof[kProfPanel], 400ull);   // ~ one idle round (scan + sleep + barriers)
          have_snap = false;
          const bool idle = pub_lo < 0 && !(chain && ub < cb);
          if (pub_lo >= 0) {
            bulk_store_wait();
            __syncthreads();
            pub_flags();
          }
          if (chain && ub < cb) {
            fold(ub, cb);
            ub = cb;
          }
          if (idle) __nanosleep(40);
          __syncthreads();   // everybody has read s_task
          continue;
        }
        BA_GT(16 + 32 * (t - g_dbg_consumer) + 4 * cb + 0, (t == g_dbg_consumer || t == g_dbg_consumer + 1) && tid == 0);
#ifdef BA_SOLVE_TRACE
        if ((t == g_dbg_consumer || t == g_dbg_consumer + 1) && tid == 0) g_dbg_time[80 + 8 * (t - g_dbg_consumer) + cb] = ce;
#endif
        if (wid == 1) cp_async_wait<0>();
        bulk_store_wait();   // the previous group has had a scan and a fetch to land in L2
        __syncthreads();   // rows of blocks cb .. ce-1 of Linv (and, first time round, Cs) are in shared memory
        pub_flags();
        if (ce < 8 && !chain && !DIST) {   // snapshot of the flags for the next scan, taken under the DMMAs below
          // (panel tasks only: a chain task gains more from the larger groups a fresh scan finds)
          if (tid == 32) {
            cp_async16(s_snap, rf);
            cp_async16(s_snap + 4, rf + 4);
            cp_async_commit();
          }
          have_snap = true;
        }
        BA_GT(16 + 32 * (t - g_dbg_consumer) + 4 * cb + 1, (t == g_dbg_consumer || t == g_dbg_consumer + 1) && tid == 0);

        // warp w: the 8x8 tiles rows 8w.., column blocks cb .. ce-1.  Tile q contracts over the
        // 8 (q + 1) columns of Linv that are non-zero in its rows: two interleaved accumulator
        // chains, operands of the next step loaded before the DMMAs of this one are issued, no
        // predication anywhere near the tensor instructions.
        // Tiles are taken in PAIRS (q, q + 1) sharing the A operand -- four accumulator chains in
        // flight per warp (tile_bench: 3529 clk for all 8 blocks against 4398 one tile at a time)
        // -- and a leftover single tile with two.
        {
          const double* pa = Cs + t4 * LDT + 8 * wid + gq;
          const int row = 8 * wid + gq;
          int q = cb;
#pragma unroll 1
          for (; q + 1 < ce; q += 2) {
            const double* pb = Bs + t4 * LDT + 8 * q + gq;
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
            {   // the eight further columns of Linv that only tile q + 1 reaches
              const double a0 = pa[kend * LDT], a1 = pa[(kend + 4) * LDT];
              const double c0 = pb[kend * LDT + 8], c1 = pb[(kend + 4) * LDT + 8];
              dmma884(g0, g1, a0, c0);
              dmma884(h0, h1, a1, c1);
            }
            e0 += f0; e1 += f1; g0 += h0; g1 += h1;
            const int col = 8 * q + 2 * t4;
            Ls[col * LDT + row] = e0;
            Ls[(col + 1) * LDT + row] = e1;
            Ls[(col + 8) * LDT + row] = g0;
            Ls[(col + 9) * LDT + row] = g1;
            if (ce == 8) {   // last group: straight out, published by a release right after the loop
              Lout[(size_t)col * ld + row] = e0;
              Lout[(size_t)(col + 1) * ld + row] = e1;
              Lout[(size_t)(col + 8) * ld + row] = g0;
              Lout[(size_t)(col + 9) * ld + row] = g1;
            }
          }
          if (q < ce) {
            const double* pb = Bs + t4 * LDT + 8 * q + gq;
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
            const int col = 8 * q + 2 * t4;
            Ls[col * LDT + row] = e0;
            Ls[(col + 1) * LDT + row] = e1;
            if (ce == 8) {
              Lout[(size_t)col * ld + row] = e0;
              Lout[(size_t)(col + 1) * ld + row] = e1;
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // Ls is read by the bulk copies below
        BA_GT(16 + 32 * (t - g_dbg_consumer) + 4 * cb + 2, (t == g_dbg_consumer || t == g_dbg_consumer + 1) && tid == 0);
        __syncthreads();   // column blocks cb .. ce-1 of L_{pi,pj} are complete in Ls
        if (ce < 8) {
          // ... and on their way to global memory, for the tasks that consume this tile column block
          // by column block.  Nobody waits here: the copies are waited for, and the flags set, one
          // scan and one fetch later, when they have all but landed.  The last group goes out
          // together with the tile flag after the loop.
          pub_issue(cb, ce);
          pub_lo = cb;
          pub_hi = ce;
        }
        last_cb = cb;
        BA_GT(16 + 32 * (t - g_dbg_consumer) + 4 * cb + 3, (t == g_dbg_consumer || t == g_dbg_consumer + 1) && tid == 0);
        cb = ce;
      }
      // tile (pi, pj) of L is complete.  The last group was stored directly; warp 1 releases it
      // (warp 0 starts the sweep of a chain task, warp 7 has the most of W left to fold; the
      // fence stalls global traffic only, and the rest of a chain task works out of shared memory).
      if (chain && split) cp_async_wait<0>();   // D_j's partial tile (Ws)
      __syncthreads();
      if (wid == 1) {
        if (lane >= last_cb && lane < 8) st_release(colflag + ((size_t)pi * T + pj) * 8 + lane, epoch);
        if (lane == 0) st_release(&g.flags[(size_t)pi * T + pj], epoch);
      }
      // DIST: the whole tile, still in Ls, goes to every peer with plain 16-byte stores: a warp store
      // is one 512-byte column, 8 columns per warp and peer, fire and forget.  (448 bulk copies per
      // tile on the TMA path took ~70 us to ISSUE at 8 ranks.)  A plain panel task does it right here
      // and raises the tile's flags on the peers; a CHAIN task keeps every remote store for its very
      // end (push_tile_to_peers after the sweep): a release at gpu scope -- the local row-block flags
      // of the sweep -- waits for the acknowledgement of every store the warp has in flight, and
      // with stores on the links that is a loaded NVLink round trip (~9 us) per release on the
      // critical path of the factorisation.
      if (DIST && !chain) push_tile_to_peers();
      if (chain && ub < 8) fold(ub, 8);
      if (chain && split) {   // (a chain task with a panel has j > 0)
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (c <= r) W.t[c][e] += Ws[(8 * c + 2 * t4 + e) * LDT + 8 * r + gq];
        bacc += vpart;
      }
      BA_TRACE(t, 6);   // panel part done
    }
    // warp 1, after a barrier behind every warp's stores of the tile to the peers: release its 8
    // column-block flags and the tile flag on every peer (system scope; the release is cumulative
    // over the other warps' stores that the barrier ordered before it)
    auto peer_tile_flags = [&]() {
      BA_PROF_T0();
      __threadfence_system();
      __syncwarp();
      for (int p = 0; p < g.world; ++p)
        if (p != g.rank && lane < 9) {
          unsigned int* pf = reinterpret_cast<unsigned int*>(g.base[p] + dl.flags);
          unsigned int* f = lane < 8 ? pf + solve_rowflag_base(T) + 8 * T + ((size_t)pi * T + pj) * 8 + lane : pf + (size_t)pi * T + pj;
          st_relaxed_sys(f, epoch);
        }
      if (lane == 0) BA_PROF_ADD(kProfPush);
    };
    if (DIST && has_panel && !chain) {
      __syncthreads();
      if (wid == 1) peer_tile_flags();
    }

    if (chain) {
      double* const LT = g.LinvT + (size_t)j * NB * NB;
      // ---- blocked sweep: 8 panel steps of 8 pivots ------------------------------------------
      // Step pb:  warp 0 (whose row block holds no trailing tiles) factors the 8x8 pivot block
      // D = W(pb,pb) by Gauss-Jordan in registers into Linv_d = L_d^{-1} and shares it through
      // shared memory;  warp r > pb forms its panel tile  Lp_r = W(r,pb) Linv_d^T,  warp pb the
      // finished row block of the inverse  Mp_c = Linv_d M(pb,c);  then the trailing tiles take
      // the rank-8 update  W(r,c) -= Lp_r Lp_c^T,  M(r,c) -= Lp_r Mp_c  on the FP64 tensor pipe.
      // M accumulates L~^{-1} exactly as a Gauss-Jordan sweep of [W | I] would, so the row
      // blocks Mp are the rows of L_jj^{-1}; L_jj itself is never needed.
      // Lookahead: the owner of row block pb+1 updates the next pivot block FIRST and hands it
      // to warp 0 through a 64-thread named barrier, so the latency-bound Gauss-Jordan of step
      // pb+1 runs concurrently with the rest of the trailing update of step pb.
      double* const LTs = buf;                          // [m][c'] = Linv[c'][m], row stride LDT
      double* const Wcol = buf + kTileDoubles;          // [8][TD]  column block pb of W
      double* const Mrow = Wcol + 8 * TD;               // [8][TD]  row block pb of M
      double* const Lp = Mrow + 8 * TD;                 // [8][TD]  panel tiles
      double* const Mp = Lp + 8 * TD;                   // [8][TD]  row block pb of Linv
      double* const Dn = Mp + 8 * TD;                   // [TD]     next pivot block, [row][col]
      double* const Ld = Dn + TD;                       // [TD]     Linv_d, [row][col]
      BA_CLK(33);
      RowTiles M;
#pragma unroll
      for (int c = 0; c < 8; ++c) M.t[c][0] = M.t[c][1] = 0.0;
      if (tid == 0) s_bad = 0;
      if (r == 0) {
        *reinterpret_cast<double2*>(Dn + gq * TS + 2 * t4) = make_double2(W.t[0][0], W.t[0][1]);
        __syncwarp();
      } else {
        *reinterpret_cast<double2*>(Wcol + r * TD + gq * TS + 2 * t4) = make_double2(W.t[0][0], W.t[0][1]);
      }
#pragma unroll
      for (int pb = 0; pb < 8; ++pb) {
        double l0 = 0.0, l1 = 0.0;   // Linv_d[gq][t4], Linv_d[gq][t4 + 4]: the DMMA operand layout
        double mc0[8], mc1[8];       // warp pb: its finished row block of the inverse, tiles c < pb
        if (r == 0) {
          if (pb > 0) asm volatile("bar.sync 1, 64;" ::: "memory");   // D_pb is in Dn
          BA_CLK0(pb * 4 + 0);
          // Lane (gq, t4) holds columns t4 and t4 + 4 of row gq of D and of the accumulated
          // inverse M.  Elimination by 2x2 pivot blocks: one reciprocal (of the block
          // determinant) sits on the dependency chain per TWO pivots.  After the four block
          // steps M D M^T is block diagonal with the 2x2 Schur complements [[a, b], [b, c]];
          // their Cholesky inverses  [[1/sqrt(a), 0], [-b/(a s), 1/s]],  s = sqrt(det / a),
          // are folded into the rows afterwards, off the chain.
          double w0 = Dn[gq * TS + t4], w1 = Dn[gq * TS + 4 + t4];
          double m0 = (gq == t4) ? 1.0 : 0.0, m1 = (gq == t4 + 4) ? 1.0 : 0.0;
          bool bad = false;
          double pa = 1.0, pb_ = 0.0, pdet = 1.0;   // this lane's row: its block's a, b, det
          const int quad = lane & 28;
#pragma unroll
          for (int p = 0; p < 8; p += 2) {
            const double wh = (p < 4) ? w0 : w1;        // the half that holds columns p, p + 1
            double a = __shfl_sync(0xffffffffu, wh, 4 * p + (p & 3));
            const double b = __shfl_sync(0xffffffffu, wh, 4 * p + ((p + 1) & 3));
            const double c = __shfl_sync(0xffffffffu, wh, 4 * (p + 1) + ((p + 1) & 3));
            const double wgp = __shfl_sync(0xffffffffu, wh, quad | (p & 3));          // D[gq][p]
            const double wgp1 = __shfl_sync(0xffffffffu, wh, quad | ((p + 1) & 3));   // D[gq][p+1]
            double det = a * c - b * b;
            if (!(a >= 1e-290) || !(det >= 1e-290 * a)) { bad = true; a = 1.0; det = 1.0; }  // warp-uniform; catches NaN
            if ((gq >> 1) == (p >> 1)) { pa = a; pb_ = b; pdet = det; }
            const double idet = pivot_rcp(det);
            const bool below = gq > p + 1;
            const double f0 = below ? (wgp * c - wgp1 * b) * idet : 0.0;
            const double f1 = below ? (wgp1 * a - wgp * b) * idet : 0.0;
            // rows p and p + 1, this lane's columns (columns <= p + 1 of W are dead, M is zero
            // beyond column p + 1: those halves are skipped statically)
            if (p < 2) {
              const double u0 = __shfl_sync(0xffffffffu, w0, 4 * p + t4), u1 = __shfl_sync(0xffffffffu, w0, 4 * (p + 1) + t4);
              w0 -= f0 * u0 + f1 * u1;
            }
            if (p < 6) {
              const double u0 = __shfl_sync(0xffffffffu, w1, 4 * p + t4), u1 = __shfl_sync(0xffffffffu, w1, 4 * (p + 1) + t4);
              w1 -= f0 * u0 + f1 * u1;
            }
            {
              const double u0 = __shfl_sync(0xffffffffu, m0, 4 * p + t4), u1 = __shfl_sync(0xffffffffu, m0, 4 * (p + 1) + t4);
              m0 -= f0 * u0 + f1 * u1;
            }
            if (p >= 4) {
              const double u0 = __shfl_sync(0xffffffffu, m1, 4 * p + t4), u1 = __shfl_sync(0xffffffffu, m1, 4 * (p + 1) + t4);
              m1 -= f0 * u0 + f1 * u1;
            }
          }
          if (bad && lane == 0) s_bad = 1;
          BA_CLK0(pb * 4 + 1);
          {
            const double isa = rsqrt(pa), isd = rsqrt(pdet);   // independent of each other
            const double is = isd * (pa * isa);                // 1 / sqrt(det / a)
            const double tb = is * pb_ * (isa * isa);
            const bool odd = gq & 1;
            const double up0 = __shfl_up_sync(0xffffffffu, m0, 4), up1 = __shfl_up_sync(0xffffffffu, m1, 4);   // row gq - 1
            l0 = odd ? is * m0 - tb * up0 : isa * m0;
            l1 = odd ? is * m1 - tb * up1 : isa * m1;
          }
          Ld[gq * TS + t4] = l0;
          Ld[gq * TS + 4 + t4] = l1;
          BA_CLK0(pb * 4 + 2);
        }
        // (X) Linv_d is in place.  Only the warps that still have work in the sweep take part in
        // its barriers (named barriers 2 and 3, participant counts known at compile time): a warp
        // whose row block is finished leaves for the barrier after the loop, so the fence it
        // issues to publish its rows of the inverse never holds the others up.
        if (r == 0 || r >= pb) {
          const int kX = 32 * (pb == 0 ? 8 : 1 + 8 - pb);
          asm volatile("bar.sync 2, %0;" ::"r"(kX) : "memory");
        }
        if (r >= pb && r > 0) {
          l0 = Ld[gq * TS + t4];
          l1 = Ld[gq * TS + 4 + t4];
        }
        if (r == pb) {
          // finished row block pb of the inverse:  Mp_c = Linv_d M(pb, c), c < pb;  Mp_pb = Linv_d.
          // All the products first (independent accumulators: the tensor pipe overlaps them), then
          // Mp, the operand of the update everybody waits for at (Y); the copies for the forward
          // substitution and for the consumers of this column are written after (Y).
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c >= pb) break;
            mc0[c] = 0.0; mc1[c] = 0.0;
            const double b0 = Mrow[c * TD + t4 * TS + gq], b1 = Mrow[c * TD + (4 + t4) * TS + gq];
            dmma884(mc0[c], mc1[c], l0, b0);
            dmma884(mc0[c], mc1[c], l1, b1);
          }
          Mp[pb * TD + gq * TS + t4] = l0;
          Mp[pb * TD + gq * TS + 4 + t4] = l1;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c >= pb) break;
            *reinterpret_cast<double2*>(Mp + c * TD + gq * TS + 2 * t4) = make_double2(mc0[c], mc1[c]);
          }
        } else if (r > pb) {
          double c0 = 0.0, c1 = 0.0;
          dmma884(c0, c1, Wcol[r * TD + gq * TS + t4], l0);
          dmma884(c0, c1, Wcol[r * TD + gq * TS + 4 + t4], l1);
          *reinterpret_cast<double2*>(Lp + r * TD + gq * TS + 2 * t4) = make_double2(c0, c1);
        }
        // (Y) Lp, Mp in place; Wcol / Mrow / Ld free again  (warps r >= pb; warp 0 only at pb = 0)
        if (r >= pb) {
          const int kY = 32 * (8 - pb);
          if (kY > 32) asm volatile("bar.sync 3, %0;" ::"r"(kY) : "memory");
          else __syncwarp();
        }
        BA_CLK(pb * 4 + 3);
        if (r == pb) {
          // LTs (forward substitution) and straight out to global memory (the tiles above the
          // diagonal of LinvT are zero from allocation and are never written)
          LTs[(8 * pb + t4) * LDT + 8 * pb + gq] = l0;
          LTs[(8 * pb + 4 + t4) * LDT + 8 * pb + gq] = l1;
          LT[(8 * pb + t4) * NB + 8 * pb + gq] = l0;
          LT[(8 * pb + 4 + t4) * NB + 8 * pb + gq] = l1;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c >= pb) break;
            LTs[(8 * c + 2 * t4) * LDT + 8 * pb + gq] = mc0[c];
            LTs[(8 * c + 2 * t4 + 1) * LDT + 8 * pb + gq] = mc1[c];
            LT[(8 * c + 2 * t4) * NB + 8 * pb + gq] = mc0[c];
            LT[(8 * c + 2 * t4 + 1) * NB + 8 * pb + gq] = mc1[c];
          }
        }
        if (r == pb && pb > 0) {
          // rows 8 pb .. 8 pb + 7 of L_jj^{-1} are final and this warp has nothing left to do in
          // the sweep: make them visible and let the consumers of this column start (the fence
          // is kept out of the X..Y window, where the whole CTA would wait for it).  Row block 0
          // belongs to warp 0, whose next Gauss-Jordan step is the critical path: warp 1
          // publishes it together with row block 1.
          __syncwarp();
          if (pb == 1 && lane == 1) st_release(&rowflag[(size_t)j * 8 + 0], epoch);
          if (lane == 0) st_release(&rowflag[(size_t)j * 8 + pb], epoch);
          BA_GT(pb, t == g_dbg_producer && lane == 0);
        }
        if (r > pb) {
          const double a0 = -Lp[r * TD + gq * TS + t4], a1 = -Lp[r * TD + gq * TS + 4 + t4];
          if (pb + 1 < 8) {
            // next pivot column first: publish it, and hand the next pivot block to warp 0
            const int c = pb + 1 < 8 ? pb + 1 : 7;   // == pb + 1 (kept in range for the unroller)
            dmma884(W.t[c][0], W.t[c][1], a0, Lp[c * TD + gq * TS + t4]);
            dmma884(W.t[c][0], W.t[c][1], a1, Lp[c * TD + gq * TS + 4 + t4]);
            if (r == pb + 1) {
              *reinterpret_cast<double2*>(Dn + gq * TS + 2 * t4) = make_double2(W.t[c][0], W.t[c][1]);
              __threadfence_block();
              asm volatile("bar.arrive 1, 64;" ::: "memory");
            } else {
              *reinterpret_cast<double2*>(Wcol + r * TD + gq * TS + 2 * t4) = make_double2(W.t[c][0], W.t[c][1]);
            }
          }
          // the rest: all operands first, then two rounds of independent DMMAs
          double b0[8], b1[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            b0[c] = b1[c] = 0.0;
            if (c == pb + 1 || c > r) continue;
            const double* q = (c > pb) ? Lp + c * TD + gq * TS + t4 : Mp + c * TD + t4 * TS + gq;
            b0[c] = q[0];
            b1[c] = q[(c > pb) ? 4 : 4 * TS];
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c == pb + 1) continue;
            if (c > pb) { if (c <= r) dmma884(W.t[c][0], W.t[c][1], a0, b0[c]); }
            else dmma884(M.t[c][0], M.t[c][1], a0, b0[c]);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c == pb + 1) continue;
            if (c > pb) { if (c <= r) dmma884(W.t[c][0], W.t[c][1], a1, b1[c]); }
            else dmma884(M.t[c][0], M.t[c][1], a1, b1[c]);
          }
          if (r == pb + 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c <= pb) *reinterpret_cast<double2*>(Mrow + c * TD + gq * TS + 2 * t4) = make_double2(M.t[c][0], M.t[c][1]);
          }
        }
        BA_CLK(36 + pb * 3);
      }
      __syncthreads();
      BA_CLK(32);
      BA_TRACE(t, 4);   // sweep done
      const bool bad = s_bad != 0;
      if (bad) {   // leave an identity behind so that dependants stay finite
        if (tid == 0 && ld_relaxed_sys(g.abort) == 0u) *g.status = 1.0;   // (an aborted launch computes garbage: it stays "timed out")
        for (int e = tid; e < NB * NB; e += kSolveThreads) {
          const double v = ((e >> 6) == (e & 63)) ? 1.0 : 0.0;
          LTs[(e >> 6) * LDT + (e & 63)] = v;
          LT[e] = v;
        }
        if (DIST && tid == 0)
          for (int p = 0; p < g.world; ++p)
            if (p != g.rank) *(g.base[p] + dl.status) = 1.0;   // every rank reports the failed pivot (flagged with the tile below)
        __threadfence();
      }
      // forward substitution: y_j = Linv (b_j - sum_{k<j} L_jk y_k); the k = j-1 term comes from
      // the panel tile this task produced itself (still in Ls)
      if (j > 0) {
        wait_flag<DIST>(g, &yflag[j - 1], epoch, s_t0, kProfYflag);
        __syncthreads();
        if (tid < NB) yk[tid] = __ldcg(g.rhs + (j - 1) * NB + tid);
        __syncthreads();
        if (tid < NB) {
          double s = 0.0;
#pragma unroll 8
          for (int m = 0; m < NB; ++m) s += Ls[m * LDT + tid] * yk[m];
          bacc += s;
        }
      }
      double yval = 0.0;
      if (tid < NB) tvec[tid] = rhs_j - bacc;
      __syncthreads();
      // (the 8x8 tiles of LTs above the block diagonal were never written: stop at the diagonal tile)
      if (tid < NB) {
        double s = 0.0;
        const int mend = 8 * ((tid >> 3) + 1);
#pragma unroll 8
        for (int m = 0; m < mend; ++m) s += LTs[m * LDT + tid] * tvec[m];
        g.rhs[j * NB + tid] = s;
        yval = s;
      }
      __syncthreads();
      if (tid == 0) st_release(&yflag[j], epoch);
      if (DIST) {
        // Everything this chain task produced goes to the peers NOW, behind the local critical path:
        // its panel tile (Ls), the inverse of the diagonal tile (lower block triangle of LTs; what
        // lies above was never written here and stays zero over there) and y_j.
        BA_PROF_T0();
        if (has_panel) push_tile_to_peers();
        for (int p = 0; p < g.world; ++p) {
          if (p == g.rank) continue;
          double* PT = g.base[p] + dl.LinvT + (size_t)j * NB * NB;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int m = 8 * wid + q;
            if ((lane >> 2) >= wid)    // columns 2 lane, 2 lane + 1 lie in a block on or below the diagonal block of row m
              *reinterpret_cast<double2*>(PT + (size_t)m * NB + 2 * lane) = *reinterpret_cast<const double2*>(LTs + m * LDT + 2 * lane);
          }
          if (tid < NB) (g.base[p] + dl.L + ld * ld)[j * NB + tid] = yval;   // every rank substitutes backwards on its own
        }
        __syncthreads();
        if (wid == 1) {
          __threadfence_system();
          __syncwarp();
          for (int p = 0; p < g.world; ++p) {
            if (p == g.rank) continue;
            unsigned int* pf = reinterpret_cast<unsigned int*>(g.base[p] + dl.flags);
            if (has_panel && lane < 8) st_relaxed_sys(pf + solve_rowflag_base(T) + 8 * T + ((size_t)pi * T + pj) * 8 + lane, epoch);
            if (has_panel && lane == 8) st_relaxed_sys(pf + (size_t)pi * T + pj, epoch);
            if (lane >= 16 && lane < 24) st_relaxed_sys(pf + solve_rowflag_base(T) + (size_t)j * 8 + (lane - 16), epoch);
            if (lane == 24) st_relaxed_sys(pf + (size_t)T * T + T + j, epoch);
          }
          if (lane == 0) BA_PROF_ADD(kProfPush);
        }
      }
    }
    BA_TRACE(t, 5);   // published
    if (tid == 0) BA_PROF_ADD(chain ? kProfChainTask : kProfTask);
  }

  // ==================================== backward substitution ==============================
  // x_k = L_kk^{-T} (y_k - sum_{i>k} L_ik^T x_i),  k = T-1 .. 0
  double* const part = buf;            // [4][NB] partial sums
  double* const accv = vec;            // [NB]
  double* const LTs = buf + 8 * NB;    // [NB][NBP] padded copy of LinvT
  const bool factor_done = !DIST && g.phase == kPhaseBackward;   // earlier launches finished the factor: nothing to wait for but the x_i
  for (;;) {
    if (!DIST && g.phase == kPhaseWindow) break;
    __syncthreads();
    if (tid == 0) s_task = (int)atomicAdd(&g.tickets[1], 1u);
    __syncthreads();
    const int bt = s_task;
    if (bt >= T) break;
    const int k = T - 1 - bt;
    BA_TRACE_SET(ntasks + bt, 0, (unsigned long long)k);
    BA_TRACE_SET(ntasks + bt, 1, (unsigned long long)blockIdx.x);
    BA_TRACE(ntasks + bt, 2);
    // everything this task reads except the x_i is long finished when it starts: wait for all of
    // it at once, fetch L_kk^{-1} and y_k, and keep the NEXT tile's share in registers so that only
    // the flag hop and the 512-byte x_i sit between x_{k+1} becoming ready and x_k going out
    if (!factor_done) {
      wait_flag<DIST>(g, &g.flags[solve_rowflag_base(T) + (size_t)k * 8 + 7], epoch, s_t0, kProfBackward);   // L_kk^{-1}
      wait_flag<DIST>(g, &g.flags[(size_t)T * T + T + k], epoch, s_t0, kProfBackward);                        // y_k
    }
    // tiles (i, k), i > k: one flag per thread, polled side by side (a single thread walking down a
    // column of 188 flags is 188 dependent L2 round trips in front of every task)
    for (int i0 = T - 1; i0 > k && !factor_done; i0 -= kSolveThreads) {
      const int i = i0 - tid;
      if (i > k) {
        const unsigned int* f = &g.flags[(size_t)i * T + k];
        unsigned int spins = 0;
        while (ldf_relaxed<DIST>(f) != epoch) {
          if (spin_expired(g, spins, s_t0)) break;
          __nanosleep(40);
        }
      }
    }
    if (!DIST) __threadfence(); else if (g.strict) __threadfence_system();   // acquire for the tiles (read with ld.cg below)
    __syncthreads();
    {
      const double* LT = g.LinvT + (size_t)k * NB * NB;
      for (int e = tid; e < NB * NB; e += kSolveThreads) LTs[(e >> 6) * NBP + (e & 63)] = __ldcg(LT + e);
    }
    const double yk_mine = (tid < NB) ? __ldcg(g.rhs + k * NB + tid) : 0.0;
    // warp w owns columns 8w .. 8w+7 of every tile; lanes span rows
    double cs[8], la[8], lb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cs[q] = 0.0;
    auto fetch_tile = [&](int i) {
      const double* Lik = g.A + (size_t)(k * NB + 8 * wid) * ld + (size_t)i * NB;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        la[q] = __ldcg(Lik + (size_t)q * ld + lane);
        lb[q] = __ldcg(Lik + (size_t)q * ld + 32 + lane);
      }
    };
    if (k < T - 1) fetch_tile(T - 1);
    for (int i = T - 1; i > k; --i) {
      // x_i is its own flag: every warp polls the 64 values until none is the "not yet" pattern
      // (8-byte stores are single-copy atomic; no fence on the producer, one round trip here,
      // no CTA barrier)
      double x0, x1;
      unsigned int xspins = 0;
      for (;;) {
        x0 = ld_relaxed_f64(g.x + i * NB + lane);
        x1 = ld_relaxed_f64(g.x + i * NB + 32 + lane);
        const bool ok = __double_as_longlong(x0) != kNotYet && __double_as_longlong(x1) != kNotYet;
        if (__all_sync(0xffffffffu, ok)) break;
        bool give_up = false;
        if (lane == 0) give_up = spin_expired(g, xspins, s_t0);
        if (__shfl_sync(0xffffffffu, (int)give_up, 0)) break;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) cs[q] += la[q] * x0 + lb[q] * x1;
      if (i - 1 > k) fetch_tile(i - 1);
    }
    BA_TRACE(ntasks + bt, 3);   // all x_i folded
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double s = warp_sum(cs[q]);
      if (lane == 0) accv[8 * wid + q] = s;
    }
    __syncthreads();
    if (tid < NB) accv[tid] = yk_mine - accv[tid];
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
    if (tid < NB) {
      double xv = part[tid] + part[NB + tid] + part[2 * NB + tid] + part[3 * NB + tid];
      if (__double_as_longlong(xv) == kNotYet) xv = __longlong_as_double(0x7ff8000000000000LL);   // (cannot happen: keeps a NaN solve from hanging)
      st_relaxed_f64(g.x + k * NB + tid, xv);
    }
    if (DIST && k == 0 && tid == 0) {   // a pivot that failed on the chain's rank fails the solve on every rank
      if (*reinterpret_cast<volatile double*>(g.base[g.rank] + dl.status) != 0.0 && *g.status == 0.0) *g.status = 1.0;
    }
    BA_TRACE(ntasks + bt, 5);
  }
  if (g.prof && tid == 0) atomicAdd(&g.prof[kProfKernel], global_ns() - s_t0);
}

// ------------------------------------------------------------------------------------------
// Tile ownership of the distributed solve (deterministic, the same on every rank).  Rank 0 keeps
// the chain tasks and the band 2 <= i - j <= band next to the diagonal: the critical path
// C_j -> (j+2, j) -> C_{j+1} never crosses a link.  What is left of row i (tasks (i, j), j < i - band)
// goes to ONE rank, so L_ij -> task (i, j+1) stays a local, progressive hand-over; rows are dealt
// longest first to the least loaded rank (cost ~ one k step per finished column + the panel phase).
std::vector<int> dist_task_list(int T, int world, int rank, int band) {
  std::vector<double> load(world, 0.0);
  std::vector<int> owner(T, 0);
  auto cost = [](int i, int j) { return (i == j) ? 2.0 * (j > 0 ? j - 1 : 0) + 8.0 : (double)j + 2.0; };
  for (int j = 0; j < T; ++j) load[0] += cost(j, j) + (j > 2 ? 0.6 * (j - 2) : 0.0);   // chain + diagonal-update tasks
  std::vector<std::pair<double, int>> rows;
  for (int i = 0; i < T; ++i) {
    double rc = 0.0;
    for (int j = 0; j + 1 < i; ++j) {
      if (i - j <= band) load[0] += cost(i, j);
      else rc += cost(i, j);
    }
    rows.push_back(std::make_pair(-rc, i));
  }
  std::sort(rows.begin(), rows.end());
  for (const auto& rw : rows) {
    int best = 0;
    for (int r = 1; r < world; ++r)
      if (load[r] < load[best]) best = r;
    owner[rw.second] = best;
    load[best] += -rw.first;
  }
  std::vector<int> mine;
  if (rank == 0) mine.push_back(0);   // C_0
  for (int jc = 0; jc + 1 < T; ++jc) {
    if (rank == 0) mine.push_back(kDiagTask | (jc + 1));          // D_{jc+1}
    if (rank == 0) mine.push_back(((jc + 1) << 16) | (jc + 1));   // C_{jc+1}
    for (int i = jc + 2; i < T; ++i) {
      const int o = (i - jc <= band) ? 0 : owner[i];
      if (o == rank) mine.push_back((i << 16) | jc);
    }
  }
  return mine;
}

// Large systems on few ranks: all-reducing the packed system and running the blocked tcgen05 solve
// (ba_solve_tc.cuh) on every rank beats the distributed DMMA solve (2,000 cameras on 2 ranks: 9.9 ms
// + the all-reduce against 15.0 ms); from 4 ranks on the distributed solve wins.
static bool tc_preferred_over_dist(const Context& c) {
  return c.tc_min_tiles > 0 && c.ld / NB >= c.tc_min_tiles && c.ld / NB > c.tc_window + 1 && c.comm_world <= c.tc_over_dist_max_world;
}

bool dist_solve_selected(const Context& c) {
  return c.comm_world > 1 && c.comm_buf && c.dist_off != 0 && c.sys_state == kSysLocal && c.dist_min_tiles > 0 &&
         c.ld / NB >= c.dist_min_tiles && c.ld / NB < 16384 && !tc_preferred_over_dist(c);
}

static cudaError_t launch_solve_dist(Context& c, bool have_mask, cudaStream_t st) {
  const int ld = c.ld, T = ld / NB;
  const DistLayout dl = dist_layout(ld);
  cudaError_t e;
  if (!c.dist_tasks) {
    const std::vector<int> mine = dist_task_list(T, c.comm_world, c.comm_rank, c.dist_band);
    if ((e = cudaMalloc((void**)&c.dist_tasks, (mine.size() + 1) * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(c.dist_tasks, mine.data(), mine.size() * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;   // `mine` goes out of scope
    c.dist_ntasks = (int)mine.size();
  }
  double* base = c.comm_buf + c.dist_off;
  // this rank's contribution, dense: masked / padding parameters get their unit pivot from rank 0 only
  expand_system_kernel<<<ld, 256, 0, st>>>(c.sys, c.n_opt_cam, c.n_sys, ld, c.cam_mask, have_mask, base + dl.contrib,
                                           base + dl.contrib + (size_t)ld * ld, c.solve_tickets, &c.scalars->status, c.dC,
                                           reinterpret_cast<unsigned int*>(base + dl.abort), base + dl.status,
                                           c.comm_rank == 0 ? 1.0 : 0.0);
  c.launches += 1;
  if (!c.dist_attr_set) {
    if ((e = cudaFuncSetAttribute(chol_dataflow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSolveSmemBytes)) != cudaSuccess) return e;
    c.dist_attr_set = true;
  }
  CholArgs g;
  memset(&g, 0, sizeof g);
  g.A = base + dl.L;
  g.rhs = g.A + (size_t)ld * ld;
  g.x = c.dC;
  g.LinvT = base + dl.LinvT;
  g.Wpart = base + dl.wpart;
  g.flags = reinterpret_cast<unsigned int*>(base + dl.flags);
  g.tickets = c.solve_tickets;
  g.status = &c.scalars->status;
  g.ld = ld; g.T = T;
  g.epoch = ++c.dist_epoch;   // collective: every rank calls the distributed solve the same number of times
  g.abort = reinterpret_cast<unsigned int*>(base + dl.abort);
  g.spin_limit_ns = (unsigned long long)(c.spin_timeout_ms * 1e6);
  g.prof = c.solve_prof_on ? c.solve_prof : nullptr;
  g.split = T >= c.split_min_tiles;
  g.strict = c.strict_flags;
  g.world = c.comm_world; g.rank = c.comm_rank;
  g.tasks = c.dist_tasks; g.ntasks = c.dist_ntasks;
  for (int p = 0; p < kMaxPeers; ++p) g.base[p] = p < c.comm_world ? c.comm_peer[p] + c.dist_off : nullptr;
#ifdef BA_SOLVE_TRACE
  g.trace = c.solve_trace;
#endif
  int grid = c.num_sms;   // every rank runs the backward substitution: at least T tasks everywhere
  if (c.solve_grid_cap > 0 && grid > c.solve_grid_cap) grid = c.solve_grid_cap;
  chol_dataflow_kernel<true><<<grid, kSolveThreads, kSolveSmemBytes, st>>>(g);
  c.launches += 1;
  return cudaGetLastError();
}

}  // namespace ba
#include "ba_solve_tc.cuh"
namespace ba {

// ------------------------------------------------------------------------------------------
// Blocked solve, trailing updates on tcgen05 (ba_solve_tc.cuh has the algorithm).
bool tc_solve_selected(const Context& c) {
  return c.tc_min_tiles > 0 && c.ld / NB >= c.tc_min_tiles && c.ld / NB > c.tc_window + 1 && !dist_solve_selected(c);
}

enum { kTcProfExpand = 0, kTcProfPanels = 1, kTcProfSlices = 2, kTcProfUpdates = 3, kTcProfBackward = 4 };

// An event behind the launches issued so far; the time since the previous event is booked on `cat`.
static void tc_mark(Context& c, int cat, cudaStream_t st) {
  if (!c.solve_prof_on) return;
  if (c.tc_ev_used == (int)c.tc_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    c.tc_ev.push_back(e);
    c.tc_ev_cat.push_back(0);
  }
  c.tc_ev_cat[c.tc_ev_used] = cat;
  cudaEventRecord(c.tc_ev[c.tc_ev_used], st);
  c.tc_ev_used += 1;
}

void tc_fold_profile(Context& c) {
  if (c.tc_ev_used > 1 && cudaEventSynchronize(c.tc_ev[c.tc_ev_used - 1]) == cudaSuccess) {
    for (int i = 1; i < c.tc_ev_used; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c.tc_ev[i - 1], c.tc_ev[i]) == cudaSuccess) c.tc_prof_ms[c.tc_ev_cat[i]] += ms;
    }
    c.tc_prof_solves += 1;
  }
  c.tc_ev_used = 0;
}

cudaError_t tc_prepare(Context& c) {
  const int ld = c.ld, w = c.tc_window, K = NB * w, S = c.tc_slices_n;
  if (c.tc_slices && c.tc_cfg[0] == S && c.tc_cfg[1] == w && c.tc_cfg[2] == c.tc_bk && c.tc_cfg[3] == ld) return cudaSuccess;
  if (S < 4 || S > 7 || w < 2 || w > tc::kMaxWindowTiles || (w & 1) || (c.tc_bk != 64 && c.tc_bk != 128)) return cudaErrorInvalidValue;
  cudaError_t e;
  if (c.tc_slices) { cudaFree(c.tc_slices); c.tc_slices = nullptr; }
  if (c.tc_scale) { cudaFree(c.tc_scale); c.tc_scale = nullptr; }
  const size_t ld_pad = ((size_t)ld + 127) / 128 * 128;
  const size_t bytes = (size_t)S * ld_pad * K;
  if ((e = cudaMalloc((void**)&c.tc_slices, bytes)) != cudaSuccess) return e;
  if ((e = cudaMemset(c.tc_slices, 0, bytes)) != cudaSuccess) return e;   // rows ld .. ld_pad stay zero for good
  if ((e = cudaMalloc((void**)&c.tc_scale, ld_pad * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMemset(c.tc_scale, 0, ld_pad * sizeof(double))) != cudaSuccess) return e;
  if (!c.tc_save && (e = cudaMalloc((void**)&c.tc_save, 64 * sizeof(double))) != cudaSuccess) return e;
  static_assert(sizeof(CUtensorMap) <= 128, "CUtensorMap");
  if (!tc::make_slice_map(reinterpret_cast<CUtensorMap*>(c.tc_map_a), c.tc_slices, K, (size_t)S * ld_pad, c.tc_bk, tc::kM) ||
      !tc::make_slice_map(reinterpret_cast<CUtensorMap*>(c.tc_map_b), c.tc_slices, K, (size_t)S * ld_pad, c.tc_bk, tc::kN)) {
    c.last_error = "cuTensorMapEncodeTiled failed (driver entry point missing or arguments rejected)";
    return cudaErrorUnknown;
  }
  c.tc_cfg[0] = S; c.tc_cfg[1] = w; c.tc_cfg[2] = c.tc_bk; c.tc_cfg[3] = ld;
  return cudaSuccess;
}

// The trailing update of one panel: A[c1.., c1..] -= L[c1.., c0..c1) L[c1.., c0..c1)^T (lower triangle), plus
// the slices, the scales and the right-hand side update that go with it.  `saved_rhs`: see window_prep_kernel.
cudaError_t launch_tc_trailing_update(Context& c, double* A, double* rhs, int c0, const double* saved_rhs, cudaStream_t st) {
  const int ld = c.ld, K = NB * c.tc_window, c1 = c0 + K;
  const int ld_pad = (ld + 127) / 128 * 128;
  cudaError_t e;
  if ((e = tc::launch_slice(c.tc_slices_n, A, ld, c0, K, rhs, saved_rhs, reinterpret_cast<int8_t*>(c.tc_slices), (size_t)ld_pad * K,
                            c.tc_scale, st)) != cudaSuccess) return e;
  c.launches += 1;
  tc_mark(c, kTcProfSlices, st);
  tc::SyrkArgs g;
  memset(&g, 0, sizeof g);
  g.A = A; g.scale = c.tc_scale; g.status = &c.scalars->status; g.abort = c.solve_abort;
  g.ld = ld; g.ld_pad = ld_pad; g.c1 = c1; g.K = K;
  g.n_nb = (ld - c1) / tc::kN;
  g.ntiles = tc::count_tiles(g.n_nb);
  g.dbg_acc = c.tc_dbg; g.dbg_ld = c.tc_dbg ? c.tc_dbg_ld : 0;
  g.dbg_skip = c.tc_dbg_skip;
  g.wait_limit_ns = (unsigned long long)(c.spin_timeout_ms * 1e6);
  g.dbg_time = c.tc_dbg_time;
  if (g.ntiles <= 0) return cudaSuccess;
  int grid = g.ntiles < c.num_sms ? g.ntiles : c.num_sms;
  if (c.solve_grid_cap > 0 && grid > c.solve_grid_cap) grid = c.solve_grid_cap;
  e = tc::launch_syrk(c.tc_slices_n, c.tc_bk, *reinterpret_cast<const CUtensorMap*>(c.tc_map_a),
                      *reinterpret_cast<const CUtensorMap*>(c.tc_map_b), g, grid, st, c.tc_attr_set);
  c.launches += 1;
  tc_mark(c, kTcProfUpdates, st);
  return e;
}

static cudaError_t launch_solve_tc(Context& c, bool have_mask, cudaStream_t st) {
  const int ld = c.ld, T = ld / NB, w = c.tc_window;
  cudaError_t e;
  if ((e = tc_prepare(c)) != cudaSuccess) return e;
  const double* packed = (c.sys_state == kSysReduced && c.comm_buf) ? c.comm_buf + comm_pad(c.sys_len) : c.sys;
  double* const A = c.Adense;
  double* const rhs = c.Adense + (size_t)ld * ld;
  if (c.solve_prof_on) {
    tc_fold_profile(c);              // (the previous profiled solve, if nobody asked for it)
    tc_mark(c, kTcProfExpand, st);   // origin of this solve's time line
  }
  expand_system_kernel<<<ld, 256, 0, st>>>(packed, c.n_opt_cam, c.n_sys, ld, c.cam_mask, have_mask, A, rhs, c.solve_tickets,
                                           &c.scalars->status, c.dC, c.solve_abort, nullptr, 1.0);
  c.launches += 1;
  tc_mark(c, kTcProfExpand, st);
  if (!c.solve_attr_set) {
    if ((e = cudaFuncSetAttribute(chol_dataflow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSolveSmemBytes)) != cudaSuccess) return e;
    c.solve_attr_set = true;
  }
  CholArgs g;
  memset(&g, 0, sizeof g);
  g.flags = c.solve_flags;
  g.tickets = c.solve_tickets;
  g.status = &c.scalars->status;
  g.ld = ld;
  g.abort = c.solve_abort;
  g.spin_limit_ns = (unsigned long long)(c.spin_timeout_ms * 1e6);
  g.strict = c.strict_flags;
  g.world = 1; g.rank = 0;
  g.x = c.dC;
  for (int j0 = 0; j0 < T; j0 += w) {
    const int Tw = T - j0;                       // tile rows of the trailing matrix
    // What is left is one window: factor it all in this launch.  (Handing the last ~80 tile rows to the
    // dataflow kernel in one launch, where it beats the blocked path on a system of its own, was
    // measured and lost 0.13 ms at 11,994^2: inside the solve its panels are already in L2.)
    const bool last = Tw <= w + 1;
    const int c0 = j0 * NB;
    // tickets of the leading w tile columns: C_0, then per column jc < w the T' - jc tasks D_{jc+1}, C_{jc+1}, (jc+2.., jc)
    const int ntasks = last ? 1 + (Tw - 1) * (Tw + 2) / 2 : 1 + w * Tw - w * (w - 1) / 2;
    double* const rhs_below = last ? nullptr : rhs + c0 + w * NB;
    tc::window_prep_kernel<<<1, 64, 0, st>>>(c.solve_tickets, rhs_below, c.tc_save);
    c.launches += 1;
    g.A = A + (size_t)c0 * ld + c0;
    g.rhs = rhs + c0;
    g.LinvT = c.LinvT + (size_t)j0 * NB * NB;
    g.Wpart = c.Wpart + (size_t)j0 * (NB * NB + NB);
    g.T = Tw;
    g.epoch = ++c.solve_epoch;
    g.phase = kPhaseWindow;
    g.window_tasks = ntasks;
    g.split = last && Tw >= c.split_min_tiles;   // (windows have k loops of at most w steps: nothing to hand over)
    g.prof = nullptr;
    int grid = ntasks < c.num_sms ? ntasks : c.num_sms;
    if (c.solve_grid_cap > 0 && grid > c.solve_grid_cap) grid = c.solve_grid_cap;
    chol_dataflow_kernel<false><<<grid, kSolveThreads, kSolveSmemBytes, st>>>(g);
    c.launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    tc_mark(c, kTcProfPanels, st);
    if (last) break;
    if ((e = launch_tc_trailing_update(c, A, rhs, c0, c.tc_save, st)) != cudaSuccess) return e;
  }
  // backward substitution over the finished factor
  g.A = A;
  g.rhs = rhs;
  g.LinvT = c.LinvT;
  g.Wpart = c.Wpart;
  g.T = T;
  g.epoch = ++c.solve_epoch;
  g.phase = kPhaseBackward;
  g.window_tasks = 0;
  tc::window_prep_kernel<<<1, 64, 0, st>>>(c.solve_tickets, nullptr, c.tc_save);
  int grid = T < c.num_sms ? T : c.num_sms;
  if (c.solve_grid_cap > 0 && grid > c.solve_grid_cap) grid = c.solve_grid_cap;
  chol_dataflow_kernel<false><<<grid, kSolveThreads, kSolveSmemBytes, st>>>(g);
  c.launches += 2;
  tc_mark(c, kTcProfBackward, st);
  return cudaGetLastError();
}

cudaError_t launch_solve(Context& c, bool have_mask, cudaStream_t st) {
  if (dist_solve_selected(c)) return launch_solve_dist(c, have_mask, st);
  if (tc_solve_selected(c)) return launch_solve_tc(c, have_mask, st);
  const int ld = c.ld, T = ld / NB;
  cudaError_t e;
  // sharded problems: factor the all-reduced copy the peers pushed (ba_comm.cu), not the local contribution
  const double* packed = (c.sys_state == kSysReduced && c.comm_buf) ? c.comm_buf + comm_pad(c.sys_len) : c.sys;
  expand_system_kernel<<<ld, 256, 0, st>>>(packed, c.n_opt_cam, c.n_sys, ld, c.cam_mask, have_mask,
                                           c.Adense, c.Adense + (size_t)ld * ld, c.solve_tickets, &c.scalars->status, c.dC,
                                           c.solve_abort, nullptr, 1.0);
  c.launches += 1;
  if (!c.solve_attr_set) {
    if ((e = cudaFuncSetAttribute(chol_dataflow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSolveSmemBytes)) != cudaSuccess) return e;
    c.solve_attr_set = true;
  }
  CholArgs g;
  memset(&g, 0, sizeof g);
  g.A = c.Adense;
  g.rhs = c.Adense + (size_t)ld * ld;
  g.x = c.dC;
  g.LinvT = c.LinvT;
  g.Wpart = c.Wpart;
  g.flags = c.solve_flags;
  g.tickets = c.solve_tickets;
  g.status = &c.scalars->status;
  g.ld = ld; g.T = T;
  g.epoch = ++c.solve_epoch;   // a fresh epoch per call: flags never need clearing
  g.abort = c.solve_abort;
  g.spin_limit_ns = (unsigned long long)(c.spin_timeout_ms * 1e6);
  g.prof = c.solve_prof_on ? c.solve_prof : nullptr;
  g.split = T >= c.split_min_tiles;
  g.strict = c.strict_flags;
  g.world = 1; g.rank = 0;
#ifdef BA_SOLVE_TRACE
  g.trace = c.solve_trace;
#endif
  const int ntasks = 1 + (T - 1) * (T + 2) / 2;
  int grid = ntasks < c.num_sms ? ntasks : c.num_sms;   // 1 CTA / SM (128 KB smem): all co-resident
  if (c.solve_grid_cap > 0 && grid > c.solve_grid_cap) grid = c.solve_grid_cap;
  chol_dataflow_kernel<false><<<grid, kSolveThreads, kSolveSmemBytes, st>>>(g);
  c.launches += 1;
  return cudaGetLastError();
}

}  // namespace ba
