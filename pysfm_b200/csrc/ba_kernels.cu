// Observation-path kernels of the bundle-adjustment inner loop (sm_100a, FP64).
//
//   linearize_eliminate_kernel   prepare_schur_complement + apply_damping +
//                                compute_schur_complement  (bundle_adjuster.py:211-278)
//   retract_cameras_kernel       update_motion on the candidate (:334-337, bundle.py:76-80)
//   backsub_cost_kernel          backsubstitute + update_structure + compute_cost(candidate)
//                                (:316-331, :340-343, :165-171)
//   cost_kernel                  compute_cost(current)            (:165-171)
//   eval_observations_kernel     Bundle.residual / Jresidual dump (bundle.py:251-277)
//
// Work decomposition: the observation arrays are point-major CSR; one warp owns one point at
// a time (grid-stride), lanes first span the point's observations (residual + Jacobians),
// then span (observation, column) slots of the Schur outer products.  Everything a point
// needs after its observations were read once stays in shared memory / registers; the only
// global traffic besides the 20-byte observation records is the per-point Vinv/bP record and
// the FP64 reductions into the L2-resident reduced system.
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include "ba_context.h"
#include "ba_peer.cuh"

namespace ba {

// Loads of the observation / point streams use __ldcs (evict-first, no L1 allocation): the few KB
// of L1 the shared-memory carve-out leaves are kept for the camera arrays, which every
// observation gathers from and which otherwise cost an L2 round trip under reduction load.
// ------------------------------------------------------------------------------------------
// shared argument block (passed by value in kernel parameter space)
struct ObsArgs {
  Intrinsics intr;
  ModelParams model;
  int n_cam, n_pt, n_obs;
  const int* __restrict__ pt_ptr;
  const int* __restrict__ obs_cam;
  const double* __restrict__ obs_uv;
  const int* __restrict__ cam_slot;
  const int* __restrict__ pt_slot;
  const double* __restrict__ cam_R;
  const double* __restrict__ cam_t;
  const double* __restrict__ pts;
};

// Deterministic grid-wide sum: every CTA deposits one partial; the last CTA to arrive adds them up
// -- all its threads load (one thread walking a few hundred partials with dependent volatile loads
// was ~20 us at the end of every kernel that carries a cost), each thread its strided share in index
// order, then a fixed shuffle tree and the warps in order: the result depends on the grid size only,
// never on timing -- and returns the total to its thread 0 (other CTAs / threads: 0, last == false).
// `ticket` must be zero on entry and is reset.
__device__ __forceinline__ double grid_sum(double warp_total, double* __restrict__ partials,
                                           unsigned int* __restrict__ ticket, bool& last) {
  __shared__ double s_warp[32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) s_warp[wid] = warp_total;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_warp[i];
    partials[blockIdx.x] = t;
    __threadfence();
    const unsigned int prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  last = s_last;
  if (!s_last) return 0.0;
  __threadfence();
  double v = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(partials + i);
  v = warp_sum(v);
  __syncthreads();            // (everybody has read s_warp's first use)
  if (lane == 0) s_warp[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nw; ++i) t += s_warp[i];
    *ticket = 0u;
  }
  return t;
}

__device__ __forceinline__ void grid_sum_store(double warp_total, double* __restrict__ partials,
                                               unsigned int* __restrict__ ticket,
                                               double* __restrict__ out) {
  bool last;
  const double t = grid_sum(warp_total, partials, ticket, last);
  if (last && threadIdx.x == 0) *out = t;
}

// grid_sum_store for the candidate cost of a SHARDED problem: the last CTA of the rank also does the
// cross-rank reduction of {cost, candidate cost} over peer memory -- push both partial sums into
// every rank's bank, flag, wait for everybody's, add in rank order -- which used to be a launch
// (peer_allreduce_costs_kernel) and a barrier of its own after every back-substitution.
__device__ __forceinline__ void grid_sum_store_peer(double warp_total, double* __restrict__ partials,
                                                    unsigned int* __restrict__ ticket, Scalars* __restrict__ sc,
                                                    const PeerArgs& g) {
  __shared__ double s_total;
  bool last;
  const double total = grid_sum(warp_total, partials, ticket, last);
  if (!last) return;
  if (threadIdx.x == 0) s_total = total;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int tid = threadIdx.x;
    const int bank = (int)(g.epoch & 1u) * 2 * kMaxPeers;
    if (tid < g.world) {
      double* c = costs_of(g.base[tid], g.sys_len) + bank + 2 * g.rank;
      c[0] = sc->cost;
      c[1] = s_total;
    }
    peer_barrier(g, 2, tid);
    __syncwarp();
    if (tid == 0) {
      const volatile double* c = costs_of(g.base[g.rank], g.sys_len) + bank;
      double a = 0.0, b = 0.0;
      for (int p = 0; p < g.world; ++p) { a += c[2 * p]; b += c[2 * p + 1]; }
      sc->cost = a;
      sc->cand_cost = b;
    }
  }
}

// ------------------------------------------------------------------------------------------
struct ElimArgs {
  ObsArgs o;
  double damping, rcond;
  int n_opt_cam, kcap;
  int gcap;                    // eliminate_group_kernel: observations a warp pass can hold (NPG * longest track)
  int probe;                   // timing probes (tools/elim_probe.py): 1 = skip the reductions, 2 = skip phase D, 3 = skip the rhs atomics
  size_t rhs_off;              // doubles of packed blocks before the right-hand side
  double* __restrict__ sys;    // packed upper 6x6 blocks (row by row), then rhs [6 n_opt_cam]
  double* __restrict__ Vinv;   // [n_pt][9]
  double* __restrict__ bP;     // [n_pt][3]
  double* __restrict__ V;      // [n_pt][9]   (blocks)
  double* __restrict__ U;      // [n_cam][36] (blocks)
  double* __restrict__ bC;     // [n_cam][6]  (blocks)
  double* __restrict__ W;      // [n_obs][18] (blocks)
  double* rec_global;          // [n_obs][56] scratch for the records of long tracks (REC_GLOBAL)
  double* __restrict__ diag_rep;   // [kDiagReplicas][n_opt_cam][36] spread copies of the diagonal blocks (folded by fold_diag_kernel)
  double* __restrict__ partials;
  unsigned int* __restrict__ ticket;
  double* __restrict__ cost_out;
};

// Per-warp shared memory (doubles):
//   stage [32][kStageStride]   one 6x6 block per lane, source of the bulk reductions
//   Jc [kcap][12] | Wv [kcap][18] | Yv [kcap][18] | jtr [kcap][6] | sb [kcap] (int2 {slot, row base})
// Strides are chosen so that the 128-bit accesses of a quarter-warp fall into distinct banks:
// 18 doubles = 9 x 16 B per W/Y record, 38 doubles = 19 x 16 B per staged block.
// The diagonal block (a, a) of the packed system takes one update per observation of camera a
// (2,500 at config 2) while an off-diagonal block takes ~100: the same-address FP64 reductions
// serialise in L2 and cost 37 % of the kernel's reduction rate (tools/microbench/red_bench.cu,
// "hot diagonal blocks": 4.2e11 adds/s against 5.7e11 for uniformly spread blocks).  The diagonal
// pairs therefore reduce into one of kDiagReplicas spread copies, chosen per warp, which
// fold_diag_kernel adds into the packed diagonal blocks (and clears) right after.
constexpr int kDiagReplicas = 16;
constexpr int kStageStride = 38;
constexpr int kStageDoubles = 32 * kStageStride;
constexpr int kObsRec = 12 + 18 + 18 + 6 + 1;
__host__ __device__ __forceinline__ int elim_warp_doubles(int kcap) {
  return (kStageDoubles + kObsRec * kcap + 1) & ~1;
}

__host__ __device__ __forceinline__ size_t packed_diag_block(int a, int nc) {   // index of block (a, a) in the packed triangle
  return (size_t)a * nc - (size_t)a * (a - 1) / 2;
}

// q-th pair (a <= b) of a point with k observations, rows a = 0..k-1 holding b = a..k-1
__device__ __forceinline__ void pair_from_index(int q, int k, int& a, int& b) {
  const float kk = 2.f * (float)k + 1.f;
  int aa = (int)((kk - sqrtf(fmaxf(kk * kk - 8.f * (float)q, 0.f))) * 0.5f);
  aa = aa < 0 ? 0 : (aa > k - 1 ? k - 1 : aa);
  while (aa > 0 && aa * k - aa * (aa - 1) / 2 > q) --aa;
  while ((aa + 1) * k - (aa + 1) * aa / 2 <= q) ++aa;
  a = aa;
  b = aa + (q - (aa * k - aa * (aa - 1) / 2));
}

// S[dst .. dst+36) += stage[0 .. 36)  as ONE asynchronous bulk reduction (TMA path, SASS UBLKRED):
// the 99 M FP64 additions of a config-2 iteration reach L2 without occupying the LSU pipe.
__device__ __forceinline__ void bulk_add_block(double* dst, const double* src_smem) {
  const unsigned int src = (unsigned int)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 288;"
               ::"l"(dst), "r"(src) : "memory");
}

// REC_GLOBAL = false: the per-observation records of the point a warp works on live in shared
// memory (tracks of up to ~430 views fit).  REC_GLOBAL = true (longer tracks somewhere in the
// scene): they live in a global scratch array indexed by observation, 56 doubles per observation
// (L2 resident while the point is processed); only the staging slots stay in shared memory.
template <bool WANT_BLOCKS, bool WANT_SCHUR, bool REC_GLOBAL>
__global__ void __launch_bounds__(256, 2)
linearize_eliminate_kernel(const ElimArgs A) {
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  double* const stage = smem + (size_t)wid * elim_warp_doubles(REC_GLOBAL ? 0 : A.kcap);
  double* Jcs = stage + kStageDoubles;
  double* Wv = Jcs + 12 * A.kcap;
  double* Yv = Wv + 18 * A.kcap;
  double* jtrs = Yv + 18 * A.kcap;
  int2* sbs = reinterpret_cast<int2*>(jtrs + 6 * A.kcap);
  double* const my_stage = stage + lane * kStageStride;
  const ObsArgs& o = A.o;
  const double damp1 = 1.0 + A.damping;
  double* __restrict__ S = A.sys;
  double* __restrict__ rhs = A.sys + A.rhs_off;
  const int rep_of_warp = (blockIdx.x * warps_per_cta + wid) % kDiagReplicas;

  double cost_acc = 0.0;
  // Two-deep software pipeline over the warp's points: the CSR range of the point after next and
  // the first 32 observation records (+ coordinates) of the next point are fetched while the
  // current point is processed, so that the dependent HBM round trips (pt_ptr -> record ->
  // camera) are off the per-point critical path.  Cameras stay L1/L2 resident.
  const int stride = gridDim.x * warps_per_cta;
  int pt = blockIdx.x * warps_per_cta + wid;
  int cur_beg = 0, cur_end = 0, nxt_beg = 0, nxt_end = 0;
  if (pt < o.n_pt) { cur_beg = __ldcs(o.pt_ptr + (pt)); cur_end = __ldcs(o.pt_ptr + (pt + 1)); }
  if (pt + stride < o.n_pt) { nxt_beg = __ldcs(o.pt_ptr + (pt + stride)); nxt_end = __ldcs(o.pt_ptr + (pt + stride + 1)); }
  int pf_cam = 0, pf_ptslot = -1;
  double2 pf_uv = make_double2(0.0, 0.0);
  double pf_x0 = 0.0, pf_x1 = 0.0, pf_x2 = 0.0;
  if (pt < o.n_pt) {
    pf_x0 = __ldcs(o.pts + (3 * pt)); pf_x1 = __ldcs(o.pts + (3 * pt + 1)); pf_x2 = __ldcs(o.pts + (3 * pt + 2));
    pf_ptslot = __ldcs(o.pt_slot + (pt));
    if (lane < cur_end - cur_beg) {
      pf_cam = __ldcs(o.obs_cam + (cur_beg + lane));
      pf_uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (cur_beg + lane));
    }
  }
  for (; pt < o.n_pt; pt += stride) {
    const int beg = cur_beg;
    const int k = cur_end - cur_beg;
    if (REC_GLOBAL) {   // this point's records: [Jc 12k | W 18k | Y 18k | jtr 6k | sb k] at 56 doubles per observation
      Jcs = A.rec_global + (size_t)beg * 56;
      Wv = Jcs + 12 * k;
      Yv = Wv + 18 * k;
      jtrs = Yv + 18 * k;
      sbs = reinterpret_cast<int2*>(jtrs + 6 * k);
    }
    const double x[3] = {pf_x0, pf_x1, pf_x2};
    const bool pt_free = pf_ptslot >= 0;
    const int cam0 = pf_cam;
    const double2 uv0 = pf_uv;
    {   // advance the pipeline
      const int p1 = pt + stride, p2 = pt + 2 * stride;
      cur_beg = nxt_beg; cur_end = nxt_end;
      if (p2 < o.n_pt) { nxt_beg = __ldcs(o.pt_ptr + (p2)); nxt_end = __ldcs(o.pt_ptr + (p2 + 1)); }
      if (p1 < o.n_pt) {
        pf_x0 = __ldcs(o.pts + (3 * p1)); pf_x1 = __ldcs(o.pts + (3 * p1 + 1)); pf_x2 = __ldcs(o.pts + (3 * p1 + 2));
        pf_ptslot = __ldcs(o.pt_slot + (p1));
        if (lane < cur_end - cur_beg) {
          pf_cam = __ldcs(o.obs_cam + (cur_beg + lane));
          pf_uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (cur_beg + lane));
        }
      }
    }
    double Vl[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};

    // ---- phase A: lanes over observations ------------------------------------------------
    for (int a = lane; a < k; a += 32) {
      const int ob = beg + a;
      const int cam = (a < 32) ? cam0 : __ldcs(o.obs_cam + (ob));
      const double2 uv = (a < 32) ? uv0 : __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (ob));
      const int slot = o.cam_slot[cam];
      double r[2], Jc[12], Jp[6];
      observe(o.intr, o.model, o.cam_R + 9 * cam, o.cam_t + 3 * cam, x, uv.x, uv.y, r, Jc, Jp);
      Vl[0] += Jp[0] * Jp[0] + Jp[3] * Jp[3];
      Vl[1] += Jp[0] * Jp[1] + Jp[3] * Jp[4];
      Vl[2] += Jp[0] * Jp[2] + Jp[3] * Jp[5];
      Vl[3] += Jp[1] * Jp[1] + Jp[4] * Jp[4];
      Vl[4] += Jp[1] * Jp[2] + Jp[4] * Jp[5];
      Vl[5] += Jp[2] * Jp[2] + Jp[5] * Jp[5];
      bl[0] += Jp[0] * r[0] + Jp[3] * r[1];
      bl[1] += Jp[1] * r[0] + Jp[4] * r[1];
      bl[2] += Jp[2] * r[0] + Jp[5] * r[1];
      if (slot >= 0 && pt_free) cost_acc += r[0] * r[0] + r[1] * r[1];
#pragma unroll
      for (int i = 0; i < 12; i += 2)
        *reinterpret_cast<double2*>(Jcs + a * 12 + i) = make_double2(Jc[i], Jc[i + 1]);
      double Wl[18];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr)
#pragma unroll
        for (int m = 0; m < 3; ++m) Wl[rr * 3 + m] = Jc[rr] * Jp[m] + Jc[6 + rr] * Jp[3 + m];
#pragma unroll
      for (int i = 0; i < 18; i += 2)
        *reinterpret_cast<double2*>(Wv + a * 18 + i) = make_double2(Wl[i], Wl[i + 1]);
      double jtr[6];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) jtr[rr] = Jc[rr] * r[0] + Jc[6 + rr] * r[1];
#pragma unroll
      for (int i = 0; i < 6; i += 2)
        *reinterpret_cast<double2*>(jtrs + a * 6 + i) = make_double2(jtr[i], jtr[i + 1]);
      // packed row base: block (slot, b) lives at index rowbase + b, b >= slot
      sbs[a] = make_int2(slot, slot >= 0 ? slot * A.n_opt_cam - slot * (slot - 1) / 2 - slot : 0);
      if (WANT_BLOCKS) {
        double* Uc = A.U + (size_t)cam * 36;
#pragma unroll
        for (int rr = 0; rr < 6; ++rr)
#pragma unroll
          for (int cc = 0; cc < 6; ++cc)
            atomicAdd(Uc + rr * 6 + cc, Jc[rr] * Jc[cc] + Jc[6 + rr] * Jc[6 + cc]);
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) atomicAdd(A.bC + (size_t)cam * 6 + rr, jtr[rr]);
        double* Wg = A.W + (size_t)ob * 18;
#pragma unroll
        for (int i = 0; i < 18; ++i) Wg[i] = Wl[i];
      }
    }
    // ---- phase B: point block ------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 6; ++i) Vl[i] = warp_sum(Vl[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) bl[i] = warp_sum(bl[i]);
    double Vf[9] = {Vl[0], Vl[1], Vl[2], Vl[1], Vl[3], Vl[4], Vl[2], Vl[4], Vl[5]};
    if (lane == 0) {
      if (WANT_BLOCKS) {
#pragma unroll
        for (int i = 0; i < 9; ++i) A.V[(size_t)pt * 9 + i] = Vf[i];
      }
      A.bP[3 * pt] = bl[0]; A.bP[3 * pt + 1] = bl[1]; A.bP[3 * pt + 2] = bl[2];
    }
    if (!WANT_SCHUR) { __syncwarp(); continue; }
    Vf[0] *= damp1; Vf[4] *= damp1; Vf[8] *= damp1;
    double Vi[9];
    sym3_pinv(Vf, A.rcond, Vi);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) A.Vinv[(size_t)pt * 9 + i] = Vi[i];
    }
    __syncwarp();
    // ---- phase C: Y = W Vinv rows, reduced right-hand side -------------------------------
    for (int task = lane; task < 6 * k; task += 32) {
      const int a = task / 6, rr = task - 6 * a;
      const double w0 = Wv[a * 18 + rr * 3], w1 = Wv[a * 18 + rr * 3 + 1], w2 = Wv[a * 18 + rr * 3 + 2];
      const double y0 = w0 * Vi[0] + w1 * Vi[3] + w2 * Vi[6];
      const double y1 = w0 * Vi[1] + w1 * Vi[4] + w2 * Vi[7];
      const double y2 = w0 * Vi[2] + w1 * Vi[5] + w2 * Vi[8];
      Yv[a * 18 + rr * 3] = y0; Yv[a * 18 + rr * 3 + 1] = y1; Yv[a * 18 + rr * 3 + 2] = y2;
      const int slot = sbs[a].x;
      if (slot >= 0 && A.probe != 3)
        atomicAdd(rhs + 6 * slot + rr, jtrs[a * 6 + rr] - (y0 * bl[0] + y1 * bl[1] + y2 * bl[2]));
    }
    __syncwarp();
    // ---- phase D: one lane per camera pair (a <= b) of the point.  The lane forms the whole 6x6
    // block in registers ( -Y_a W_b^T, plus the damped Jc^T Jc on the diagonal pairs), parks it
    // in its staging slot and hands it to the bulk-reduction engine.  Slots ascend inside a
    // point, so (a, b) is always in the stored upper block triangle. --------------------------
    // Pair order: the k diagonal pairs (a, a) first, then the k(k-1)/2 pairs a < b row by row,
    // so that only the first round(s) carry the Jc^T Jc term.
    const int npairs = (A.probe == 2) ? 0 : k * (k + 1) / 2;
    for (int q0 = 0; q0 < npairs; q0 += 32) {
      const int q = q0 + lane;
      // the previous round's (or point's) reduction must have finished READING this slot
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (q < npairs) {
        int a, b;
        if (q < k) {
          a = b = q;
        } else {
          pair_from_index(q - k, k - 1, a, b);
          b += 1;
        }
        const int2 sa = sbs[a];
        if (sa.x >= 0) {
          const int slot_b = sbs[b].x;
          double Wb[18];
#pragma unroll
          for (int i = 0; i < 18; i += 2) {
            const double2 w = *reinterpret_cast<const double2*>(Wv + b * 18 + i);
            Wb[i] = w.x; Wb[i + 1] = w.y;
          }
          const double* Ya = Yv + a * 18;
          const double* Jca = Jcs + a * 12;
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) {
            const double y0 = Ya[rr * 3], y1 = Ya[rr * 3 + 1], y2 = Ya[rr * 3 + 2];
            double v[6];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc)
              v[cc] = -(y0 * Wb[cc * 3] + y1 * Wb[cc * 3 + 1] + y2 * Wb[cc * 3 + 2]);
            if (q0 < k && a == b) {   // diagonal pair: + damped Jc^T Jc
              const double j0 = Jca[rr], j1 = Jca[6 + rr];
#pragma unroll
              for (int cc = 0; cc < 6; ++cc) {
                const double d = j0 * Jca[cc] + j1 * Jca[6 + cc];
                v[cc] += (rr == cc) ? d * damp1 : d;
              }
            }
#pragma unroll
            for (int cc = 0; cc < 6; cc += 2)
              *reinterpret_cast<double2*>(my_stage + rr * 6 + cc) = make_double2(v[cc], v[cc + 1]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          double* dst = (a == b) ? A.diag_rep + ((size_t)rep_of_warp * A.n_opt_cam + sa.x) * 36
                                 : S + (size_t)(sa.y + slot_b) * 36;
          if (A.probe != 1) bulk_add_block(dst, my_stage);
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    __syncwarp();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  grid_sum_store(warp_sum(cost_acc), A.partials, A.ticket, A.cost_out);
}

// S(a, a) += sum over the replicas, replicas cleared for the next iteration.  One CTA per camera.
__global__ void __launch_bounds__(64) fold_diag_kernel(double* __restrict__ S, double* __restrict__ diag_rep, int n_opt_cam) {
  const int a = blockIdx.x, e = threadIdx.x;
  if (e >= 36) return;
  double v[kDiagReplicas];
#pragma unroll
  for (int r = 0; r < kDiagReplicas; ++r) v[r] = __ldcg(diag_rep + ((size_t)r * n_opt_cam + a) * 36 + e);
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < kDiagReplicas; ++r) {
    s += v[r];
    diag_rep[((size_t)r * n_opt_cam + a) * 36 + e] = 0.0;
  }
  S[packed_diag_block(a, n_opt_cam) * 36 + e] += s;
}


// ------------------------------------------------------------------------------------------
// Elimination with SEVERAL POINTS PER WARP PASS (short tracks, the common case): with one point per
// warp, 10 of 32 lanes work in the per-observation phases at 10 observations per point, and those
// phases are bound by FP64 issue.  Here a warp takes NPG consecutive points whose observations
// fill its lanes (host guarantees NPG * longest track <= 32):
//   A  lane = observation: residual, Jc, Jp, W = Jc^T Jp, Jc^T r            (registers)
//   B  V, bP of each point by masked warp sums; damped V^-1 in every lane of the point
//   C  Y = W V^-1 and the reduced right-hand side, lane-local; W, Y parked in shared memory
//   D  one lane per camera pair of the group: the n diagonal pairs first (round 0, lane = own
//      observation, operands still in registers), then the off-diagonal pairs point by point;
//      each 6x6 block goes out as one 288-byte bulk reduction, as in the kernel above.
// Same results as linearize_eliminate_kernel<false, true, false> up to summation order.
__host__ __device__ __forceinline__ int elim_group_warp_doubles(int cap) {
  return (kStageDoubles + 37 * cap + 1) & ~1;   // stage | W [cap][18] | Y [cap][18] | sb [cap] (int2)
}

template <int NPG>
__global__ void __launch_bounds__(192, 2) eliminate_group_kernel(const ElimArgs A) {
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int cap = A.gcap;
  double* const stage = smem + (size_t)wid * elim_group_warp_doubles(cap);
  double* const Wv = stage + kStageDoubles;
  double* const Yv = Wv + 18 * cap;
  int2* const sbs = reinterpret_cast<int2*>(Yv + 18 * cap);
  double* const my_stage = stage + lane * kStageStride;
  const ObsArgs& o = A.o;
  const double damp1 = 1.0 + A.damping;
  double* __restrict__ S = A.sys;
  double* __restrict__ rhs = A.sys + A.rhs_off;

  double cost_acc = 0.0;
  const int n_groups = (o.n_pt + NPG - 1) / NPG;
  const int stride = gridDim.x * warps_per_cta;
  for (int g = blockIdx.x * warps_per_cta + wid; g < n_groups; g += stride) {
    const int p0 = g * NPG;
    // observations e[j] .. e[j+1] belong to point p0 + j (points past the end are empty)
    int e[NPG + 1];
#pragma unroll
    for (int j = 0; j <= NPG; ++j) e[j] = __ldcs(o.pt_ptr + (p0 + j < o.n_pt ? p0 + j : o.n_pt));
    const int n = e[NPG] - e[0];
    const int ob = e[0] + lane;
    const bool active = lane < n;
    int pi = 0;
#pragma unroll
    for (int j = 1; j < NPG; ++j) pi += (ob >= e[j]) ? 1 : 0;
    int pbeg = e[0];
#pragma unroll
    for (int j = 1; j < NPG; ++j)
      if (pi == j) pbeg = e[j];
    const int pt = p0 + pi;

    // ---- phase A: lane = observation ---------------------------------------------------------
    double Vl[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
    double Jc[12], W[18], jtr[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) Jc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 18; ++i) W[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) jtr[i] = 0.0;
    int slot = -1, rowbase = 0;
    if (active) {
      const int cam = __ldcs(o.obs_cam + ob);
      const double2 uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + ob);
      const double x[3] = {__ldcs(o.pts + 3 * pt), __ldcs(o.pts + 3 * pt + 1), __ldcs(o.pts + 3 * pt + 2)};
      const bool pt_free = __ldcs(o.pt_slot + pt) >= 0;
      slot = o.cam_slot[cam];
      double r[2], Jp[6];
      observe(o.intr, o.model, o.cam_R + 9 * cam, o.cam_t + 3 * cam, x, uv.x, uv.y, r, Jc, Jp);
      Vl[0] = Jp[0] * Jp[0] + Jp[3] * Jp[3];
      Vl[1] = Jp[0] * Jp[1] + Jp[3] * Jp[4];
      Vl[2] = Jp[0] * Jp[2] + Jp[3] * Jp[5];
      Vl[3] = Jp[1] * Jp[1] + Jp[4] * Jp[4];
      Vl[4] = Jp[1] * Jp[2] + Jp[4] * Jp[5];
      Vl[5] = Jp[2] * Jp[2] + Jp[5] * Jp[5];
      bl[0] = Jp[0] * r[0] + Jp[3] * r[1];
      bl[1] = Jp[1] * r[0] + Jp[4] * r[1];
      bl[2] = Jp[2] * r[0] + Jp[5] * r[1];
      if (slot >= 0 && pt_free) cost_acc += r[0] * r[0] + r[1] * r[1];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) {
#pragma unroll
        for (int m = 0; m < 3; ++m) W[rr * 3 + m] = Jc[rr] * Jp[m] + Jc[6 + rr] * Jp[3 + m];
        jtr[rr] = Jc[rr] * r[0] + Jc[6 + rr] * r[1];
      }
      // packed row base: block (slot, b) lives at index rowbase + b, b >= slot
      rowbase = slot >= 0 ? slot * A.n_opt_cam - slot * (slot - 1) / 2 - slot : 0;
    }

    // ---- phase B: point blocks (masked warp sums, one point at a time) -----------------------
    double Vs[6] = {0, 0, 0, 0, 0, 0}, bs[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < NPG; ++j) {
      if (e[j + 1] == e[j]) continue;   // warp-uniform
      const bool mine = active && pi == j;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double t = warp_sum(mine ? Vl[i] : 0.0);
        if (mine) Vs[i] = t;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double t = warp_sum(mine ? bl[i] : 0.0);
        if (mine) bs[i] = t;
      }
    }
    double Vf[9] = {Vs[0], Vs[1], Vs[2], Vs[1], Vs[3], Vs[4], Vs[2], Vs[4], Vs[5]};
    Vf[0] *= damp1; Vf[4] *= damp1; Vf[8] *= damp1;
    if (!active) { Vf[0] = Vf[4] = Vf[8] = 1.0; }   // idle lanes stay on the fast path of the inverse
    double Vi[9];
    sym3_pinv(Vf, A.rcond, Vi);
    if (active && ob == pbeg) {   // first lane of its point
      A.bP[3 * pt] = bs[0]; A.bP[3 * pt + 1] = bs[1]; A.bP[3 * pt + 2] = bs[2];
#pragma unroll
      for (int i = 0; i < 9; ++i) A.Vinv[(size_t)pt * 9 + i] = Vi[i];
    }
    // points without observations: zero gradient, pinv of the zero block
#pragma unroll
    for (int j = 0; j < NPG; ++j) {
      if (e[j + 1] != e[j] || p0 + j >= o.n_pt) continue;   // warp-uniform
      if (lane == j) {
        const double Z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        double Zi[9];
        sym3_pinv(Z, A.rcond, Zi);
        A.bP[3 * (p0 + j)] = 0.0; A.bP[3 * (p0 + j) + 1] = 0.0; A.bP[3 * (p0 + j) + 2] = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) A.Vinv[(size_t)(p0 + j) * 9 + i] = Zi[i];
      }
    }

    // ---- phase C: Y = W Vinv, reduced right-hand side (lane-local) ----------------------------
    double Y[18];
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) {
      const double w0 = W[rr * 3], w1 = W[rr * 3 + 1], w2 = W[rr * 3 + 2];
      Y[rr * 3] = w0 * Vi[0] + w1 * Vi[3] + w2 * Vi[6];
      Y[rr * 3 + 1] = w0 * Vi[1] + w1 * Vi[4] + w2 * Vi[7];
      Y[rr * 3 + 2] = w0 * Vi[2] + w1 * Vi[5] + w2 * Vi[8];
    }
    if (active) {
      if (slot >= 0) {
#pragma unroll
        for (int rr = 0; rr < 6; ++rr)
          atomicAdd(rhs + 6 * slot + rr, jtr[rr] - (Y[rr * 3] * bs[0] + Y[rr * 3 + 1] * bs[1] + Y[rr * 3 + 2] * bs[2]));
      }
#pragma unroll
      for (int i = 0; i < 18; i += 2) {
        *reinterpret_cast<double2*>(Wv + lane * 18 + i) = make_double2(W[i], W[i + 1]);
        *reinterpret_cast<double2*>(Yv + lane * 18 + i) = make_double2(Y[i], Y[i + 1]);
      }
      sbs[lane] = make_int2(slot, rowbase);
    }
    __syncwarp();

    // ---- phase D: one lane per camera pair of the group --------------------------------------
    int cpairs[NPG];
    int npairs = n;
#pragma unroll
    for (int j = 0; j < NPG; ++j) {
      const int kj = e[j + 1] - e[j];
      cpairs[j] = kj * (kj - 1) / 2;
      npairs += cpairs[j];
    }
    for (int q0 = 0; q0 < npairs; q0 += 32) {
      const int q = q0 + lane;
      // the previous round's (or pass's) reduction must have finished READING this slot
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (q < npairs) {
        if (q < n) {
          // round 0 only (n <= 32): the diagonal pair of this lane's own observation,
          // -Y W^T + damped Jc^T Jc, straight from registers
          if (slot >= 0) {
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) {
              const double y0 = Y[rr * 3], y1 = Y[rr * 3 + 1], y2 = Y[rr * 3 + 2];
              const double j0 = Jc[rr], j1 = Jc[6 + rr];
              double v[6];
#pragma unroll
              for (int cc = 0; cc < 6; ++cc) {
                const double d = j0 * Jc[cc] + j1 * Jc[6 + cc];
                v[cc] = -(y0 * W[cc * 3] + y1 * W[cc * 3 + 1] + y2 * W[cc * 3 + 2]) + ((rr == cc) ? d * damp1 : d);
              }
#pragma unroll
              for (int cc = 0; cc < 6; cc += 2)
                *reinterpret_cast<double2*>(my_stage + rr * 6 + cc) = make_double2(v[cc], v[cc + 1]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_add_block(A.diag_rep + ((size_t)((blockIdx.x * warps_per_cta + wid) % kDiagReplicas) * A.n_opt_cam + slot) * 36, my_stage);
          }
        } else {
          int idx = q - n, j = 0;
#pragma unroll
          for (int jj = 0; jj + 1 < NPG; ++jj)
            if (j == jj && idx >= cpairs[jj]) { idx -= cpairs[jj]; j = jj + 1; }
          int start = 0, kj = e[1] - e[0];
#pragma unroll
          for (int jj = 1; jj < NPG; ++jj)
            if (j == jj) { start = e[jj] - e[0]; kj = e[jj + 1] - e[jj]; }
          int a, b;
          pair_from_index(idx, kj - 1, a, b);
          b += 1;
          const int2 sa = sbs[start + a];
          if (sa.x >= 0) {
            const int slot_b = sbs[start + b].x;
            double Wb[18];
#pragma unroll
            for (int i = 0; i < 18; i += 2) {
              const double2 w = *reinterpret_cast<const double2*>(Wv + (start + b) * 18 + i);
              Wb[i] = w.x; Wb[i + 1] = w.y;
            }
            const double* Ya = Yv + (start + a) * 18;
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) {
              const double y0 = Ya[rr * 3], y1 = Ya[rr * 3 + 1], y2 = Ya[rr * 3 + 2];
              double v[6];
#pragma unroll
              for (int cc = 0; cc < 6; ++cc)
                v[cc] = -(y0 * Wb[cc * 3] + y1 * Wb[cc * 3 + 1] + y2 * Wb[cc * 3 + 2]);
#pragma unroll
              for (int cc = 0; cc < 6; cc += 2)
                *reinterpret_cast<double2*>(my_stage + rr * 6 + cc) = make_double2(v[cc], v[cc + 1]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_add_block(S + (size_t)(sa.y + slot_b) * 36, my_stage);
          }
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    __syncwarp();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  grid_sum_store(warp_sum(cost_acc), A.partials, A.ticket, A.cost_out);
}

// ------------------------------------------------------------------------------------------
// cand cameras = state cameras (+) (-dC)      (update = -solution, bundle_adjuster.py:208)
__global__ void retract_cameras_kernel(int n_cam, const int* __restrict__ cam_slot,
                                       const double* __restrict__ R, const double* __restrict__ t,
                                       const double* __restrict__ delta, double sign,
                                       double* __restrict__ Rc, double* __restrict__ tc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cam) return;
  const int slot = cam_slot[i];
  if (slot < 0 || delta == nullptr) {
#pragma unroll
    for (int j = 0; j < 9; ++j) Rc[9 * i + j] = R[9 * i + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) tc[3 * i + j] = t[3 * i + j];
    return;
  }
  double d[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) d[j] = sign * delta[6 * slot + j];
  camera_retract(R + 9 * i, t + 3 * i, d, Rc + 9 * i, tc + 3 * i);
}

__global__ void retract_points_kernel(int n_pt, const int* __restrict__ pt_slot,
                                      const double* __restrict__ x, const double* __restrict__ delta,
                                      double* __restrict__ xc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pt) return;
  const int slot = pt_slot[i];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    xc[3 * i + j] = x[3 * i + j] + ((slot >= 0 && delta) ? delta[3 * slot + j] : 0.0);
}

// ------------------------------------------------------------------------------------------
struct BacksubArgs {
  ObsArgs o;               // o.cam_R/cam_t/pts = CURRENT state (linearisation point)
  const double* __restrict__ cand_R;   // candidate cameras (read; SC = false)
  const double* __restrict__ cand_t;
  double* cand_R_out;                  // candidate cameras (written by CTA 0; SC = true)
  double* cand_t_out;
  double* __restrict__ cand_pts;
  const double* __restrict__ dC;    // [6 n_opt_cam]
  const double* __restrict__ Vinv;
  const double* __restrict__ bP;
  double* __restrict__ dP;          // [n_pt][3]
  double* __restrict__ partials;
  unsigned int* __restrict__ ticket;
  double* __restrict__ cost_out;
  Scalars* sc;               // the handle's scalar record (cost_out == &sc->cand_cost)
  PeerArgs peer;             // peer.world > 1: the cost reduction over the ranks is fused into the kernel's epilogue
};

__device__ __forceinline__ void backsub_cost_epilogue(double warp_total, const BacksubArgs& A) {
  if (A.peer.world > 1) grid_sum_store_peer(warp_total, A.partials, A.ticket, A.sc, A.peer);
  else grid_sum_store(warp_total, A.partials, A.ticket, A.cost_out);
}

// sum over the G-lane group a lane belongs to (G = 8, 16 or 32, groups are lane-aligned)
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// G lanes per point: with tracks of <= 16 (<= 8) observations a warp works on 2 (4) points at a
// time, which halves (quarters) the number of dependent load -> compute -> reduce round trips
// each warp goes through; this kernel is bound by that latency chain, not by bytes.
// SC = cameras staged in shared memory: every CTA retracts ALL cameras itself at start
// (R exp(-dC[:3]), t - dC[3:]; a few hundred cameras are cheaper to recompute per CTA than to
// launch a separate kernel for and gather from L2), keeps the linearisation-point and the
// candidate cameras in shared memory for its two passes, and CTA 0 writes the candidate cameras
// out.  SC = false (too many cameras for shared memory): retract_cameras_kernel ran before.
template <int G, bool SC>
__global__ void __launch_bounds__(256, 3) backsub_cost_kernel(const BacksubArgs A) {
  // SC: [n_cam][12] state {R, t} | [n_cam][12] candidate | [n_cam][6] dC of the camera (0 if fixed) | [n_cam] slot
  extern __shared__ __align__(16) double cam_sm[];
  const ObsArgs& o = A.o;
  const double* const dC_sm = cam_sm + 24 * o.n_cam;
  int* const slot_sm = reinterpret_cast<int*>(cam_sm + 30 * o.n_cam);
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  constexpr int NG = 32 / G;
  const int sub = lane / G, gl = lane % G;
  double cost_acc = 0.0;
  if (SC) {
    for (int i = threadIdx.x; i < o.n_cam; i += blockDim.x) {
      double R[9], t[3], Rc[9], tc[3];
#pragma unroll
      for (int j = 0; j < 9; ++j) R[j] = o.cam_R[9 * i + j];
#pragma unroll
      for (int j = 0; j < 3; ++j) t[j] = o.cam_t[3 * i + j];
      const int slot = o.cam_slot[i];
      slot_sm[i] = slot;
      if (slot >= 0) {
        double d[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const double v = A.dC[6 * slot + j];
          cam_sm[24 * o.n_cam + 6 * i + j] = v;   // dC_sm[6 i + j]
          d[j] = -v;
        }
        camera_retract(R, t, d, Rc, tc);
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) cam_sm[24 * o.n_cam + 6 * i + j] = 0.0;
#pragma unroll
        for (int j = 0; j < 9; ++j) Rc[j] = R[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) tc[j] = t[j];
      }
      double* s0 = cam_sm + 12 * i;
      double* s1 = cam_sm + 12 * (o.n_cam + i);
#pragma unroll
      for (int j = 0; j < 9; ++j) { s0[j] = R[j]; s1[j] = Rc[j]; }
#pragma unroll
      for (int j = 0; j < 3; ++j) { s0[9 + j] = t[j]; s1[9 + j] = tc[j]; }
      if (blockIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 9; ++j) A.cand_R_out[9 * i + j] = Rc[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) A.cand_t_out[3 * i + j] = tc[j];
      }
    }
    __syncthreads();
  }
  for (int base = (blockIdx.x * warps_per_cta + wid) * NG; base < o.n_pt; base += gridDim.x * warps_per_cta * NG) {
    const int pt = base + sub;
    const bool live = pt < o.n_pt;
    const int beg = live ? __ldcs(o.pt_ptr + (pt)) : 0;
    const int k = live ? __ldcs(o.pt_ptr + (pt + 1)) - beg : 0;
    double x[3] = {0.0, 0.0, 0.0};
    int pslot = -1;
    if (live) {
      x[0] = __ldcs(o.pts + (3 * pt)); x[1] = __ldcs(o.pts + (3 * pt + 1)); x[2] = __ldcs(o.pts + (3 * pt + 2));
      pslot = __ldcs(o.pt_slot + (pt));
    }
    double xc[3] = {x[0], x[1], x[2]};
    // sum_j W_j^T dC_j  ==  sum_j Jp_j^T (Jc_j dC_j)
    double acc[3] = {0, 0, 0};
    // the lane's first observation stays in registers for the candidate pass below (tracks no
    // longer than the lane group -- the usual case -- then read their records exactly once)
    int cam_first = -1;
    double2 uv_first = make_double2(0.0, 0.0);
    if (pslot >= 0) {
      for (int a = gl; a < k; a += G) {
        const int ob = beg + a;
        const int cam = __ldcs(o.obs_cam + (ob));
        const int slot = SC ? slot_sm[cam] : o.cam_slot[cam];
        if (slot < 0) continue;
        const double2 uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (ob));
        if (a == gl) { cam_first = cam; uv_first = uv; }
        double r[2], Jc[12], Jp[6];
        if (SC) observe(o.intr, o.model, cam_sm + 12 * cam, cam_sm + 12 * cam + 9, x, uv.x, uv.y, r, Jc, Jp);
        else observe(o.intr, o.model, o.cam_R + 9 * cam, o.cam_t + 3 * cam, x, uv.x, uv.y, r, Jc, Jp);
        const double* d = SC ? dC_sm + 6 * cam : A.dC + 6 * slot;
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) { q0 += Jc[j] * d[j]; q1 += Jc[6 + j] * d[j]; }
#pragma unroll
        for (int m = 0; m < 3; ++m) acc[m] += Jp[m] * q0 + Jp[3 + m] * q1;
      }
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) acc[m] = group_sum<G>(acc[m]);
    if (pslot >= 0) {
      const double g0 = A.bP[3 * pt] - acc[0], g1 = A.bP[3 * pt + 1] - acc[1], g2 = A.bP[3 * pt + 2] - acc[2];
      const double* Vi = A.Vinv + (size_t)pt * 9;
      double dp[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) dp[m] = Vi[3 * m] * g0 + Vi[3 * m + 1] * g1 + Vi[3 * m + 2] * g2;
#pragma unroll
      for (int m = 0; m < 3; ++m) xc[m] = x[m] - dp[m];
      if (gl == 0) {
#pragma unroll
        for (int m = 0; m < 3; ++m) A.dP[3 * pt + m] = dp[m];
      }
    } else if (live && gl == 0) {
#pragma unroll
      for (int m = 0; m < 3; ++m) A.dP[3 * pt + m] = 0.0;
    }
    if (live && gl == 0) {
#pragma unroll
      for (int m = 0; m < 3; ++m) A.cand_pts[3 * pt + m] = xc[m];
    }
    if (pslot >= 0) {
      for (int a = gl; a < k; a += G) {
        int cam;
        double2 uv;
        if (a == gl) {   // kept from the first pass (cam_first < 0: no observation in an optimised camera)
          if (cam_first < 0) continue;
          cam = cam_first;
          uv = uv_first;
        } else {
          const int ob = beg + a;
          cam = __ldcs(o.obs_cam + (ob));
          if ((SC ? slot_sm[cam] : o.cam_slot[cam]) < 0) continue;
          uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (ob));
        }
        double r[2];
        if (SC) residual_only(o.intr, o.model, cam_sm + 12 * (o.n_cam + cam), cam_sm + 12 * (o.n_cam + cam) + 9, xc, uv.x, uv.y, r);
        else residual_only(o.intr, o.model, A.cand_R + 9 * cam, A.cand_t + 3 * cam, xc, uv.x, uv.y, r);
        cost_acc += r[0] * r[0] + r[1] * r[1];
      }
    }
  }
  backsub_cost_epilogue(warp_sum(cost_acc), A);
}


// ------------------------------------------------------------------------------------------
// Back-substitution, TILE-STREAMED and OBSERVATION-PARALLEL (the default whenever the cameras fit in
// shared memory and no track is longer than 64 views).  backsub_cost_kernel above walks a dependent
// chain of global loads per point (pt_ptr -> obs_cam -> slot -> obs_uv -> Vinv / bP) with 10 of 16
// lanes busy at 10 observations per point, and ran at 6.6 % of the HBM peak.  Here a CTA owns
// CHUNKS of P consecutive points holding at most kTileObs = 256 observations (one per thread):
//   * the point-major CSR makes every array of a chunk one contiguous range, so the whole chunk --
//     CSR offsets, points, slots, Vinv, bP and the observation records -- is fetched with
//     independent, coalesced loads into registers while the PREVIOUS chunk is being computed, and
//     parked in shared memory at the top of the next iteration (the bounds of the chunk after next
//     travel one iteration further ahead);
//   * pass 1, thread = observation: Jacobians at the linearisation point, Jp^T (Jc dC) into shared
//     memory;  pass 2, thread = point: sum of the point's terms in observation order (no atomics:
//     deterministic), dP = Vinv (bP - sum), candidate point;  pass 3, thread = observation: residual
//     of the candidate.  All 256 lanes work in the two expensive passes.
// Cameras, their candidate poses, dC and slots live in shared memory as in the SC variant above.
constexpr int kTilePts = 32;
constexpr int kTileObs = 256;

struct TileStage {     // shared-memory image of one chunk
  int ptr[kTilePts + 1];
  int slot[kTilePts];
  double x[3 * kTilePts];
  double xc[3 * kTilePts];
  double Vi[9 * kTilePts];
  double bP[3 * kTilePts];
  int cam[kTileObs];
  double2 uv[kTileObs];
  double c[3 * kTileObs];          // pass 1 -> pass 2: Jp^T (Jc dC) of every observation
  unsigned char pt_of[kTileObs];   // point (position in the chunk) of every observation
};

__global__ void __launch_bounds__(256, 2) backsub_tile_kernel(const BacksubArgs A, int P) {
  extern __shared__ __align__(16) double cam_sm[];   // cameras as in backsub_cost_kernel<G, true>, then the stage
  const ObsArgs& o = A.o;
  const double* const dC_sm = cam_sm + 24 * o.n_cam;
  int* const slot_sm = reinterpret_cast<int*>(cam_sm + 30 * o.n_cam);
  TileStage& S = *reinterpret_cast<TileStage*>(cam_sm + ((31 * o.n_cam + 2) & ~1));
  const int tid = threadIdx.x;
  const int n_chunks = (o.n_pt + P - 1) / P;

  // ---- chunk pipeline: bounds two chunks ahead, data one chunk ahead (registers) -------------
  auto bounds = [&](int c, int& p0, int& np, int& ob0, int& nob) {
    p0 = c * P;
    np = 0; ob0 = 0; nob = 0;
    if (c < n_chunks) {
      np = (o.n_pt - p0 < P) ? o.n_pt - p0 : P;
      ob0 = __ldg(o.pt_ptr + p0);
      nob = __ldg(o.pt_ptr + p0 + np) - ob0;
    }
  };
  int r_ptr = 0, r_slot = -1, r_cam = 0;
  double r_x = 0.0, r_Vi0 = 0.0, r_Vi1 = 0.0, r_bP = 0.0;
  double2 r_uv = make_double2(0.0, 0.0);
  auto fetch = [&](int p0, int np, int ob0, int nob) {   // independent loads of one chunk, this thread's share
    if (np > 0 && tid <= np) r_ptr = __ldcs(o.pt_ptr + p0 + tid);   // (np == 0: past the last chunk, nothing to fetch)
    if (tid < np) r_slot = __ldcs(o.pt_slot + p0 + tid);
    if (tid < 3 * np) { r_x = __ldcs(o.pts + 3 * (size_t)p0 + tid); r_bP = __ldcs(A.bP + 3 * (size_t)p0 + tid); }
    if (tid < 9 * np) r_Vi0 = __ldcs(A.Vinv + 9 * (size_t)p0 + tid);
    if (tid + 256 < 9 * np) r_Vi1 = __ldcs(A.Vinv + 9 * (size_t)p0 + tid + 256);
    if (tid < nob) { r_cam = __ldcs(o.obs_cam + ob0 + tid); r_uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + ob0 + tid); }
  };
  int c_cur = blockIdx.x;
  int p0, np, ob0, nob;            // chunk whose data sits in registers
  bounds(c_cur, p0, np, ob0, nob);
  int n_p0, n_np, n_ob0, n_nob;    // bounds of the chunk after it
  bounds(c_cur + gridDim.x, n_p0, n_np, n_ob0, n_nob);
  fetch(p0, np, ob0, nob);

  // ---- cameras: every CTA retracts all of them itself (see backsub_cost_kernel) -------------
  for (int i = tid; i < o.n_cam; i += blockDim.x) {
    double R[9], t[3], Rc[9], tc[3];
#pragma unroll
    for (int j = 0; j < 9; ++j) R[j] = o.cam_R[9 * i + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) t[j] = o.cam_t[3 * i + j];
    const int slot = o.cam_slot[i];
    slot_sm[i] = slot;
    if (slot >= 0) {
      double d[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double v = A.dC[6 * slot + j];
        cam_sm[24 * o.n_cam + 6 * i + j] = v;
        d[j] = -v;
      }
      camera_retract(R, t, d, Rc, tc);
    } else {
#pragma unroll
      for (int j = 0; j < 6; ++j) cam_sm[24 * o.n_cam + 6 * i + j] = 0.0;
#pragma unroll
      for (int j = 0; j < 9; ++j) Rc[j] = R[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) tc[j] = t[j];
    }
    double* s0 = cam_sm + 12 * i;
    double* s1 = cam_sm + 12 * (o.n_cam + i);
#pragma unroll
    for (int j = 0; j < 9; ++j) { s0[j] = R[j]; s1[j] = Rc[j]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) { s0[9 + j] = t[j]; s1[9 + j] = tc[j]; }
    if (blockIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < 9; ++j) A.cand_R_out[9 * i + j] = Rc[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) A.cand_t_out[3 * i + j] = tc[j];
    }
  }

  double cost_acc = 0.0;
  for (; c_cur < n_chunks; c_cur += gridDim.x) {
    __syncthreads();   // the previous chunk's readers are done with the stage (first time round: the cameras are in place)
    if (tid <= np) S.ptr[tid] = r_ptr;
    if (tid < np) S.slot[tid] = r_slot;
    if (tid < 3 * np) { S.x[tid] = r_x; S.bP[tid] = r_bP; }
    if (tid < 9 * np) S.Vi[tid] = r_Vi0;
    if (tid + 256 < 9 * np) S.Vi[tid + 256] = r_Vi1;
    if (tid < nob) { S.cam[tid] = r_cam; S.uv[tid] = r_uv; }
    const int c_p0 = p0, c_np = np, c_ob0 = ob0, c_nob = nob;
    // next chunk's loads go out now and land while this chunk is computed
    p0 = n_p0; np = n_np; ob0 = n_ob0; nob = n_nob;
    bounds(c_cur + 2 * gridDim.x, n_p0, n_np, n_ob0, n_nob);
    fetch(p0, np, ob0, nob);
    __syncthreads();
    if (tid < c_np)    // the point of every observation (tracks are short: a few stores per point)
      for (int a = S.ptr[tid] - c_ob0, e = S.ptr[tid + 1] - c_ob0; a < e; ++a) S.pt_of[a] = (unsigned char)tid;
    __syncthreads();

    // ---- pass 1: thread = observation ------------------------------------------------------
    int lp = 0, cam = 0;
    bool use = false;
    double2 uv = make_double2(0.0, 0.0);
    if (tid < c_nob) {
      lp = S.pt_of[tid];
      cam = S.cam[tid];
      uv = S.uv[tid];
      use = S.slot[lp] >= 0 && slot_sm[cam] >= 0;
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
      if (use) {
        const double x[3] = {S.x[3 * lp], S.x[3 * lp + 1], S.x[3 * lp + 2]};
        double r[2], Jc[12], Jp[6];
        observe(o.intr, o.model, cam_sm + 12 * cam, cam_sm + 12 * cam + 9, x, uv.x, uv.y, r, Jc, Jp);
        const double* d = dC_sm + 6 * cam;
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) { q0 += Jc[j] * d[j]; q1 += Jc[6 + j] * d[j]; }
        c0 = Jp[0] * q0 + Jp[3] * q1;
        c1 = Jp[1] * q0 + Jp[4] * q1;
        c2 = Jp[2] * q0 + Jp[5] * q1;
      }
      S.c[3 * tid] = c0; S.c[3 * tid + 1] = c1; S.c[3 * tid + 2] = c2;
    }
    __syncthreads();
    // ---- pass 2: thread = point ---------------------------------------------------------------
    if (tid < c_np) {
      const int pt = c_p0 + tid;
      const double x0 = S.x[3 * tid], x1 = S.x[3 * tid + 1], x2 = S.x[3 * tid + 2];
      double d0 = 0.0, d1 = 0.0, d2 = 0.0;
      if (S.slot[tid] >= 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int a = S.ptr[tid] - c_ob0, e = S.ptr[tid + 1] - c_ob0; a < e; ++a) { a0 += S.c[3 * a]; a1 += S.c[3 * a + 1]; a2 += S.c[3 * a + 2]; }
        const double g0 = S.bP[3 * tid] - a0, g1 = S.bP[3 * tid + 1] - a1, g2 = S.bP[3 * tid + 2] - a2;
        const double* Vi = S.Vi + 9 * tid;
        d0 = Vi[0] * g0 + Vi[1] * g1 + Vi[2] * g2;
        d1 = Vi[3] * g0 + Vi[4] * g1 + Vi[5] * g2;
        d2 = Vi[6] * g0 + Vi[7] * g1 + Vi[8] * g2;
      }
      S.xc[3 * tid] = x0 - d0; S.xc[3 * tid + 1] = x1 - d1; S.xc[3 * tid + 2] = x2 - d2;
      A.dP[3 * (size_t)pt] = d0; A.dP[3 * (size_t)pt + 1] = d1; A.dP[3 * (size_t)pt + 2] = d2;
      A.cand_pts[3 * (size_t)pt] = x0 - d0; A.cand_pts[3 * (size_t)pt + 1] = x1 - d1; A.cand_pts[3 * (size_t)pt + 2] = x2 - d2;
    }
    __syncthreads();
    // ---- pass 3: thread = observation, residual of the candidate -----------------------------
    if (use) {
      const double xc[3] = {S.xc[3 * lp], S.xc[3 * lp + 1], S.xc[3 * lp + 2]};
      double r[2];
      residual_only(o.intr, o.model, cam_sm + 12 * (o.n_cam + cam), cam_sm + 12 * (o.n_cam + cam) + 9, xc, uv.x, uv.y, r);
      cost_acc += r[0] * r[0] + r[1] * r[1];
    }
  }
  backsub_cost_epilogue(warp_sum(cost_acc), A);
}

// ------------------------------------------------------------------------------------------
struct CostArgs {
  ObsArgs o;
  double* __restrict__ partials;
  unsigned int* __restrict__ ticket;
  double* __restrict__ cost_out;
};

__global__ void __launch_bounds__(256) cost_kernel(const CostArgs A) {
  const ObsArgs& o = A.o;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  double cost_acc = 0.0;
  for (int pt = blockIdx.x * warps_per_cta + wid; pt < o.n_pt; pt += gridDim.x * warps_per_cta) {
    if (__ldcs(o.pt_slot + (pt)) < 0) continue;
    const int beg = __ldcs(o.pt_ptr + (pt));
    const int k = __ldcs(o.pt_ptr + (pt + 1)) - beg;
    const double x[3] = {__ldcs(o.pts + (3 * pt)), __ldcs(o.pts + (3 * pt + 1)), __ldcs(o.pts + (3 * pt + 2))};
    for (int a = lane; a < k; a += 32) {
      const int ob = beg + a;
      const int cam = __ldcs(o.obs_cam + (ob));
      if (o.cam_slot[cam] < 0) continue;
      const double2 uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (ob));
      double r[2];
      residual_only(o.intr, o.model, o.cam_R + 9 * cam, o.cam_t + 3 * cam, x, uv.x, uv.y, r);
      cost_acc += r[0] * r[0] + r[1] * r[1];
    }
  }
  grid_sum_store(warp_sum(cost_acc), A.partials, A.ticket, A.cost_out);
}

// ------------------------------------------------------------------------------------------
struct EvalArgs {
  ObsArgs o;
  double* __restrict__ r;   // [n_obs][2]
  double* __restrict__ Jc;  // [n_obs][12]
  double* __restrict__ Jp;  // [n_obs][6]
};

__global__ void __launch_bounds__(256) eval_observations_kernel(const EvalArgs A) {
  const ObsArgs& o = A.o;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  for (int pt = blockIdx.x * warps_per_cta + wid; pt < o.n_pt; pt += gridDim.x * warps_per_cta) {
    const int beg = __ldcs(o.pt_ptr + (pt));
    const int k = __ldcs(o.pt_ptr + (pt + 1)) - beg;
    const double x[3] = {__ldcs(o.pts + (3 * pt)), __ldcs(o.pts + (3 * pt + 1)), __ldcs(o.pts + (3 * pt + 2))};
    for (int a = lane; a < k; a += 32) {
      const int ob = beg + a;
      const int cam = __ldcs(o.obs_cam + (ob));
      const double2 uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + (ob));
      double r[2], Jc[12], Jp[6];
      observe(o.intr, o.model, o.cam_R + 9 * cam, o.cam_t + 3 * cam, x, uv.x, uv.y, r, Jc, Jp);
      A.r[2 * ob] = r[0]; A.r[2 * ob + 1] = r[1];
#pragma unroll
      for (int i = 0; i < 12; ++i) A.Jc[(size_t)ob * 12 + i] = Jc[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) A.Jp[(size_t)ob * 6 + i] = Jp[i];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Linear multi-view triangulation of every point (triangulate.py:6-18, Bundle.triangulate_all
// bundle.py:313-321): the two algebraic constraints per observation
//     (K0 - u K2)(R x + t) = 0,   (K1 - v K2)(R x + t) = 0
// are accumulated as 3x3 normal equations by a warp per point and solved with the symmetric
// pseudo-inverse (eigen-directions below 1e-13 of the largest eigenvalue are dropped, which is
// numpy.linalg.lstsq's minimum-norm answer for a track that is seen by a single camera).
struct TriArgs {
  ObsArgs o;
  double* __restrict__ out;   // [n_pt][3]
};

__global__ void __launch_bounds__(256) triangulate_kernel(const TriArgs A) {
  const ObsArgs& o = A.o;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const double* K = o.intr.K;
  for (int pt = blockIdx.x * warps_per_cta + wid; pt < o.n_pt; pt += gridDim.x * warps_per_cta) {
    const int beg = __ldcs(o.pt_ptr + pt);
    const int k = __ldcs(o.pt_ptr + pt + 1) - beg;
    double M[6] = {0, 0, 0, 0, 0, 0}, v[3] = {0, 0, 0};
    for (int a = lane; a < k; a += 32) {
      const int ob = beg + a;
      const int cam = __ldcs(o.obs_cam + ob);
      const double2 uv = __ldcs(reinterpret_cast<const double2*>(o.obs_uv) + ob);
      const double* R = o.cam_R + 9 * cam;
      const double* t = o.cam_t + 3 * cam;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double m = e ? uv.y : uv.x;
        const double c0 = K[3 * e] - m * K[6], c1 = K[3 * e + 1] - m * K[7], c2 = K[3 * e + 2] - m * K[8];
        const double r0 = c0 * R[0] + c1 * R[3] + c2 * R[6];
        const double r1 = c0 * R[1] + c1 * R[4] + c2 * R[7];
        const double r2 = c0 * R[2] + c1 * R[5] + c2 * R[8];
        const double bb = -(c0 * t[0] + c1 * t[1] + c2 * t[2]);
        M[0] += r0 * r0; M[1] += r0 * r1; M[2] += r0 * r2;
        M[3] += r1 * r1; M[4] += r1 * r2; M[5] += r2 * r2;
        v[0] += r0 * bb; v[1] += r1 * bb; v[2] += r2 * bb;
      }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) M[i] = warp_sum(M[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
      const double Mf[9] = {M[0], M[1], M[2], M[1], M[3], M[4], M[2], M[4], M[5]};
      double Mi[9];
      sym3_pinv(Mf, 1e-13, Mi);
#pragma unroll
      for (int i = 0; i < 3; ++i) A.out[3 * pt + i] = Mi[3 * i] * v[0] + Mi[3 * i + 1] * v[1] + Mi[3 * i + 2] * v[2];
    }
  }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
static ObsArgs make_obs_args(const Context& c, const ParamSet& ps) {
  ObsArgs o;
  o.intr = c.intr;
  o.model = c.model;
  o.n_cam = c.n_cam; o.n_pt = c.n_pt; o.n_obs = c.n_obs;
  o.pt_ptr = c.pt_ptr; o.obs_cam = c.obs_cam; o.obs_uv = c.obs_uv;
  o.cam_slot = c.cam_slot; o.pt_slot = c.pt_slot;
  o.cam_R = ps.cam_R; o.cam_t = ps.cam_t; o.pts = ps.pts;
  return o;
}

static int point_grid(const Context& c, int warps_per_cta, int ctas_per_sm) {
  const int want = (c.n_pt + warps_per_cta - 1) / warps_per_cta;
  const int cap = c.num_sms * ctas_per_sm;
  int g = want < cap ? want : cap;
  if (g < 1) g = 1;
  if (g > c.partials_cap) g = c.partials_cap;
  return g;
}

cudaError_t launch_linearize_eliminate(Context& c, double damping, double rcond, int flags,
                                       cudaStream_t st) {
  const bool blocks = flags & 1, schur = flags & 2;
  cudaError_t e;
  const int probe = (flags >> 4) & 3;
  if (schur) {
    e = cudaMemsetAsync(c.sys, 0, c.sys_len * sizeof(double), st);
    if (e != cudaSuccess) return e;
  }
  if (blocks) {
    if ((e = cudaMemsetAsync(c.U, 0, (size_t)c.n_cam * 36 * sizeof(double), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c.bC, 0, (size_t)c.n_cam * 6 * sizeof(double), st)) != cudaSuccess) return e;
  }
  ElimArgs A;
  A.o = make_obs_args(c, c.state);
  A.damping = damping; A.rcond = rcond; A.n_opt_cam = c.n_opt_cam;
  A.probe = probe;
  A.rhs_off = c.sys_len - (size_t)c.n_sys;
  int kcap = c.max_track_len < 1 ? 1 : c.max_track_len;
  A.kcap = kcap;
  A.sys = c.sys; A.Vinv = c.Vinv; A.bP = c.bP; A.V = c.V; A.U = c.U; A.bC = c.bC; A.W = c.W;
  A.partials = c.partials; A.ticket = c.counters; A.cost_out = &c.scalars->cost;
  // Short tracks, Schur path only: several points per warp pass (eliminate_group_kernel).  OFF by
  // default -- parity-clean over the whole GPU suite, but 0.320 ms against 0.279 ms at config 2:
  // the kernel is bound by the drain of its reductions, which 12 warps per SM overlap worse than
  // 16, not by the per-observation phases this variant speeds up.  Opt in with
  // PYSFM_B200_ELIM_GROUP=1 (experiments only).
  {
    static const bool group_on = getenv("PYSFM_B200_ELIM_GROUP") && getenv("PYSFM_B200_ELIM_GROUP")[0] == '1';
    const int npg = 3 * kcap <= 32 ? 3 : (2 * kcap <= 32 ? 2 : 1);
    if (schur && !blocks && probe == 0 && npg > 1 && group_on) {
      A.gcap = (npg * kcap + 1) & ~1;
      const int gw = 6;   // warps per CTA, 2 CTAs per SM
      const size_t gsmem = (size_t)elim_group_warp_doubles(A.gcap) * sizeof(double) * gw;
      typedef void (*GroupKernel)(const ElimArgs);
      GroupKernel gk = npg == 3 ? eliminate_group_kernel<3> : eliminate_group_kernel<2>;
      bool& attr = c.elim_group_attr_set[npg - 2];
      if (!attr) {
        if ((e = cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024)) != cudaSuccess) return e;
        attr = true;
      }
      const int n_groups = (c.n_pt + npg - 1) / npg;
      int grid = (n_groups + gw - 1) / gw;
      if (grid > 2 * c.num_sms) grid = 2 * c.num_sms;
      if (grid < 1) grid = 1;
      if (grid > c.partials_cap) grid = c.partials_cap;
      A.diag_rep = c.diag_rep;
      gk<<<grid, gw * 32, gsmem, st>>>(A);
      c.launches += 1;
      if (c.n_opt_cam > 0) {
        fold_diag_kernel<<<c.n_opt_cam, 64, 0, st>>>(c.sys, c.diag_rep, c.n_opt_cam);
        c.launches += 1;
      }
      return cudaGetLastError();
    }
  }
  const size_t budget = 200 * 1024;
  // records in shared memory while at least two warps fit a CTA; otherwise (a track of more than
  // ~210 views somewhere) in the global scratch array, which is allocated on first use
  const bool rec_global = (size_t)elim_warp_doubles(kcap) * sizeof(double) * 2 > budget;
  if (rec_global && !c.rec_scratch) {
    if ((e = cudaMalloc((void**)&c.rec_scratch, (size_t)(c.n_obs > 0 ? c.n_obs : 1) * 56 * sizeof(double))) != cudaSuccess) return e;
  }
  A.rec_global = c.rec_scratch;
  A.diag_rep = c.diag_rep;
  const size_t per_warp = (size_t)elim_warp_doubles(rec_global ? 0 : kcap) * sizeof(double);
  int warps = (int)(budget / per_warp);
  if (warps > 8) warps = 8;
  const size_t smem = per_warp * warps;
  int ctas_per_sm = (int)((budget + 24 * 1024) / (smem + 1024));
  if (ctas_per_sm > 8 / warps * 4) ctas_per_sm = 8 / warps * 4;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 2) ctas_per_sm = 2;   // __launch_bounds__(256, 2)
  const int grid = point_grid(c, warps, ctas_per_sm);
  typedef void (*ElimKernel)(const ElimArgs);
  static const ElimKernel table[8] = {
      linearize_eliminate_kernel<false, false, false>, linearize_eliminate_kernel<false, true, false>,
      linearize_eliminate_kernel<true, false, false>,  linearize_eliminate_kernel<true, true, false>,
      linearize_eliminate_kernel<false, false, true>,  linearize_eliminate_kernel<false, true, true>,
      linearize_eliminate_kernel<true, false, true>,   linearize_eliminate_kernel<true, true, true>};
  const int variant = (rec_global ? 4 : 0) + (blocks ? 2 : 0) + (schur ? 1 : 0);
  ElimKernel kern = table[variant];
  if (!c.elim_attr_set[variant]) {   // opt in to the full shared-memory carve-out once
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
    c.elim_attr_set[variant] = true;
  }
  kern<<<grid, warps * 32, smem, st>>>(A);
  c.launches += 1;
  if (schur && c.n_opt_cam > 0) {
    fold_diag_kernel<<<c.n_opt_cam, 64, 0, st>>>(c.sys, c.diag_rep, c.n_opt_cam);
    c.launches += 1;
  }
  return cudaGetLastError();
}

cudaError_t launch_backsub_retract_cost(Context& c, cudaStream_t st) {
  // cameras in shared memory (and retracted inside the kernel) while 24 doubles per camera fit
  // three resident CTAs per SM; otherwise the stand-alone retraction kernel and global gathers
  const size_t cam_smem = (size_t)c.n_cam * (30 * sizeof(double) + sizeof(int));   // {R,t} state + candidate, dC, slot
  const bool sc = cam_smem <= 64 * 1024;
  if (!sc) {
    const int tb = 128;
    retract_cameras_kernel<<<(c.n_cam + tb - 1) / tb, tb, 0, st>>>(
        c.n_cam, c.cam_slot, c.state.cam_R, c.state.cam_t, c.dC, -1.0, c.cand.cam_R, c.cand.cam_t);
    c.launches += 1;
  }
  BacksubArgs A;
  A.o = make_obs_args(c, c.state);
  A.cand_R = c.cand.cam_R; A.cand_t = c.cand.cam_t; A.cand_pts = c.cand.pts;
  A.cand_R_out = c.cand.cam_R; A.cand_t_out = c.cand.cam_t;
  A.dC = c.dC; A.Vinv = c.Vinv; A.bP = c.bP; A.dP = c.dP;
  A.partials = c.partials; A.ticket = c.counters + 1; A.cost_out = &c.scalars->cand_cost;
  A.sc = c.scalars;
  A.peer.world = 0;
  c.costs_reduced = false;
  if (c.fuse_cost_reduction && c.comm_buf && c.comm_world > 1) {
    bool connected = true;
    for (int p = 0; p < c.comm_world; ++p) connected = connected && c.comm_peer[p] != nullptr;
    if (connected) {
      A.peer = make_peer_args(c);       // (one epoch of the handle's collectives, like ba_allreduce_costs)
      c.costs_reduced = true;           // ba_allreduce_costs has nothing left to do for this trial
    }
  }
  // lanes per point: the narrowest group that holds the longest track in at most two passes
  // (a group of 8 with two lanes doing a second observation beats a group of 16 with six idle
  // lanes at 10 observations per point: 46 -> 43 us -- the kernel is bound by loads in flight)
  const int kmax = c.max_track_len < 1 ? 32 : c.max_track_len;
  // (a software-pipelined single-pass variant at 2 CTAs/SM measured 73 us against 56 us for this
  // one at 3 CTAs/SM: occupancy beats a shorter dependency chain here)
  const int g = kmax <= 12 ? 8 : (kmax <= 24 ? 16 : 32);
  cudaError_t e = cudaSuccess;
  // tile-streamed, observation-parallel variant: cameras in shared memory and chunks of >= 4 points
  // within kTileObs observations (one per thread)
  {
    static const bool tile_off = getenv("PYSFM_B200_BACKSUB_TILE") && getenv("PYSFM_B200_BACKSUB_TILE")[0] == '0';
    int P = kTileObs / kmax;
    if (P > kTilePts) P = kTilePts;
    const size_t tile_smem = (((size_t)31 * c.n_cam + 2) & ~(size_t)1) * sizeof(double) + sizeof(TileStage);
    if (sc && !tile_off && P >= 4 && c.max_track_len >= 1 && tile_smem <= 100 * 1024) {
      if (!c.backsub_tile_attr_set[0]) {
        if ((e = cudaFuncSetAttribute(backsub_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
        c.backsub_tile_attr_set[0] = true;
      }
      const int n_chunks = (c.n_pt + P - 1) / P;
      int grid = n_chunks < 2 * c.num_sms ? n_chunks : 2 * c.num_sms;
      if (grid > c.partials_cap) grid = c.partials_cap;
      backsub_tile_kernel<<<grid, 256, tile_smem, st>>>(A, P);
      c.launches += 1;
      return cudaGetLastError();
    }
  }
  const int grid = point_grid(c, 8 * (32 / g), 3);
#define BA_LAUNCH_BACKSUB(G)                                                                          \
  do {                                                                                                \
    if (sc) {                                                                                         \
      if (!c.backsub_attr_set[G == 8 ? 0 : (G == 16 ? 1 : 2)]) {                                      \
        e = cudaFuncSetAttribute(backsub_cost_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
        if (e != cudaSuccess) return e;                                                               \
        c.backsub_attr_set[G == 8 ? 0 : (G == 16 ? 1 : 2)] = true;                                    \
      }                                                                                               \
      backsub_cost_kernel<G, true><<<grid, 256, cam_smem, st>>>(A);                                   \
    } else {                                                                                          \
      backsub_cost_kernel<G, false><<<grid, 256, 0, st>>>(A);                                         \
    }                                                                                                 \
  } while (0)
  if (g == 8) BA_LAUNCH_BACKSUB(8);
  else if (g == 16) BA_LAUNCH_BACKSUB(16);
  else BA_LAUNCH_BACKSUB(32);
#undef BA_LAUNCH_BACKSUB
  c.launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_cost(Context& c, cudaStream_t st) {
  CostArgs A;
  A.o = make_obs_args(c, c.state);
  A.partials = c.partials; A.ticket = c.counters; A.cost_out = &c.scalars->cost;
  const int grid = point_grid(c, 8, 8);
  cost_kernel<<<grid, 256, 0, st>>>(A);
  c.launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_eval_observations(Context& c, cudaStream_t st) {
  EvalArgs A;
  A.o = make_obs_args(c, c.state);
  A.r = c.obs_r; A.Jc = c.obs_Jc; A.Jp = c.obs_Jp;
  const int grid = point_grid(c, 8, 8);
  eval_observations_kernel<<<grid, 256, 0, st>>>(A);
  c.launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_triangulate(Context& c, cudaStream_t st) {
  TriArgs A;
  A.o = make_obs_args(c, c.state);
  A.out = c.state.pts;
  const int grid = point_grid(c, 8, 8);
  triangulate_kernel<<<grid, 256, 0, st>>>(A);
  c.launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_retract(Context& c, bool have_cam, bool have_pt, cudaStream_t st) {
  const int tb = 128;
  retract_cameras_kernel<<<(c.n_cam + tb - 1) / tb, tb, 0, st>>>(
      c.n_cam, c.cam_slot, c.state.cam_R, c.state.cam_t, have_cam ? c.delta_cam : nullptr, 1.0,
      c.cand.cam_R, c.cand.cam_t);
  retract_points_kernel<<<(c.n_pt + tb - 1) / tb, tb, 0, st>>>(
      c.n_pt, c.pt_slot, c.state.pts, have_pt ? c.delta_pt : nullptr, c.cand.pts);
  c.launches += 2;
  return cudaGetLastError();
}

}  // namespace ba
