// Peer-memory collectives for points sharded over the GPUs of ONE node.
//
// Every rank owns one IPC-exported device buffer:
//     contrib [sys_len]   its own packed reduced system (the elimination kernel reduces into it)
//     reduced [sys_len]   the all-reduced system (peers push their slices here; the solver reads it)
//     costs   [2][world][2]  {cost, candidate cost} of every rank (pushed by the ranks), two banks
//                         used in turn (epoch parity): a fast rank's next reduction cannot
//                         overwrite values a slower rank has not summed yet
//     flags   [3][world]  arrival flags of the three barriers (epoch valued, never reset)
// and maps the buffers of all peers (NVLink 5 / NVSwitch: every peer at full bandwidth).
//
//   peer_allreduce_system_kernel   ONE kernel per LM iteration instead of an NCCL all-reduce:
//       barrier A (everybody's elimination is complete)  ->  rank r sums slice r of all ranks'
//       contributions over NVLink, in rank order, and pushes the result into EVERY rank's
//       `reduced`  ->  barrier B (all slices have landed).  2 (N-1)/N |sys| bytes per rank cross
//       the links, the same as a ring all-reduce, but in two hops and one launch; each element is
//       reduced by exactly one rank, so all ranks factor bit-identical systems.
//   peer_allreduce_costs_kernel    the two cost scalars, summed in rank order on every rank.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_context.h"
#include "ba_peer.cuh"

namespace ba {

__global__ void __launch_bounds__(256) peer_allreduce_system_kernel(const PeerArgs g) {
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  // ---- barrier A: every rank's contribution is complete (its elimination kernel precedes this
  // launch in its stream); CTA 0 signals, every CTA waits on the local flags -------------------
  if (blockIdx.x == 0 && tid < g.world) {
    __threadfence_system();
    st_release_sys(flags_of(g.base[tid], g.sys_len) + 0 * kMaxPeers + g.rank, g.epoch);
  }
  if (tid < g.world) wait_arrival(g, flags_of(g.base[g.rank], g.sys_len) + 0 * kMaxPeers + tid);
  __syncthreads();
  // ---- reduce my slice over all ranks (rank order), push it to everybody -----------------------
  // all peers' loads of an element are in flight together (NVLink round trip ~1 us), then the sum
  const size_t n2 = (g.hi - g.lo) / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = g.lo + 2 * i;
    double2 v[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < g.world) v[p] = ld_peer2(contrib_of(g.base[p]) + e);
    double2 s = v[0];
#pragma unroll
    for (int p = 1; p < kMaxPeers; ++p)
      if (p < g.world) { s.x += v[p].x; s.y += v[p].y; }
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < g.world) *reinterpret_cast<double2*>(reduced_of(g.base[p], g.sys_len) + e) = s;
  }
  // ---- barrier B: the last CTA of this rank tells everybody that its slice has landed, then
  // waits for everybody else's ----------------------------------------------------------------
  __threadfence_system();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(g.done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (tid == 0) *g.done = 0u;
    peer_barrier(g, 1, tid);
  }
}

__global__ void peer_allreduce_costs_kernel(const PeerArgs g, Scalars* sc) {
  const int tid = threadIdx.x;
  const int bank = (int)(g.epoch & 1u) * 2 * kMaxPeers;
  if (tid < g.world) {
    double* c = costs_of(g.base[tid], g.sys_len) + bank + 2 * g.rank;
    c[0] = sc->cost;
    c[1] = sc->cand_cost;
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

PeerArgs make_peer_args(Context& c) {
  PeerArgs g;
  for (int p = 0; p < kMaxPeers; ++p) g.base[p] = p < c.comm_world ? c.comm_peer[p] : nullptr;
  g.world = c.comm_world; g.rank = c.comm_rank; g.sys_len = c.sys_len;
  size_t chunk = (c.sys_len + c.comm_world - 1) / c.comm_world;
  chunk = (chunk + 1) & ~(size_t)1;
  const size_t padded = (c.sys_len + 1) & ~(size_t)1;   // the buffer is padded to an even length
  g.lo = chunk * c.comm_rank < padded ? chunk * c.comm_rank : padded;
  g.hi = g.lo + chunk < padded ? g.lo + chunk : padded;
  g.epoch = ++c.comm_epoch;
  g.done = c.comm_done;
  g.status = &c.scalars->status;
  g.spin_limit_ns = (unsigned long long)(c.spin_timeout_ms * 1e6);
  return g;
}

cudaError_t launch_peer_allreduce_system(Context& c, cudaStream_t st) {
  const PeerArgs g = make_peer_args(c);
  const size_t n2 = (g.hi - g.lo) / 2;
  int grid = (int)((n2 + 255) / 256);
  if (grid > 4 * c.num_sms) grid = 4 * c.num_sms;
  if (grid < 1) grid = 1;
  peer_allreduce_system_kernel<<<grid, 256, 0, st>>>(g);
  c.launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_peer_allreduce_costs(Context& c, cudaStream_t st) {
  const PeerArgs g = make_peer_args(c);
  peer_allreduce_costs_kernel<<<1, 32, 0, st>>>(g, c.scalars);
  c.launches += 1;
  return cudaGetLastError();
}

}  // namespace ba
