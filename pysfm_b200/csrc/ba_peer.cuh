// Device-side pieces of the peer-memory collectives shared by ba_comm.cu (the stand-alone
// all-reduce kernels) and ba_kernels.cu (the cost reduction fused into the back-substitution).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_context.h"

namespace ba {

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 ld_peer2(const double* p) {   // never cached: peers rewrite it every iteration
  double2 v;
  asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}

struct PeerArgs {
  double* base[kMaxPeers];   // comm buffer of every rank (own rank included), peer-mapped
  int world, rank;
  size_t sys_len;            // doubles
  size_t lo, hi;             // this rank's slice [lo, hi) of the packed system (even bounds)
  unsigned int epoch;
  unsigned int* done;        // local CTA counter (last-CTA-done)
  double* status;            // local scalar status word: 2 when a barrier ran past its deadline
  unsigned long long spin_limit_ns;
};

__device__ __forceinline__ unsigned long long comm_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait for *f to reach the epoch; gives up (status 2 -> BA_ERR_TIMEOUT) when a peer never arrives
__device__ __forceinline__ void wait_arrival(const PeerArgs& g, const unsigned int* f) {
  const unsigned long long t0 = comm_ns();
  unsigned int spins = 0;
  while ((int)(ld_acquire_sys(f) - g.epoch) < 0) {
    if ((++spins & 255u) == 0u && comm_ns() - t0 > g.spin_limit_ns) {
      *g.status = 2.0;
      break;
    }
    __nanosleep(40);
  }
}

__device__ __forceinline__ double* contrib_of(double* base) { return base; }
__device__ __forceinline__ double* reduced_of(double* base, size_t sys_len) { return base + comm_pad(sys_len); }
__device__ __forceinline__ double* costs_of(double* base, size_t sys_len) { return base + 2 * comm_pad(sys_len); }
__device__ __forceinline__ unsigned int* flags_of(double* base, size_t sys_len) {
  return reinterpret_cast<unsigned int*>(base + 2 * comm_pad(sys_len) + 4 * kMaxPeers);
}

// signal barrier `which` to every rank, then wait until every rank has signalled us
__device__ __forceinline__ void peer_barrier(const PeerArgs& g, int which, int lane_in_block) {
  if (lane_in_block < g.world) {
    __threadfence_system();
    st_release_sys(flags_of(g.base[lane_in_block], g.sys_len) + which * kMaxPeers + g.rank, g.epoch);
    wait_arrival(g, flags_of(g.base[g.rank], g.sys_len) + which * kMaxPeers + lane_in_block);
  }
}


// host: the argument block of the next collective (bumps the handle's epoch)
PeerArgs make_peer_args(Context& c);

}  // namespace ba
