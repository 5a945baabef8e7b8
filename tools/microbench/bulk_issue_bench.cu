// How expensive is ISSUING cp.reduce.async.bulk (UBLKRED) ops?  One warp stages 32 blocks of 288 B and
// reduces them into random 288-byte slots of an L2-resident array, round after round.
//   V0: every lane issues its own op (the compiler serialises: ELECT / R2UR x3 / UBLKRED / branch)
//   V1: lane 0 issues all 32 ops, destination indices read back from shared memory, unrolled by 8
//   V2: lanes 0, 8, 16, 24 issue 8 ops each
// Reports clocks per op per warp and chip-wide adds/s for 1, 4, 8, 16 warps per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ void bulk_add(double* dst, const double* src_smem) {
  const unsigned int src = (unsigned int)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 288;" ::"l"(dst), "r"(src) : "memory");
}

template <int V>
__global__ void __launch_bounds__(512) k_issue(double* S, int nblocks, int rounds, long long* clk) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double* stage = sm + (size_t)wid * (32 * 38 + 32);
  int* dsti = reinterpret_cast<int*>(stage + 32 * 38);
  for (int i = lane; i < 32 * 38; i += 32) stage[i] = 1.0;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const uint32_t blk = hash32(warp * 7919u + r * 104729u + lane * 13u) % nblocks;
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (V == 0) {
      bulk_add(S + (size_t)blk * 36, stage + lane * 38);
    } else {
      dsti[lane] = (int)blk;
      __syncwarp();
      if (V == 1) {
        if (lane == 0) {
#pragma unroll 8
          for (int i = 0; i < 32; ++i) bulk_add(S + (size_t)dsti[i] * 36, stage + i * 38);
        }
      } else {
        if ((lane & 7) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bulk_add(S + (size_t)dsti[lane + i] * 36, stage + (lane + i) * 38);
        }
      }
      __syncwarp();
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  long long t1 = clock64();
  if (lane == 0 && warp == 0) clk[0] = t1 - t0;
}

int main() {
  const int nblocks = 199 * 200 / 2;
  double* S; long long* clk;
  CK(cudaMalloc(&S, (size_t)nblocks * 36 * 8)); CK(cudaMemset(S, 0, (size_t)nblocks * 36 * 8));
  CK(cudaMalloc(&clk, 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int rounds = 200;
  for (int wps = 1; wps <= 16; wps *= 2) {
    for (int v = 0; v < 3; ++v) {
      const size_t smem = (size_t)wps * (32 * 38 + 32) * 8;
      auto kern = v == 0 ? k_issue<0> : v == 1 ? k_issue<1> : k_issue<2>;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<148, wps * 32, smem>>>(S, nblocks, rounds, clk);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      long long h; CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
      const double ops = 148.0 * wps * 32 * rounds;
      printf("warps/SM=%2d  V%d  %8.3f ms  %6.1f clk/op/warp  %.3g adds/s\n", wps, v, best, (double)h / (32.0 * rounds), ops * 36 / (best * 1e-3));
    }
  }
  return 0;
}
