// Microbenchmark: how fast can 6x6 FP64 blocks be accumulated into an L2-resident block array?
//   mode 0: scalar red.global.add.f64, each lane 1 contiguous element of a 36-double block (36 of 64 lanes used over 2 instr)
//   mode 1: scalar REDG in the r1a pattern (lane = column, 6 strided rows)  [dense ld layout]
//   mode 2: cp.reduce.async.bulk .add.f64, one 288-byte op per block, issued by one lane per block
//   mode 3: as 2 but two 6x6 blocks per op where (576 B) -- upper bound for bigger ops
//   modes 5-8 (r1q, questions left by tools/elim_probe.py: the elimination kernel drains its
//   reductions at 63 % of mode 2's rate): 288-byte bulk reductions with the kernel's own
//   constraints -- ONE staging slot per lane (wait_group.read 0 before every round) and/or the
//   kernel's address distribution (10 of the 55 blocks of a point hit one of the 199 DIAGONAL
//   blocks, 2,500 updates each per iteration).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// every warp processes `groups` groups of 55 blocks (one "point" with k=10)
__global__ void __launch_bounds__(256) k_scalar_blockmajor(double* S, int nblocks, int groups) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int g = 0; g < groups; ++g) {
    for (int b = 0; b < 55; ++b) {
      const uint32_t blk = hash32(warp * 7919u + g * 104729u + b * 13u) % nblocks;
      double* p = S + (size_t)blk * 36;
      atomicAdd(p + lane, 1.0);
      if (lane < 4) atomicAdd(p + 32 + lane, 1.0);
    }
  }
}

// r1a pattern: dense matrix ld; for 10 obs: lanes span (b,cc), loop a<=b, 6 rows
__global__ void __launch_bounds__(256) k_scalar_dense(double* S, int ncam, int ld, int groups) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int g = 0; g < groups; ++g) {
    int cams[10];
    uint32_t base = hash32(warp * 7919u + g * 104729u) % (ncam - 10 * 19);
    for (int i = 0; i < 10; ++i) cams[i] = base + i * 19 + (hash32(warp + g * 31 + i) % 19);
    for (int s0 = 0; s0 < 60; s0 += 32) {
      const int s = s0 + lane;
      const bool valid = s < 60;
      const int b = valid ? s / 6 : 0, cc = valid ? s % 6 : 0;
      for (int a = 0; a < 10; ++a) {
        if (valid && a <= b) {
          double* p = S + (size_t)(6 * cams[a]) * ld + 6 * cams[b] + cc;
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) atomicAdd(p + (size_t)rr * ld, 1.0);
        }
      }
    }
  }
}

template <int BYTES>
__global__ void __launch_bounds__(256) k_bulk(double* S, int nblocks, int groups) {
  extern __shared__ __align__(128) double stage[];  // per warp: 32 * BYTES
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double* my = stage + (size_t)wid * 32 * (BYTES / 8);
  for (int i = lane; i < 32 * (BYTES / 8); i += 32) my[i] = 1.0;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  constexpr int per_group = 55 * 288 / BYTES;  // same number of bytes per group
  for (int g = 0; g < groups; ++g) {
    for (int b0 = 0; b0 < per_group; b0 += 32) {
      const int b = b0 + lane;
      if (b < per_group) {
        const uint32_t blk = hash32(warp * 7919u + g * 104729u + b * 13u) % nblocks;
        double* dst = S + (size_t)blk * (BYTES / 8);
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(my + (size_t)lane * (BYTES / 8));
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                     :: "l"(dst), "r"(src), "n"(BYTES) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// HOT: 10 of the 55 blocks of a group go to diagonal blocks (a, a), as in the elimination kernel;
// DEPTH: bulk groups a lane may have in flight before it reuses its staging slot (kernel: 0).
template <bool HOT, int DEPTH>
__global__ void __launch_bounds__(256) k_bulk_like_kernel(double* S, int ncam, int nblocks, int groups) {
  extern __shared__ __align__(128) double stage[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double* my = stage + (size_t)wid * 32 * 36;
  for (int i = lane; i < 32 * 36; i += 32) my[i] = 1.0;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  for (int g = 0; g < groups; ++g) {
    for (int b0 = 0; b0 < 55; b0 += 32) {
      const int b = b0 + lane;
      if (DEPTH == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (b < 55) {
        uint32_t blk = hash32(warp * 7919u + g * 104729u + b * 13u) % nblocks;
        if (HOT && b < 10) {   // diagonal block (a, a) of a random camera: packed index a*ncam - a(a-1)/2
          const uint32_t a = hash32(warp * 31u + g * 17u + b) % ncam;
          blk = a * ncam - a * (a - 1) / 2;
        }
        double* dst = S + (size_t)blk * 36;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(my + (size_t)lane * 36);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 288;" ::"l"(dst), "r"(src) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (DEPTH == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int ncam = 199, ld = 1216;
  const int nblocks = ncam * (ncam + 1) / 2;
  const int total_groups = 50000;
  double *S, *D;
  CK(cudaMalloc(&S, (size_t)nblocks * 36 * 2 * sizeof(double)));
  CK(cudaMalloc(&D, (size_t)ld * ld * sizeof(double)));
  CK(cudaMemset(S, 0, (size_t)nblocks * 36 * 2 * sizeof(double)));
  CK(cudaMemset(D, 0, (size_t)ld * ld * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(k_bulk_like_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 288));
  CK(cudaFuncSetAttribute(k_bulk_like_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 288));
  CK(cudaFuncSetAttribute(k_bulk_like_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 288));
  CK(cudaFuncSetAttribute(k_bulk_like_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 288));
  for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
    const int grid = 148 * ctas_per_sm, warps = grid * 8;
    const int groups = (total_groups + warps - 1) / warps;
    const double nblk_ops = (double)groups * warps * 55;
    for (int mode = 0; mode < 9; ++mode) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) k_scalar_blockmajor<<<grid, 256>>>(S, nblocks, groups);
        else if (mode == 1) k_scalar_dense<<<grid, 256>>>(D, ncam, ld, groups);
        else if (mode == 2) { CK(cudaFuncSetAttribute(k_bulk<288>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 288)); k_bulk<288><<<grid, 256, 8 * 32 * 288>>>(S, nblocks, groups); }
        else if (mode == 3) { CK(cudaFuncSetAttribute(k_bulk<576>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 576)); k_bulk<576><<<grid, 256, 8 * 32 * 576>>>(S, nblocks / 2, groups); }
        else if (mode == 5) k_bulk_like_kernel<false, 1><<<grid, 256, 8 * 32 * 288>>>(S, ncam, nblocks, groups);
        else if (mode == 6) k_bulk_like_kernel<false, 0><<<grid, 256, 8 * 32 * 288>>>(S, ncam, nblocks, groups);
        else if (mode == 7) k_bulk_like_kernel<true, 1><<<grid, 256, 8 * 32 * 288>>>(S, ncam, nblocks, groups);
        else if (mode == 8) k_bulk_like_kernel<true, 0><<<grid, 256, 8 * 32 * 288>>>(S, ncam, nblocks, groups);
        else { CK(cudaFuncSetAttribute(k_bulk<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 96)); k_bulk<96><<<grid, 256, 8 * 32 * 96>>>(S, nblocks * 3, groups); }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      const char* names[] = {"scalar REDG, block-major contiguous", "scalar REDG, r1a dense pattern",
                             "bulk reduce 288 B", "bulk reduce 576 B", "bulk reduce 96 B",
                             "288 B, uniform blocks, 2 groups in flight", "288 B, uniform blocks, 1 slot (kernel)",
                             "288 B, hot diagonal blocks, 2 in flight", "288 B, hot diagonal blocks, 1 slot (kernel)"};
      printf("ctas/sm=%d  %-38s  %8.3f ms   %.3g block-equivalents/s  (%.3g f64 adds/s)\n", ctas_per_sm, names[mode], best,
             nblk_ops / (best * 1e-3), nblk_ops * 36 / (best * 1e-3));
    }
  }
  double h[4];
  CK(cudaMemcpy(h, S, sizeof h, cudaMemcpyDeviceToHost));
  printf("check %g %g\n", h[0], h[1]);
  return 0;
}
