// FP64 throughput on one SM and on the chip: DFMA vs mma.sync.m8n8k4.f64 (DMMA).
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__global__ void k_dfma(double* out, int iters) {
  double a[8], x = threadIdx.x * 1e-3, y = 1.0000001;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, x);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
  double c[8][2], a = threadIdx.x * 1e-3, b = 1.0000001;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chains: latency
__global__ void k_dfma_lat(double* out, int iters) {
  double a = threadIdx.x, y = 1.0000001, x = 1e-3;
  for (int it = 0; it < iters; ++it) a = fma(a, y, x);
  out[threadIdx.x] = a;
}
__global__ void k_dmma_lat(double* out, int iters) {
  double c0 = 1, c1 = 2, a = threadIdx.x * 1e-3, b = 1.0000001;
  for (int it = 0; it < iters; ++it)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  out[threadIdx.x] = c0 + c1;
}
__global__ void k_div_lat(double* out, int iters) {
  double a = 1.0 + threadIdx.x;
  for (int it = 0; it < iters; ++it) a = 1.0 / a + 0.5;
  out[threadIdx.x] = a;
}
__global__ void k_rsqrt_lat(double* out, int iters) {
  double a = 1.0 + threadIdx.x;
  for (int it = 0; it < iters; ++it) a = rsqrt(a) + 0.5;
  out[threadIdx.x] = a;
}

int main() {
  double* out; CK(cudaMalloc(&out, 148 * 1024 * 8 * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_dfma<<<148, warps * 32>>>(out, iters); else k_dmma<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const double fma_per_thread = mode == 0 ? 8.0 * iters : 8.0 * iters * 256 / 32;
      const double total = fma_per_thread * warps * 32 * 148;
      printf("%s warps/SM=%2d: %.3f ms, %.2f TFLOP/s chip, %.1f FMA/clk/SM @%d MHz nominal\n", mode ? "DMMA" : "DFMA", warps, best,
             2 * total / (best * 1e-3) / 1e12, total / 148 / (best * 1e-3) / (clk * 1e3), clk / 1000);
    }
  }
  float ms;
  const char* names[] = {"DFMA", "DMMA", "1/x+0.5", "rsqrt+0.5"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_dfma_lat<<<1, 32>>>(out, iters);
      else if (mode == 1) k_dmma_lat<<<1, 32>>>(out, iters);
      else if (mode == 2) k_div_lat<<<1, 32>>>(out, iters);
      else k_rsqrt_lat<<<1, 32>>>(out, iters);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("%s dependent chain: %.1f ns/op = %.1f cycles @%d MHz nominal\n", names[mode], ms * 1e6 / iters, ms * 1e-3 / iters * clk * 1e3, clk / 1000);
  }
  return 0;
}
