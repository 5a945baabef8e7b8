// Latency microbenchmarks behind the diagonal-tile sweep of ba_solve.cu (one CTA, clock64 deltas):
//   DMMA issue/latency from a single warp, FP64 op latencies, shuffle, MUFU.RCP64H, barrier cost.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_lat(long long* out, double* sink, int nwarps_active) {
  __shared__ double sm[2048];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-3;
  __syncthreads();
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  double a = 1.0 + lane * 1e-3, b = 1.0000001;
  long long t[16];
  const bool act = wid < nwarps_active;
  // 1: 8 independent DMMAs
  t[0] = clock64();
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  t[1] = clock64();
  // 2: 16 DMMAs = 8 independent chains of 2
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], b, a);
  }
  t[2] = clock64();
  // 3: 8 dependent DMMAs
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[0][0], c[0][1], a, b);
  }
  t[3] = clock64();
  // 4: 16 dependent DFMA
  double x = c[0][0];
#pragma unroll
  for (int i = 0; i < 16; ++i) x = fma(x, b, a);
  t[4] = clock64();
  // 5: 16 dependent shuffles (double = 2 SHFL)
#pragma unroll
  for (int i = 0; i < 16; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 4) & 31) + 1.0;
  t[5] = clock64();
  // 6: 16 dependent (rcp.approx + 3 fma)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    x = fma(r, e, r) + 1.5;
  }
  t[6] = clock64();
  // 7: 16 x (16 dependent DADD removed) -> 16 dependent smem load->add
#pragma unroll
  for (int i = 0; i < 16; ++i) x = sm[((int)x + lane + i) & 2047] + 1.0;
  t[7] = clock64();
  // 8: 16 __syncthreads
#pragma unroll
  for (int i = 0; i < 16; ++i) __syncthreads();
  t[8] = clock64();
  // 9: 16 dependent rsqrt
#pragma unroll
  for (int i = 0; i < 16; ++i) x = rsqrt(x) + 1.5;
  t[9] = clock64();
  // 10: 16 dependent 1/x
#pragma unroll
  for (int i = 0; i < 16; ++i) x = 1.0 / x + 1.5;
  t[10] = clock64();
  // 11: DSETP+FSEL chain
#pragma unroll
  for (int i = 0; i < 16; ++i) x = (x > 1.7) ? x * 0.9 : x + 0.3;
  t[11] = clock64();
  double s = x;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  sink[blockIdx.x * blockDim.x + tid] = s;
  if (lane == 0)
    for (int i = 0; i < 11; ++i) out[wid * 16 + i] = t[i + 1] - t[i];
}

int main() {
  long long* d; double* sink;
  CK(cudaMalloc(&d, 8 * 16 * 8)); CK(cudaMalloc(&sink, 256 * 8));
  const char* names[] = {"8 indep DMMA", "16 DMMA (8 chains of 2)", "8 dependent DMMA", "16 dep DFMA", "16 dep SHFL64+DADD",
                         "16 dep rcp.approx+3FMA+DADD", "16 dep LDS+DADD(+cvt)", "16 __syncthreads", "16 dep rsqrt+DADD", "16 dep 1/x+DADD", "16 dep DSETP/sel+DMUL"};
  for (int nact = 1; nact <= 8; nact *= 2) {
    for (int rep = 0; rep < 2; ++rep) k_lat<<<1, 256>>>(d, sink, nact);
    CK(cudaDeviceSynchronize());
    long long h[8 * 16];
    CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    printf("active DMMA warps = %d (8 warps resident)\n", nact);
    for (int i = 0; i < 11; ++i) printf("  %-32s warp0 %6lld clk   warp%d %6lld clk\n", names[i], h[i], nact - 1, h[(nact - 1) * 16 + i]);
  }
  return 0;
}
