// Stand-alone harness for the DISTRIBUTED reduced-system solve (pysfm_b200/csrc/ba_solve.cu,
// chol_dataflow_kernel<true>) with N "virtual ranks" on ONE GPU: N contexts in one process whose
// peer pointers point at each other's sections (peer memory is just memory here), N launches on
// N streams with 148 / N CTAs each so that all of them are co-resident.  Exercises the whole
// protocol -- task ownership, contribution sums, tile / Linv / y pushes, flags, start barrier,
// redundant backward substitution -- without a multi-GPU box; the real NVLink path is covered by
// tests/test_multi_gpu.py.
//     dist_solve_bench [n_opt_cam] [virtual ranks] [reps] [band] [strict]
#include "../../pysfm_b200/csrc/ba_solve.cu"

#include <cmath>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

int main(int argc, char** argv) {
  const int nc = argc > 1 ? atoi(argv[1]) : 199;
  const int world = argc > 2 ? (atoi(argv[2]) < 2 ? 2 : atoi(argv[2])) : 2;
  const int reps = argc > 3 ? atoi(argv[3]) : 5;
  const int band = argc > 4 ? atoi(argv[4]) : 2;
  const int strict = argc > 5 ? atoi(argv[5]) : 0;
  const int n = 6 * nc, ld = (n + 63) / 64 * 64, T = ld / 64;
  const size_t nblk = (size_t)nc * (nc + 1) / 2, sys_len = nblk * 36 + n;
  std::vector<double> G((size_t)n * 64), A((size_t)n * n), b(n), packed(sys_len);
  srand(1);
  for (auto& v : G) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0;
      for (int m = 0; m < 64; ++m) s += G[(size_t)i * 64 + m] * G[(size_t)j * 64 + m];
      s = s / 8 + (i == j ? 5.0 : 0.0);
      A[(size_t)i * n + j] = A[(size_t)j * n + i] = s;
    }
  for (int i = 0; i < n; ++i) b[i] = rand() / (double)RAND_MAX - 0.5;
  for (int a = 0; a < nc; ++a)
    for (int bb = a; bb < nc; ++bb) {
      const size_t blk = ba::packed_block(a, bb, nc);
      for (int rr = 0; rr < 6; ++rr)
        for (int cc = 0; cc < 6; ++cc) packed[blk * 36 + rr * 6 + cc] = A[(size_t)(6 * a + rr) * n + 6 * bb + cc];
    }
  for (int i = 0; i < n; ++i) packed[nblk * 36 + i] = b[i];

  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const ba::DistLayout dl = ba::dist_layout(ld);
  const size_t comm_len = (ba::comm_doubles(sys_len) + 31) & ~(size_t)31;
  std::vector<ba::Context> ctx(world);
  std::vector<cudaStream_t> streams(world);
  // rank r contributes w_r * (A, b), sum w_r = 1 (dyadic weights: the sum is exact in FP64)
  std::vector<double> w(world, 0.0);
  {
    double left = 1.0;
    for (int r = 0; r + 1 < world; ++r) { w[r] = left / 2; left -= w[r]; }
    w[world - 1] = left;
  }
  for (int r = 0; r < world; ++r) {
    ba::Context& c = ctx[r];
    c.n_opt_cam = nc; c.n_sys = n; c.ld = ld; c.sys_len = sys_len; c.num_sms = sms;
    c.solve_grid_cap = sms / world;
    c.comm_world = world; c.comm_rank = r; c.dist_off = comm_len; c.dist_min_tiles = 1; c.dist_band = band;
    c.spin_timeout_ms = 4000.0;
    c.strict_flags = strict;
    CK(cudaMalloc(&c.comm_buf, (comm_len + dl.total) * 8));
    CK(cudaMemset(c.comm_buf, 0, (comm_len + dl.total) * 8));
    c.sys = c.comm_buf;
    std::vector<double> mine(sys_len);
    for (size_t i = 0; i < sys_len; ++i) mine[i] = w[r] * packed[i];
    CK(cudaMemcpy(c.sys, mine.data(), sys_len * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c.solve_tickets, 8));
    CK(cudaMalloc(&c.dC, ld * 8));
    CK(cudaMalloc(&c.cam_mask, ld));
    CK(cudaMalloc(&c.scalars, sizeof(ba::Scalars))); CK(cudaMemset(c.scalars, 0, sizeof(ba::Scalars)));
    CK(cudaMalloc(&c.solve_prof, 16 * 8)); CK(cudaMemset(c.solve_prof, 0, 16 * 8));
    c.solve_prof_on = true;
    CK(cudaStreamCreate(&streams[r]));
  }
  for (int r = 0; r < world; ++r)
    for (int p = 0; p < world; ++p) ctx[r].comm_peer[p] = ctx[p].comm_buf;
  std::vector<cudaEvent_t> e0(world), e1(world);
  for (int r = 0; r < world; ++r) { cudaEventCreate(&e0[r]); cudaEventCreate(&e1[r]); }
  float best = 1e30f;
  for (int it = 0; it < reps; ++it) {
    CK(cudaDeviceSynchronize());
    for (int r = 0; r < world; ++r) {
      CK(cudaEventRecord(e0[r], streams[r]));
      CK(ba::launch_solve(ctx[r], false, streams[r]));
      CK(cudaEventRecord(e1[r], streams[r]));
    }
    CK(cudaDeviceSynchronize());
    float worst = 0;
    for (int r = 0; r < world; ++r) { float ms; cudaEventElapsedTime(&ms, e0[r], e1[r]); worst = std::max(worst, ms); }
    best = std::min(best, worst);
  }
  {
    unsigned long long prof[8][16];
    for (int r = 0; r < world; ++r) CK(cudaMemcpy(prof[r], ctx[r].solve_prof, 16 * 8, cudaMemcpyDeviceToHost));
    const char* names[] = {"panel tasks", "wait_k", "last step spin", "panel idle", "diag flag", "tile push", "contrib tail", "y flag",
                           "backward waits", "barrier", "kernel x CTAs", "chain tasks", "diag tasks"};
    printf("wait profile, ms summed over CTAs and %d launches (per launch per CTA in brackets, us):\n", reps);
    for (int r = 0; r < world; ++r) {
      printf("  rank %d:", r);
      for (int s2 = 0; s2 < 13; ++s2) printf(" %s %.2f (%.1f);", names[s2], prof[r][s2] * 1e-6, prof[r][s2] * 1e-3 / reps / (sms / world));
      printf("\n");
    }
  }
  int fails = 0;
  std::vector<double> x0(ld);
  for (int r = 0; r < world; ++r) {
    std::vector<double> x(ld);
    CK(cudaMemcpy(x.data(), ctx[r].dC, ld * 8, cudaMemcpyDeviceToHost));
    double rmax = 0, bmax = 0;
    for (int i = 0; i < n; ++i) {
      double s = -b[i];
      for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * x[j];
      rmax = std::max(rmax, fabs(s)); bmax = std::max(bmax, fabs(b[i]));
    }
    ba::Scalars sc; CK(cudaMemcpy(&sc, ctx[r].scalars, sizeof sc, cudaMemcpyDeviceToHost));
    bool same = true;
    if (r == 0) x0 = x; else same = memcmp(x0.data(), x.data(), (size_t)n * 8) == 0;
    printf("rank %d/%d: %d of %d tile tasks, residual %.3e (rel %.3e), status %g, bits %s rank 0\n", r, world, ctx[r].dist_ntasks,
           1 + (T - 1) * (T + 2) / 2, rmax, rmax / bmax, sc.status, same ? "==" : "!=");
    if (!(rmax / bmax < 1e-9) || sc.status != 0.0 || !same) ++fails;
  }
  printf("nc=%d n=%d T=%d world=%d band=%d grid=%d/rank: distributed solve %.3f ms (best of %d, max over ranks), %.2f GFLOP/s  %s\n", nc, n, T,
         world, band, sms / world, best, reps, (double)n * n * n / 3 / (best * 1e-3) / 1e9, fails ? "FAIL" : "OK");
  return fails ? 1 : 0;
}
