// Stand-alone harness for the reduced-system solver (pysfm_b200/csrc/ba_solve.cu): builds a random
// SPD system in the packed block layout, runs launch_solve, checks the residual on the host and
// prints a per-task timeline.   solve_bench [n_opt_cam] [reps]
#define BA_SOLVE_TRACE 1
#include "../../pysfm_b200/csrc/ba_solve.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

int main(int argc, char** argv) {
  const int nc = argc > 1 ? atoi(argv[1]) : 199;
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  const int n = 6 * nc, ld = (n + 63) / 64 * 64, T = ld / 64;
  const size_t nblk = (size_t)nc * (nc + 1) / 2, sys_len = nblk * 36 + n;
  // dense SPD: A = G G^T / n + 5 I with G random
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

  ba::Context c;
  c.n_opt_cam = nc; c.n_sys = n; c.ld = ld; c.sys_len = sys_len;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); c.num_sms = sms;
  const int ntasks = 1 + (T - 1) * (T + 2) / 2;
  CK(cudaMalloc(&c.sys, sys_len * 8)); CK(cudaMemcpy(c.sys, packed.data(), sys_len * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c.Adense, ((size_t)ld * ld + ld) * 8));
  CK(cudaMalloc(&c.LinvT, (size_t)T * 4096 * 8)); CK(cudaMemset(c.LinvT, 0, (size_t)T * 4096 * 8));
  CK(cudaMalloc(&c.Wpart, (size_t)T * (4096 + 64) * 8));
  CK(cudaMalloc(&c.solve_flags, ba::solve_flag_count(T) * 4)); CK(cudaMemset(c.solve_flags, 0, ba::solve_flag_count(T) * 4));
  CK(cudaMalloc(&c.solve_tickets, 8));
  CK(cudaMalloc(&c.solve_abort, 8)); CK(cudaMemset(c.solve_abort, 0, 8));
  CK(cudaMalloc(&c.solve_prof, 16 * 8)); CK(cudaMemset(c.solve_prof, 0, 16 * 8));
  c.solve_prof_on = true;
  CK(cudaMalloc(&c.dC, ld * 8));
  CK(cudaMalloc(&c.cam_mask, ld));
  CK(cudaMalloc(&c.scalars, sizeof(ba::Scalars))); CK(cudaMemset(c.scalars, 0, sizeof(ba::Scalars)));
  CK(cudaMalloc(&c.solve_trace, (size_t)(ntasks + T) * 8 * 8)); CK(cudaMemset(c.solve_trace, 0, (size_t)(ntasks + T) * 64));
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  float best = 1e30f, best_exp = 0, sum = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    CK(ba::launch_solve(c, false, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
    if (r > 0) sum += ms;
  }
  const float mean = reps > 1 ? sum / (reps - 1) : best;
  (void)e2; (void)best_exp;
  std::vector<double> x(ld);
  CK(cudaMemcpy(x.data(), c.dC, ld * 8, cudaMemcpyDeviceToHost));
  double rmax = 0, bmax = 0;
  for (int i = 0; i < n; ++i) {
    double s = -b[i];
    for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * x[j];
    rmax = std::max(rmax, fabs(s)); bmax = std::max(bmax, fabs(b[i]));
  }
  ba::Scalars sc; CK(cudaMemcpy(&sc, c.scalars, sizeof sc, cudaMemcpyDeviceToHost));
  printf("nc=%d n=%d T=%d tasks=%d: solve %.3f ms (best of %d; mean %.3f), residual %.3e (rel %.3e), status %g, %.2f GFLOP/s\n", nc, n, T,
         ntasks, best, reps, mean, rmax, rmax / bmax, sc.status, (double)n * n * n / 3 / (best * 1e-3) / 1e9);
  {
    unsigned long long prof[8][16];
    CK(cudaMemcpy(prof[0], c.solve_prof, 16 * 8, cudaMemcpyDeviceToHost));
    const char* names[] = {"panel tasks", "wait_k", "last step spin", "panel idle", "diag flag", "tile push", "contrib tail", "y flag",
                           "backward waits", "barrier", "kernel x CTAs", "chain tasks", "diag tasks"};
    printf("wait profile (us per launch per CTA):");
    for (int s2 = 0; s2 < 13; ++s2) printf(" %s %.1f;", names[s2], prof[0][s2] * 1e-3 / reps / sms);
    printf("\n");
  }
  // timeline of the last run
  std::vector<unsigned long long> tr((size_t)(ntasks + T) * 8);
  CK(cudaMemcpy(tr.data(), c.solve_trace, tr.size() * 8, cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull;
  for (int t = 0; t < ntasks + T; ++t) if (tr[(size_t)t * 8 + 2]) t0 = std::min(t0, tr[(size_t)t * 8 + 2]);
  const int show = argc > 3 ? atoi(argv[3]) : 60;
  printf("task  (i,j)  cta  grab_us  kloop_us  panel_done_us  sweep_done_us  publish_us   (chain tasks show (j,j))\n");
  for (int t = 0; t < ntasks + T; ++t) {
    const unsigned long long* r = &tr[(size_t)t * 8];
    const int i = (int)(r[0] >> 32), j = (int)(r[0] & 0xffffffffu);
    const bool interesting = t >= ntasks || i == j || i == j + 2 || i == T - 1;
    if (!interesting || (t > show && t < ntasks - 8)) continue;
    printf("%4d (%3d,%3d) %4llu %9.2f %9.2f %9.2f %9.2f %9.2f\n", t, i, j, r[1], (r[2] - t0) * 1e-3, (r[3] - t0) * 1e-3,
           r[6] ? (r[6] - t0) * 1e-3 : 0.0, r[4] ? (r[4] - t0) * 1e-3 : 0.0, (r[5] - t0) * 1e-3);
  }
  {
    unsigned long long dt[256];
    CK(cudaMemcpyFromSymbol(dt, ba::g_dbg_time, sizeof dt));
    printf("row-block flags of chain task C_3 (ticket 36 at T = 19) released at (us):");
    for (int q = 0; q < 8; ++q) printf(" %.2f", (dt[q] - t0) * 1e-3);
    for (int tk = 0; tk < 2; ++tk) {
      printf("\nconsumer ticket %d groups: ", tk + 52);
      for (int q = 0; q < 8; ++q) {
        const unsigned long long* d = dt + 16 + 32 * tk + 4 * q;
        if (!d[0]) continue;
        printf("\n   cb=%d ce=%llu: polled %.2f fetched %.2f panel %.2f update %.2f", q, dt[80 + 8 * tk + q], (d[0] - t0) * 1e-3, (d[1] - t0) * 1e-3,
               (d[2] - t0) * 1e-3, (d[3] - t0) * 1e-3);
      }
    }
    printf("\n");
  }
  long long clk[64];
  CK(cudaMemcpyFromSymbol(clk, ba::g_sweep_clk, sizeof clk));
  printf("sweep of task 0 (warp 7), clocks since k-loop end: ");
  for (int pb = 0; pb < 8; ++pb)
    printf("\n  pb=%d  gj_start %lld  gj_done %lld  Ld_out %lld  postY(w7) %lld  update_done(w7) %lld", pb, clk[pb * 4] - clk[33], clk[pb * 4 + 1] - clk[33],
           clk[pb * 4 + 2] - clk[33], clk[pb * 4 + 3] - clk[33], clk[36 + pb * 3] - clk[33]);
  printf("\n  end %lld\n", clk[32] - clk[33]);
  return 0;
}
