// Stand-alone harness for the blocked tcgen05 solve (pysfm_b200/csrc/ba_solve_tc.cuh).
//
//   tc_bench syrk  [ld] [window] [slices] [bk]     one trailing update on a random panel: INT8 digits, scales and the
//                                                  right-hand side against a host restatement (digits bit for bit), the raw
//                                                  INT32 level sums out of TMEM bit for bit, the updated matrix bit for bit
//   tc_bench perf  [ld] [window] [slices] [bk] [reps] [skip]   time of one trailing update at a given size (no verification)
//   tc_bench solve [n_opt_cam] [window] [slices] [bk] [reps] [min_tiles]
//                                                  random SPD system: blocked tcgen05 solve against the DMMA dataflow
//                                                  solve (same library code path as ba_solve): solutions, residuals, times
#include "../../pysfm_b200/csrc/ba_solve.cu"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e__), __LINE__); exit(1);} } while (0)

static double urand() { return rand() / (double)RAND_MAX - 0.5; }

static void alloc_context(ba::Context& c, int nc) {
  const int n = 6 * nc, ld = (n + 63) / 64 * 64, T = ld / 64;
  c.n_opt_cam = nc; c.n_sys = n; c.ld = ld; c.sys_len = (size_t)nc * (nc + 1) / 2 * 36 + n;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); c.num_sms = sms;
  CK(cudaMalloc(&c.sys, c.sys_len * 8));
  CK(cudaMalloc(&c.Adense, ((size_t)ld * ld + ld) * 8));
  CK(cudaMalloc(&c.LinvT, (size_t)T * 4096 * 8)); CK(cudaMemset(c.LinvT, 0, (size_t)T * 4096 * 8));
  CK(cudaMalloc(&c.Wpart, (size_t)T * (4096 + 64) * 8));
  CK(cudaMalloc(&c.solve_flags, ba::solve_flag_count(T) * 4)); CK(cudaMemset(c.solve_flags, 0, ba::solve_flag_count(T) * 4));
  CK(cudaMalloc(&c.solve_tickets, 8));
  CK(cudaMalloc(&c.solve_abort, 8)); CK(cudaMemset(c.solve_abort, 0, 8));
  CK(cudaMalloc(&c.solve_prof, 16 * 8)); CK(cudaMemset(c.solve_prof, 0, 16 * 8));
  CK(cudaMalloc(&c.dC, ld * 8));
  CK(cudaMalloc(&c.cam_mask, ld));
  CK(cudaMalloc(&c.scalars, sizeof(ba::Scalars))); CK(cudaMemset(c.scalars, 0, sizeof(ba::Scalars)));
  c.spin_timeout_ms = 5000.0;
}

// ---------------------------------------------------------------------------------------------
static int run_syrk(int ld, int w, int S, int bk) {
  const int K = 64 * w, c0 = 0, c1 = K, ld_pad = (ld + 127) / 128 * 128;
  printf("== syrk: ld %d (pad %d), window %d (K %d), slices %d, bk %d\n", ld, ld_pad, w, K, S, bk);
  ba::Context c;
  alloc_context(c, ld / 6);   // (only the solver workspace matters here)
  c.ld = ld;
  cudaFree(c.Adense); CK(cudaMalloc(&c.Adense, ((size_t)ld * ld + ld) * 8));
  c.tc_window = w; c.tc_slices_n = S; c.tc_bk = bk; c.tc_min_tiles = 1;
  CK(ba::tc_prepare(c));
  // random lower matrix; panel rows get different magnitudes, one row is zero, one holds a single huge entry
  std::vector<double> A((size_t)ld * ld + ld);
  srand(7);
  for (int q = 0; q < ld; ++q)
    for (int p = 0; p < ld; ++p) {
      double v = urand();
      if (q < c1 && p >= c1) {
        v *= std::ldexp(1.0, (p * 7) % 23 - 11);
        if (p == c1 + 5) v = 0.0;
        if (p == c1 + 9 && q == 3) v = 12345.678;
      }
      A[(size_t)q * ld + p] = v;
    }
  for (int i = 0; i < ld; ++i) A[(size_t)ld * ld + i] = urand();
  std::vector<double> saved(64);
  for (int i = 0; i < 64; ++i) saved[i] = urand();
  CK(cudaMemcpy(c.Adense, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tc_save, saved.data(), 64 * 8, cudaMemcpyHostToDevice));
  const size_t dbg_n = (size_t)S * ld_pad * ld;
  CK(cudaMalloc(&c.tc_dbg, dbg_n * 4)); CK(cudaMemset(c.tc_dbg, 0xff, dbg_n * 4));
  c.tc_dbg_ld = ld;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaEventRecord(e0));
  CK(ba::launch_tc_trailing_update(c, c.Adense, c.Adense + (size_t)ld * ld, c0, c.tc_save, 0));
  CK(cudaEventRecord(e1));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("FAIL: kernel error %s\n", cudaGetErrorString(se)); return 1; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  ba::Scalars sc; CK(cudaMemcpy(&sc, c.scalars, sizeof sc, cudaMemcpyDeviceToHost));
  unsigned int ab; CK(cudaMemcpy(&ab, c.solve_abort, 4, cudaMemcpyDeviceToHost));
  printf("   slice + syrk %.3f ms, status %g, abort %u\n", ms, sc.status, ab);
  int bad = (sc.status != 0.0 || ab != 0);

  // ---- host restatement of the slices ----
  std::vector<int8_t> dig((size_t)S * ld_pad * K), hdig((size_t)S * ld_pad * K, 0);
  std::vector<double> scale(ld_pad), hscale(ld_pad, 0.0), rhs(ld), hrhs(ld);
  CK(cudaMemcpy(dig.data(), c.tc_slices, dig.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(scale.data(), c.tc_scale, ld_pad * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rhs.data(), c.Adense + (size_t)ld * ld, ld * 8, cudaMemcpyDeviceToHost));
  long long dig_bad = 0, scale_bad = 0; double rhs_err = 0;
  for (int p = c1; p < ld; ++p) {
    double mx = 0, dot = 0, mag = 0;
    for (int k = 0; k < K; ++k) {
      const double v = A[(size_t)(c0 + k) * ld + p], y = A[(size_t)ld * ld + c0 + k];
      mx = std::fmax(mx, std::fabs(v)); dot += v * y; mag += std::fabs(v * y);
    }
    const int e = mx > 0 ? std::ilogb(mx) + 1 : 0;
    hscale[p] = std::ldexp(1.0, e - 6);
    const double b = (p < c1 + 64) ? saved[p - c1] : A[(size_t)ld * ld + p];
    hrhs[p] = b - dot;
    rhs_err = std::fmax(rhs_err, std::fabs(hrhs[p] - rhs[p]) / (mag + std::fabs(b) + 1e-300));
    if (hscale[p] != scale[p]) ++scale_bad;
    for (int k = 0; k < K; ++k) {
      double t = A[(size_t)(c0 + k) * ld + p] * std::ldexp(1.0, 6 - e);
      for (int s = 0; s < S; ++s) {
        const double d = std::nearbyint(t);
        hdig[((size_t)s * ld_pad + p) * K + k] = (int8_t)(int)d;
        if (dig[((size_t)s * ld_pad + p) * K + k] != (int8_t)(int)d) ++dig_bad;
        t = (t - d) * 128.0;
      }
    }
  }
  printf("   slices: %lld digit mismatches, %lld scale mismatches, rhs max rel err %.2e\n", dig_bad, scale_bad, rhs_err);
  bad |= (dig_bad != 0 || scale_bad != 0 || rhs_err > 1e-12);

  // ---- level sums out of TMEM, bit for bit (from the DEVICE's digits: isolates the tensor path) ----
  std::vector<int> dacc(dbg_n);
  CK(cudaMemcpy(dacc.data(), c.tc_dbg, dbg_n * 4, cudaMemcpyDeviceToHost));
  std::vector<double> Aout((size_t)ld * ld);
  CK(cudaMemcpy(Aout.data(), c.Adense, (size_t)ld * ld * 8, cudaMemcpyDeviceToHost));
  long long acc_bad = 0, a_bad = 0, shown = 0, untouched_bad = 0;
  double a_err = 0;
  std::vector<long long> lvl_bad(S, 0);
  std::vector<int> hacc(S);
  for (int i = c1; i < ld; ++i)
    for (int j = c1; j <= i; ++j) {
      for (int l = 0; l < S; ++l) {
        int s = 0;
        for (int p = 0; p <= l; ++p) {
          const int8_t* a = &dig[((size_t)p * ld_pad + i) * K];
          const int8_t* b = &dig[((size_t)(l - p) * ld_pad + j) * K];
          int t = 0;
          for (int k = 0; k < K; ++k) t += (int)a[k] * (int)b[k];
          s += t;
        }
        hacc[l] = s;
        const int d = dacc[((size_t)l * ld_pad + i) * ld + j];
        if (d != s) {
          ++acc_bad; ++lvl_bad[l];
          if (shown < 12) { printf("   level %d (%d,%d): device %d host %d\n", l, i, j, d, s); ++shown; }
        }
      }
      double v = 0;
      for (int l = S - 1; l >= 0; --l) v = v * 0.0078125 + (double)hacc[l];
      const double want = A[(size_t)j * ld + i] - (v * scale[i]) * scale[j];
      const double got = Aout[(size_t)j * ld + i];
      if (want != got) { ++a_bad; a_err = std::fmax(a_err, std::fabs(want - got) / (std::fabs(want) + 1e-300)); }
    }
  // nothing outside the trailing lower triangle may change
  for (int q = 0; q < ld; ++q)
    for (int p = 0; p < ld; ++p)
      if (!(p >= c1 && q >= c1 && p >= q) && Aout[(size_t)q * ld + p] != A[(size_t)q * ld + p]) ++untouched_bad;
  printf("   level sums: %lld mismatches (per level:", acc_bad);
  for (int l = 0; l < S; ++l) printf(" %lld", lvl_bad[l]);
  printf("); updated A: %lld of %lld entries differ (max rel %.2e); %lld entries outside the trailing triangle changed\n", a_bad,
         (long long)(ld - c1) * (ld - c1 + 1) / 2, a_err, untouched_bad);
  // accuracy of the whole update against plain FP64
  double upd_err = 0, upd_max = 0;
  for (int i = c1; i < ld; i += 7)
    for (int j = c1; j <= i; j += 5) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += A[(size_t)(c0 + k) * ld + i] * A[(size_t)(c0 + k) * ld + j];
      const double got = A[(size_t)j * ld + i] - Aout[(size_t)j * ld + i];
      upd_err = std::fmax(upd_err, std::fabs(got - s) / (scale[i] * scale[j] * 4096.0 * K));
      upd_max = std::fmax(upd_max, std::fabs(s));
    }
  printf("   L L^T against FP64: max error %.2e of (row scale x column scale x K)   [2^-%d = %.1e]\n", upd_err, 7 * S, std::ldexp(1.0, -7 * S));
  bad |= (acc_bad != 0 || a_bad != 0 || untouched_bad != 0);
  printf(bad ? "FAIL syrk\n" : "PASS syrk\n");
  return bad;
}

// ---------------------------------------------------------------------------------------------
static int run_solve(int nc, int w, int S, int bk, int reps, int min_tiles) {
  const int n = 6 * nc, ld = (n + 63) / 64 * 64, T = ld / 64;
  printf("== solve: nc %d n %d T %d, window %d, slices %d, bk %d\n", nc, n, T, w, S, bk);
  const size_t nblk = (size_t)nc * (nc + 1) / 2, sys_len = nblk * 36 + n;
  std::vector<double> G((size_t)n * 64), b(n), packed(sys_len);
  srand(1);
  for (auto& v : G) v = urand();
  for (int i = 0; i < n; ++i) b[i] = urand();
  // A = G G^T / 8 + 5 I, written straight into the packed upper blocks
  auto Aij = [&](int i, int j) {
    double s = 0;
    for (int m = 0; m < 64; ++m) s += G[(size_t)i * 64 + m] * G[(size_t)j * 64 + m];
    return s / 8 + (i == j ? 5.0 : 0.0);
  };
  for (int a = 0; a < nc; ++a)
    for (int bb = a; bb < nc; ++bb) {
      const size_t blk = ba::packed_block(a, bb, nc);
      for (int rr = 0; rr < 6; ++rr)
        for (int cc = 0; cc < 6; ++cc) packed[blk * 36 + rr * 6 + cc] = Aij(6 * a + rr, 6 * bb + cc);
    }
  for (int i = 0; i < n; ++i) packed[nblk * 36 + i] = b[i];
  ba::Context c;
  alloc_context(c, nc);
  CK(cudaMemcpy(c.sys, packed.data(), sys_len * 8, cudaMemcpyHostToDevice));
  c.tc_window = w; c.tc_slices_n = S; c.tc_bk = bk;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<double> x[2];
  float best[2] = {1e30f, 1e30f};
  int bad = 0;
  for (int mode = 0; mode < 2; ++mode) {   // 0: DMMA dataflow solve, 1: blocked tcgen05 solve
    c.tc_min_tiles = mode ? min_tiles : 0;   // (also the size below which the tail goes to the dataflow kernel)
    if (mode && !ba::tc_solve_selected(c)) { printf("   (system too small for the blocked path)\n"); return 0; }
    for (int r = 0; r < reps; ++r) {
      CK(cudaEventRecord(e0));
      CK(ba::launch_solve(c, false, 0));
      CK(cudaEventRecord(e1));
      cudaError_t se = cudaEventSynchronize(e1);
      if (se != cudaSuccess) { printf("FAIL: kernel error %s\n", cudaGetErrorString(se)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best[mode]) best[mode] = ms;
    }
    x[mode].resize(ld);
    CK(cudaMemcpy(x[mode].data(), c.dC, ld * 8, cudaMemcpyDeviceToHost));
    ba::Scalars sc; CK(cudaMemcpy(&sc, c.scalars, sizeof sc, cudaMemcpyDeviceToHost));
    // residual from the packed blocks (symmetric)
    std::vector<double> rres(n);
    for (int i = 0; i < n; ++i) rres[i] = -b[i];
    for (int a = 0; a < nc; ++a)
      for (int bb = a; bb < nc; ++bb) {
        const double* blkp = &packed[ba::packed_block(a, bb, nc) * 36];
        for (int rr = 0; rr < 6; ++rr)
          for (int cc = 0; cc < 6; ++cc) {
            rres[6 * a + rr] += blkp[rr * 6 + cc] * x[mode][6 * bb + cc];
            if (bb != a) rres[6 * bb + cc] += blkp[rr * 6 + cc] * x[mode][6 * a + rr];
          }
      }
    double rmax = 0, bmax = 0;
    for (int i = 0; i < n; ++i) { rmax = std::fmax(rmax, std::fabs(rres[i])); bmax = std::fmax(bmax, std::fabs(b[i])); }
    const double flops = (double)n * n * n / 3;
    printf("   %s: %.3f ms best of %d (%.1f TFLOP/s FP64-equivalent), residual rel %.3e, status %g, launches %lld\n",
           mode ? "tcgen05 blocked" : "DMMA dataflow  ", best[mode], reps, flops / (best[mode] * 1e-3) / 1e12, rmax / bmax, sc.status, c.launches);
    if (!(rmax / bmax < 1e-8) || sc.status != 0.0) bad = 1;
  }
  double dmax = 0, xmax = 0;
  for (int i = 0; i < n; ++i) { dmax = std::fmax(dmax, std::fabs(x[0][i] - x[1][i])); xmax = std::fmax(xmax, std::fabs(x[0][i])); }
  printf("   solutions differ by %.3e relative; speed-up %.2fx\n", dmax / xmax, best[0] / best[1]);
  if (!(dmax / xmax < 1e-8)) bad = 1;
  printf(bad ? "FAIL solve\n" : "PASS solve\n");
  return bad;
}

// ---------------------------------------------------------------------------------------------
// time of one trailing update (slices + products) at a given size, no verification
static int run_perf(int ld, int w, int S, int bk, int reps, int skip) {
  const int K = 64 * w;
  ba::Context c;
  alloc_context(c, 64);
  c.ld = ld;
  cudaFree(c.Adense); CK(cudaMalloc(&c.Adense, ((size_t)ld * ld + ld) * 8));
  c.tc_window = w; c.tc_slices_n = S; c.tc_bk = bk; c.tc_min_tiles = 1; c.tc_dbg_skip = skip;
  CK(ba::tc_prepare(c));
  {
    std::vector<double> col((size_t)ld);
    srand(3);
    for (int q = 0; q <= K; ++q) {   // the panel columns (and the right-hand side behind the matrix)
      for (int p = 0; p < ld; ++p) col[p] = urand();
      CK(cudaMemcpy(q < K ? c.Adense + (size_t)q * ld : c.Adense + (size_t)ld * ld, col.data(), (size_t)ld * 8, cudaMemcpyHostToDevice));
    }
  }
  CK(cudaMalloc(&c.tc_dbg_time, 8 * 8)); CK(cudaMemset(c.tc_dbg_time, 0, 64));
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  const int n_nb = (ld - K) / 64, ntiles = ba::tc::count_tiles(n_nb);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    CK(ba::launch_tc_trailing_update(c, c.Adense, c.Adense + (size_t)ld * ld, 0, nullptr, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = (double)(ld - K) * (ld - K) * K;   // lower triangle: m^2/2 * K * 2
  ba::Scalars sc; CK(cudaMemcpy(&sc, c.scalars, sizeof sc, cudaMemcpyDeviceToHost));
  printf("perf: skip %d ld %d K %d S %d bk %d: slices + update %.3f ms (best of %d), %d tiles of 128x64 -> %.2f us per tile per SM, %.1f TFLOP/s FP64-equivalent, status %g\n",
         skip, ld, K, S, bk, best, reps, ntiles, best * 1e3 / ((ntiles + c.num_sms - 1) / c.num_sms), flops / (best * 1e-3) / 1e12, sc.status);
  {
    unsigned long long tm[8];
    CK(cudaMemcpy(tm, c.tc_dbg_time, 64, cudaMemcpyDeviceToHost));
    const double per = 1.0 / ((double)ntiles * reps);   // clocks per tile
    printf("      clocks per tile: producer waits for a stage %.0f | MMA issuer waits for the epilogue %.0f, for operands %.0f | epilogue waits for the products %.0f, drains + combines %.0f, updates A %.0f\n",
           tm[0] * per, tm[1] * per, tm[2] * per, tm[3] * per, tm[4] * per, tm[5] * per);
  }
  return 0;
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "syrk";
  auto arg = [&](int i, int d) { return argc > i ? atoi(argv[i]) : d; };
  if (!strcmp(mode, "syrk")) return run_syrk(arg(2, 640), arg(3, 2), arg(4, 6), arg(5, 64));
  if (!strcmp(mode, "solve")) return run_solve(arg(2, 199), arg(3, 8), arg(4, 6), arg(5, 64), arg(6, 3), arg(7, 1));
  if (!strcmp(mode, "perf")) return run_perf(arg(2, 12032), arg(3, 8), arg(4, 6), arg(5, 64), arg(6, 3), arg(7, 0));
  printf("usage: tc_bench syrk|solve|perf ...\n");
  return 2;
}
