// Internal state behind the opaque ba_handle (see include/ba_b200.h).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "ba_math.cuh"

namespace ba {

constexpr int kSolveTile = 64;  // Cholesky tile; the reduced system is padded to a multiple
constexpr int kMaxPeers = 8;    // GPUs of one node (ba_comm.cu)

// Flags of the dataflow solver (u32, epoch valued): [T*T] tile (i,j) | [T] spare | [T] y_k |
// pad to a multiple of 4 | [8T] rows 8b.. of Linv_kk | [8T*T] columns 8b.. of L_ij.
__host__ __device__ inline size_t solve_rowflag_base(int T) { return ((size_t)T * T + 2 * (size_t)T + 3) & ~(size_t)3; }
inline size_t solve_flag_count(int T) { return solve_rowflag_base(T) + 8 * (size_t)T + 8 * (size_t)T * T; }

// doubles reserved per section of the peer-visible comm buffer (even, so that sections stay
// 16-byte aligned): [contrib | reduced | costs 2 banks x 2*kMaxPeers | flags 3*kMaxPeers u32]
__host__ __device__ inline size_t comm_pad(size_t sys_len) { return (sys_len + 2 + 31) & ~(size_t)31; }
inline size_t comm_doubles(size_t sys_len) { return 2 * comm_pad(sys_len) + 4 * kMaxPeers + (3 * kMaxPeers + 1) / 2 + 2; }

// Distributed reduced solve (ba_solve.cu, DIST): peer-visible section of every rank, placed behind
// the comm sections in the same IPC-exported allocation.  Offsets in doubles from the section base:
//   contrib [ld*ld + ld]   this rank's dense lower-triangular contribution + its rhs (expand_system_kernel)
//   L       [ld*ld + ld]   the factor, replicated: every tile is computed by ONE rank and pushed to all; rhs -> y
//   LinvT   [T][64*64]     transposed inverses of the diagonal tiles (pushed row block by row block)
//   wpart   [T][64*64+64]  partial diagonal tiles / forward-substitution sums (diagonal-update tasks -> chain tasks)
//   flags   [solve_flag_count(T)] u32, same layout as the single-GPU solver's, written by the tile owners
//   bar     [kMaxPeers] u32  start barrier (epoch valued)  | abort u32 | pad | status f64
struct DistLayout {
  size_t contrib, L, LinvT, wpart, flags, bar, abort, status, total;
};
__host__ __device__ inline DistLayout dist_layout(int ld) {
  const size_t T = (size_t)ld / kSolveTile;
  const size_t dense = (size_t)ld * ld + ld;
  const size_t nflags = solve_rowflag_base((int)T) + 8 * T + 8 * T * T;
  DistLayout d;
  d.contrib = 0;
  d.L = d.contrib + ((dense + 1) & ~(size_t)1);
  d.LinvT = d.L + ((dense + 1) & ~(size_t)1);
  d.wpart = d.LinvT + T * kSolveTile * kSolveTile;     // D_j -> C_j hand-over, local to the chain's rank
  d.flags = d.wpart + T * (kSolveTile * kSolveTile + kSolveTile);
  d.bar = d.flags + (((nflags + 3) / 2) & ~(size_t)1);
  d.abort = d.bar + kMaxPeers / 2;
  d.status = d.abort + 2;
  d.total = d.status + 2;
  return d;
}

// state of the packed reduced system of a handle
enum { kSysLocal = 0,     // this rank's contribution (fresh from the elimination kernel)
       kSysReduced = 1,   // all-reduced by ba_allreduce_system: the `reduced` section of the comm buffer
       kSysUploaded = 2   // overwritten by the caller (ba_upload_system): solved as is, locally
};

struct ParamSet {   // caller-owned device arrays
  double* cam_R = nullptr;  // [n_cam][9]
  double* cam_t = nullptr;  // [n_cam][3]
  double* pts = nullptr;    // [n_pt][3]
};

// Scalar record kept in device memory; mirrored on the host by ba_read_scalars.
struct Scalars {
  double cost;        // compute_cost(current)
  double cand_cost;   // compute_cost(candidate)
  double status;      // 0 ok, 1 non-positive pivot in the reduced solve
  double spare;
};

struct Context {
  int device = 0;
  int n_cam = 0, n_pt = 0, n_obs = 0, n_opt_cam = 0, n_opt_pt = 0;
  int n_sys = 0;      // 6 * n_opt_cam
  int ld = 0;         // padded leading dimension of the dense copy the solver factors
  size_t sys_len = 0; // doubles in the packed system buffer (blocks + rhs)
  int num_sms = 148;
  int max_track_len = 0;  // longest track, fetched lazily from pt_ptr (sizes shared memory)

  Intrinsics intr{};
  ModelParams model{};

  // bound, caller-owned
  const int* pt_ptr = nullptr;
  const int* obs_cam = nullptr;
  const double* obs_uv = nullptr;
  const int* cam_slot = nullptr;
  const int* pt_slot = nullptr;
  ParamSet state, cand;
  double* sys = nullptr;  // packed reduced system: upper 6x6 blocks, row by row, then rhs

  // library-owned workspace
  double* Vinv = nullptr;   // [n_pt][9]   HPP_invs
  double* bP = nullptr;     // [n_pt][3]   bPs
  double* V = nullptr;      // [n_pt][9]   HPPs (undamped)       (BA_WANT_BLOCKS)
  double* U = nullptr;      // [n_cam][36] HCCs (undamped)       (BA_WANT_BLOCKS)
  double* bC = nullptr;     // [n_cam][6]  bCs                   (BA_WANT_BLOCKS)
  double* W = nullptr;      // [n_obs][18] HCPs, lazily allocated (BA_WANT_BLOCKS)
  double* io_out = nullptr; // [4 + ld + 3 n_pt]  {scalars | dC | dP} in one block (one D2H per host-driven trial)
  double* dC = nullptr;     // [ld]        reduced solution (inside io_out)
  double* Adense = nullptr; // [ld*ld + ld] dense lower-triangular copy + rhs, factored in place
  double* LinvT = nullptr;  // [ld/64][64*64] transposed inverses of the diagonal Cholesky tiles
  double* Wpart = nullptr;  // [ld/64][64*64 + 64] partial diagonal tiles + forward-substitution sums (D_j -> C_j)
  unsigned int* solve_flags = nullptr;    // [solve_flag_count(T)] tile / y_k / Linv row-block / L column-block ready flags (epoch valued)
  unsigned int* solve_tickets = nullptr;  // [2] task tickets of the dataflow solver
  unsigned long long* solve_trace = nullptr;  // debug timeline (BA_SOLVE_TRACE builds only)
  unsigned int solve_epoch = 0;
  bool solve_attr_set = false;
  bool elim_attr_set[8] = {false, false, false, false, false, false, false, false};
  bool elim_group_attr_set[2] = {false, false};
  double* rec_scratch = nullptr;   // [n_obs][56] records of long tracks (elimination kernel, REC_GLOBAL)
  double* diag_rep = nullptr;      // [16][n_opt_cam][36] spread copies of the diagonal blocks (elimination kernel; zero between launches)
  bool backsub_attr_set[3] = {false, false, false};
  bool backsub_tile_attr_set[3] = {false, false, false};
  double* dP = nullptr;     // [n_pt][3]   point update (rows of non-updated tracks are zero; inside io_out)
  double* obs_r = nullptr;  // [n_obs][2]  lazily allocated (ba_eval_observations)
  double* obs_Jc = nullptr; // [n_obs][12]
  double* obs_Jp = nullptr; // [n_obs][6]
  double* delta_cam = nullptr;  // [n_opt_cam][6] staging for ba_retract
  double* delta_pt = nullptr;   // [n_pt][3]
  unsigned char* cam_mask = nullptr;  // [ld] 1 = free parameter
  double* partials = nullptr;   // per-CTA cost partial sums
  unsigned int* counters = nullptr;  // last-CTA-done tickets of the grid-wide cost sums
  Scalars* scalars = nullptr;
  int partials_cap = 0;

  // peer-memory collectives (points sharded over the GPUs of one node)
  int comm_world = 0, comm_rank = 0;
  double* comm_buf = nullptr;               // own IPC-exported buffer (== comm_peer[comm_rank])
  double* comm_peer[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned int* comm_done = nullptr;        // last-CTA-done counter of the all-reduce kernel
  unsigned int comm_epoch = 0;
  bool fuse_cost_reduction = true;          // BA_OPT_FUSE_COST_REDUCTION: backsub's last CTA also reduces the costs over the ranks
  bool costs_reduced = false;               // the scalars already hold the sums over all ranks
  int sys_state = kSysLocal;                // what the bound system holds (which copy the solver reads)
  // distributed reduced solve (tiles owned by ranks, operands exchanged over peer memory)
  size_t dist_off = 0;                      // doubles from comm_buf to the rank's DistLayout section (0 = none)
  int* dist_tasks = nullptr;                // this rank's tile tasks in global ticket order ((i << 16) | j; chain C_j = (j, j))
  int dist_ntasks = 0;
  unsigned int dist_epoch = 0;
  int dist_min_tiles = 32;                  // BA_OPT_DIST_SOLVE_MIN_TILES: distributed solve when ld/64 >= this (0 = never)
  int dist_band = 6;                        // BA_OPT_DIST_BAND: tiles with i - j <= band stay on rank 0 with the chain tasks
  bool dist_attr_set = false;
  // robustness knobs of the spin-waits (solver flags, peer barriers)
  double spin_timeout_ms = 10000.0;         // BA_OPT_SPIN_TIMEOUT_MS
  int strict_flags = 0;                     // BA_OPT_STRICT_FLAGS: release/acquire flag publication in the solver
  int solve_grid_cap = 0;                   // BA_OPT_SOLVE_GRID_CAP: at most this many solver CTAs (0 = one per SM)
  int split_min_tiles = 32;                 // diagonal-update tasks take over part of the chain tasks from this many tile rows on
  // blocked solve with the tcgen05 trailing update (ba_solve_tc.cuh; single-GPU handles, large systems)
  int tc_min_tiles = 80;                    // BA_OPT_TC_MIN_TILES: tile rows from which ba_solve takes this path (0 = never)
  int tc_slices_n = 6;                      // BA_OPT_TC_SLICES: INT8 slices per FP64 operand (4 .. 7)
  int tc_window = 8;                        // BA_OPT_TC_WINDOW: tile columns per panel (even, <= 16): K = 64 * window
  int tc_over_dist_max_world = 2;           // BA_OPT_TC_OVER_DIST_MAX_WORLD: sharded handles of at most this many ranks all-reduce and run the
                                            // blocked solve on every rank instead of the distributed solve (same value on every rank)
  int tc_bk = 64;                           // BA_OPT_TC_BK: bytes of K per pipeline stage = swizzle span (64 or 128)
  signed char* tc_slices = nullptr;         // [slices][ld_pad][64 * window] INT8 digits of the current panel, K-major
  double* tc_scale = nullptr;               // [ld_pad] power-of-two row scales of the current panel
  double* tc_save = nullptr;                // [64] right-hand side block the window launch's last chain task overwrites
  int* tc_dbg = nullptr;                    // tests: raw level sums of the next trailing update ([slices][ld_pad][tc_dbg_ld])
  int tc_dbg_ld = 0;
  unsigned long long* tc_dbg_time = nullptr;  // microbenchmarks: role timers of the trailing-update kernel (SyrkArgs::dbg_time)
  int tc_dbg_skip = 0;                      // microbenchmarks: roles of the trailing-update kernel switched off (SyrkArgs::dbg_skip)
  bool tc_attr_set[16] = {false, false, false, false, false, false, false, false, false, false, false, false, false, false, false, false};
  // CUDA-event breakdown of the blocked solve (BA_OPT_SOLVER_PROFILE; ba_tc_solve_profile): one event behind every
  // launch, tagged with what the time since the previous event was spent on
  std::vector<cudaEvent_t> tc_ev;
  std::vector<int> tc_ev_cat;
  int tc_ev_used = 0;
  double tc_prof_ms[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // expand | panels | slices | trailing updates | backward substitution
  long long tc_prof_solves = 0;
  int tc_cfg[4] = {0, 0, 0, 0};             // (slices, window, bk, ld) the buffers and tensor maps were built for
  alignas(64) unsigned char tc_map_a[128];  // CUtensorMap: box 128 rows x bk bytes
  alignas(64) unsigned char tc_map_b[128];  // CUtensorMap: box  64 rows x bk bytes
  unsigned int* solve_abort = nullptr;      // [1] set by a spin-wait that ran past the deadline
  unsigned long long* solve_prof = nullptr; // [16] wait-time profile of the solver (ba_solver_profile)
  bool solve_prof_on = false;               // BA_OPT_SOLVER_PROFILE

  long long launches = 0;
  std::string last_error;
};

// ---- kernel launchers (ba_kernels.cu / ba_solve.cu) ----
cudaError_t launch_linearize_eliminate(Context& c, double damping, double rcond, int flags,
                                       cudaStream_t st);
cudaError_t launch_backsub_retract_cost(Context& c, cudaStream_t st);
cudaError_t launch_cost(Context& c, cudaStream_t st);
cudaError_t launch_eval_observations(Context& c, cudaStream_t st);
cudaError_t launch_retract(Context& c, bool have_cam, bool have_pt, cudaStream_t st);
cudaError_t launch_triangulate(Context& c, cudaStream_t st);
cudaError_t launch_peer_allreduce_system(Context& c, cudaStream_t st);
cudaError_t launch_peer_allreduce_costs(Context& c, cudaStream_t st);
cudaError_t launch_solve(Context& c, bool have_mask, cudaStream_t st);
bool tc_solve_selected(const Context& c);
void tc_fold_profile(Context& c);     // folds the pending events of the last profiled blocked solve into tc_prof_ms (synchronises on them)
cudaError_t tc_prepare(Context& c);   // slice buffers + tensor maps for (c.ld, tc_slices_n, tc_window, tc_bk)
cudaError_t launch_tc_trailing_update(Context& c, double* A, double* rhs, int c0, const double* saved_rhs, cudaStream_t st);
bool dist_solve_selected(const Context& c);

}  // namespace ba
