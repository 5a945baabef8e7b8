/*
 * ba_b200.h -- C ABI of the B200-native bundle-adjustment inner loop (libba_b200.so).
 *
 * The reference (alexflint/pysfm) has no FFI: its boundary for this path is the Python
 * class bundle_adjuster.BundleAdjuster.  Each entry point below replaces one stage of that
 * class and cites the reference code it stands in for (paths relative to the pysfm tree).
 * The Python drop-in (pysfm_b200/bundle_adjuster.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a pysfm maintainer would add.
 *
 * Conventions
 *   - every entry returns an int status (BA_OK, ...); nothing throws across the boundary;
 *   - plain pointers and sizes only, no torch/C++ types;
 *   - "dev" pointers are CUDA device pointers owned by the CALLER (the Python side holds them
 *     in torch tensors); the library owns only the opaque handle and its internal workspace;
 *   - all kernels are enqueued on the caller's stream (cudaStream_t passed as void*), nothing
 *     synchronises except ba_read_scalars / ba_get_* / ba_sync;
 *   - one handle per GPU (rank); handles are not thread-safe; there is no global state;
 *   - all arithmetic is IEEE FP64, indices are int32.
 *
 * Data layout in HBM (SoA: one array per attribute, never an array of camera/track objects)
 *   cam_R   [n_cam][9]   row-major rotation            cam_t [n_cam][3]
 *   pts     [n_pt][3]
 *   pt_ptr  [n_pt+1]     CSR offsets into the observation arrays (observations are
 *                        point-major; inside one point they are sorted by camera slot)
 *   obs_cam [n_obs]      camera position (0..n_cam-1) of each observation
 *   obs_uv  [n_obs][2]   measured pixel
 *   cam_slot[n_cam]      position of the camera in the reduced system (0..n_opt_cam-1),
 *                        or -1 when the camera is held fixed
 *   pt_slot [n_pt]       position of the point in the structure update, or -1 when the
 *                        track is eliminated but not updated (track_mask semantics)
 *   sys     [ba_system_size(n_opt_cam)]  reduced camera system, PACKED: the upper block
 *                        triangle of S as 6x6 row-major blocks (a, b), a <= b, block rows back
 *                        to back (block index a*nc' - a(a-1)/2 + (b-a), 36 doubles each,
 *                        diagonal blocks stored in full), followed by the right-hand side
 *                        b [6 nc'].  This is the one buffer that is all-reduced across ranks
 *                        when points are sharded (half the bytes of a dense matrix).  The
 *                        solver expands it into a library-owned dense lower-triangular copy
 *                        with leading dimension ba_system_ld(6 nc') before factoring.
 */
#ifndef BA_B200_H_
#define BA_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ba_context* ba_handle;

enum {
  BA_OK = 0,
  BA_ERR_ILLCONDITIONED = 1, /* reduced system not positive definite -> Python raises
                                NormalEquationsIllconditioned (bundle_adjuster.py:27,302-305) */
  BA_ERR_CUDA = 2,
  BA_ERR_NCCL = 3,           /* reserved: collectives are issued by the host side */
  BA_ERR_BAD_ARGUMENT = 4,
  BA_ERR_NOT_BOUND = 5,
  BA_ERR_NONFINITE = 6,      /* solve_status of ba_read_scalars: a cost came out NaN/Inf (the
                                reference's drivers run under numpy.seterr(all='raise'),
                                window_slam.py:70) */
  BA_ERR_TIMEOUT = 7         /* solve_status: a spin-wait of the solver or of a peer barrier ran
                                past BA_OPT_SPIN_TIMEOUT_MS (a peer died or a kernel faulted) */
};

/* options for ba_set_option */
enum {
  BA_OPT_SPIN_TIMEOUT_MS = 0,      /* waiting budget of one solver / collective launch (default 10000) */
  BA_OPT_STRICT_FLAGS = 1,         /* 1: the single-GPU solver publishes its column-block flags with
                                      release/acquire instead of relaxed stores (default 0) */
  BA_OPT_DIST_SOLVE_MIN_TILES = 2, /* sharded handles: ba_solve runs the distributed solve when the
                                      reduced system has at least this many 64-wide tile rows
                                      (default 32, i.e. >= 2048 camera parameters; 0 = never) */
  BA_OPT_DIST_BAND = 3,            /* distributed solve: tiles with i - j <= band stay on rank 0 */
  BA_OPT_SOLVE_GRID_CAP = 4,       /* at most this many solver CTAs (0 = one per SM) */
  BA_OPT_SOLVER_PROFILE = 5,       /* 1: the solver accumulates its wait-time profile (ba_solver_profile) */
  BA_OPT_FUSE_COST_REDUCTION = 6,  /* sharded handles over peer memory (default 1): ba_backsub_retract_cost reduces
                                      {cost, candidate cost} over the ranks in its own epilogue and the following
                                      ba_allreduce_costs is a no-op; set the same value on every rank */
  BA_OPT_TC_MIN_TILES = 7,         /* ba_solve takes the blocked factorisation whose trailing updates run as INT8
                                      products on tcgen05 (Ozaki slices of the FP64 panel, exact INT32 sums in
                                      tensor memory, FP64 recombination) when the reduced system has at least this
                                      many 64-wide tile rows and is not solved distributed (default 80, i.e. >= 854
                                      cameras; 0 = never: FP64 DMMA only) */
  BA_OPT_TC_SLICES = 8,            /* INT8 slices per FP64 operand, 4 .. 7 (default 6: trailing updates to 2^-42 of
                                      the row scales; 7 = FP64 level; each step fewer is ~25 % faster and 128x coarser) */
  BA_OPT_TC_WINDOW = 9,            /* tile columns per panel of the blocked factorisation (even, 2 .. 16; default 8) */
  BA_OPT_TC_BK = 10,               /* bytes of the contraction per pipeline stage = TMA/UMMA swizzle span (64 or 128) */
  BA_OPT_TC_OVER_DIST_MAX_WORLD = 11 /* sharded handles of at most this many ranks (default 2) prefer "all-reduce + blocked
                                      tcgen05 solve on every rank" over the distributed solve when both qualify
                                      (ba_dist_solve_active then answers 0); set the same value on every rank */
};

enum { BA_MODEL_GAUSSIAN = 0, BA_MODEL_CAUCHY = 1 }; /* sensor_model.py:7-32 / :37-72 */

/* flags for ba_linearize_eliminate */
enum {
  BA_WANT_BLOCKS = 1, /* also emit HCCs/HPPs/HCPs/bCs/bPs (prepare_schur_complement outputs) */
  BA_WANT_SCHUR = 2   /* damp, invert point blocks, accumulate S and b */
};

/* selectors for ba_get_array */
enum {
  BA_ARR_HCC = 0,   /* [n_cam][6][6]   bundle_adjuster.py:106 */
  BA_ARR_HPP = 1,   /* [n_pt][3][3]    :108 */
  BA_ARR_HCP = 2,   /* [n_obs][6][3]   :107, sparse: one block per observation */
  BA_ARR_BC = 3,    /* [n_cam][6]      :110 */
  BA_ARR_BP = 4,    /* [n_pt][3]       :111 */
  BA_ARR_HPP_INV = 5, /* [n_pt][3][3]  :109 */
  BA_ARR_DC = 6,    /* [n_opt_cam][6]  solution of the reduced system (before negation) */
  BA_ARR_DP = 7,    /* [n_pt][3]       back-substituted point update (before negation); rows of
                                       tracks that are eliminated but not updated are zero */
  BA_ARR_RESIDUAL = 8, /* [n_obs][2]   bundle.py:251 (ba_eval_observations) */
  BA_ARR_JC = 9,    /* [n_obs][2][6]   bundle.py:255-277 */
  BA_ARR_JP = 10    /* [n_obs][2][3] */
};

const char* ba_version(void);
const char* ba_last_error(ba_handle h); /* text of the last CUDA error seen by this handle */

/* Leading dimension used for an n x n reduced system (n padded up to the solver tile). */
int ba_system_ld(int n);
/* Number of doubles in the packed reduced-system buffer for n_opt_cam optimised cameras. */
size_t ba_system_size(int n_opt_cam);

/* Problem sizes are fixed per handle (set_bundle, bundle_adjuster.py:54-114). */
int ba_create(int device, int n_cam, int n_pt, int n_obs, int n_opt_cam, int n_opt_pt,
              ba_handle* out);
int ba_destroy(ba_handle h);

/* K (bundle.py:138) and the sensor model (sensor_model.py) -- HOST pointers, copied.
 * params: Gaussian -> row-major 2x2 L; Cauchy -> {sigma, sigma^2, linear_window, 0}. */
int ba_set_intrinsics(ba_handle h, const double* K9_host);
int ba_set_sensor_model(ba_handle h, int kind, const double* params4_host);

/* Visibility structure (Track.measurements, bundle.py:95-111), constant across iterations. */
int ba_bind_structure(ba_handle h, const int* pt_ptr_dev, const int* obs_cam_dev,
                      const double* obs_uv_dev, const int* cam_slot_dev, const int* pt_slot_dev);
/* Current estimate and the candidate ("bnext", bundle_adjuster.py:143) parameter buffers. */
int ba_bind_state(ba_handle h, double* cam_R_dev, double* cam_t_dev, double* pts_dev);
int ba_bind_candidate(ba_handle h, double* cam_R_dev, double* cam_t_dev, double* pts_dev);
/* Reduced system buffer, ba_system_size(n_opt_cam) doubles, see layout above. */
int ba_bind_system(ba_handle h, double* sys_dev);
/* Overwrite the reduced system with a caller-supplied packed system (HOST, ba_system_size
 * doubles): what solve_motion_normal_eqns(S, b, mask) does with its arguments
 * (bundle_adjuster.py:281-290).  The next ba_solve factors exactly this system, on this rank. */
int ba_upload_system(ba_handle h, const double* packed_host, void* stream);
/* The reduced system as compute_schur_complement returns it (bundle_adjuster.py:247-278), packed,
 * to HOST memory (synchronises): after ba_allreduce_system the all-reduced copy, otherwise the
 * bound buffer. */
int ba_get_system(ba_handle h, double* packed_host, size_t count, void* stream);
/* Tuning / robustness knobs (BA_OPT_*). */
int ba_set_option(ba_handle h, int option, double value);

/* prepare_schur_complement + apply_damping + compute_schur_complement in one pass over the
 * observations (bundle_adjuster.py:211-234, :238-242, :247-278).  Zeroes and accumulates the
 * bound system buffer; also leaves the current cost (compute_cost, :165-171) in the scalars.
 * pinv_rcond < 0 selects the plain inverse (SCHUR_COMPLIMENT_PINV_THRESHOLD = None, :253). */
int ba_linearize_eliminate(ba_handle h, double damping, double pinv_rcond, int flags,
                           void* stream);

/* solve_motion_normal_eqns (bundle_adjuster.py:281-312): mirrors nothing, factors the
 * packed system (Cholesky of a dense copy; the bound buffer is left intact) and solves for dC.  cam_param_mask_host is NULL or
 * 6*n_opt_cam bytes (0 = parameter frozen, :296-309).  A non-positive pivot is recorded in
 * the scalars and reported by ba_read_scalars as BA_ERR_ILLCONDITIONED. */
int ba_solve(ba_handle h, const unsigned char* cam_param_mask_host, void* stream);

/* backsubstitute (:316-331) + update_motion/update_structure on the candidate (:334-343,
 * bundle.py:76-80: R <- R exp(-dC[:3]), t <- t - dC[3:], x <- x - dP) + compute_cost of the
 * candidate (:146). */
int ba_backsub_retract_cost(ba_handle h, void* stream);

/* compute_cost(bundle) (:165-171) of the CURRENT state only; result in the scalars. */
int ba_cost(ba_handle h, void* stream);

/* Accept the candidate (bundle_adjuster.py:151): swaps the state/candidate bindings. */
int ba_accept(ba_handle h);

/* Blocks until the stream is idle, then returns cost, candidate cost and the solver status
 * (BA_OK, BA_ERR_ILLCONDITIONED, BA_ERR_TIMEOUT, or BA_ERR_NONFINITE when a cost is NaN/Inf).
 * The only synchronising call on the product path. */
int ba_read_scalars(ba_handle h, double* cost, double* cand_cost, int* solve_status,
                    void* stream);
/* One whole LM trial driven from HOST buffers -- what BundleAdjuster.compute_update
 * (bundle_adjuster.py:176-208) plus the candidate cost of optimize (:143-146) amount to when the
 * bundle lives in host memory: H2D of the current estimate (cam_R [n_cam][9], cam_t [n_cam][3],
 * pts [n_pt][3]; any may be NULL = keep the device copy), ba_linearize_eliminate(BA_WANT_SCHUR),
 * ba_solve, ba_backsub_retract_cost, D2H of dC [6 n_opt_cam] and dP [n_pt][3] (either may be
 * NULL) and of the scalars, ONE stream synchronisation at the end.  Single-GPU handles only
 * (sharded problems need the host-side all-reduce between the stages).  Pass pinned host memory
 * for the copies to be asynchronous. */
int ba_trial_host(ba_handle h, const double* cam_R_host, const double* cam_t_host,
                  const double* pts_host, double damping, double pinv_rcond,
                  const unsigned char* cam_param_mask_host, double* dC_host, double* dP_host,
                  double* cost, double* cand_cost, int* solve_status, void* stream);

/* The same trial with ONE copy in each direction.  in_host = [cam_R 9 n_cam | cam_t 3 n_cam |
 * pts 3 n_pt] (NULL keeps the device estimate; a single H2D when the bound state arrays are
 * contiguous in that order, three otherwise); out_host receives
 * [cost, candidate cost, status (0 / 1 = ill-conditioned), spare | dC ba_system_ld(6 n_opt_cam) |
 * dP 3 n_pt] in a single D2H. */
int ba_trial_host_packed(ba_handle h, const double* in_host, double damping, double pinv_rcond,
                         const unsigned char* cam_param_mask_host, double* out_host, void* stream);

/* Device address of the 4-double scalar record {cost, cand_cost, status, spare} so the host
 * side can all-reduce the two costs when points are sharded over ranks. */
int ba_scalars_ptr(ba_handle h, double** scalars_dev);

/* ---- points sharded over the GPUs of ONE node --------------------------------
 * Peer-memory collectives over NVLink: each rank (one handle per GPU/process) exports one device
 * buffer through CUDA IPC and maps the buffers of its peers.
 *   ba_comm_create    allocates the buffer, returns its 64-byte IPC handle and REBINDS the system
 *                     buffer of the handle to it (ba_comm_system_ptr gives the device pointer);
 *   ba_comm_connect   takes the handles of all ranks (world x 64 bytes, rank order; exchanged by
 *                     the host side, e.g. torch.distributed.all_gather_object);
 *   ba_allreduce_system  ONE kernel: barrier, each rank sums its slice of all ranks' packed systems
 *                     (rank order) and pushes it to every rank, barrier; the next ba_solve factors
 *                     the reduced copy.  Replaces the NCCL all-reduce of the reduced camera system;
 *   ba_allreduce_costs   sums {cost, candidate cost} over the ranks into every rank's scalars (a no-op
 *                     right after a ba_backsub_retract_cost that already did it, BA_OPT_FUSE_COST_REDUCTION).
 * All ranks must issue the same sequence of these two calls. */
int ba_comm_create(ba_handle h, int rank, int world, unsigned char* ipc_handle_out64);
int ba_comm_connect(ba_handle h, const unsigned char* ipc_handles_all);
/* Unmaps the peers' buffers.  Tear-down is collective: every rank disconnects, the host side
 * synchronises the ranks, and only then may a rank destroy its handle (which frees the buffer its
 * peers had mapped). */
int ba_comm_disconnect(ba_handle h);
int ba_comm_system_ptr(ba_handle h, double** sys_dev);
int ba_allreduce_system(ba_handle h, void* stream);
int ba_allreduce_costs(ba_handle h, void* stream);
/* Large reduced systems on sharded handles: 1 when the next ba_solve on a fresh local contribution
 * (i.e. WITHOUT ba_allreduce_system in between) will run the DISTRIBUTED solve -- one launch per
 * rank that sums the ranks' contributions tile by tile over peer memory (reduce-scatter), factors
 * the tiles it owns, pushes each finished tile of L to every rank (all-gather) and substitutes
 * backwards on its own copy, so every rank ends up with the same dC.  Replaces "all-reduce, then
 * every rank factors the whole system" of SURVEY section 8e when the solve dominates (config 4).
 * All ranks must then call ba_solve collectively. */
int ba_dist_solve_active(ba_handle h);

/* 1 when ba_solve on this handle, as configured now, takes the blocked factorisation whose trailing
 * updates run on tcgen05 (BA_OPT_TC_MIN_TILES; large reduced systems that are not solved
 * distributed), 0 when it runs the FP64 dataflow kernel alone.  Both solve the same system
 * (solve_motion_normal_eqns, bundle_adjuster.py:281-312); they differ by the slice truncation
 * documented at BA_OPT_TC_SLICES. */
int ba_tc_solve_active(ba_handle h);

/* Per-observation residuals and Jacobians of the current state (bundle.py:251, :255-277),
 * for Bundle.residuals()/Jresiduals() and stage-wise parity checks. */
int ba_eval_observations(ba_handle h, void* stream);

/* Copies one internal array to HOST memory (synchronises).  count = number of doubles. */
int ba_get_array(ba_handle h, int which, double* dst_host, size_t count, void* stream);

/* Stand-alone retraction used by update_motion/update_structure on host bundles
 * (bundle_adjuster.py:334-343): cand <- state (+) (delta_cam, delta_pt), deltas are HOST
 * arrays [n_opt_cam][6] / [n_opt_pt][3] (either may be NULL = no change). */
int ba_retract(ba_handle h, const double* delta_cam_host, const double* delta_pt_host,
               void* stream);

/* Linear (algebraic) multi-view triangulation of EVERY bound point from the current cameras
 * and the measurements (triangulate.algebraic_lsq, triangulate.py:6-18, as called by
 * Bundle.triangulate_all, bundle.py:313-321); overwrites the bound state points [n_pt][3]. */
int ba_triangulate(ba_handle h, void* stream);

/* Device-side scene packer (set_bundle's selection, bundle_adjuster.py:54-101, applied to the
 * measurements bundle_io.load read, bundle_io.py:10-27): turns the RAW observation list of the whole
 * bundle -- raw_track[o], raw_cam[o] (ids), raw_uv[o][2], resident on the device -- plus two look-up
 * tables (track id -> position in the selection or -1, camera id -> position or -1) and cam_slot
 * into the point-major CSR arrays of the layout above: pt_ptr [n_pt + 1], obs_cam / obs_track [n_obs],
 * obs_uv [n_obs][2] (capacity n_raw each), every track's observations ordered by (cam_slot, camera
 * position).  scratch = 2 n_pt ints.  All pointers are DEVICE pointers except n_obs_out (host; the
 * call synchronises the stream to return it).  Stateless: no handle.  A sliding-window driver
 * (window_slam.py:30-39) calls this once per window with new tables; the observation list is
 * uploaded once. */
int ba_pack_observations(int device, int n_raw, const int* raw_track_dev, const int* raw_cam_dev,
                         const double* raw_uv_dev, int n_pt, const int* track_lut_dev,
                         const int* cam_lut_dev, const int* cam_slot_dev, int* pt_ptr_dev,
                         int* obs_cam_dev, double* obs_uv_dev, int* obs_track_dev, int* scratch_dev,
                         int* n_obs_out, void* stream);

/* Overwrite the reduced solution dC from a HOST array [n_opt_cam][6] -- lets the staged
 * Python API call backsubstitute(dC) with a caller-supplied camera update (:316). */
int ba_set_solution(ba_handle h, const double* dC_host, void* stream);

int ba_sync(ba_handle h, void* stream);

/* Diagnostics (the reference has none; SURVEY section 5 lists stage timers as the natural
 * counterpart of its print statements): nanoseconds, summed over the CTAs of all solver launches
 * since the last reset, in 16 slots: [0] plain panel tasks, [1] waiting for k-loop operands,
 * [2] last-step polls, [3] idle rounds waiting for rows of the diagonal inverse, [4] waiting for
 * the diagonal-update task, [5] raising flags on peers, [6] fetching the peers' contributions,
 * [7] waiting for y_k, [8] backward-substitution waits, [9] start barrier, [10] whole launch,
 * [11] chain tasks, [12] diagonal-update tasks.  Needs BA_OPT_SOLVER_PROFILE = 1.  Synchronises. */
int ba_solver_profile(ba_handle h, unsigned long long* out16_host, int reset, void* stream);

/* CUDA-event breakdown of the blocked reduced solve (BA_OPT_TC_MIN_TILES) since the last reset, in
 * milliseconds summed over the solves run with BA_OPT_SOLVER_PROFILE = 1: [0] expand, [1] panels (FP64
 * dataflow kernel on the leading tile columns), [2] INT8 slices + right-hand side update, [3] trailing
 * updates on tcgen05, [4] backward substitution; [5] = number of solves.  Synchronises on the last
 * profiled solve.  Diagnostics (the reference has none; SURVEY section 5: stage timers). */
int ba_tc_solve_profile(ba_handle h, double* out6_host, int reset);

/* Test entry of the tcgen05 trailing update that the blocked factorisation of large reduced
 * systems (BA_OPT_TC_MIN_TILES) applies after every panel -- one step of the Cholesky behind
 * solve_motion_normal_eqns (bundle_adjuster.py:281-312, numpy.linalg.solve there).  Stateless;
 * everything lives in HOST memory; synchronises.  A_host is a dense [ld x ld] column-major matrix
 * whose first K = 64 * window columns hold a factored panel L; the call performs, on the device,
 *     A[i, j] -= sum_k L[i, k] L[j, k]      for K <= j <= i < ld   (lower triangle of the trailing block)
 *     rhs[i]  -= sum_k L[i, k] rhs[k]        for K <= i < ld        (rows K .. K+63 start from saved_rhs
 *                                                                   when it is not NULL)
 * through `slices` (4 .. 7) signed 7-bit INT8 digit planes per FP64 entry, exact INT32 level sums
 * on the tensor cores, FP64 recombination (ba_solve_tc.cuh).  Optional outputs expose every integer
 * intermediate so that a CPU restatement can be held to it bit for bit:
 *   digits_host     [slices][ld_pad][K]  (ld_pad = ld rounded up to 128; rows < K are not written)
 *   scale_host      [ld_pad]             2^(e_i - 6), x_ik = L_ik 2^-e_i in (-1, 1)
 *   level_sums_host [slices][ld_pad][ld] sum over p + q = level, k of digit_p[i, k] digit_q[j, k]
 * ld: multiple of 64, > K; window even, 2 .. 16; bk 64 or 128.  Returns BA_ERR_TIMEOUT if a wait
 * inside the kernel ran past its deadline. */
int ba_tc_trailing_update_host(int device, int ld, int window, int slices, int bk, double* A_host,
                               double* rhs_host, const double* saved_rhs_host,
                               signed char* digits_host, double* scale_host, int* level_sums_host);

/* Number of kernel launches issued through this handle since creation (bench evidence). */
long long ba_launch_count(ba_handle h);

#ifdef __cplusplus
}
#endif
#endif /* BA_B200_H_ */
