// C ABI of libba_b200.so -- see include/ba_b200.h for the contract and the reference
// call sites each entry point replaces.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/ba_b200.h"
#include "ba_context.h"

using ba::Context;

struct ba_context : public Context {};

namespace {

int fail_cuda(Context* c, cudaError_t e, const char* where) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", where, cudaGetErrorString(e));
  if (c) c->last_error = buf;
  return BA_ERR_CUDA;
}

#define BA_CUDA(c, call)                                     \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return fail_cuda((c), e__, #call); \
  } while (0)

// The entry points run on the handle's device and leave the caller's current device as they found
// it (torch follows cudaGetDevice: a handle on cuda:1 must not move the process to GPU 1).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
      err = cudaSetDevice(dev);
      switched = (err == cudaSuccess) && prev >= 0;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define BA_ON_DEVICE(h)                 \
  DeviceGuard guard__((h)->device);     \
  if (guard__.err != cudaSuccess) return fail_cuda((h), guard__.err, "cudaSetDevice")

template <typename T>
cudaError_t dev_alloc(T** p, size_t count) {
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemset(*p, 0, count * sizeof(T));
}

// Wait for a stream by polling (the host-driven trial is ~0.6 ms long and its caller wants the result
// NOW: a blocking cudaStreamSynchronize adds the wake-up latency of the driver's interrupt path, tens
// of microseconds, to every trial).
cudaError_t spin_sync(cudaStream_t st) {
  cudaError_t e;
  while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
  }
  return e;
}

bool bound(const Context& c) {
  return c.pt_ptr && c.obs_cam && c.obs_uv && c.cam_slot && c.pt_slot && c.state.cam_R &&
         c.state.cam_t && c.state.pts;
}

int ensure_track_len(Context& c, cudaStream_t st) {
  if (c.max_track_len > 0) return BA_OK;
  std::vector<int> h((size_t)c.n_pt + 1);
  BA_CUDA(&c, cudaMemcpyAsync(h.data(), c.pt_ptr, h.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  BA_CUDA(&c, cudaStreamSynchronize(st));
  int m = 1;
  for (int i = 0; i < c.n_pt; ++i) {
    const int k = h[i + 1] - h[i];
    if (k < 0) { c.last_error = "pt_ptr is not monotone"; return BA_ERR_BAD_ARGUMENT; }
    if (k > m) m = k;
  }
  if (h[c.n_pt] != c.n_obs) { c.last_error = "pt_ptr[n_pt] != n_obs"; return BA_ERR_BAD_ARGUMENT; }
  c.max_track_len = m;
  return BA_OK;
}

}  // namespace

extern "C" {

const char* ba_version(void) { return "ba_b200 0.1 (sm_100a, fp64)"; }

const char* ba_last_error(ba_handle h) { return h ? h->last_error.c_str() : "null handle"; }

int ba_system_ld(int n) {
  if (n < 1) n = 1;
  return (n + ba::kSolveTile - 1) / ba::kSolveTile * ba::kSolveTile;
}

size_t ba_system_size(int n_opt_cam) {
  if (n_opt_cam < 0) n_opt_cam = 0;
  const size_t nc = (size_t)n_opt_cam;
  return nc * (nc + 1) / 2 * 36 + 6 * nc;
}

int ba_create(int device, int n_cam, int n_pt, int n_obs, int n_opt_cam, int n_opt_pt,
              ba_handle* out) {
  if (!out || n_cam < 1 || n_pt < 1 || n_obs < 0 || n_opt_cam < 0 || n_opt_cam > n_cam ||
      n_opt_pt < 0 || n_opt_pt > n_pt)
    return BA_ERR_BAD_ARGUMENT;
  *out = nullptr;
  DeviceGuard guard__(device);
  if (guard__.err != cudaSuccess) return BA_ERR_CUDA;
  ba_context* c = new (std::nothrow) ba_context();
  if (!c) return BA_ERR_CUDA;
  c->device = device;
  c->n_cam = n_cam; c->n_pt = n_pt; c->n_obs = n_obs;
  c->n_opt_cam = n_opt_cam; c->n_opt_pt = n_opt_pt;
  c->n_sys = 6 * n_opt_cam;
  c->ld = ba_system_ld(c->n_sys);
  c->sys_len = ba_system_size(n_opt_cam);
  const size_t T = (size_t)c->ld / ba::kSolveTile;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
    c->num_sms = sms;
  c->partials_cap = c->num_sms * 32;
  c->intr.K[0] = c->intr.K[4] = c->intr.K[8] = 1.0;  // Bundle default K = I (bundle.py:138)
  c->model.kind = BA_MODEL_GAUSSIAN;                 // GaussianModel(1.) (bundle.py:139)
  c->model.p[0] = 1.0; c->model.p[3] = 1.0;
  bool ok = dev_alloc(&c->Vinv, (size_t)n_pt * 9) == cudaSuccess &&
            dev_alloc(&c->bP, (size_t)n_pt * 3) == cudaSuccess &&
            dev_alloc(&c->V, (size_t)n_pt * 9) == cudaSuccess &&
            dev_alloc(&c->U, (size_t)n_cam * 36) == cudaSuccess &&
            dev_alloc(&c->bC, (size_t)n_cam * 6) == cudaSuccess &&
            dev_alloc(&c->io_out, (size_t)4 + c->ld + (size_t)n_pt * 3) == cudaSuccess &&
            dev_alloc(&c->Adense, (size_t)c->ld * c->ld + c->ld) == cudaSuccess &&
            dev_alloc(&c->LinvT, T * ba::kSolveTile * ba::kSolveTile) == cudaSuccess &&
            dev_alloc(&c->Wpart, T * (ba::kSolveTile * ba::kSolveTile + ba::kSolveTile)) == cudaSuccess &&
            dev_alloc(&c->solve_flags, ba::solve_flag_count((int)T)) == cudaSuccess &&
            dev_alloc(&c->solve_tickets, (size_t)2) == cudaSuccess &&
            dev_alloc(&c->solve_abort, (size_t)2) == cudaSuccess &&
            dev_alloc(&c->diag_rep, (size_t)16 * 36 * (size_t)n_opt_cam) == cudaSuccess &&
            dev_alloc(&c->solve_prof, (size_t)16) == cudaSuccess &&
            dev_alloc(&c->delta_cam, (size_t)n_cam * 6) == cudaSuccess &&
            dev_alloc(&c->delta_pt, (size_t)n_pt * 3) == cudaSuccess &&
            dev_alloc(&c->cam_mask, (size_t)c->ld) == cudaSuccess &&
            dev_alloc(&c->partials, (size_t)c->partials_cap) == cudaSuccess &&
            dev_alloc(&c->counters, (size_t)8) == cudaSuccess &&
            true;
  if (!ok) {
    ba_destroy(c);
    return BA_ERR_CUDA;
  }
  // results of a trial live in ONE block so that a host-driven trial reads them back in one copy
  static_assert(sizeof(ba::Scalars) == 4 * sizeof(double), "scalar record is 4 doubles");
  c->scalars = reinterpret_cast<ba::Scalars*>(c->io_out);
  c->dC = c->io_out + 4;
  c->dP = c->io_out + 4 + c->ld;
  *out = c;
  return BA_OK;
}

int ba_destroy(ba_handle h) {
  if (!h) return BA_OK;
  DeviceGuard guard__(h->device);
  void* ptrs[] = {h->Vinv, h->bP, h->V, h->U, h->bC, h->W, h->io_out, h->obs_r, h->obs_Jc,
                  h->obs_Jp, h->delta_cam, h->delta_pt, h->cam_mask, h->partials, h->counters,
                  h->Adense, h->LinvT, h->Wpart, h->solve_flags, h->solve_tickets, h->solve_abort, h->solve_prof, h->dist_tasks, h->diag_rep,
                  h->tc_slices, h->tc_scale, h->tc_save};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (cudaEvent_t e : h->tc_ev) cudaEventDestroy(e);
  for (int p = 0; p < ba::kMaxPeers; ++p)
    if (h->comm_peer[p] && p != h->comm_rank) cudaIpcCloseMemHandle(h->comm_peer[p]);
  if (h->rec_scratch) cudaFree(h->rec_scratch);
  if (h->comm_buf) cudaFree(h->comm_buf);
  if (h->comm_done) cudaFree(h->comm_done);
  delete h;
  return BA_OK;
}

// ---- peer-memory collectives (ba_comm.cu) -----------------------------------------------------
int ba_comm_create(ba_handle h, int rank, int world, unsigned char* ipc_handle_out) {
  if (!h || !ipc_handle_out || world < 2 || world > ba::kMaxPeers || rank < 0 || rank >= world)
    return BA_ERR_BAD_ARGUMENT;
  if (h->comm_buf) { h->last_error = "ba_comm_create called twice"; return BA_ERR_BAD_ARGUMENT; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  BA_ON_DEVICE(h);
  // [contrib | reduced | costs | flags] of the collectives, then the section of the distributed solve
  const size_t comm_len = (ba::comm_doubles(h->sys_len) + 31) & ~(size_t)31;
  BA_CUDA(h, dev_alloc(&h->comm_buf, comm_len + ba::dist_layout(h->ld).total));
  h->dist_off = comm_len;
  BA_CUDA(h, dev_alloc(&h->comm_done, (size_t)1));
  cudaIpcMemHandle_t mh;
  BA_CUDA(h, cudaIpcGetMemHandle(&mh, h->comm_buf));
  memcpy(ipc_handle_out, &mh, sizeof mh);
  h->comm_rank = rank;
  h->comm_world = world;
  h->comm_peer[rank] = h->comm_buf;
  h->sys = h->comm_buf;   // the rank's own contribution lives at the head of the exported buffer
  BA_CUDA(h, cudaDeviceSynchronize());
  return BA_OK;
}

int ba_comm_connect(ba_handle h, const unsigned char* ipc_handles_all) {
  if (!h || !ipc_handles_all || !h->comm_buf) return BA_ERR_BAD_ARGUMENT;
  BA_ON_DEVICE(h);
  for (int p = 0; p < h->comm_world; ++p) {
    if (p == h->comm_rank) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, ipc_handles_all + (size_t)p * sizeof mh, sizeof mh);
    void* ptr = nullptr;
    BA_CUDA(h, cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->comm_peer[p] = static_cast<double*>(ptr);
  }
  return BA_OK;
}

int ba_comm_disconnect(ba_handle h) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  BA_ON_DEVICE(h);
  BA_CUDA(h, cudaDeviceSynchronize());
  for (int p = 0; p < ba::kMaxPeers; ++p)
    if (h->comm_peer[p] && p != h->comm_rank) {
      cudaIpcCloseMemHandle(h->comm_peer[p]);
      h->comm_peer[p] = nullptr;
    }
  return BA_OK;
}

int ba_comm_system_ptr(ba_handle h, double** sys_dev) {
  if (!h || !sys_dev || !h->comm_buf) return BA_ERR_BAD_ARGUMENT;
  *sys_dev = h->comm_buf;
  return BA_OK;
}

int ba_allreduce_system(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!h->comm_buf) return BA_ERR_NOT_BOUND;
  for (int p = 0; p < h->comm_world; ++p)
    if (!h->comm_peer[p]) return BA_ERR_NOT_BOUND;
  BA_ON_DEVICE(h);
  BA_CUDA(h, ba::launch_peer_allreduce_system(*h, (cudaStream_t)stream));
  h->sys_state = ba::kSysReduced;
  return BA_OK;
}

int ba_allreduce_costs(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!h->comm_buf) return BA_ERR_NOT_BOUND;
  if (h->costs_reduced) {      // ba_backsub_retract_cost already reduced them in its epilogue
    h->costs_reduced = false;
    return BA_OK;
  }
  BA_ON_DEVICE(h);
  BA_CUDA(h, ba::launch_peer_allreduce_costs(*h, (cudaStream_t)stream));
  return BA_OK;
}

int ba_set_intrinsics(ba_handle h, const double* K9) {
  if (!h || !K9) return BA_ERR_BAD_ARGUMENT;
  memcpy(h->intr.K, K9, 9 * sizeof(double));
  return BA_OK;
}

int ba_set_sensor_model(ba_handle h, int kind, const double* p4) {
  if (!h || !p4 || (kind != BA_MODEL_GAUSSIAN && kind != BA_MODEL_CAUCHY)) return BA_ERR_BAD_ARGUMENT;
  h->model.kind = kind;
  memcpy(h->model.p, p4, 4 * sizeof(double));
  return BA_OK;
}

int ba_bind_structure(ba_handle h, const int* pt_ptr, const int* obs_cam, const double* obs_uv,
                      const int* cam_slot, const int* pt_slot) {
  if (!h || !pt_ptr || !obs_cam || !obs_uv || !cam_slot || !pt_slot) return BA_ERR_BAD_ARGUMENT;
  if (((size_t)obs_uv & 15) != 0) { h->last_error = "obs_uv must be 16-byte aligned"; return BA_ERR_BAD_ARGUMENT; }
  h->pt_ptr = pt_ptr; h->obs_cam = obs_cam; h->obs_uv = obs_uv;
  h->cam_slot = cam_slot; h->pt_slot = pt_slot;
  h->max_track_len = 0;
  return BA_OK;
}

int ba_bind_state(ba_handle h, double* R, double* t, double* pts) {
  if (!h || !R || !t || !pts) return BA_ERR_BAD_ARGUMENT;
  h->state.cam_R = R; h->state.cam_t = t; h->state.pts = pts;
  return BA_OK;
}

int ba_bind_candidate(ba_handle h, double* R, double* t, double* pts) {
  if (!h || !R || !t || !pts) return BA_ERR_BAD_ARGUMENT;
  h->cand.cam_R = R; h->cand.cam_t = t; h->cand.pts = pts;
  return BA_OK;
}

int ba_bind_system(ba_handle h, double* sys) {
  if (!h || !sys) return BA_ERR_BAD_ARGUMENT;
  if (((size_t)sys & 15) != 0) { h->last_error = "sys must be 16-byte aligned"; return BA_ERR_BAD_ARGUMENT; }
  h->sys = sys;
  h->sys_state = ba::kSysLocal;
  return BA_OK;
}

int ba_upload_system(ba_handle h, const double* packed_host, void* stream) {
  if (!h || !packed_host) return BA_ERR_BAD_ARGUMENT;
  if (!h->sys) return BA_ERR_NOT_BOUND;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  if (h->sys_len)
    BA_CUDA(h, cudaMemcpyAsync(h->sys, packed_host, h->sys_len * sizeof(double), cudaMemcpyHostToDevice, st));
  BA_CUDA(h, cudaStreamSynchronize(st));
  h->sys_state = ba::kSysUploaded;   // solved as is, on this rank (not the stale all-reduced copy)
  return BA_OK;
}

int ba_get_system(ba_handle h, double* packed_host, size_t count, void* stream) {
  if (!h || !packed_host) return BA_ERR_BAD_ARGUMENT;
  if (!h->sys) return BA_ERR_NOT_BOUND;
  if (count != h->sys_len) { h->last_error = "ba_get_system: wrong element count"; return BA_ERR_BAD_ARGUMENT; }
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  const double* src = (h->sys_state == ba::kSysReduced && h->comm_buf) ? h->comm_buf + ba::comm_pad(h->sys_len) : h->sys;
  if (count) BA_CUDA(h, cudaMemcpyAsync(packed_host, src, count * sizeof(double), cudaMemcpyDeviceToHost, st));
  BA_CUDA(h, cudaStreamSynchronize(st));
  return BA_OK;
}

int ba_set_option(ba_handle h, int option, double value) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  switch (option) {
    case BA_OPT_SPIN_TIMEOUT_MS: if (!(value > 0.0)) return BA_ERR_BAD_ARGUMENT; h->spin_timeout_ms = value; break;
    case BA_OPT_STRICT_FLAGS: h->strict_flags = value != 0.0; break;
    case BA_OPT_DIST_SOLVE_MIN_TILES: if (value < 0.0) return BA_ERR_BAD_ARGUMENT; h->dist_min_tiles = (int)value; break;
    case BA_OPT_DIST_BAND:
      if (value < 1.0 || value > 64.0) return BA_ERR_BAD_ARGUMENT;
      if (h->dist_tasks) { h->last_error = "BA_OPT_DIST_BAND must be set before the first distributed solve"; return BA_ERR_BAD_ARGUMENT; }
      h->dist_band = (int)value;
      break;
    case BA_OPT_SOLVE_GRID_CAP: if (value < 0.0) return BA_ERR_BAD_ARGUMENT; h->solve_grid_cap = (int)value; break;
    case BA_OPT_SOLVER_PROFILE: h->solve_prof_on = value != 0.0; break;
    case BA_OPT_FUSE_COST_REDUCTION: h->fuse_cost_reduction = value != 0.0; break;
    case BA_OPT_TC_MIN_TILES: if (value < 0.0) return BA_ERR_BAD_ARGUMENT; h->tc_min_tiles = (int)value; break;
    case BA_OPT_TC_SLICES: if (value < 4.0 || value > 7.0) return BA_ERR_BAD_ARGUMENT; h->tc_slices_n = (int)value; break;
    case BA_OPT_TC_WINDOW: {
      const int w = (int)value;
      if (w < 2 || w > 16 || (w & 1)) return BA_ERR_BAD_ARGUMENT;
      h->tc_window = w;
      break;
    }
    case BA_OPT_TC_BK: if (value != 64.0 && value != 128.0) return BA_ERR_BAD_ARGUMENT; h->tc_bk = (int)value; break;
    case BA_OPT_TC_OVER_DIST_MAX_WORLD: if (value < 0.0) return BA_ERR_BAD_ARGUMENT; h->tc_over_dist_max_world = (int)value; break;
    default: return BA_ERR_BAD_ARGUMENT;
  }
  return BA_OK;
}

int ba_dist_solve_active(ba_handle h) {
  if (!h) return 0;
  ba::Context probe = *h;          // the question is about a FRESH local contribution
  probe.sys_state = ba::kSysLocal;
  for (int p = 0; p < probe.comm_world; ++p)
    if (!probe.comm_peer[p]) return 0;
  return ba::dist_solve_selected(probe) ? 1 : 0;
}

int ba_tc_solve_active(ba_handle h) {
  if (!h) return 0;
  return ba::tc_solve_selected(*h) ? 1 : 0;
}

int ba_linearize_eliminate(ba_handle h, double damping, double pinv_rcond, int flags, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h) || ((flags & BA_WANT_SCHUR) && !h->sys)) return BA_ERR_NOT_BOUND;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  int rc = ensure_track_len(*h, st);
  if (rc != BA_OK) return rc;
  if ((flags & BA_WANT_BLOCKS) && !h->W) BA_CUDA(h, dev_alloc(&h->W, (size_t)h->n_obs * 18));
  if (flags & BA_WANT_SCHUR) h->sys_state = ba::kSysLocal;   // a fresh local contribution
  h->costs_reduced = false;                                   // ... and a fresh local cost
  cudaError_t e = ba::launch_linearize_eliminate(*h, damping, pinv_rcond, flags, st);
  BA_CUDA(h, e);
  return BA_OK;
}

int ba_solve(ba_handle h, const unsigned char* mask_host, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!h->sys) return BA_ERR_NOT_BOUND;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  if (mask_host && h->n_sys > 0)
    BA_CUDA(h, cudaMemcpyAsync(h->cam_mask, mask_host, (size_t)h->n_sys, cudaMemcpyHostToDevice, st));
  BA_CUDA(h, ba::launch_solve(*h, mask_host != nullptr, st));
  return BA_OK;
}

int ba_backsub_retract_cost(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h) || !h->cand.cam_R || !h->cand.cam_t || !h->cand.pts) return BA_ERR_NOT_BOUND;
  BA_ON_DEVICE(h);
  BA_CUDA(h, ba::launch_backsub_retract_cost(*h, (cudaStream_t)stream));
  return BA_OK;
}

int ba_cost(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h)) return BA_ERR_NOT_BOUND;
  h->costs_reduced = false;
  BA_ON_DEVICE(h);
  BA_CUDA(h, ba::launch_cost(*h, (cudaStream_t)stream));
  return BA_OK;
}

int ba_accept(ba_handle h) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!h->cand.cam_R) return BA_ERR_NOT_BOUND;
  ba::ParamSet t = h->state;
  h->state = h->cand;
  h->cand = t;
  return BA_OK;
}

int ba_read_scalars(ba_handle h, double* cost, double* cand_cost, int* solve_status, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  ba::Scalars s;
  BA_CUDA(h, cudaMemcpyAsync(&s, h->scalars, sizeof s, cudaMemcpyDeviceToHost, st));
  BA_CUDA(h, spin_sync(st));
  if (cost) *cost = s.cost;
  if (cand_cost) *cand_cost = s.cand_cost;
  if (solve_status) {
    // 2 = a spin-wait gave up; 1 = non-positive pivot; a NaN/Inf cost is reported as well (the
    // reference's drivers run under numpy.seterr(all='raise'), window_slam.py:70)
    *solve_status = (s.status == 2.0) ? BA_ERR_TIMEOUT : (s.status != 0.0) ? BA_ERR_ILLCONDITIONED :
                    !(isfinite(s.cost) && isfinite(s.cand_cost)) ? BA_ERR_NONFINITE : BA_OK;
  }
  return BA_OK;
}

int ba_trial_host(ba_handle h, const double* cam_R_host, const double* cam_t_host,
                  const double* pts_host, double damping, double pinv_rcond,
                  const unsigned char* cam_param_mask_host, double* dC_host, double* dP_host,
                  double* cost, double* cand_cost, int* solve_status, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h) || !h->sys || !h->cand.cam_R || !h->cand.cam_t || !h->cand.pts) return BA_ERR_NOT_BOUND;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  if (cam_R_host)
    BA_CUDA(h, cudaMemcpyAsync(h->state.cam_R, cam_R_host, (size_t)h->n_cam * 9 * sizeof(double), cudaMemcpyHostToDevice, st));
  if (cam_t_host)
    BA_CUDA(h, cudaMemcpyAsync(h->state.cam_t, cam_t_host, (size_t)h->n_cam * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  if (pts_host)
    BA_CUDA(h, cudaMemcpyAsync(h->state.pts, pts_host, (size_t)h->n_pt * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = ba_linearize_eliminate(h, damping, pinv_rcond, BA_WANT_SCHUR, stream);
  if (rc != BA_OK) return rc;
  if ((rc = ba_solve(h, cam_param_mask_host, stream)) != BA_OK) return rc;
  if ((rc = ba_backsub_retract_cost(h, stream)) != BA_OK) return rc;
  if (dC_host && h->n_sys)
    BA_CUDA(h, cudaMemcpyAsync(dC_host, h->dC, (size_t)h->n_sys * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (dP_host)
    BA_CUDA(h, cudaMemcpyAsync(dP_host, h->dP, (size_t)h->n_pt * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  return ba_read_scalars(h, cost, cand_cost, solve_status, stream);
}

int ba_trial_host_packed(ba_handle h, const double* in_host, double damping, double pinv_rcond,
                         const unsigned char* cam_param_mask_host, double* out_host, void* stream) {
  if (!h || !out_host) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h) || !h->sys || !h->cand.cam_R || !h->cand.cam_t || !h->cand.pts) return BA_ERR_NOT_BOUND;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  const size_t nR = (size_t)h->n_cam * 9, nt = (size_t)h->n_cam * 3, nx = (size_t)h->n_pt * 3;
  if (in_host) {
    if (h->state.cam_t == h->state.cam_R + nR && h->state.pts == h->state.cam_t + nt) {
      BA_CUDA(h, cudaMemcpyAsync(h->state.cam_R, in_host, (nR + nt + nx) * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
      BA_CUDA(h, cudaMemcpyAsync(h->state.cam_R, in_host, nR * sizeof(double), cudaMemcpyHostToDevice, st));
      BA_CUDA(h, cudaMemcpyAsync(h->state.cam_t, in_host + nR, nt * sizeof(double), cudaMemcpyHostToDevice, st));
      BA_CUDA(h, cudaMemcpyAsync(h->state.pts, in_host + nR + nt, nx * sizeof(double), cudaMemcpyHostToDevice, st));
    }
  }
  int rc = ba_linearize_eliminate(h, damping, pinv_rcond, BA_WANT_SCHUR, stream);
  if (rc != BA_OK) return rc;
  if ((rc = ba_solve(h, cam_param_mask_host, stream)) != BA_OK) return rc;
  if ((rc = ba_backsub_retract_cost(h, stream)) != BA_OK) return rc;
  BA_CUDA(h, cudaMemcpyAsync(out_host, h->io_out, ((size_t)4 + h->ld + nx) * sizeof(double), cudaMemcpyDeviceToHost, st));
  BA_CUDA(h, spin_sync(st));
  return BA_OK;
}

int ba_scalars_ptr(ba_handle h, double** p) {
  if (!h || !p) return BA_ERR_BAD_ARGUMENT;
  *p = reinterpret_cast<double*>(h->scalars);
  return BA_OK;
}

int ba_eval_observations(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h)) return BA_ERR_NOT_BOUND;
  BA_ON_DEVICE(h);
  if (!h->obs_r) {
    BA_CUDA(h, dev_alloc(&h->obs_r, (size_t)h->n_obs * 2));
    BA_CUDA(h, dev_alloc(&h->obs_Jc, (size_t)h->n_obs * 12));
    BA_CUDA(h, dev_alloc(&h->obs_Jp, (size_t)h->n_obs * 6));
  }
  BA_CUDA(h, ba::launch_eval_observations(*h, (cudaStream_t)stream));
  return BA_OK;
}

int ba_get_array(ba_handle h, int which, double* dst, size_t count, void* stream) {
  if (!h || !dst) return BA_ERR_BAD_ARGUMENT;
  const double* src = nullptr;
  size_t n = 0;
  switch (which) {
    case BA_ARR_HCC: src = h->U; n = (size_t)h->n_cam * 36; break;
    case BA_ARR_HPP: src = h->V; n = (size_t)h->n_pt * 9; break;
    case BA_ARR_HCP: src = h->W; n = (size_t)h->n_obs * 18; break;
    case BA_ARR_BC: src = h->bC; n = (size_t)h->n_cam * 6; break;
    case BA_ARR_BP: src = h->bP; n = (size_t)h->n_pt * 3; break;
    case BA_ARR_HPP_INV: src = h->Vinv; n = (size_t)h->n_pt * 9; break;
    case BA_ARR_DC: src = h->dC; n = (size_t)h->n_sys; break;
    case BA_ARR_DP: src = h->dP; n = (size_t)h->n_pt * 3; break;
    case BA_ARR_RESIDUAL: src = h->obs_r; n = (size_t)h->n_obs * 2; break;
    case BA_ARR_JC: src = h->obs_Jc; n = (size_t)h->n_obs * 12; break;
    case BA_ARR_JP: src = h->obs_Jp; n = (size_t)h->n_obs * 6; break;
    default: return BA_ERR_BAD_ARGUMENT;
  }
  if (!src) return BA_ERR_NOT_BOUND;
  if (count != n) { h->last_error = "ba_get_array: wrong element count"; return BA_ERR_BAD_ARGUMENT; }
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  if (n) BA_CUDA(h, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  BA_CUDA(h, cudaStreamSynchronize(st));
  return BA_OK;
}

int ba_retract(ba_handle h, const double* delta_cam_host, const double* delta_pt_host, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h) || !h->cand.cam_R) return BA_ERR_NOT_BOUND;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  if (delta_cam_host && h->n_opt_cam)
    BA_CUDA(h, cudaMemcpyAsync(h->delta_cam, delta_cam_host, (size_t)h->n_opt_cam * 6 * sizeof(double),
                               cudaMemcpyHostToDevice, st));
  if (delta_pt_host && h->n_opt_pt)
    BA_CUDA(h, cudaMemcpyAsync(h->delta_pt, delta_pt_host, (size_t)h->n_opt_pt * 3 * sizeof(double),
                               cudaMemcpyHostToDevice, st));
  BA_CUDA(h, ba::launch_retract(*h, delta_cam_host != nullptr, delta_pt_host != nullptr, st));
  return BA_OK;
}

int ba_triangulate(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  if (!bound(*h)) return BA_ERR_NOT_BOUND;
  BA_ON_DEVICE(h);
  BA_CUDA(h, ba::launch_triangulate(*h, (cudaStream_t)stream));
  return BA_OK;
}

int ba_set_solution(ba_handle h, const double* dC_host, void* stream) {
  if (!h || !dC_host) return BA_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  if (h->n_sys)
    BA_CUDA(h, cudaMemcpyAsync(h->dC, dC_host, (size_t)h->n_sys * sizeof(double), cudaMemcpyHostToDevice, st));
  BA_CUDA(h, cudaStreamSynchronize(st));
  return BA_OK;
}

int ba_sync(ba_handle h, void* stream) {
  if (!h) return BA_ERR_BAD_ARGUMENT;
  BA_ON_DEVICE(h);
  BA_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return BA_OK;
}

int ba_solver_profile(ba_handle h, unsigned long long* out16, int reset, void* stream) {
  if (!h || !out16) return BA_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  BA_ON_DEVICE(h);
  BA_CUDA(h, cudaMemcpyAsync(out16, h->solve_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (reset) BA_CUDA(h, cudaMemsetAsync(h->solve_prof, 0, 16 * sizeof(unsigned long long), st));
  BA_CUDA(h, cudaStreamSynchronize(st));
  return BA_OK;
}

int ba_tc_solve_profile(ba_handle h, double* out6_host, int reset) {
  if (!h || !out6_host) return BA_ERR_BAD_ARGUMENT;
  BA_ON_DEVICE(h);
  ba::tc_fold_profile(*h);
  for (int i = 0; i < 5; ++i) out6_host[i] = h->tc_prof_ms[i];
  out6_host[5] = (double)h->tc_prof_solves;
  if (reset) {
    for (int i = 0; i < 5; ++i) h->tc_prof_ms[i] = 0.0;
    h->tc_prof_solves = 0;
  }
  return BA_OK;
}

long long ba_launch_count(ba_handle h) { return h ? h->launches : 0; }

int ba_tc_trailing_update_host(int device, int ld, int window, int slices, int bk, double* A_host, double* rhs_host,
                               const double* saved_rhs_host, signed char* digits_host, double* scale_host, int* level_sums_host) {
  if (ld < 128 || (ld % 64) != 0 || !A_host || !rhs_host || ld <= 64 * window) return BA_ERR_BAD_ARGUMENT;
  DeviceGuard guard__(device);
  ba::Context c;   // a scratch context: only the solver workspace of the trailing update exists
  c.device = device;
  c.ld = ld;
  c.tc_window = window; c.tc_slices_n = slices; c.tc_bk = bk;
  cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, device);
  const size_t ld_pad = ((size_t)ld + 127) / 128 * 128, K = (size_t)64 * window;
  const size_t dense = (size_t)ld * ld + ld;
  int rc = BA_OK;
  bool ok = dev_alloc(&c.Adense, dense) == cudaSuccess && dev_alloc(&c.scalars, 1) == cudaSuccess &&
            dev_alloc(&c.solve_abort, 2) == cudaSuccess &&
            cudaMemset(c.scalars, 0, sizeof(ba::Scalars)) == cudaSuccess && cudaMemset(c.solve_abort, 0, 8) == cudaSuccess &&
            ba::tc_prepare(c) == cudaSuccess;
  if (ok && level_sums_host) {
    ok = dev_alloc(&c.tc_dbg, (size_t)slices * ld_pad * ld) == cudaSuccess &&
         cudaMemset(c.tc_dbg, 0, (size_t)slices * ld_pad * ld * sizeof(int)) == cudaSuccess;
    c.tc_dbg_ld = ld;
  }
  if (ok) {
    ok = cudaMemcpy(c.Adense, A_host, (size_t)ld * ld * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(c.Adense + (size_t)ld * ld, rhs_host, (size_t)ld * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && saved_rhs_host) ok = cudaMemcpy(c.tc_save, saved_rhs_host, 64 * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (ok) {
    ok = ba::launch_tc_trailing_update(c, c.Adense, c.Adense + (size_t)ld * ld, 0, saved_rhs_host ? c.tc_save : nullptr, 0) == cudaSuccess &&
         cudaDeviceSynchronize() == cudaSuccess;
  }
  if (ok) {
    ba::Scalars sc;
    ok = cudaMemcpy(&sc, c.scalars, sizeof sc, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(A_host, c.Adense, (size_t)ld * ld * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(rhs_host, c.Adense + (size_t)ld * ld, (size_t)ld * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok && digits_host) ok = cudaMemcpy(digits_host, c.tc_slices, (size_t)slices * ld_pad * K, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok && scale_host) ok = cudaMemcpy(scale_host, c.tc_scale, ld_pad * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok && level_sums_host)
      ok = cudaMemcpy(level_sums_host, c.tc_dbg, (size_t)slices * ld_pad * ld * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok && sc.status != 0.0) rc = BA_ERR_TIMEOUT;
  }
  if (!ok) rc = (rc == BA_OK) ? BA_ERR_CUDA : rc;
  void* ptrs[] = {c.Adense, c.scalars, c.solve_abort, c.tc_slices, c.tc_scale, c.tc_save, c.tc_dbg};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (rc == BA_ERR_CUDA) cudaGetLastError();
  return rc;
}

}  // extern "C"
