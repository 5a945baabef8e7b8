// Device-side scene packer: observation list -> point-major CSR in the layout of include/ba_b200.h.
//
// The reference keeps measurements as a dict per Track object (bundle.py:95-111, filled from text
// by bundle_io.load, bundle_io.py:10-27) and BundleAdjuster.set_bundle selects cameras / tracks by
// id lists (bundle_adjuster.py:54-101).  Here the raw observation list of the WHOLE bundle
// (track id, camera id, pixel) is uploaded once; a selection is two small look-up tables
// (camera id -> position, track id -> position, -1 = not selected), and this file turns list +
// tables into the packed sub-problem without the data leaving the device:
//
//   count_kernel      observations per selected track (integer atomics)
//   scan_kernel       exclusive prefix sum -> pt_ptr      (one CTA, chunked block scan)
//   scatter_kernel    each kept observation into its track's segment (atomic cursor)
//   sort_kernel       every segment ordered by (reduced-system slot, camera position): fixed cameras
//                     first, then ascending slot -- the order linearize_eliminate_kernel relies on
//                     for its upper-triangular block pairs; a camera appears at most once per
//                     track, so the result is unique whatever order the scatter produced.
//
// A sliding-window driver (window_slam.py:30-39) re-runs only this with new tables: the
// observation list is never uploaded again.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ba_b200.h"

namespace ba {

__global__ void __launch_bounds__(256) pack_count_kernel(const int* __restrict__ raw_track, const int* __restrict__ raw_cam, int n_raw,
                                                         const int* __restrict__ track_lut, const int* __restrict__ cam_lut,
                                                         int* __restrict__ counts) {
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_raw; o += gridDim.x * blockDim.x) {
    const int tp = track_lut[raw_track[o]];
    const int cp = cam_lut[raw_cam[o]];
    if (tp >= 0 && cp >= 0) atomicAdd(counts + tp, 1);
  }
}

// pt_ptr[i] = sum of counts[0 .. i), pt_ptr[n] = total; cursor[i] = pt_ptr[i].  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) pack_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ pt_ptr,
                                                         int* __restrict__ cursor) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + (wid > 0 ? s_warp[wid - 1] : 0) + x - v;
    if (i < n) { pt_ptr[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) pt_ptr[n] = s_carry;
}

__global__ void __launch_bounds__(256) pack_scatter_kernel(const int* __restrict__ raw_track, const int* __restrict__ raw_cam,
                                                           const double* __restrict__ raw_uv, int n_raw,
                                                           const int* __restrict__ track_lut, const int* __restrict__ cam_lut,
                                                           int* __restrict__ cursor, int* __restrict__ obs_cam,
                                                           double* __restrict__ obs_uv, int* __restrict__ obs_track) {
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_raw; o += gridDim.x * blockDim.x) {
    const int tp = track_lut[raw_track[o]];
    const int cp = cam_lut[raw_cam[o]];
    if (tp < 0 || cp < 0) continue;
    const int dst = atomicAdd(cursor + tp, 1);
    obs_cam[dst] = cp;
    obs_track[dst] = tp;
    reinterpret_cast<double2*>(obs_uv)[dst] = reinterpret_cast<const double2*>(raw_uv)[o];
  }
}

// One thread per track: insertion sort of its segment by (cam_slot, camera position).  Tracks are a
// few to a few hundred observations long; this is set-up work, run once per selection.
__global__ void __launch_bounds__(128) pack_sort_kernel(const int* __restrict__ pt_ptr, int n_pt, const int* __restrict__ cam_slot,
                                                        int* __restrict__ obs_cam, double* __restrict__ obs_uv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  const int beg = pt_ptr[p], end = pt_ptr[p + 1];
  double2* uv = reinterpret_cast<double2*>(obs_uv);
  for (int i = beg + 1; i < end; ++i) {
    const int c = obs_cam[i];
    const double2 z = uv[i];
    const long long key = ((long long)cam_slot[c] << 32) | (unsigned int)c;
    int j = i - 1;
    while (j >= beg) {
      const int cj = obs_cam[j];
      const long long kj = ((long long)cam_slot[cj] << 32) | (unsigned int)cj;
      if (kj <= key) break;
      obs_cam[j + 1] = cj;
      uv[j + 1] = uv[j];
      --j;
    }
    obs_cam[j + 1] = c;
    uv[j + 1] = z;
  }
}

}  // namespace ba

extern "C" int ba_pack_observations(int device, int n_raw, const int* raw_track_dev, const int* raw_cam_dev,
                                    const double* raw_uv_dev, int n_pt, const int* track_lut_dev, const int* cam_lut_dev,
                                    const int* cam_slot_dev, int* pt_ptr_dev, int* obs_cam_dev, double* obs_uv_dev,
                                    int* obs_track_dev, int* scratch_dev, int* n_obs_out, void* stream) {
  if (n_raw < 0 || n_pt < 1 || !track_lut_dev || !cam_lut_dev || !cam_slot_dev || !pt_ptr_dev || !scratch_dev || !n_obs_out)
    return BA_ERR_BAD_ARGUMENT;
  if (n_raw > 0 && (!raw_track_dev || !raw_cam_dev || !raw_uv_dev || !obs_cam_dev || !obs_uv_dev || !obs_track_dev))
    return BA_ERR_BAD_ARGUMENT;
  if ((((size_t)raw_uv_dev | (size_t)obs_uv_dev) & 15) != 0) return BA_ERR_BAD_ARGUMENT;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device && cudaSetDevice(device) != cudaSuccess) return BA_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  int* counts = scratch_dev;            // [n_pt]
  int* cursor = scratch_dev + n_pt;     // [n_pt]
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)n_pt * sizeof(int), st);
  const int grid = n_raw > 0 ? (n_raw + 255) / 256 < 148 * 8 ? (n_raw + 255) / 256 : 148 * 8 : 1;
  if (e == cudaSuccess && n_raw > 0) ba::pack_count_kernel<<<grid, 256, 0, st>>>(raw_track_dev, raw_cam_dev, n_raw, track_lut_dev, cam_lut_dev, counts);
  if (e == cudaSuccess) ba::pack_scan_kernel<<<1, 1024, 0, st>>>(counts, n_pt, pt_ptr_dev, cursor);
  if (e == cudaSuccess && n_raw > 0) {
    ba::pack_scatter_kernel<<<grid, 256, 0, st>>>(raw_track_dev, raw_cam_dev, raw_uv_dev, n_raw, track_lut_dev, cam_lut_dev, cursor,
                                                  obs_cam_dev, obs_uv_dev, obs_track_dev);
    ba::pack_sort_kernel<<<(n_pt + 127) / 128, 128, 0, st>>>(pt_ptr_dev, n_pt, cam_slot_dev, obs_cam_dev, obs_uv_dev);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(n_obs_out, pt_ptr_dev + n_pt, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}
