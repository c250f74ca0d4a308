// Motion LOADER on the device: on-disk clip format -> the tables the hot path reads.
//
// Replaces, for all clips of a (re)load at once, the per-clip CPU work of
//   MotionLibSMPL.load_motion_with_skeleton   phc/utils/motion_lib_smpl.py:101-180   (heading randomisation :141-149)
//   SkeletonState.local_rotation / global_transformation  poselib/poselib/skeleton/skeleton3d.py:390-461  (FK)
//   SkeletonMotion._compute_velocity / _compute_angular_velocity  skeleton3d.py:1100-1121  (np.gradient, angle-axis of
//       frame-to-frame rotations, scipy gaussian_filter1d(sigma = 2, mode = "nearest"))
//   compute_motion_dof_vels                    phc/utils/motion_lib_base.py:47-70
// and writes gts / grs / lrs / gvs / gavs / dvs as float32 (the `.float()` of motion_lib_base.py:300-307); phc_motion_pack
// then builds the per-frame records.  The reference does this on CPU worker processes every `shape_resampling_interval`
// epochs and re-uploads ~1 GB.
//
// Numerics follow the reference's own precision choices: float64 everywhere (the clips are float64 on disk) EXCEPT
// (a) the local rotations, which poselib assembles in a float32 buffer (quat_identity_like, skeleton3d.py:449-459), so the
// FK chain and the dof velocities see float32-rounded local rotations, and (b) the dof velocities, which therefore run
// through phc's quat_mul / quat_to_angle_axis in float32 (phc_math.cuh, reference operation order, no FMA contraction:
// this file is compiled with -fmad=false).
//
// Kernel 1 (motion_fk_kernel): one warp per frame, lane = body (bodies > 32: strided).  Frame f and f+1 rotations are staged in
// shared memory; the tree is walked level by level (depth from the parent table), one __syncwarp per level.
// Kernel 2 (motion_filter_kernel): one thread per (frame, body): 17-tap gaussian over np.gradient of the float64 positions
// and over the raw angular velocities, indices clamped to the clip ("nearest").
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/phc_b200.h"
#include "phc_math.cuh"

extern "C" void phc_set_error(const char* msg);
extern "C" int phc_check_cuda(cudaError_t e, const char* what);
extern "C" void phc_count_launches(int n);

namespace phc {
namespace load {

constexpr int kWarps = 4;
constexpr int kRadius = 8;                      // int(4.0 * sigma + 0.5), sigma = 2
struct Taps { double w[2 * kRadius + 1]; };

struct D4 { double x, y, z, w; };
struct D3 { double x, y, z; };

__device__ __forceinline__ D4 dmul(D4 a, D4 b) {          // Hamilton product (core/rotation3d.py:15-27)
  D4 r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}
__device__ __forceinline__ D4 dconj(D4 q) { D4 r = {-q.x, -q.y, -q.z, q.w}; return r; }
__device__ __forceinline__ D4 dunit(D4 q) {
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  D4 r = {q.x / n, q.y / n, q.z / n, q.w / n};
  return r;
}
// quat_normalize = unit(quat_pos(q)) (core/rotation3d.py:31-39, :93-99)
__device__ __forceinline__ D4 dnormpos(D4 q) {
  if (q.w < 0.0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  return dunit(q);
}
// quat_rotate: imaginary part of q * (v, 0) * conj(q) (core/rotation3d.py:206-212)
__device__ __forceinline__ D3 drot(D4 q, D3 v) {
  D4 vq = {v.x, v.y, v.z, 0.0};
  const D4 r = dmul(dmul(q, vq), dconj(q));
  D3 o = {r.x, r.y, r.z};
  return o;
}
__device__ __forceinline__ D4 ld4(const double* p) { D4 r = {p[0], p[1], p[2], p[3]}; return r; }
__device__ __forceinline__ void st4(double* p, D4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

// binary search: largest c with starts[c] <= row
__device__ __forceinline__ int find_clip(const int64_t* starts, int M, int64_t row) {
  int lo = 0, hi = M - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (starts[mid] <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

struct FkArgs {
  const double* quat;      // [F, J, 4] pose_quat_global (xyzw)
  const double* trans;     // [F, 3]    root_trans_offset
  const double* offsets;   // [M, J, 3] skeleton_tree.local_translation of every clip's skeleton
  const int32_t* parents;  // [J]
  const double* heading;   // [M] or NULL
  const int64_t* starts;   // [M]
  const int64_t* nframes;  // [M]
  const double* fps;       // [M]
  int64_t F;
  int32_t M, J;
  float *gts, *grs, *lrs, *dvs;
  double* pos64;           // [F, J, 3]
  double* rawang;          // [F, J, 3]
  int32_t* frame_clip;     // [F]
};

__global__ void __launch_bounds__(kWarps * 32) motion_fk_kernel(const FkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = a.J;
  // per-warp carve-up: g0[J][4], g1[J][4], ch[J][4], ps[J][3] doubles; l0[J][4], l1[J][4] floats; depth[J] ints
  const size_t per_warp = (size_t)J * (15 * sizeof(double) + 8 * sizeof(float) + sizeof(int));
  unsigned char* base = smem_raw + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
  double* g0 = reinterpret_cast<double*>(base);
  double* g1 = g0 + 4 * J;
  double* ch = g1 + 4 * J;
  double* ps = ch + 4 * J;
  float* l0 = reinterpret_cast<float*>(ps + 3 * J);
  float* l1 = l0 + 4 * J;
  int* depth = reinterpret_cast<int*>(l1 + 4 * J);

  const int64_t f = (int64_t)blockIdx.x * kWarps + warp;
  if (f >= a.F) return;
  int c = 0;
  if (lane == 0) c = find_clip(a.starts, a.M, f);
  c = __shfl_sync(0xffffffffu, c, 0);
  const int64_t start = a.starts[c], nf = a.nframes[c];
  const int64_t fl = f - start;
  const bool has_next = fl + 1 < nf;
  const double dt = 1 / a.fps[c];                 // time_delta = 1 / fps
  const float dt32 = (float)(1.0 / a.fps[c]);     // compute_motion_dof_vels: float32 tensor / python float
  if (lane == 0) a.frame_clip[f] = c;

  bool rot = a.heading != nullptr;
  D4 hq = {0.0, 0.0, 0.0, 1.0};
  double hc = 1.0, hs = 0.0;
  if (rot) {
    const double th = a.heading[c];
    hq.z = sin(0.5 * th); hq.w = cos(0.5 * th);
    hc = cos(th); hs = sin(th);
  }
  // ---- stage the (heading-rotated) global rotations of frame f and f+1, tree depth of every body ---------------------
  for (int j = lane; j < J; j += 32) {
    D4 q0 = ld4(a.quat + ((size_t)f * J + j) * 4);
    D4 q1 = has_next ? ld4(a.quat + ((size_t)(f + 1) * J + j) * 4) : q0;
    if (rot) {       // Rotation.from_quat normalises; random_heading_rot * R; as_quat (no sign canonicalisation)
      q0 = dunit(dmul(hq, dunit(q0)));
      q1 = dunit(dmul(hq, dunit(q1)));
    }
    st4(g0 + 4 * j, q0);
    st4(g1 + 4 * j, q1);
    float* o = a.grs + ((size_t)f * J + j) * 4;
    o[0] = (float)q0.x; o[1] = (float)q0.y; o[2] = (float)q0.z; o[3] = (float)q0.w;
    int d = 0;
    for (int p = a.parents[j]; p >= 0 && d < J; p = a.parents[p]) ++d;      // bounded: a malformed (cyclic) table cannot hang
    depth[j] = d;
  }
  __syncwarp();
  // ---- local rotations (float32 buffer), raw angular velocity, dof velocity ------------------------------------------
  int max_depth = 0;
  for (int j = lane; j < J; j += 32) {
    const int p = a.parents[j];
    const D4 q0 = ld4(g0 + 4 * j), q1 = ld4(g1 + 4 * j);
    D4 lr0 = q0, lr1 = q1;
    if (p >= 0) {
      lr0 = dnormpos(dmul(dconj(ld4(g0 + 4 * p)), q0));
      lr1 = dnormpos(dmul(dconj(ld4(g1 + 4 * p)), q1));
    }
    const Q4 f0 = q4((float)lr0.x, (float)lr0.y, (float)lr0.z, (float)lr0.w);
    const Q4 f1 = q4((float)lr1.x, (float)lr1.y, (float)lr1.z, (float)lr1.w);
    l0[4 * j] = f0.x; l0[4 * j + 1] = f0.y; l0[4 * j + 2] = f0.z; l0[4 * j + 3] = f0.w;
    l1[4 * j] = f1.x; l1[4 * j + 1] = f1.y; l1[4 * j + 2] = f1.z; l1[4 * j + 3] = f1.w;
    float* o = a.lrs + ((size_t)f * J + j) * 4;
    o[0] = f0.x; o[1] = f0.y; o[2] = f0.z; o[3] = f0.w;
    // angular velocity before filtering: quat_angle_axis(quat_mul_norm(r[t+1], conj(r[t]))); last frame = identity -> 0
    D3 av = {0.0, 0.0, 0.0};
    if (has_next) {
      const D4 dq = dnormpos(dmul(q1, dconj(q0)));
      double s = 2.0 * (dq.w * dq.w) - 1.0;
      s = fmin(fmax(s, -1.0), 1.0);
      const double ang = acos(s);
      const double n = fmax(sqrt(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z), 1e-9);
      av.x = (dq.x / n) * ang / dt; av.y = (dq.y / n) * ang / dt; av.z = (dq.z / n) * ang / dt;
    }
    double* ra = a.rawang + ((size_t)f * J + j) * 3;
    ra[0] = av.x; ra[1] = av.y; ra[2] = av.z;
    // dof velocity: float32, phc quat_mul / quat_to_angle_axis; the last frame repeats the previous one
    if (j >= 1 && has_next) {
      const V3 e = quat_to_exp_map(qmul(qconj(f0), f1));
      const float vx = e.x / dt32, vy = e.y / dt32, vz = e.z / dt32;
      float* dv = a.dvs + ((size_t)f * (J - 1) + (j - 1)) * 3;
      dv[0] = vx; dv[1] = vy; dv[2] = vz;
      if (fl + 2 == nf) { dv += (size_t)(J - 1) * 3; dv[0] = vx; dv[1] = vy; dv[2] = vz; }
    }
    if (j >= 1 && nf < 2) {      // degenerate one-frame clip (the reference cannot load it): defined output, zeros
      float* dv = a.dvs + ((size_t)f * (J - 1) + (j - 1)) * 3;
      dv[0] = dv[1] = dv[2] = 0.f;
    }
    max_depth = max(max_depth, depth[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_depth = max(max_depth, __shfl_xor_sync(0xffffffffu, max_depth, o));
  __syncwarp();
  // ---- forward kinematics, level by level (transform_mul chain over the float32-rounded local rotations) ---------------
  const double* tr = a.trans + (size_t)f * 3;
  for (int d = 0; d <= max_depth; ++d) {
    for (int j = lane; j < J; j += 32) {
      if (depth[j] != d) continue;
      const int p = a.parents[j];
      const D4 lr = {(double)l0[4 * j], (double)l0[4 * j + 1], (double)l0[4 * j + 2], (double)l0[4 * j + 3]};
      if (p < 0) {
        st4(ch + 4 * j, lr);
        ps[3 * j] = hc * tr[0] - hs * tr[1];      // trans @ R.T for the rotation about z
        ps[3 * j + 1] = hs * tr[0] + hc * tr[1];
        ps[3 * j + 2] = tr[2];
      } else {
        const D4 cp = ld4(ch + 4 * p);
        st4(ch + 4 * j, dnormpos(dmul(cp, lr)));
        const double* of = a.offsets + ((size_t)c * J + j) * 3;
        D3 ov = {of[0], of[1], of[2]};
        const D3 r = drot(cp, ov);
        ps[3 * j] = r.x + ps[3 * p]; ps[3 * j + 1] = r.y + ps[3 * p + 1]; ps[3 * j + 2] = r.z + ps[3 * p + 2];
      }
    }
    __syncwarp();
  }
  for (int i = lane; i < 3 * J; i += 32) {
    a.pos64[(size_t)f * 3 * J + i] = ps[i];
    a.gts[(size_t)f * 3 * J + i] = (float)ps[i];
  }
}

struct FilterArgs {
  const double* pos64;
  const double* rawang;
  const int32_t* frame_clip;
  const int64_t* starts;
  const int64_t* nframes;
  const double* fps;
  int64_t F;
  int32_t J;
  float *gvs, *gavs;
  Taps taps;
};

__global__ void __launch_bounds__(256) motion_filter_kernel(const FilterArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.F * a.J) return;
  const int64_t f = i / a.J;
  const int j = (int)(i - f * a.J);
  const int c = a.frame_clip[f];
  const int64_t start = a.starts[c], nf = a.nframes[c];
  const int64_t fl = f - start;
  const double dt = 1 / a.fps[c];
  const size_t stride = (size_t)a.J * 3;
  const double* P = a.pos64 + (size_t)start * stride + (size_t)j * 3;
  const double* A = a.rawang + (size_t)start * stride + (size_t)j * 3;
  double v[3] = {0.0, 0.0, 0.0}, w[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
  for (int k = -kRadius; k <= kRadius; ++k) {
    int64_t t = fl + k;
    t = t < 0 ? 0 : (t > nf - 1 ? nf - 1 : t);                 // mode = "nearest"
    const double wk = a.taps.w[k + kRadius];
    // np.gradient (unit spacing): one-sided at the ends, central inside; then / time_delta
    int64_t lo, hi;
    double div = 1.0;
    if (nf < 2) { lo = hi = 0; }                                  // one-frame clip (the reference cannot load it): zero velocity
    else if (t == 0) { lo = 0; hi = 1; }
    else if (t == nf - 1) { lo = nf - 2; hi = nf - 1; }
    else { lo = t - 1; hi = t + 1; div = 2.0; }
    const double* ph = P + (size_t)hi * stride;
    const double* pl = P + (size_t)lo * stride;
    const double* pa = A + (size_t)t * stride;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      v[x] += wk * (((ph[x] - pl[x]) / div) / dt);
      w[x] += wk * pa[x];
    }
  }
  float* gv = a.gvs + (size_t)i * 3;
  float* ga = a.gavs + (size_t)i * 3;
#pragma unroll
  for (int x = 0; x < 3; ++x) { gv[x] = (float)v[x]; ga[x] = (float)w[x]; }
}

}  // namespace load
}  // namespace phc

extern "C" int64_t phc_motion_load_workspace_bytes(int64_t num_frames_total, int32_t num_bodies) {
  if (num_frames_total < 0 || num_bodies < 1) return 0;
  const int64_t bf = num_frames_total * num_bodies;
  return bf * 6 * (int64_t)sizeof(double) + ((num_frames_total * (int64_t)sizeof(int32_t) + 15) & ~(int64_t)15);
}

extern "C" int phc_motion_load(const double* pose_quat_global, const double* root_trans, const double* offsets,
                               const int32_t* parents, const double* heading, const int64_t* length_starts,
                               const int64_t* num_frames, const double* fps, int64_t num_frames_total, int32_t num_motions,
                               int32_t num_bodies, float* gts, float* grs, float* lrs, float* gvs, float* gavs, float* dvs,
                               void* workspace, void* stream) {
  using namespace phc::load;
  if (num_frames_total == 0) return PHC_OK;
  if (!pose_quat_global || !root_trans || !offsets || !parents || !length_starts || !num_frames || !fps || !gts || !grs ||
      !lrs || !gvs || !gavs || !dvs || !workspace) {
    phc_set_error("phc_motion_load: a required pointer is NULL (only heading may be NULL)");
    return PHC_ERR_INVALID_ARG;
  }
  if (num_frames_total < 0 || num_motions < 1 || num_bodies < 2) {
    phc_set_error("phc_motion_load: bad sizes (num_frames_total >= 0, num_motions >= 1, num_bodies >= 2)");
    return PHC_ERR_INVALID_ARG;
  }
  if (num_bodies > PHC_LOAD_MAX_BODIES) {
    phc_set_error("phc_motion_load: more than PHC_LOAD_MAX_BODIES bodies");
    return PHC_ERR_UNSUPPORTED;
  }
  if (reinterpret_cast<uintptr_t>(workspace) & 15) { phc_set_error("phc_motion_load: workspace must be 16-byte aligned"); return PHC_ERR_INVALID_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int J = num_bodies;
  const int64_t bf = num_frames_total * J;
  double* pos64 = static_cast<double*>(workspace);
  double* rawang = pos64 + bf * 3;
  int32_t* frame_clip = reinterpret_cast<int32_t*>(rawang + bf * 3);

  FkArgs fa;
  fa.quat = pose_quat_global; fa.trans = root_trans; fa.offsets = offsets; fa.parents = parents; fa.heading = heading;
  fa.starts = length_starts; fa.nframes = num_frames; fa.fps = fps; fa.F = num_frames_total; fa.M = num_motions; fa.J = J;
  fa.gts = gts; fa.grs = grs; fa.lrs = lrs; fa.dvs = dvs; fa.pos64 = pos64; fa.rawang = rawang; fa.frame_clip = frame_clip;
  const size_t per_warp = ((size_t)J * (15 * sizeof(double) + 8 * sizeof(float) + sizeof(int)) + 15) & ~(size_t)15;
  const size_t smem = per_warp * kWarps;        // J = 64: 39.9 KB, under the 48 KB default limit
  const int64_t grid1 = (num_frames_total + kWarps - 1) / kWarps;
  if (grid1 > 0x7fffffff) { phc_set_error("phc_motion_load: too many frames for one launch"); return PHC_ERR_UNSUPPORTED; }
  motion_fk_kernel<<<(unsigned)grid1, kWarps * 32, smem, st>>>(fa);
  phc_count_launches(1);
  int rc = phc_check_cuda(cudaGetLastError(), "motion_fk_kernel launch");
  if (rc) return rc;

  FilterArgs fl;
  fl.pos64 = pos64; fl.rawang = rawang; fl.frame_clip = frame_clip; fl.starts = length_starts; fl.nframes = num_frames;
  fl.fps = fps; fl.F = num_frames_total; fl.J = J; fl.gvs = gvs; fl.gavs = gavs;
  {   // scipy.ndimage._filters._gaussian_kernel1d(sigma = 2, order = 0, radius = 8)
    double sum = 0.0;
    for (int k = -kRadius; k <= kRadius; ++k) { fl.taps.w[k + kRadius] = exp(-0.5 / (2.0 * 2.0) * (double)k * (double)k); sum += fl.taps.w[k + kRadius]; }
    for (int k = 0; k <= 2 * kRadius; ++k) fl.taps.w[k] /= sum;
  }
  const int64_t grid2 = (bf + 255) / 256;
  if (grid2 > 0x7fffffff) { phc_set_error("phc_motion_load: too many frames for one launch"); return PHC_ERR_UNSUPPORTED; }
  motion_filter_kernel<<<(unsigned)grid2, 256, 0, st>>>(fl);
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "motion_filter_kernel launch");
}
