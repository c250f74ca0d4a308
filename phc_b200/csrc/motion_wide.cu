// Motion-library queries for characters with more than 32 bodies / hinge dofs (Unitree G1: 38 + 1 bodies, 37 dofs; SMPL-X: 52
// bodies): the strided counterparts of motion.cu's motion_state / amp_demo / set_env_state kernels.  One warp per query as
// there, but the bodies (and hinge dofs) are strided over the lanes (j = lane, lane + 32, ...) and the root record every lane
// needs for the AMP observation is sampled by the lane itself instead of being broadcast -- no warp collectives at all.
// The per-body arithmetic is the shared sample_motion() of motion_sample.cuh.  Entry points: phc_motion_state /
// phc_amp_obs_demo / phc_set_env_state dispatch here when J + E > 32 or D > 32.
//
// STATUS: validated against the goldens of the unmodified reference (g1.npz, h1.npz, motion.npz, envstep.npz demo, reset.npz)
// through the CPU emulation of this very source (tests/test_motion_emu_cpu.py); not run on a GPU yet (GPU tests opt-in:
// PHC_TEST_WIDE=1).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/phc_b200.h"
#include "phc_common.cuh"
#include "phc_math.cuh"
#include "motion_sample.cuh"

namespace phc {
namespace wide {

__global__ void __launch_bounds__(128)
motion_state_wide_kernel(const __grid_constant__ PhcMotionLib lib, const int64_t* __restrict__ ids, const float* __restrict__ times,
                         const float* __restrict__ offset, int64_t n, const __grid_constant__ PhcMotionStateOut out) {
  const int64_t qi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qi >= n) return;
  const int J = lib.num_bodies, JE = J + lib.num_ext_bodies, D = lib.num_dofs;
  const V3 off = offset ? v3(offset[3 * qi], offset[3 * qi + 1], offset[3 * qi + 2]) : v3(0.f, 0.f, 0.f);
  const bool want_joint = (out.dof_pos != nullptr) || (out.dof_vel != nullptr);
  const int last = JE > D ? JE : D;
  for (int jj = lane; jj < last; jj += 32) {
    MotionSample s = sample_motion(lib, ids[qi], times[qi], v3(0.f, 0.f, 0.f), jj, want_joint);
    if (offset) s.body.p = s.body.p + off;          // the reference adds the offset only when one is given
    if (jj < JE) {                                   // robots: all J + E bodies (the *_t outputs)
      const int64_t bt = qi * JE + jj;
      if (out.rg_pos_t) st3g(out.rg_pos_t + 3 * bt, s.body.p);
      if (out.rg_rot_t) st4g(out.rg_rot_t + 4 * bt, s.body.q);
      if (out.body_vel_t) st3g(out.body_vel_t + 3 * bt, s.body.v);
      if (out.body_ang_vel_t) st3g(out.body_ang_vel_t + 3 * bt, s.body.w);
    }
    if (jj < J) {
      const int64_t bj = qi * J + jj;
      if (out.rg_pos) st3g(out.rg_pos + 3 * bj, s.body.p);
      if (out.rb_rot) st4g(out.rb_rot + 4 * bj, s.body.q);
      if (out.body_vel) st3g(out.body_vel + 3 * bj, s.body.v);
      if (out.body_ang_vel) st3g(out.body_ang_vel + 3 * bj, s.body.w);
    }
    if (D > 0) {
      if (jj < D) {
        if (out.dof_pos) out.dof_pos[qi * D + jj] = s.dof_pos.x;
        if (out.dof_vel) out.dof_vel[qi * D + jj] = s.dof_vel.x;
      }
    } else if (jj > 0 && jj < J) {
      const int64_t dj = qi * (J - 1) + (jj - 1);
      if (out.dof_pos) st3g(out.dof_pos + 3 * dj, s.dof_pos);
      if (out.dof_vel) st3g(out.dof_vel + 3 * dj, s.dof_vel);
    }
    if (jj == 0) {
      if (out.root_pos) st3g(out.root_pos + 3 * qi, s.body.p);
      if (out.root_rot) st4g(out.root_rot + 4 * qi, s.body.q);
      if (out.root_vel) st3g(out.root_vel + 3 * qi, s.body.v);
      if (out.root_ang_vel) st3g(out.root_ang_vel + 3 * qi, s.body.w);
    }
  }
}

struct AmpDemoWideArgs {
  PhcMotionLib lib;
  const int64_t* ids;
  const float* times0;
  int64_t n;
  int32_t first_step, num_steps;
  float dt;
  uint32_t flags;
  int32_t key_bodies[PHC_MAX_KEY_BODIES];
  int32_t num_key_bodies;
  int32_t amp_joints[PHC_MAX_AMP_JOINTS];
  int32_t num_amp_joints;
  float* out;
  int64_t out_stride;
  const int64_t* only_where;
  int32_t slot_offset;
  const int32_t* slot_offset_dev;
};

// warp per (sample, history step): motion sample at t0 - (first_step + k) dt, then build_amp_observations_smpl / _robot
__global__ void __launch_bounds__(128) amp_demo_wide_kernel(const __grid_constant__ AmpDemoWideArgs a) {
  const int64_t wi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wi >= a.n * a.num_steps) return;
  const int64_t si = wi / a.num_steps;
  if (a.only_where && a.only_where[si] == 0) return;
  const int k = (int)(wi - si * a.num_steps);
  const int J = a.lib.num_bodies, D = a.lib.num_dofs;
  const float t = a.times0[si] + (-a.dt * (float)(k + a.first_step));      // humanoid_amp.py:257-261 / :577-582
  const BodyS r = sample_motion(a.lib, a.ids[si], t, v3(0.f, 0.f, 0.f), 0, false).body;     // the root, in every lane

  const bool upright = a.flags & PHC_FLAG_UPRIGHT, has_h = a.flags & PHC_FLAG_ROOT_HEIGHT_OBS;
  const Q4 root_q = upright ? r.q : strip_base_rot(r.q);
  const Q4 hinv = quat_about_z(-heading_angle(root_q));
  const int nj = a.num_amp_joints, nk = a.num_key_bodies;
  const int kp = (k + (a.slot_offset_dev ? *a.slot_offset_dev : a.slot_offset)) % a.num_steps;
  const int row = has_h + 12 + (D > 0 ? 2 * D : 9 * nj) + 3 * nk;
  float* o = a.out + si * a.out_stride + (int64_t)kp * row + (has_h ? 1 : 0);
  if (lane == 0) {
    if (has_h) o[-1] = r.p.z;
    const TanNorm tn = tan_norm((a.flags & PHC_FLAG_LOCAL_ROOT_OBS) ? qmul(hinv, root_q) : root_q);
    st3g(o, tn.t); st3g(o + 3, tn.n);
    st3g(o + 6, qrot_z(hinv, r.v));
    st3g(o + 9, qrot_z(hinv, r.w));
  }
  const int last = J > D ? J : D;
  for (int jj = lane; jj < last; jj += 32) {
    const MotionSample s = sample_motion(a.lib, a.ids[si], t, v3(0.f, 0.f, 0.f), jj, true);
    if (D > 0) {       // build_amp_observations_robot (humanoid_amp.py:1062-1104): raw hinge angles and velocities
      if (jj < D) { o[12 + jj] = s.dof_pos.x; o[12 + D + jj] = s.dof_vel.x; }
      if (jj < J)
        for (int kk = 0; kk < nk; ++kk)
          if (a.key_bodies[kk] == jj) st3g(o + 12 + 2 * D + 3 * kk, qrot_z(hinv, s.body.p - r.p));
      continue;
    }
    if (jj >= J) continue;
    if (jj > 0)
      for (int kk = 0; kk < nj; ++kk)
        if (a.amp_joints[kk] == jj - 1) {
          const TanNorm tn = tan_norm(exp_map_to_quat(s.dof_pos));
          st3g(o + 12 + 6 * kk, tn.t); st3g(o + 12 + 6 * kk + 3, tn.n);
          st3g(o + 12 + 6 * nj + 3 * kk, s.dof_vel);
        }
    for (int kk = 0; kk < nk; ++kk)
      if (a.key_bodies[kk] == jj) st3g(o + 12 + 9 * nj + 3 * kk, qrot_z(hinv, s.body.p - r.p));
  }
}

// HumanoidAMP._set_env_state (humanoid_amp.py:605-637) for the envs with mask != 0: rigid-body rows + dof pos / vel
__global__ void __launch_bounds__(128)
set_env_state_wide_kernel(const __grid_constant__ PhcMotionLib lib, const int64_t* __restrict__ ids, const float* __restrict__ times,
                          const float* __restrict__ offset, const int64_t* __restrict__ only_where, int64_t n,
                          float* __restrict__ body_state, int bpe, float* __restrict__ dof_state) {
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (env >= n) return;
  if (only_where && only_where[env] == 0) return;
  const int J = lib.num_bodies, D = lib.num_dofs;
  const V3 off = offset ? v3(offset[3 * env], offset[3 * env + 1], offset[3 * env + 2]) : v3(0.f, 0.f, 0.f);
  const int last = J > D ? J : D;
  for (int jj = lane; jj < last; jj += 32) {
    MotionSample s = sample_motion(lib, ids[env], times[env], v3(0.f, 0.f, 0.f), jj, dof_state != nullptr);
    if (offset) s.body.p = s.body.p + off;
    if (D > 0 && dof_state && jj < D) {                       // hinge joints: [D, 2] (pos, vel)
      float* d = dof_state + ((size_t)env * D + jj) * 2;
      d[0] = s.dof_pos.x; d[1] = s.dof_vel.x;
    }
    if (jj >= J) continue;
    float* o = body_state + ((size_t)env * bpe + jj) * kRec;
    o[0] = s.body.p.x; o[1] = s.body.p.y; o[2] = s.body.p.z;
    o[3] = s.body.q.x; o[4] = s.body.q.y; o[5] = s.body.q.z; o[6] = s.body.q.w;
    o[7] = s.body.v.x; o[8] = s.body.v.y; o[9] = s.body.v.z;
    o[10] = s.body.w.x; o[11] = s.body.w.y; o[12] = s.body.w.z;
    if (D == 0 && dof_state && jj > 0) {
      float* d = dof_state + ((size_t)env * (J - 1) + (jj - 1)) * 6;      // [D, 2] interleaved (pos, vel)
      d[0] = s.dof_pos.x; d[1] = s.dof_vel.x; d[2] = s.dof_pos.y; d[3] = s.dof_vel.y; d[4] = s.dof_pos.z; d[5] = s.dof_vel.z;
    }
  }
}

}  // namespace wide
}  // namespace phc

extern "C" int phc_check_cuda(cudaError_t e, const char* what);
extern "C" void phc_count_launches(int n);

// called by the entry points of motion.cu after their argument validation
extern "C" int phc_motion_state_wide_launch(const PhcMotionLib* lib, const int64_t* ids, const float* times, const float* offset,
                                            int64_t n, const PhcMotionStateOut* out, void* stream) {
  const int wpb = 4;
  phc::wide::motion_state_wide_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(*lib, ids, times, offset, n, *out);
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "motion_state_wide_kernel launch");
}

extern "C" int phc_amp_obs_demo_wide_launch(const PhcMotionLib* lib, const int64_t* ids, const float* times0, int64_t n, int32_t first_step,
                                            int32_t num_steps, float dt, uint32_t flags, const int32_t* key_bodies, int32_t nk,
                                            const int32_t* amp_joints, int32_t nj, float* out, int64_t out_stride,
                                            const int64_t* only_where, int32_t slot_offset, const int32_t* slot_offset_dev, void* stream) {
  phc::wide::AmpDemoWideArgs a;
  a.lib = *lib; a.ids = ids; a.times0 = times0; a.n = n; a.first_step = first_step; a.num_steps = num_steps; a.dt = dt;
  a.flags = flags; a.num_key_bodies = nk; a.num_amp_joints = nj; a.out = out; a.out_stride = out_stride; a.only_where = only_where;
  for (int i = 0; i < PHC_MAX_AMP_JOINTS; ++i) a.amp_joints[i] = i < nj ? amp_joints[i] : -1;
  for (int i = 0; i < PHC_MAX_KEY_BODIES; ++i) a.key_bodies[i] = i < nk ? key_bodies[i] : -1;
  a.slot_offset = ((slot_offset % num_steps) + num_steps) % num_steps;
  a.slot_offset_dev = slot_offset_dev;
  const int wpb = 4;
  const int64_t warps = n * num_steps;
  phc::wide::amp_demo_wide_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "amp_demo_wide_kernel launch");
}

extern "C" int phc_set_env_state_wide_launch(const PhcMotionLib* lib, const int64_t* ids, const float* times, const float* offset,
                                             const int64_t* only_where, int64_t n, float* body_state, int32_t bodies_per_env,
                                             float* dof_state, void* stream) {
  const int wpb = 4;
  phc::wide::set_env_state_wide_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      *lib, ids, times, offset, only_where, n, body_state, bodies_per_env, dof_state);
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "set_env_state_wide_kernel launch");
}
