// The reference-motion sample shared by the motion kernels (motion.cu: one body per lane; motion_wide.cu: bodies strided over
// the lanes for more than 32 bodies): frame bracket, lerp / slerp of one body record, the joint (dof) part of joint j.
// Reference: phc/utils/motion_lib_base.py:437-520 (get_motion_state), phc/utils/motion_lib_real.py:236-361 (hinge joints).
#pragma once
#include <stdint.h>

#include "../../include/phc_b200.h"
#include "phc_math.cuh"

namespace phc {

constexpr int kRec = 13;

struct BodyS { V3 p; Q4 q; V3 v; V3 w; };
__device__ __forceinline__ BodyS ld_body(const float* s) {
  BodyS b;
  b.p = v3(s[0], s[1], s[2]); b.q = q4(s[3], s[4], s[5], s[6]); b.v = v3(s[7], s[8], s[9]); b.w = v3(s[10], s[11], s[12]);
  return b;
}

// One reference-motion sample for lane `j`: blended body record + joint (dof) position/velocity of joint j-1.
struct MotionSample { BodyS body; V3 dof_pos; V3 dof_vel; };

__device__ __forceinline__ MotionSample sample_motion(const PhcMotionLib& lib, int64_t mid, float time, V3 off, int j,
                                                      bool want_joint) {
  const Bracket b = frame_bracket(time, lib.motion_len[mid], lib.motion_num_frames[mid], lib.motion_dt[mid]);
  const int64_t r0 = lib.length_starts[mid] + b.i0, r1 = lib.length_starts[mid] + b.i1;
  const float bl = b.blend, omb = 1.0f - bl;
  const int jb = j < lib.num_bodies + lib.num_ext_bodies ? j : 0;     // lanes that only carry a dof read body 0 (unused)
  const BodyS a0 = ld_body(lib.frames_body + r0 * lib.body_stride + jb * kRec);
  const BodyS a1 = ld_body(lib.frames_body + r1 * lib.body_stride + jb * kRec);
  MotionSample s;
  s.body.p = lerp3(a0.p, a1.p, omb, bl) + off;
  s.body.v = lerp3(a0.v, a1.v, omb, bl);
  s.body.w = lerp3(a0.w, a1.w, omb, bl);
  s.body.q = slerp(a0.q, a1.q, bl);
  s.dof_pos = v3(0.f, 0.f, 0.f);
  s.dof_vel = v3(0.f, 0.f, 0.f);
  if (want_joint && lib.frames_joint && lib.num_dofs > 0) {
    // hinge-joint robot: lane j carries dof j; dof_pos and dof_vel are both interpolated linearly (motion_lib_real.py:283-285)
    if (j < lib.num_dofs) {
      const float* j0 = lib.frames_joint + r0 * lib.joint_stride;
      const float* j1 = lib.frames_joint + r1 * lib.joint_stride;
      s.dof_pos.x = lerp1(j0[j], j1[j], omb, bl);
      s.dof_vel.x = lerp1(j0[lib.num_dofs + j], j1[lib.num_dofs + j], omb, bl);
    }
  } else if (want_joint && lib.frames_joint) {
    const int J = lib.num_bodies;
    const float* j0 = lib.frames_joint + r0 * lib.joint_stride;
    const float* j1 = lib.frames_joint + r1 * lib.joint_stride;
    const Q4 l0 = q4(j0[4 * j], j0[4 * j + 1], j0[4 * j + 2], j0[4 * j + 3]);
    const Q4 l1 = q4(j1[4 * j], j1[4 * j + 1], j1[4 * j + 2], j1[4 * j + 3]);
    const Q4 lq = slerp(l0, l1, bl);
    if (j > 0) {
      s.dof_pos = quat_to_exp_map(lq);
      const float* d0 = j0 + 4 * J + 3 * (j - 1);
      const float* d1 = j1 + 4 * J + 3 * (j - 1);
      s.dof_vel = lerp3(v3(d0[0], d0[1], d0[2]), v3(d1[0], d1[1], d1[2]), omb, bl);
    }
  }
  return s;
}

__device__ __forceinline__ void st3g(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
__device__ __forceinline__ void st4g(float* d, Q4 q) { d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w; }

}  // namespace phc
