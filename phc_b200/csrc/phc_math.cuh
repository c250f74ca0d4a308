// Quaternion / exp-map arithmetic of the PHC hot path, one scalar op per reference tensor op.
//
// Reference: phc/utils/isaacgym_torch_utils.py (quat_mul :25-45, normalize :49-50, quat_conjugate :93-96,
// quat_from_angle_axis :104-108, normalize_angle :111-112) and phc/utils/torch_utils.py (my_quat_rotate :46-55,
// quat_to_angle_axis :58-78, quat_to_tan_norm :101-113, exp_map_to_angle_axis :147-166, slerp :176-197,
// calc_heading* :200-240).  Quaternions are xyzw.
//
// Numerics contract: every expression keeps the reference's operation ORDER, IEEE div/sqrt and the accurate libm entry
// points (acosf, atan2f, sinf, cosf -- never the __ fast intrinsics).  motion.cu / ppo_scalars.cu are compiled with
// -fmad=false (no FMA contraction at all).  env_step.cu allows contraction in the well-conditioned arithmetic (quaternion
// products, rotations, lerps: an FMA moves a result by <= 1 ulp of its operands, ~1e-7 of an O(1) value) and pins the
// ill-conditioned spots to the reference's separately rounded mul / add with the PHC_MUL / PHC_ADD / PHC_SUB intrinsics,
// which the compiler never contracts: the frame-bracket / blend arithmetic (a half-ulp change of i0*dt moves the blend by
// 1e-5), slerp's dot product and 1 - c*c (sin(acos c) cancels catastrophically for neighbouring frames: the weights would
// move at the 1e-4 level) and the 1 - w*w of the angle extraction.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PHC_HD __host__ __device__ __forceinline__
#else
#define PHC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PHC_MUL(a, b) __fmul_rn((a), (b))
#define PHC_ADD(a, b) __fadd_rn((a), (b))
#define PHC_SUB(a, b) __fsub_rn((a), (b))
#else
#define PHC_MUL(a, b) ((a) * (b))
#define PHC_ADD(a, b) ((a) + (b))
#define PHC_SUB(a, b) ((a) - (b))
#endif

namespace phc {

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };

PHC_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PHC_HD Q4 q4(float x, float y, float z, float w) { Q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PHC_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
PHC_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }

// (1-b)*a0 + b*a1, the reference's lerp expression (motion_lib_base.py:474-480)
// (pinned: the interpolated reference pose is bit-identical at every call site, in every kernel, cached or recomputed)
PHC_HD float lerp1(float a0, float a1, float omb, float b) { return PHC_ADD(PHC_MUL(omb, a0), PHC_MUL(b, a1)); }
PHC_HD V3 lerp3(V3 a0, V3 a1, float omb, float b) {
  return v3(lerp1(a0.x, a1.x, omb, b), lerp1(a0.y, a1.y, omb, b), lerp1(a0.z, a1.z, omb, b));
}

PHC_HD Q4 qmul(Q4 a, Q4 b) {
  const float ww = (a.z + a.x) * (b.x + b.y);
  const float yy = (a.w - a.y) * (b.w + b.z);
  const float zz = (a.w + a.y) * (b.w - b.z);
  const float xx = ww + yy + zz;
  const float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
  Q4 r;
  r.w = qq - ww + (a.z - a.y) * (b.y - b.z);
  r.x = qq - xx + (a.x + a.w) * (b.x + b.w);
  r.y = qq - yy + (a.w - a.x) * (b.y + b.z);
  r.z = qq - zz + (a.z + a.y) * (b.w - b.x);
  return r;
}

PHC_HD Q4 qconj(Q4 q) { return q4(-q.x, -q.y, -q.z, q.w); }

// qmul with a quaternion about z on the left (a.x == a.y == 0: the heading quaternions).  Every "+ 0" / "- 0" of the general
// expression is dropped (x + 0 == x exactly), so for finite inputs the result equals qmul(a, b) operation by operation.
PHC_HD Q4 qmul_zl(Q4 a, Q4 b) {
  const float ww = a.z * (b.x + b.y);
  const float yy = a.w * (b.w + b.z);
  const float zz = a.w * (b.w - b.z);
  const float xx = ww + yy + zz;
  const float qq = 0.5f * (xx + a.z * (b.x - b.y));
  Q4 r;
  r.w = qq - ww + a.z * (b.y - b.z);
  r.x = qq - xx + a.w * (b.x + b.w);
  r.y = qq - yy + a.w * (b.y + b.z);
  r.z = qq - zz + a.z * (b.w - b.x);
  return r;
}

// ... and on the right (b.x == b.y == 0): ww = (a.z + a.x) * 0 and (a.z - a.x) * 0 vanish, b.x + b.w == b.w etc.
PHC_HD Q4 qmul_zr(Q4 a, Q4 b) {
  const float yy = (a.w - a.y) * (b.w + b.z);
  const float zz = (a.w + a.y) * (b.w - b.z);
  const float xx = yy + zz;
  const float qq = 0.5f * xx;
  Q4 r;
  r.w = qq + (a.z - a.y) * (-b.z);
  r.x = qq - xx + (a.x + a.w) * b.w;
  r.y = qq - yy + (a.w - a.x) * b.z;
  r.z = qq - zz + (a.z + a.y) * b.w;
  return r;
}

// v*(2w^2-1) + (qv x v)*w*2 + qv*(qv.v)*2
PHC_HD V3 qrot(Q4 q, V3 v) {
  const float s = 2.0f * (q.w * q.w) - 1.0f;
  const float cx = q.y * v.z - q.z * v.y;
  const float cy = q.z * v.x - q.x * v.z;
  const float cz = q.x * v.y - q.y * v.x;
  const float d = q.x * v.x + q.y * v.y + q.z * v.z;
  V3 r;
  r.x = (v.x * s + cx * q.w * 2.0f) + q.x * d * 2.0f;
  r.y = (v.y * s + cy * q.w * 2.0f) + q.y * d * 2.0f;
  r.z = (v.z * s + cz * q.w * 2.0f) + q.z * d * 2.0f;
  return r;
}

// heading = atan2 of the rotated x axis; quat_from_angle_axis(+-heading, z) incl. its final re-normalisation
PHC_HD float heading_angle(Q4 q) {
  const float s = 2.0f * (q.w * q.w) - 1.0f;            // x and y of qrot(q, (1,0,0)), zero terms folded
  const float dx = s + q.x * q.x * 2.0f;
  const float dy = q.z * q.w * 2.0f + q.y * q.x * 2.0f;
  return atan2f(dy, dx);
}

PHC_HD void sin_cos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);          // one shared range reduction, same 2-ulp accuracy class as sinf / cosf
#else
  *s = sinf(x); *c = cosf(x);
#endif
}

PHC_HD Q4 quat_about_z(float angle) {
  const float th = angle / 2.0f;
  float s, c;
  sin_cos(th, &s, &c);
  float n = sqrtf(s * s + c * c);
  n = n < 1e-9f ? 1e-9f : n;
  const float inv = 1.0f / n;               // one IEEE division shared by both components (n == 1 up to rounding)
  return q4(0.0f, 0.0f, s * inv, c * inv);
}

// qrot(h, v) for a quaternion about z (h.x == h.y == 0: the heading quaternions).  Every product with an exact zero and
// every "+ 0" of the general expression is dropped; for finite inputs the result is bit-identical to qrot (only the
// sign of a zero result can differ), at half the instructions.  The compiler may not do this itself under IEEE rules.
PHC_HD V3 qrot_z(Q4 h, V3 v) {
  const float s = 2.0f * (h.w * h.w) - 1.0f;
  const float cx = -(h.z * v.y);
  const float cy = h.z * v.x;
  const float d = h.z * v.z;
  V3 r;
  r.x = v.x * s + cx * h.w * 2.0f;
  r.y = v.y * s + cy * h.w * 2.0f;
  r.z = v.z * s + h.z * d * 2.0f;
  return r;
}

// quat_to_tan_norm: q * (1,0,0) and q * (0,0,1) with the same zero-folding (bit-identical to two general qrot calls)
struct TanNorm { V3 t, n; };
PHC_HD TanNorm tan_norm(Q4 q) {
  const float s = 2.0f * (q.w * q.w) - 1.0f;
  TanNorm r;
  r.t.x = s + q.x * q.x * 2.0f;
  r.t.y = q.z * q.w * 2.0f + q.y * q.x * 2.0f;
  r.t.z = -(q.y * q.w) * 2.0f + q.z * q.x * 2.0f;
  r.n.x = q.y * q.w * 2.0f + q.x * q.z * 2.0f;
  r.n.y = -(q.x * q.w) * 2.0f + q.y * q.z * 2.0f;
  r.n.z = s + q.z * q.z * 2.0f;
  return r;
}

PHC_HD float wrap_angle(float x) { return atan2f(sinf(x), cosf(x)); }

// normalize_angle for an argument already known to lie in [0, 2*pi]: atan2(sin x, cos x) = x (x <= pi) or x - 2*pi.
// Used on the per-step path (2*acos(w) and joint angles), where it replaces sinf + cosf + atan2f (~100 instructions per
// lane) by one compare; it agrees with wrap_angle to fp32 rounding of the result (the reference's own atan2(sin, cos)
// carries the same ~1e-7 relative rounding).
PHC_HD float wrap_angle_0_2pi(float x) { return x > 3.14159265358979323846f ? x - 6.28318530717958647692f : x; }

// quat_to_angle_axis: angle only (what the tracking reward reads) ...
PHC_HD float quat_angle(Q4 q) {
  const float s = sqrtf(PHC_SUB(1.0f, PHC_MUL(q.w, q.w)));
  const float ang = wrap_angle_0_2pi(2.0f * acosf(q.w));
  return (fabsf(s) > 1e-5f) ? ang : 0.0f;        // NaN (|w|>1) compares false -> 0, like torch.where(mask, ...)
}

// ... and the full exp map angle*axis (dof_pos of the reference pose)
PHC_HD V3 quat_to_exp_map(Q4 q) {
  const float s = sqrtf(PHC_SUB(1.0f, PHC_MUL(q.w, q.w)));
  const float ang = wrap_angle(2.0f * acosf(q.w));
  if (fabsf(s) > 1e-5f) return v3(ang * (q.x / s), ang * (q.y / s), ang * (q.z / s));
  return v3(0.0f * 0.0f, 0.0f * 0.0f, 0.0f * 1.0f);
}

// exp_map_to_quat (exp_map_to_angle_axis then quat_from_angle_axis with both normalisations)
PHC_HD Q4 exp_map_to_quat(V3 e) {
  const float n0 = sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
  const float inv0 = 1.0f / n0;             // shared reciprocals: 3 IEEE divisions in this function instead of 10
  V3 ax = v3(e.x * inv0, e.y * inv0, e.z * inv0);
  float ang = (n0 <= 6.28318530717958647692f) ? wrap_angle_0_2pi(n0) : wrap_angle(n0);   // joint angles are < 2*pi
  if (!(fabsf(ang) > 1e-5f)) { ang = 0.0f; ax = v3(0.0f, 0.0f, 1.0f); }
  float an = sqrtf(ax.x * ax.x + ax.y * ax.y + ax.z * ax.z);
  an = an < 1e-9f ? 1e-9f : an;
  const float th = ang / 2.0f;
  float s, c;
  sin_cos(th, &s, &c);
  const float inva = 1.0f / an;
  const float x = (ax.x * inva) * s, y = (ax.y * inva) * s, z = (ax.z * inva) * s;
  float qn = sqrtf(x * x + y * y + z * z + c * c);
  qn = qn < 1e-9f ? 1e-9f : qn;
  const float invq = 1.0f / qn;
  return q4(x * invq, y * invq, z * invq, c * invq);
}

PHC_HD Q4 slerp(Q4 q0, Q4 q1, float t) {
  float c = PHC_ADD(PHC_ADD(PHC_ADD(PHC_MUL(q0.x, q1.x), PHC_MUL(q0.y, q1.y)), PHC_MUL(q0.z, q1.z)), PHC_MUL(q0.w, q1.w));
  if (c < 0.0f) q1 = q4(-q1.x, -q1.y, -q1.z, -q1.w);
  c = fabsf(c);
  const float half = acosf(c);
  const float s = sqrtf(PHC_SUB(1.0f, PHC_MUL(c, c)));
  const float inv_s = 1.0f / s;             // one IEEE division for both weights
  const float ra = sinf((1.0f - t) * half) * inv_s;
  const float rb = sinf(t * half) * inv_s;
  Q4 r = q4(lerp1(q0.x, q1.x, ra, rb), lerp1(q0.y, q1.y, ra, rb), lerp1(q0.z, q1.z, ra, rb), lerp1(q0.w, q1.w, ra, rb));
  if (fabsf(s) < 0.001f) r = q4(lerp1(q0.x, q1.x, 0.5f, 0.5f), lerp1(q0.y, q1.y, 0.5f, 0.5f), lerp1(q0.z, q1.z, 0.5f, 0.5f), lerp1(q0.w, q1.w, 0.5f, 0.5f));
  if (fabsf(c) >= 1.0f) r = q0;
  return r;
}

// remove_base_rot (humanoid.py:1935-1939): q * conj([.5,.5,.5,.5])
PHC_HD Q4 strip_base_rot(Q4 q) { return qmul(q, q4(-0.5f, -0.5f, -0.5f, 0.5f)); }

// _calc_frame_blend (motion_lib_base.py:549-559): fp32 order time/len -> clip -> *(nf-1) -> trunc
struct Bracket { int64_t i0, i1; float blend; };
PHC_HD Bracket frame_bracket(float time, float len, int64_t nf, float dt) {
  float phase = time / len;
  phase = fminf(fmaxf(phase, 0.0f), 1.0f);
  if (time < 0.0f) time = 0.0f;
  Bracket b;
  b.i0 = (int64_t)(phase * (float)(nf - 1));
  b.i1 = (b.i0 + 1 < nf - 1) ? b.i0 + 1 : nf - 1;
  float bl = PHC_SUB(time, PHC_MUL((float)b.i0, dt)) / dt;
  b.blend = fminf(fmaxf(bl, 0.0f), 1.0f);
  return b;
}

// Same bracket with 32-bit frame indices (every clip has far fewer than 2^31 frames): avoids the emulated 64-bit
// float<->int conversions on the per-step path.  Identical results.
struct Bracket32 { int i0, i1; float blend; };
PHC_HD Bracket32 frame_bracket32(float time, float len, int nf, float dt) {
  float phase = time / len;
  phase = fminf(fmaxf(phase, 0.0f), 1.0f);
  if (time < 0.0f) time = 0.0f;
  Bracket32 b;
  b.i0 = (int)(phase * (float)(nf - 1));
  b.i1 = (b.i0 + 1 < nf - 1) ? b.i0 + 1 : nf - 1;
  float bl = PHC_SUB(time, PHC_MUL((float)b.i0, dt)) / dt;
  b.blend = fminf(fmaxf(bl, 0.0f), 1.0f);
  return b;
}

}  // namespace phc
