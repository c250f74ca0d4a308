// TEST INFRASTRUCTURE: compiles the kernels' own arithmetic header (phc_b200/csrc/phc_math.cuh -- every function is
// PHC_HD = host + device) with g++ and exposes it through a C ABI, so tests/test_math_header_cpu.py can pin the very source
// the CUDA kernels are built from against the goldens of the unmodified reference WITHOUT a GPU.  Built with
// -ffp-contract=off: the host then evaluates each expression with separately rounded operations, i.e. the reference's
// operation order; the device build may additionally contract well-conditioned mul+add pairs (phc_math.cuh header).
// Not part of the product library.
#include <cstdint>

#include "../phc_b200/csrc/phc_math.cuh"

using namespace phc;

extern "C" {

void h_qmul(const float* a, const float* b, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const Q4 r = qmul(q4(a[4*i], a[4*i+1], a[4*i+2], a[4*i+3]), q4(b[4*i], b[4*i+1], b[4*i+2], b[4*i+3])); o[4*i] = r.x; o[4*i+1] = r.y; o[4*i+2] = r.z; o[4*i+3] = r.w; }
}
// the two zero-folded products against a general operand: out_l = qmul_zl(z(a), b), out_r = qmul_zr(b, z(a)), and the general
// qmul on the same (zeroed) operands, z(a) = (0, 0, a.z, a.w)
void h_qmul_z(const float* a, const float* b, float* out_l, float* ref_l, float* out_r, float* ref_r, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const Q4 z = q4(0.0f, 0.0f, a[4*i+2], a[4*i+3]), g = q4(b[4*i], b[4*i+1], b[4*i+2], b[4*i+3]);
    const Q4 l = qmul_zl(z, g), lr = qmul(z, g), r = qmul_zr(g, z), rr = qmul(g, z);
    out_l[4*i] = l.x; out_l[4*i+1] = l.y; out_l[4*i+2] = l.z; out_l[4*i+3] = l.w;
    ref_l[4*i] = lr.x; ref_l[4*i+1] = lr.y; ref_l[4*i+2] = lr.z; ref_l[4*i+3] = lr.w;
    out_r[4*i] = r.x; out_r[4*i+1] = r.y; out_r[4*i+2] = r.z; out_r[4*i+3] = r.w;
    ref_r[4*i] = rr.x; ref_r[4*i+1] = rr.y; ref_r[4*i+2] = rr.z; ref_r[4*i+3] = rr.w;
  }
}
void h_qrot(const float* q, const float* v, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const V3 r = qrot(q4(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3]), v3(v[3*i], v[3*i+1], v[3*i+2])); o[3*i] = r.x; o[3*i+1] = r.y; o[3*i+2] = r.z; }
}
// qrot_z(z(q), v) and the general qrot on the same operands
void h_qrot_z(const float* q, const float* v, float* o, float* ref, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const Q4 z = q4(0.0f, 0.0f, q[4*i+2], q[4*i+3]);
    const V3 x = v3(v[3*i], v[3*i+1], v[3*i+2]), r = qrot_z(z, x), g = qrot(z, x);
    o[3*i] = r.x; o[3*i+1] = r.y; o[3*i+2] = r.z; ref[3*i] = g.x; ref[3*i+1] = g.y; ref[3*i+2] = g.z;
  }
}
void h_tan_norm(const float* q, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const TanNorm t = tan_norm(q4(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3])); o[6*i] = t.t.x; o[6*i+1] = t.t.y; o[6*i+2] = t.t.z; o[6*i+3] = t.n.x; o[6*i+4] = t.n.y; o[6*i+5] = t.n.z; }
}
void h_quat_angle(const float* q, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) o[i] = quat_angle(q4(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3]));
}
void h_quat_to_exp_map(const float* q, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const V3 r = quat_to_exp_map(q4(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3])); o[3*i] = r.x; o[3*i+1] = r.y; o[3*i+2] = r.z; }
}
void h_exp_map_to_quat(const float* e, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const Q4 r = exp_map_to_quat(v3(e[3*i], e[3*i+1], e[3*i+2])); o[4*i] = r.x; o[4*i+1] = r.y; o[4*i+2] = r.z; o[4*i+3] = r.w; }
}
void h_slerp(const float* a, const float* b, const float* t, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const Q4 r = slerp(q4(a[4*i], a[4*i+1], a[4*i+2], a[4*i+3]), q4(b[4*i], b[4*i+1], b[4*i+2], b[4*i+3]), t[i]); o[4*i] = r.x; o[4*i+1] = r.y; o[4*i+2] = r.z; o[4*i+3] = r.w; }
}
void h_heading(const float* q, float* ang, float* hq, float* hinv, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const float h = heading_angle(q4(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3]));
    const Q4 z = quat_about_z(h), zi = quat_about_z(-h);
    ang[i] = h;
    hq[4*i] = z.x; hq[4*i+1] = z.y; hq[4*i+2] = z.z; hq[4*i+3] = z.w;
    hinv[4*i] = zi.x; hinv[4*i+1] = zi.y; hinv[4*i+2] = zi.z; hinv[4*i+3] = zi.w;
  }
}
void h_strip_base_rot(const float* q, float* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const Q4 r = strip_base_rot(q4(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3])); o[4*i] = r.x; o[4*i+1] = r.y; o[4*i+2] = r.z; o[4*i+3] = r.w; }
}
void h_frame_bracket(const float* time, const float* len, const int64_t* nf, const float* dt, int64_t* i0, int64_t* i1, float* blend,
                     int32_t* i0_32, int32_t* i1_32, float* blend32, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const Bracket b = frame_bracket(time[i], len[i], nf[i], dt[i]);
    const Bracket32 c = frame_bracket32(time[i], len[i], (int)nf[i], dt[i]);
    i0[i] = b.i0; i1[i] = b.i1; blend[i] = b.blend; i0_32[i] = c.i0; i1_32[i] = c.i1; blend32[i] = c.blend;
  }
}

}  // extern "C"
