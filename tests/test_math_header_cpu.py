"""The kernels' arithmetic header on the CPU: phc_b200/csrc/phc_math.cuh is host+device code, so the very source the CUDA
kernels are built from is compiled here with g++ (tests/math_header_host.cpp, -ffp-contract=off) and pinned against the goldens
of the unmodified reference (tests/golden/quat.npz = phc/utils/torch_utils.py on seeded inputs, motion.npz = _calc_frame_blend)
without a GPU.  Also proves the "zero-folded" variants (qmul_zl / qmul_zr / qrot_z) equal the general expressions bit for bit
when nothing is contracted.  Tolerances: rtol 1e-5 / atol 1e-6 like the GPU parity tests (glibc's libm vs torch's differ by ulps)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
f32p, i64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("mh") / "libphc_math_host.so")
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", os.path.join(HERE, "math_header_host.cpp"), "-o", so, "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def fp(a):
    return a.ctypes.data_as(f32p)


def golden(name):
    z = np.load(os.path.join(HERE, "golden", name))
    return {k: np.ascontiguousarray(z[k]) for k in z.files}


def close(a, b, rtol=1e-5, atol=1e-6, what=""):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


def test_quaternion_functions_match_the_reference_goldens(host):
    g = golden("quat.npz")
    a, b, v, t, e = g["a"], g["b"], g["v"], np.ascontiguousarray(g["t"].reshape(-1)), g["e"]
    n = a.shape[0]
    out4, out3, out6, out1 = np.empty((n, 4), np.float32), np.empty((n, 3), np.float32), np.empty((n, 6), np.float32), np.empty(n, np.float32)
    host.h_qmul(fp(a), fp(b), fp(out4), C.c_int64(n))
    close(out4, g["quat_mul"], what="quat_mul")
    host.h_qrot(fp(a), fp(v), fp(out3), C.c_int64(n))
    close(out3, g["my_quat_rotate"], atol=2e-6, what="my_quat_rotate")
    host.h_tan_norm(fp(a), fp(out6), C.c_int64(n))
    close(out6, g["quat_to_tan_norm"], what="quat_to_tan_norm")
    host.h_quat_angle(fp(a), fp(out1), C.c_int64(n))
    close(out1, g["angle"], what="quat_to_angle_axis angle")
    small = g["small"]
    o_s = np.empty(small.shape[0], np.float32)
    host.h_quat_angle(fp(small), fp(o_s), C.c_int64(small.shape[0]))
    close(o_s, g["angle_small"], rtol=1e-4, atol=2e-6, what="angle of near-identity rotations (2*acos(w) near 1)")
    host.h_quat_to_exp_map(fp(a), fp(out3), C.c_int64(n))
    close(out3, g["quat_to_exp_map"], rtol=1e-4, atol=2e-6, what="quat_to_exp_map")
    host.h_exp_map_to_quat(fp(e), fp(out4), C.c_int64(n))
    close(out4, g["exp_map_to_quat"], what="exp_map_to_quat")
    host.h_slerp(fp(a), fp(b), fp(t), fp(out4), C.c_int64(n))
    close(out4, g["slerp"], rtol=1e-4, atol=1e-5, what="slerp (sin(acos c) near c = 1)")
    ang, hq, hinv = np.empty(n, np.float32), np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
    host.h_heading(fp(a), fp(ang), fp(hq), fp(hinv), C.c_int64(n))
    close(ang, g["calc_heading"], what="calc_heading")
    close(hq, g["calc_heading_quat"], what="calc_heading_quat")
    close(hinv, g["calc_heading_quat_inv"], what="calc_heading_quat_inv")
    assert np.array_equal(hinv[:, 2], -hq[:, 2]) and np.array_equal(hinv[:, 3], hq[:, 3])      # the kernel uses the exact conjugate
    host.h_strip_base_rot(fp(a), fp(out4), C.c_int64(n))
    close(out4, g["remove_base_rot"], what="remove_base_rot")


def test_zero_folded_variants_are_bit_identical_to_the_general_expressions(host):
    rng = np.random.default_rng(3)
    n = 4096
    a = rng.standard_normal((n, 4)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.standard_normal((n, 4)).astype(np.float32)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    v = (rng.standard_normal((n, 3)) * 3).astype(np.float32)
    ol, rl, orr, rr = (np.empty((n, 4), np.float32) for _ in range(4))
    host.h_qmul_z(fp(a), fp(b), fp(ol), fp(rl), fp(orr), fp(rr), C.c_int64(n))
    assert np.array_equal(ol, rl), "qmul_zl differs from qmul"
    assert np.array_equal(orr, rr), "qmul_zr differs from qmul"
    o, r = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    host.h_qrot_z(fp(a), fp(v), fp(o), fp(r), C.c_int64(n))
    assert np.array_equal(o, r), "qrot_z differs from qrot"


def test_frame_bracket_matches_calc_frame_blend(host):
    g = golden("motion.npz")
    ids = g["ids"]
    time, ln = g["times"].astype(np.float32), g["tab_lengths"][ids].astype(np.float32)
    nf, dt = np.ascontiguousarray(g["tab_num_frames"][ids].astype(np.int64)), g["tab_dts"][ids].astype(np.float32)
    n = len(ids)
    i0, i1, bl = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float32)
    j0, j1, b32 = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32)
    host.h_frame_bracket(fp(time), fp(ln), nf.ctypes.data_as(i64p), fp(dt), i0.ctypes.data_as(i64p), i1.ctypes.data_as(i64p), fp(bl),
                         j0.ctypes.data_as(i32p), j1.ctypes.data_as(i32p), fp(b32), C.c_int64(n))
    assert np.array_equal(i0, g["idx0"]) and np.array_equal(i1, g["idx1"])            # frame indices: bit-exact
    assert np.array_equal(bl, g["blend"])                                              # same fp32 operation order -> same bits
    assert np.array_equal(j0, i0) and np.array_equal(j1, i1) and np.array_equal(b32, bl)   # the 32-bit bracket of the step kernel
