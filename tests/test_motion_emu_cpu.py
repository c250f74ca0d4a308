"""The motion-library kernels on the CPU: motion.cu (one body per lane) and motion_wide.cu (bodies strided over the lanes, for
more than 32 bodies) are compiled verbatim against the warp emulation of tests/emu/ and checked against the goldens of the
unmodified reference: MotionLibBase.get_motion_state (motion.npz), MotionLibReal.get_motion_state incl. the *_t outputs
(h1.npz, g1.npz), build_amp_obs_demo (envstep.npz, h1.npz, g1.npz), _init_amp_obs_ref (reset.npz).  The strided kernels run on
EVERY golden (they are generic in the body count); the lane-per-body kernels on the shapes they take (<= 32 bodies)."""
import ctypes as C
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

from phc_b200 import _lib, synthetic as syn          # noqa: E402
from tests.helpers import close, load                 # noqa: E402

P = C.c_void_p


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import build_emu
    lib = C.CDLL(build_emu.build_motion(str(tmp_path_factory.mktemp("memu"))))
    lib.emu_motion_state.argtypes = [C.POINTER(_lib.PhcMotionLib), P, P, P, C.c_int64, C.POINTER(_lib.PhcMotionStateOut), C.c_int]
    lib.emu_amp_obs_demo.argtypes = [C.POINTER(_lib.PhcMotionLib), P, P, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_uint32, P, C.c_int32, P,
                                     C.c_int32, P, C.c_int64, P, C.c_int32, C.c_int]
    lib.emu_set_env_state.argtypes = [C.POINTER(_lib.PhcMotionLib), P, P, P, P, C.c_int64, P, C.c_int32, P, C.c_int]
    return lib


def round4(x):
    return (x + 3) & ~3


class HostLib:
    """phc_motion_pack / phc_motion_pack_dofs on the host (layouts of include/phc_b200.h)."""

    def __init__(self, gts, grs, gvs, gavs, lengths, num_frames, dts, length_starts, lrs=None, dvs=None, num_ext=0, dof_pos=None, dof_vel=None):
        F, JE = gts.shape[0], gts.shape[1]
        J = JE - num_ext
        bs = round4(13 * JE)
        self.fb = torch.zeros(F, bs)
        self.fb[:, :13 * JE] = torch.cat((gts, grs, gvs, gavs), -1).reshape(F, 13 * JE)
        D = 0 if dof_pos is None else dof_pos.shape[1]
        if D:
            js = round4(2 * D)
            self.fj = torch.zeros(F, js)
            self.fj[:, :D], self.fj[:, D:2 * D] = dof_pos, dof_vel
        else:
            js = round4(4 * J + 3 * (J - 1))
            self.fj = torch.zeros(F, js)
            self.fj[:, :4 * J] = lrs.reshape(F, 4 * J)
            self.fj[:, 4 * J:4 * J + 3 * (J - 1)] = dvs.reshape(F, 3 * (J - 1))
        self.t = [lengths.float().contiguous(), dts.float().contiguous(), num_frames.long().contiguous(), length_starts.long().contiguous()]
        self.c = _lib.PhcMotionLib(self.fb.data_ptr(), self.fj.data_ptr(), self.t[0].data_ptr(), self.t[1].data_ptr(), self.t[2].data_ptr(),
                                   self.t[3].data_ptr(), F, int(lengths.shape[0]), J, bs, js, num_ext, D)
        self.J, self.JE, self.D = J, JE, D


def smpl_lib(g, prefix="tab_"):
    f = lambda k: g[prefix + k]
    return HostLib(f("gts"), f("grs"), f("gvs"), f("gavs"), f("lengths"), f("num_frames"), f("dts"), f("length_starts"), lrs=f("lrs"), dvs=f("dvs"))


def robot_lib(g, num_ext):
    f = lambda k: g["tab_" + k]
    return HostLib(f("gts_t"), f("grs_t"), f("gvs_t"), f("gavs_t"), f("lengths"), f("num_frames"), f("dts"), f("length_starts"), num_ext=num_ext,
                   dof_pos=f("dof_pos"), dof_vel=f("dvs"))


def motion_state(emu, lib, ids, times, offset, wide):
    n, J, JE, D = len(ids), lib.J, lib.JE, lib.D
    dofs = D if D else 3 * (J - 1)
    out = dict(rg_pos=torch.zeros(n, J, 3), rb_rot=torch.zeros(n, J, 4), body_vel=torch.zeros(n, J, 3), body_ang_vel=torch.zeros(n, J, 3),
               dof_pos=torch.zeros(n, dofs), dof_vel=torch.zeros(n, dofs), root_pos=torch.zeros(n, 3), root_rot=torch.zeros(n, 4),
               root_vel=torch.zeros(n, 3), root_ang_vel=torch.zeros(n, 3), rg_pos_t=torch.zeros(n, JE, 3), rg_rot_t=torch.zeros(n, JE, 4),
               body_vel_t=torch.zeros(n, JE, 3), body_ang_vel_t=torch.zeros(n, JE, 3))
    co = _lib.PhcMotionStateOut(**{k: out[k].data_ptr() for k, _ in _lib.PhcMotionStateOut._fields_})
    ids, times = ids.long().contiguous(), times.float().contiguous()
    off = None if offset is None else offset.float().contiguous()
    assert emu.emu_motion_state(C.byref(lib.c), ids.data_ptr(), times.data_ptr(), None if off is None else off.data_ptr(), n, C.byref(co), int(wide)) == 0
    return out


def amp_demo(emu, lib, ids, t0, key_bodies, amp_joints, wide, first_step=0, num_steps=10, flags=None):
    flags = (_lib.PHC_FLAG_UPRIGHT | _lib.PHC_FLAG_LOCAL_ROOT_OBS | _lib.PHC_FLAG_ROOT_HEIGHT_OBS) if flags is None else flags
    nk, nj = len(key_bodies), len(amp_joints)
    A = 13 + (2 * lib.D if lib.D else 9 * nj) + 3 * nk
    n = len(ids)
    out = torch.zeros(n, num_steps, A)
    kb = (C.c_int32 * max(nk, 1))(*key_bodies)
    aj = (C.c_int32 * max(nj, 1))(*amp_joints)
    ids, t0 = ids.long().contiguous(), t0.float().contiguous()
    assert emu.emu_amp_obs_demo(C.byref(lib.c), ids.data_ptr(), t0.data_ptr(), n, first_step, num_steps, 1.0 / 30.0, flags, kb, nk, aj, nj,
                                out.data_ptr(), num_steps * A, None, 0, int(wide)) == 0
    return out


SMPL_JOINTS = [d // 3 for d in syn.SMPL_DOF_SUBSET[::3]]


@pytest.mark.parametrize("wide", [0, 1])
def test_motion_state_smpl(emu, wide):
    g = load("motion.npz")
    lib = smpl_lib(g)
    out = motion_state(emu, lib, g["ids"], g["times"], g["offset"], wide)
    for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_vel", "rg_pos", "rb_rot", "body_vel", "body_ang_vel"):
        close(out[k], g["out_" + k], what=f"{k} wide={wide}")
    close(out["dof_pos"], g["out_dof_pos"], rtol=1e-4, atol=2e-5, what="dof_pos (acos near identity)")
    close(motion_state(emu, lib, g["ids"], g["times"], None, wide)["rg_pos"], g["out_noffset_rg_pos"], what="rg_pos no offset")


@pytest.mark.parametrize("name,ext,wide", [("h1.npz", 3, 0), ("h1.npz", 3, 1), ("g1.npz", 1, 1)])
def test_motion_state_robot(emu, name, ext, wide):
    g = load(name)
    lib = robot_lib(g, ext)
    out = motion_state(emu, lib, g["ms_ids"], g["ms_times"], g["ms_offset"], wide)
    for k in ("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "rg_pos", "rb_rot", "body_vel", "body_ang_vel",
              "rg_pos_t", "rg_rot_t", "body_vel_t", "body_ang_vel_t"):
        close(out[k], g["ms_out_" + k], what=f"{name} {k} wide={wide}")


@pytest.mark.parametrize("wide", [0, 1])
def test_amp_demo_smpl_and_history_init(emu, wide):
    g = load("envstep.npz")
    demo = amp_demo(emu, smpl_lib(g), g["demo_ids"], g["demo_t0"], syn.SMPL_KEY_BODIES, SMPL_JOINTS, wide)
    close(demo, g["demo_out"], rtol=1e-4, atol=2e-5, what=f"build_amp_obs_demo wide={wide}")
    r = load("reset.npz")
    hist = amp_demo(emu, smpl_lib(r), r["motion_ids"], r["t0"], syn.SMPL_KEY_BODIES, SMPL_JOINTS, wide, first_step=1, num_steps=9)
    close(hist, r["hist_after"][r["env_ids"]], rtol=1e-4, atol=2e-5, what=f"_init_amp_obs_ref wide={wide}")


@pytest.mark.parametrize("name,ext,keys,wide", [("h1.npz", 3, syn.H1_KEY_BODIES, 0), ("h1.npz", 3, syn.H1_KEY_BODIES, 1), ("g1.npz", 1, syn.G1_KEY_BODIES, 1)])
def test_amp_demo_robot(emu, name, ext, keys, wide):
    g = load(name)
    demo = amp_demo(emu, robot_lib(g, ext), g["demo_ids"], g["demo_t0"], keys, [], wide)
    close(demo, g["demo_out"], rtol=1e-4, atol=2e-5, what=f"{name} amp demo wide={wide}")


@pytest.mark.parametrize("name,ext,wide", [("h1.npz", 3, 0), ("g1.npz", 1, 1)])
def test_set_env_state_writes_the_motion_state(emu, name, ext, wide):
    """_set_env_state = the reference pose at (id, time) written into the simulator rows of the masked envs: equal to the
    get_motion_state golden of the same queries, other rows untouched."""
    g = load(name)
    lib = robot_lib(g, ext)
    n, J, D = len(g["ms_ids"]), lib.J, lib.D
    bpe = J + 2
    body = torch.full((n, bpe, 13), 9.0)
    dof = torch.full((n, D, 2), 9.0)
    mask = torch.ones(n, dtype=torch.int64)
    mask[::4] = 0
    ids, times, off = g["ms_ids"].long().contiguous(), g["ms_times"].float().contiguous(), g["ms_offset"].float().contiguous()
    assert emu.emu_set_env_state(C.byref(lib.c), ids.data_ptr(), times.data_ptr(), off.data_ptr(), mask.data_ptr(), n, body.data_ptr(), bpe,
                                 dof.data_ptr(), int(wide)) == 0
    m = mask.bool()
    close(body[m][:, :J, 0:3], g["ms_out_rg_pos"][m], what="body pos")
    close(body[m][:, :J, 3:7], g["ms_out_rb_rot"][m], what="body rot")
    close(body[m][:, :J, 7:10], g["ms_out_body_vel"][m], what="body vel")
    close(body[m][:, :J, 10:13], g["ms_out_body_ang_vel"][m], what="body ang vel")
    close(dof[m][..., 0], g["ms_out_dof_pos"][m], what="dof pos")
    close(dof[m][..., 1], g["ms_out_dof_vel"][m], what="dof vel")
    assert bool((body[~m] == 9.0).all()) and bool((dof[~m] == 9.0).all()) and bool((body[:, J:] == 9.0).all())


# ---- motion_load.cu (the loader) on the CPU against the golden of the real MotionLibSMPL.load_motion_with_skeleton ---------
def test_loader_kernels_vs_reference_golden(tmp_path):
    import shutil
    import numpy as np
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import build_emu
    lib = C.CDLL(build_emu.build_load(str(tmp_path)))
    lib.emu_motion_load.argtypes = [P] * 8 + [C.c_int64, C.c_int32, C.c_int32] + [P] * 9
    z = np.load(os.path.join(HERE, "golden", "load.npz"))
    q, t = np.ascontiguousarray(z["pose_quat_global"], np.float64), np.ascontiguousarray(z["root_trans"], np.float64)
    off, par = np.ascontiguousarray(z["offsets"], np.float64), np.ascontiguousarray(z["parents"], np.int32)
    head, nf, fps = np.ascontiguousarray(z["heading"], np.float64), np.ascontiguousarray(z["num_frames"], np.int64), np.ascontiguousarray(z["fps"], np.float64)
    starts = np.ascontiguousarray(np.cumsum(nf) - nf, np.int64)
    F, J, M = q.shape[0], q.shape[1], len(nf)
    out = {k: np.zeros((F, J if k != "dvs" else J - 1, w), np.float32) for k, w in (("gts", 3), ("grs", 4), ("lrs", 4), ("gvs", 3), ("gavs", 3), ("dvs", 3))}
    pos64, rawang, clip = np.zeros((F, J, 3)), np.zeros((F, J, 3)), np.zeros(F, np.int32)
    ptr = lambda a: a.ctypes.data_as(P)
    assert lib.emu_motion_load(ptr(q), ptr(t), ptr(off), ptr(par), ptr(head), ptr(starts), ptr(nf), ptr(fps), F, M, J, ptr(out["gts"]), ptr(out["grs"]),
                               ptr(out["lrs"]), ptr(out["gvs"]), ptr(out["gavs"]), ptr(out["dvs"]), ptr(pos64), ptr(rawang), ptr(clip)) == 0
    assert np.array_equal(clip, np.repeat(np.arange(M), nf))
    for k in ("gts", "grs", "lrs", "gvs", "gavs"):
        close(torch.from_numpy(out[k]), torch.from_numpy(z[k]), rtol=1e-6, atol=1e-6, what=f"loader {k}")
    close(torch.from_numpy(out["dvs"]), torch.from_numpy(z["dvs"]), rtol=1e-5, atol=2e-5, what="loader dvs")
