"""GPU parity of the device-side motion loader (phc_motion_load, csrc/motion_load.cu) -- SURVEY.md 8(f) rank 1.

Golden: tests/golden/load.npz = outputs of the UNMODIFIED MotionLibSMPL.load_motion_with_skeleton (make_golden.py:gen_load).
Oracle: oracle/motion_load_oracle.py (numpy restatement pinned to the same golden) at a size the CPU finishes in seconds.
Tolerances: positions / rotations / filtered velocities come out of a float64 pipeline (rtol 1e-6 after the float32 cast);
dof velocities run in float32 through 2*acos(w) of a near-identity rotation (conditioning ~1/sin(angle/2)): rtol 1e-5, atol 2e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import motion_load_oracle as ML
from phc_b200 import ops, synthetic as syn
from tests.helpers import close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(z, heading=True):
    d = lambda k, dt: torch.from_numpy(np.ascontiguousarray(z[k])).to(dt).to(DEV)
    return ops.load_motion_tables(d("pose_quat_global", torch.float64), d("root_trans", torch.float64), d("offsets", torch.float64),
                                  d("parents", torch.int32), d("num_frames", torch.int64), d("fps", torch.float64),
                                  d("heading", torch.float64) if heading else None)


def _check(out, exp):
    for k in ("gts", "grs", "lrs", "gvs", "gavs"):
        close(out[k].cpu(), torch.from_numpy(exp[k]), rtol=1e-6, atol=1e-6, what=k)
    close(out["dvs"].cpu(), torch.from_numpy(exp["dvs"]), rtol=1e-5, atol=2e-5, what="dvs")


def test_loader_matches_reference_golden():
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "load.npz")))
    out = _run(z)
    _check(out, z)
    nf = torch.from_numpy(z["num_frames"])
    assert torch.equal(out["length_starts"].cpu(), torch.cumsum(nf, 0) - nf)
    close(out["lengths"].cpu(), torch.from_numpy(((z["num_frames"] - 1) / z["fps"]).astype(np.float32)), what="lengths")
    close(out["dts"].cpu(), torch.from_numpy((1.0 / z["fps"]).astype(np.float32)), what="dts")


def _random_clips(M, J, seed, min_f=2, max_f=120):
    rng = np.random.default_rng(seed)
    if J == 24:
        parents = np.array(syn.SMPL_PARENTS, dtype=np.int32)
    else:
        parents = np.array([-1] + [int(rng.integers(0, j)) for j in range(1, J)], dtype=np.int32)
    nf = rng.integers(min_f, max_f + 1, size=M)
    F = int(nf.sum())
    q = rng.standard_normal((F, J, 4))
    # smooth the rotations inside a clip so frame-to-frame angles look like motion data; keep arbitrary signs / norms ~1
    s = 0
    for n in nf:
        w = np.cumsum(rng.standard_normal((n, J, 4)) * 0.05, 0) + rng.standard_normal((1, J, 4))
        q[s:s + n] = w / np.linalg.norm(w, axis=-1, keepdims=True)
        s += n
    t = np.cumsum(rng.standard_normal((F, 3)) * 0.02, 0)
    off = rng.standard_normal((M, J, 3)) * 0.15
    off[:, 0] = 0
    fps = rng.choice([30.0, 60.0, 120.0], size=M)
    heading = np.pi * (2 * rng.random(M) - 1)
    return dict(pose_quat_global=q, root_trans=t, offsets=off, parents=parents, num_frames=nf.astype(np.int64), fps=fps, heading=heading)


@pytest.mark.parametrize("M,J", [(40, 24), (9, 52), (5, 3)])
def test_loader_matches_oracle(M, J):
    z = _random_clips(M, J, seed=M + J)
    exp = ML.load_clips(z["pose_quat_global"], z["root_trans"], z["num_frames"], z["fps"], z["parents"], z["offsets"], z["heading"])
    _check(_run(z), exp)


def test_loader_without_heading_randomisation():
    z = _random_clips(6, 24, seed=3)
    exp = ML.load_clips(z["pose_quat_global"], z["root_trans"], z["num_frames"], z["fps"], z["parents"], z["offsets"], None)
    _check(_run(z, heading=False), exp)


def test_loader_properties_at_scale():
    """4096 clips x 60-300 frames (the faithful one-clip-per-env load): size-independent properties instead of the oracle --
    unit quaternions, root position = rotated root translation, the last dof-velocity row repeats the previous one, and a
    constant-velocity clip gives exactly that velocity after the (normalised) gaussian filter."""
    M, J = 4096, 24
    z = _random_clips(M, J, seed=11, min_f=60, max_f=300)
    # clip 0: pure translation at constant velocity, constant rotation
    n0 = int(z["num_frames"][0])
    z["pose_quat_global"][:n0] = z["pose_quat_global"][0]
    v = np.array([0.3, -0.2, 0.1])
    z["root_trans"][:n0] = np.arange(n0)[:, None] * v / z["fps"][0]
    out = _run(z)
    torch.cuda.synchronize()
    grs, lrs = out["grs"], out["lrs"]
    assert torch.allclose(grs.norm(dim=-1), torch.ones_like(grs[..., 0]), atol=1e-6)
    assert torch.allclose(lrs.norm(dim=-1), torch.ones_like(lrs[..., 0]), atol=1e-6)
    th = torch.from_numpy(z["heading"]).to(DEV)
    clip = torch.repeat_interleave(torch.arange(M, device=DEV), out["num_frames"])
    t = torch.from_numpy(z["root_trans"]).to(DEV)
    c, s = torch.cos(th)[clip], torch.sin(th)[clip]
    root = torch.stack([c * t[:, 0] - s * t[:, 1], s * t[:, 0] + c * t[:, 1], t[:, 2]], -1).float()
    close(out["gts"][:, 0].cpu(), root.cpu(), rtol=1e-6, atol=1e-6, what="root position")
    last = out["length_starts"] + out["num_frames"] - 1
    assert torch.equal(out["dvs"][last], out["dvs"][last - 1])
    c0, s0 = np.cos(z["heading"][0]), np.sin(z["heading"][0])
    v_rot = torch.tensor([c0 * v[0] - s0 * v[1], s0 * v[0] + c0 * v[1], v[2]], dtype=torch.float32)
    close(out["gvs"][:n0].cpu(), v_rot.expand(n0, J, 3), rtol=1e-5, atol=1e-6, what="constant velocity")
    assert float(out["gavs"][:n0].abs().max()) < 1e-6 and float(out["dvs"][:n0].abs().max()) < 1e-3      # float32 identity noise, as in the reference


def test_motion_lib_smpl_end_to_end():
    """MotionLibSMPL mirror: load -> tables -> packed records -> get_motion_state agrees with the oracle's interpolation of the
    oracle-loaded tables; the instance plugs into HumanoidIm as cfg['motion_data']."""
    from types import SimpleNamespace
    from oracle import phc_oracle as O
    from phc_b200.motion_lib import MotionLibSMPL
    z = _random_clips(12, 24, seed=5, min_f=20, max_f=60)
    clips, s = {}, 0
    for i, n in enumerate(z["num_frames"]):
        clips[f"clip{i}"] = {"pose_quat_global": z["pose_quat_global"][s:s + n], "root_trans_offset": torch.from_numpy(z["root_trans"][s:s + n]),
                             "pose_aa": np.zeros((n, 72)), "fps": float(z["fps"][i])}
        s += n
    lib = MotionLibSMPL(SimpleNamespace(motion_file=clips, device=DEV, fix_height=0, min_length=-1, max_length=-1, im_eval=False,
                                        multi_thread=False, smpl_type="smpl"))
    trees = [SimpleNamespace(local_translation=z["offsets"][i], parent_indices=z["parents"], node_names=[f"b{j}" for j in range(24)])
             for i in range(12)]
    lib.load_motions(trees, [torch.zeros(17)] * 12, [np.zeros(10)] * 12, random_sample=False, heading=z["heading"])
    exp = ML.load_clips(z["pose_quat_global"], z["root_trans"], z["num_frames"], z["fps"], z["parents"], z["offsets"], z["heading"])
    _check({k: getattr(lib, k) for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs")}, exp)
    assert lib.num_motions() == 12 and lib.num_bodies == 24
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 12, (64,), generator=g)
    times = torch.rand(64, generator=g) * lib.get_motion_length().cpu()[ids]
    st = lib.get_motion_state(ids.to(DEV), times.to(DEV))
    tab = O.MotionTables(*[torch.from_numpy(exp[k]) for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs")], lib.lengths.cpu(),
                         lib.num_frames.cpu(), lib.dts.cpu(), lib.length_starts.cpu())
    ref = O.motion_state(tab, ids, times)
    for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel", "dof_vel"):
        close(st[k].cpu(), ref[k], rtol=1e-5, atol=2e-5, what=f"get_motion_state {k}")
    assert st["motion_aa"].shape == (64, 72) and st["motion_bodies"].shape == (64, 17)
    t = lib.sample_time_interval(ids.to(DEV))
    assert bool(((t * 30).round() - t * 30).abs().max() < 1e-3) and bool((t <= lib.get_motion_length(ids.to(DEV))).all())
