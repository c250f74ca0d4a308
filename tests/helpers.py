"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from oracle import phc_oracle as O
from phc_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def tables_from(d, prefix="tab_"):
    f = lambda k: d[prefix + k]
    return O.MotionTables(f("gts"), f("grs"), f("lrs"), f("gvs"), f("gavs"), f("dvs"), f("lengths"),
                          f("num_frames"), f("dts"), f("length_starts"))


def motion_data_from(d, prefix="tab_"):
    return syn.MotionData(**{k: d[prefix + k] for k in syn.MotionData.__dataclass_fields__})


def oracle_tables(m: syn.MotionData):
    return O.MotionTables(m.gts, m.grs, m.lrs, m.gvs, m.gavs, m.dvs, m.lengths, m.num_frames, m.dts, m.length_starts)


def env_state_from(d, tag):
    return syn.EnvState(**{k: d[f"{tag}_in_{k}"] for k in syn.EnvState.__dataclass_fields__})


def smpl_step_config(**kw):
    base = dict(key_bodies=syn.SMPL_KEY_BODIES, reset_bodies=syn.SMPL_RESET_BODIES,
                dof_subset=torch.tensor(syn.SMPL_DOF_SUBSET))
    base.update(kw)
    return O.StepConfig(**base)


def close(a, b, rtol=1e-5, atol=1e-6, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.dtype.is_floating_point:
        bad = ~torch.isclose(a.double(), b.double(), rtol=rtol, atol=atol, equal_nan=True)
        if bad.any():
            i = bad.nonzero()[0].tolist()
            err = (a.double() - b.double()).abs()
            raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {err[bad].max():.3e}, "
                                 f"first at {i}: {a[tuple(i)].item()} vs {b[tuple(i)].item()}")
    else:
        assert torch.equal(a, b), f"{what}: integer mismatch ({int((a != b).sum())} elements)"


# ---- hinge-joint robot (H1) fixtures: tests/golden/h1.npz ---------------------------------------------------------
def robot_tables_from(d, prefix="tab_"):
    f = lambda k: d[prefix + k]
    return O.RobotTables(f("gts_t"), f("grs_t"), f("gvs_t"), f("gavs_t"), f("dof_pos"), f("dvs"), f("lengths"), f("num_frames"),
                         f("dts"), f("length_starts"), syn.H1_NUM_BODIES)


def robot_motion_data_from(d, prefix="tab_"):
    f = lambda k: d[prefix + k]
    return syn.RobotMotionData(gts_t=f("gts_t"), grs_t=f("grs_t"), gvs_t=f("gvs_t"), gavs_t=f("gavs_t"), dof_pos=f("dof_pos"), dvs=f("dvs"),
                               lengths=f("lengths"), num_frames=f("num_frames"), dts=f("dts"), length_starts=f("length_starts"),
                               num_bodies=syn.H1_NUM_BODIES)


def h1_step_config(**kw):
    base = dict(key_bodies=syn.H1_KEY_BODIES, reset_bodies=None, dof_subset=None)
    base.update(kw)
    return O.StepConfig(**base)
