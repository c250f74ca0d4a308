"""Host-side logic of the MotionLibSMPL mirror that needs no GPU (phc/utils/motion_lib_base.py:131-172, :350-389): clip-list
filtering / ordering in load_data, the Auto-PMCP sampling weights, and the refusal to load without a CUDA device (no CPU
fallback).  The device-side load is covered by tests/test_gpu_motion_load.py."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from phc_b200 import ops
from phc_b200.motion_lib import FixHeightMode, MotionLibSMPL


def _clips(lengths):
    return {f"clip{i}": {"pose_quat_global": np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n, 24, 1)), "root_trans_offset": torch.zeros(n, 3),
                         "pose_aa": np.zeros((n, 72)), "fps": 30} for i, n in enumerate(lengths)}


def _cfg(clips, **kw):
    base = dict(motion_file=clips, device="cpu", fix_height=FixHeightMode.no_fix, min_length=-1, max_length=-1, im_eval=False,
                multi_thread=False, smpl_type="smpl")
    base.update(kw)
    return SimpleNamespace(**base)


def test_load_data_filters_and_orders_like_the_reference():
    clips = _clips([10, 40, 25, 5])
    lib = MotionLibSMPL(_cfg(clips))
    assert list(lib._motion_data_keys) == ["clip0", "clip1", "clip2", "clip3"] and lib._num_unique_motions == 4
    lib = MotionLibSMPL(_cfg(clips, min_length=20))                     # motion_lib_base.py:143-144
    assert list(lib._motion_data_keys) == ["clip1", "clip2"]
    lib = MotionLibSMPL(_cfg(clips, im_eval=True))                      # longest first (:145-146)
    assert list(lib._motion_data_keys) == ["clip1", "clip2", "clip0", "clip3"]
    assert torch.allclose(lib._sampling_prob, torch.full((4,), 0.25))


def test_auto_pmcp_sampling_weights():
    lib = MotionLibSMPL(_cfg(_clips([10, 12, 14, 16, 18])))
    lib.update_hard_sampling_weight(["clip1", "clip3"])                 # only the failed clips (:350-362)
    assert torch.allclose(lib._sampling_prob, torch.tensor([0.0, 0.5, 0.0, 0.5, 0.0]))
    lib.update_hard_sampling_weight([])
    assert torch.allclose(lib._sampling_prob, torch.full((5,), 0.2))
    lib.update_soft_sampling_weight(["clip0"])                          # termination history (:364-378)
    lib.update_soft_sampling_weight(["clip0", "clip4"])
    assert torch.allclose(lib._sampling_prob, torch.tensor([2 / 3, 0.0, 0.0, 0.0, 1 / 3]))
    assert lib.update_sampling_prob(torch.zeros(5)) is False and lib.update_sampling_prob(torch.ones(3)) is False


def test_load_motions_has_no_cpu_path():
    lib = MotionLibSMPL(_cfg(_clips([10, 12])))
    tree = SimpleNamespace(local_translation=np.zeros((24, 3)), parent_indices=np.arange(-1, 23), node_names=[str(i) for i in range(24)])
    with pytest.raises(ops.PhcError):
        lib.load_motions([tree, tree], [torch.zeros(17)] * 2, [np.zeros(10)] * 2, random_sample=False)
    with pytest.raises(ops.PhcError):
        lib.to("cpu")
