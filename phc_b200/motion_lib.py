"""Host-side mirror of the reference's motion library for SMPL humanoids, with the LOAD done on the device.

Mirrors `MotionLibBase` / `MotionLibSMPL` (phc/utils/motion_lib_base.py:109-567, phc/utils/motion_lib_smpl.py:31-200): same
constructor config fields (`motion_file`, `device`, `fix_height`, `min_length`, `max_length`, `im_eval`, `multi_thread`,
`smpl_type`), same methods (`load_data`, `load_motions`, `get_motion_state`, `sample_motions`, `sample_time`,
`sample_time_interval`, `get_motion_length`, `num_motions`, `get_total_length`, `update_*_sampling_weight`) and the same table
attributes after a load (`gts grs lrs gvs gavs dvs _motion_lengths _motion_fps _motion_dt _motion_num_frames length_starts
motion_ids num_bodies`).  What differs is where the work happens: the reference runs forward kinematics, the gaussian-filtered
finite differences and the dof velocities clip by clip on CPU worker processes and uploads ~1 GB of tables every
`shape_resampling_interval` epochs (amp_agent.py:511-515); here the raw clip arrays go to the GPU once per load and
`phc_motion_load` (csrc/motion_load.cu) produces all tables in two launches, `phc_motion_pack` the per-frame records the fused
env step reads.  An instance can be passed as `cfg["motion_data"]` of `phc_b200.env.humanoid_im.HumanoidIm`.

No CPU path: `load_motions` needs a CUDA device.  `fix_trans_height` (needs the SMPL mesh model, smpl_sim) is taken as in the
reference when `data/smpl` is missing: `mesh_parsers = None`, no height fix (motion_lib_smpl.py:56-58, :151-154).
"""
from __future__ import annotations

import glob
import os.path as osp
import random
from enum import Enum
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops


class FixHeightMode(Enum):          # motion_lib_base.py:28-31
    no_fix = 0
    full_fix = 1
    ankle_fix = 2


class MotionlibMode(Enum):          # motion_lib_base.py:103-105
    file = 1
    directory = 2


class MotionLibSMPL:
    def __init__(self, motion_lib_cfg):
        cfg = motion_lib_cfg if not isinstance(motion_lib_cfg, dict) else SimpleNamespace(**motion_lib_cfg)
        self.m_cfg = cfg
        self._device = torch.device(getattr(cfg, "device", "cuda"))
        self._sim_fps = 1 / getattr(cfg, "step_dt", 1 / 30)
        self.im_eval = bool(getattr(cfg, "im_eval", False))
        self.test = bool(getattr(cfg, "test", False))              # reference: the global flags.test / flags.im_eval
        self.mesh_parsers = None
        self.load_data(cfg.motion_file, min_length=getattr(cfg, "min_length", -1), im_eval=self.im_eval)
        self.setup_constants(fix_height=getattr(cfg, "fix_height", FixHeightMode.no_fix), multi_thread=getattr(cfg, "multi_thread", False))

    # ---- motion_lib_base.py:131-158 ---------------------------------------------------------------------------------
    def load_data(self, motion_file, min_length=-1, im_eval=False):
        if isinstance(motion_file, dict):                            # already-loaded {key: clip dict}
            self.mode = MotionlibMode.file
            self._motion_data_load = motion_file
        elif osp.isfile(motion_file):
            import joblib
            self.mode = MotionlibMode.file
            self._motion_data_load = joblib.load(motion_file)
        else:
            self.mode = MotionlibMode.directory
            self._motion_data_load = glob.glob(osp.join(motion_file, "*.pkl"))
        if self.mode == MotionlibMode.file:
            if min_length != -1:
                data_list = {k: v for k, v in self._motion_data_load.items() if len(v["pose_quat_global"]) >= min_length}
            elif im_eval:
                data_list = dict(sorted(self._motion_data_load.items(), key=lambda e: len(e[1]["pose_quat_global"]), reverse=True))
            else:
                data_list = self._motion_data_load
            self._motion_data_list = np.empty(len(data_list), dtype=object)
            self._motion_data_list[:] = list(data_list.values())
            self._motion_data_keys = np.array(list(data_list.keys()))
        else:
            self._motion_data_list = np.array(self._motion_data_load)
            self._motion_data_keys = np.array(self._motion_data_load)
        self._num_unique_motions = len(self._motion_data_list)

    # ---- motion_lib_base.py:160-172 ---------------------------------------------------------------------------------
    def setup_constants(self, fix_height=FixHeightMode.full_fix, multi_thread=True):
        self.fix_height = fix_height
        self.multi_thread = multi_thread
        n, dev = self._num_unique_motions, self._device
        self._curr_motion_ids = None
        self._termination_history = torch.zeros(n, device=dev)
        self._success_rate = torch.zeros(n, device=dev)
        self._sampling_history = torch.zeros(n, device=dev)
        self._sampling_prob = torch.ones(n, device=dev) / n
        self._sampling_batch_prob = None

    # ---- motion_lib_base.py:181-326 + motion_lib_smpl.py:101-180, on the device -----------------------------------------
    def load_motions(self, skeleton_trees: Sequence, gender_betas: Sequence, limb_weights: Sequence, random_sample: bool = True,
                     start_idx: int = 0, max_len: int = -1, heading: Optional[np.ndarray] = None):
        """Loads len(skeleton_trees) clips (one per humanoid).  `skeleton_trees[i]` needs `.local_translation` [J,3],
        `.parent_indices` [J] and `.node_names` (poselib's SkeletonTree has them).  `heading` overrides the random heading
        angles (radians, one per clip; tests)."""
        if self._device.type != "cuda":
            raise ops.PhcError("MotionLibSMPL.load_motions runs on a CUDA device only; there is no CPU fallback")
        dev = self._device
        num = len(skeleton_trees)
        self.num_joints = len(skeleton_trees[0].node_names)
        if random_sample:
            sample_idxes = torch.multinomial(self._sampling_prob, num_samples=num, replacement=True).to(dev)
        else:
            sample_idxes = torch.remainder(torch.arange(num) + start_idx, self._num_unique_motions).to(dev)
        self._curr_motion_ids = sample_idxes
        self.one_hot_motions = torch.nn.functional.one_hot(sample_idxes, num_classes=self._num_unique_motions).to(dev)
        idx_np = sample_idxes.cpu().numpy()
        self.curr_motion_keys = self._motion_data_keys[idx_np]
        self._sampling_batch_prob = self._sampling_prob[sample_idxes] / self._sampling_prob[sample_idxes].sum()

        max_length = getattr(self.m_cfg, "max_length", -1)
        quats, transs, nfs, fpss, aas, bodies = [], [], [], [], [], []
        for f, clip in enumerate(self._motion_data_list[idx_np]):
            if not isinstance(clip, dict) and osp.isfile(clip):
                import joblib
                key = clip.split("/")[-1].split(".")[0]
                clip = joblib.load(clip)[key]
            seq_len = clip["root_trans_offset"].shape[0]
            if max_length == -1 or seq_len < max_length:
                start, end = 0, seq_len
            else:
                start = random.randint(0, seq_len - max_length)
                end = start + max_length
            q = np.asarray(clip["pose_quat_global"][start:end], dtype=np.float64)
            t = clip["root_trans_offset"][start:end]
            t = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
            if q.shape[0] < 2:
                raise ValueError(f"clip {self.curr_motion_keys[f]!r} has fewer than 2 frames (np.gradient needs 2)")
            quats.append(q)
            transs.append(t.astype(np.float64))
            nfs.append(q.shape[0])
            fpss.append(float(clip.get("fps", 30)))
            if "beta" in clip:
                aas.append(np.asarray(clip["pose_aa"][start:end]).reshape(-1, self.num_joints * 3))
                bodies.append(torch.as_tensor(gender_betas[f], dtype=torch.float32))
            else:
                aas.append(np.zeros((q.shape[0], self.num_joints * 3)))
                bodies.append(torch.zeros(17))
        if heading is None and not self.im_eval and not self.test:       # motion_lib_smpl.py:141-149
            heading = np.pi * (2 * np.random.random(num) - 1.0)
        J = quats[0].shape[1]
        offsets = np.stack([np.asarray(t.local_translation, dtype=np.float64).reshape(J, 3) for t in skeleton_trees])
        parents = np.asarray(skeleton_trees[0].parent_indices, dtype=np.int32).reshape(J)

        def up(a, dtype):      # one pinned staging copy per array
            h = torch.from_numpy(np.ascontiguousarray(a)).to(dtype).pin_memory()
            return h.to(dev, non_blocking=True)

        tabs = ops.load_motion_tables(up(np.concatenate(quats), torch.float64), up(np.concatenate(transs), torch.float64),
                                      up(offsets, torch.float64), up(parents, torch.int32), up(np.asarray(nfs), torch.int64),
                                      up(np.asarray(fpss), torch.float64), None if heading is None else up(np.asarray(heading), torch.float64))
        self.gts, self.grs, self.lrs = tabs["gts"], tabs["grs"], tabs["lrs"]
        self.gvs, self.gavs, self.dvs = tabs["gvs"], tabs["gavs"], tabs["dvs"]
        self.grvs, self.gravs = self.gvs[:, 0], self.gavs[:, 0]          # global_root_velocity / global_root_angular_velocity
        self._motion_lengths = tabs["lengths"]
        self._motion_fps = torch.tensor(fpss, device=dev, dtype=torch.float32)
        self._motion_dt = tabs["dts"]
        self._motion_num_frames = tabs["num_frames"]
        self._motion_bodies = torch.stack(bodies).to(dev).float()
        self._motion_aa = torch.tensor(np.concatenate(aas), device=dev, dtype=torch.float32)
        self._motion_limb_weights = torch.tensor(np.array(limb_weights), device=dev, dtype=torch.float32)
        self._num_motions = num
        self.length_starts = tabs["length_starts"]
        self.motion_ids = torch.arange(num, dtype=torch.long, device=dev)
        self.num_bodies = J
        self._packed = ops.pack_motion_lib(self.gts, self.grs, self.gvs, self.gavs, self.lrs, self.dvs, self._motion_lengths,
                                           self._motion_num_frames, self._motion_dt, self.length_starts)
        return self

    # ---- the attribute names phc_b200.env.humanoid_im.HumanoidIm reads from cfg["motion_data"] -------------------------
    @property
    def lengths(self):
        return self._motion_lengths

    @property
    def num_frames(self):
        return self._motion_num_frames

    @property
    def dts(self):
        return self._motion_dt

    @property
    def packed(self) -> ops.PackedMotionLib:
        return self._packed

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise ops.PhcError("MotionLibSMPL tables live on the CUDA device they were loaded on")
        return self

    # ---- queries (motion_lib_base.py:328-520) ---------------------------------------------------------------------------
    def num_motions(self):
        return self._num_motions

    def get_total_length(self):
        return float(self._motion_lengths.sum())

    def get_motion_state(self, motion_ids, motion_times, offset=None) -> Dict[str, torch.Tensor]:
        out = ops.motion_state(self._packed, motion_ids, motion_times, offset)
        out["motion_aa"] = self._motion_aa[self._frame_index(motion_ids, motion_times)]
        out["motion_bodies"] = self._motion_bodies[motion_ids]
        out["motion_limb_weights"] = self._motion_limb_weights[motion_ids]
        return out

    def _frame_index(self, motion_ids, motion_times):
        """frame_idx0 + length_starts of _calc_frame_blend (motion_lib_base.py:549-559), for the `motion_aa` row."""
        ln, nf = self._motion_lengths[motion_ids], self._motion_num_frames[motion_ids]
        phase = torch.clip(motion_times / ln, 0.0, 1.0)
        return (phase * (nf - 1)).long() + self.length_starts[motion_ids]

    def sample_motions(self, n):
        return torch.multinomial(self._sampling_batch_prob, num_samples=n, replacement=True).to(self._device)

    def sample_time(self, motion_ids, truncate_time=None):
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len = motion_len - truncate_time
        return phase * motion_len

    def sample_time_interval(self, motion_ids, truncate_time=None):
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len = motion_len - truncate_time
        curr_fps = 1 / 30
        return ((phase * motion_len) / curr_fps).long() * curr_fps

    def get_motion_length(self, motion_ids=None):
        return self._motion_lengths if motion_ids is None else self._motion_lengths[motion_ids]

    def get_motion_num_steps(self, motion_ids=None):
        nf = self._motion_num_frames if motion_ids is None else self._motion_num_frames[motion_ids]
        fps = self._motion_fps if motion_ids is None else self._motion_fps[motion_ids]
        return (nf * self._sim_fps / fps).ceil().int()

    # ---- Auto-PMCP sampling weights (motion_lib_base.py:350-389) -------------------------------------------------------
    def update_hard_sampling_weight(self, failed_keys: List):
        if len(failed_keys) > 0:
            all_keys = self._motion_data_keys.tolist()
            indexes = [all_keys.index(k) for k in failed_keys]
            self._sampling_prob[:] = 0
            self._sampling_prob[indexes] = 1 / len(indexes)
        else:
            self._sampling_prob = torch.ones(self._num_unique_motions, device=self._device) / self._num_unique_motions

    def update_soft_sampling_weight(self, failed_keys: List):
        if len(failed_keys) > 0:
            all_keys = self._motion_data_keys.tolist()
            indexes = [all_keys.index(k) for k in failed_keys]
            self._termination_history[indexes] += 1
            self.update_sampling_prob(self._termination_history)
        else:
            self._sampling_prob = torch.ones(self._num_unique_motions, device=self._device) / self._num_unique_motions

    def update_sampling_prob(self, termination_history):
        if len(termination_history) == len(self._termination_history) and termination_history.sum() > 0:
            self._sampling_prob[:] = termination_history / termination_history.sum()
            self._termination_history = termination_history
            return True
        return False
