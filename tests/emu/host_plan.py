"""TEST INFRASTRUCTURE: run `phc_b200.ops.EnvStepPlan`'s own argument assembly on HOST tensors and hand the resulting PhcStepArgs
to the CPU emulation of the kernel (tests/emu/build_emu.py).  Only what needs a device is replaced: the motion-library pack
(numpy here, the layout of phc_motion_pack), the per-env motion records (phc_env_motion_gather) and the launch itself."""
import contextlib
import ctypes as C

import numpy as np
import torch

from phc_b200 import _lib, ops


def round4(x):
    return (x + 3) & ~3


def host_pack(gts, grs, gvs, gavs, lengths, num_frames, dts, length_starts, num_ext=0, num_dofs=0):
    """phc_motion_pack on the host: frames_body[F, round4(13 * JE)] = per body pos3 rot4 vel3 angvel3 (include/phc_b200.h)."""
    F, JE = gts.shape[0], gts.shape[1]
    bs = round4(13 * JE)
    fb = torch.zeros(F, bs, dtype=torch.float32)
    rec = torch.cat((gts, grs, gvs, gavs), dim=-1).reshape(F, 13 * JE)
    fb[:, :13 * JE] = rec
    J = JE - num_ext
    js = round4(2 * num_dofs) if num_dofs else round4(4 * J + 3 * (J - 1))
    lengths, dts = lengths.float().contiguous(), dts.float().contiguous()
    num_frames, length_starts = num_frames.long().contiguous(), length_starts.long().contiguous()
    c = _lib.PhcMotionLib(fb.data_ptr(), None, lengths.data_ptr(), dts.data_ptr(), num_frames.data_ptr(), length_starts.data_ptr(),
                          F, int(lengths.shape[0]), J, bs, js, num_ext, num_dofs)
    return ops.PackedMotionLib(fb, None, lengths, dts, num_frames, length_starts, J, c, num_ext, num_dofs)


@contextlib.contextmanager
def host_mode():
    """ops._req without the CUDA requirement; per-env motion records gathered with torch on the host."""
    real_req, real_refresh = ops._req, ops.EnvStepPlan.refresh_motion_params

    def req(t, dtype, name, device=None):
        assert torch.is_tensor(t) and t.dtype == dtype and t.is_contiguous(), name
        return t

    def refresh(self):
        ids = self._keep["motion_ids"]
        m = self.mlib
        em = self._env_motion
        em[:, 0] = m.lengths[ids].view(torch.int32)
        em[:, 1] = m.dts[ids].view(torch.int32)
        em[:, 2] = m.num_frames[ids].to(torch.int32)
        em[:, 3] = m.length_starts[ids].to(torch.int32)

    ops._req, ops.EnvStepPlan.refresh_motion_params = req, refresh
    try:
        yield
    finally:
        ops._req, ops.EnvStepPlan.refresh_motion_params = real_req, real_refresh


class Emu:
    VARIANTS = {"smpl": 0, "fast": 1, "getup": 2, "generic": 3, "fut": 4, "getup_generic": 5, "wide": 6, "fastk": 7}

    def __init__(self, so_path):
        self.lib = C.CDLL(so_path)
        self.lib.emu_env_step.restype = C.c_int
        self.lib.emu_env_step.argtypes = [C.POINTER(_lib.PhcStepArgs)] + [C.c_int] * 6
        self.lib.emu_fast_flags.restype = C.c_uint32

    def run(self, plan, variant):
        a = plan.args
        real = _lib.load()
        J, T = a.lib.num_bodies, a.time_steps
        E, DR = a.lib.num_ext_bodies, a.lib.num_dofs
        n_shape = a.num_shape if a.shape_params else 0
        n_limb = a.num_limb if a.limb_weights else 0
        self_dim = real.phc_self_obs_dim(J, a.flags) + n_shape + n_limb
        obs_dim = self_dim + real.phc_task_obs_dim(a.num_track if a.num_track > 0 else J, T)
        amp_dim = 0 if not a.amp_out else (real.phc_amp_obs_dim_robot(DR, a.num_key_bodies, a.flags) if DR > 0
                                           else real.phc_amp_obs_dim(a.num_amp_joints, a.num_key_bodies, a.flags))
        # the launcher's derived arguments (phc_env_step in env_step.cu)
        alias_obs = 2 * a.lib.body_stride + round4(J * 13) >= round4(obs_dim)
        state_bulk_ok = (a.body_state % 16 == 0) and ((a.bodies_per_env * 13) % 4 == 0) and ((J * 13) % 4 == 0)
        rc = self.lib.emu_env_step(C.byref(a), obs_dim, self_dim, amp_dim, int(alias_obs), int(state_bulk_ok), self.VARIANTS[variant])
        assert rc == 0
