"""Deterministic synthetic inputs for the PHC hot path.

There is no simulator (Isaac Gym) and no AMASS data on the build or GPU boxes, so the bench, the
parity tests and smoke() drive the path with seeded synthetic data of the real shapes:

* motion clips: random-walk local joint rotations pushed through forward kinematics on the 24-body
  SMPL-MuJoCo tree (same body order as phc/data/assets/mjcf/smpl_humanoid.xml), 30 fps, with
  finite-difference velocities -- laid out as the flat frame tables MotionLibBase keeps
  (reference: phc/utils/motion_lib_base.py:300-307);
* simulator state [N, J, 13]: reference pose + noise, 5 % of bodies exactly on the reference rotation
  (exercises the sin(theta) <= 1e-5 mask) and 5 % fully random rotations (exercises angle wrap / slerp flip).

Everything is generated with torch on the CPU from an explicit seed and then moved to the target device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

SMPL_BODY_NAMES = ["Pelvis", "L_Hip", "L_Knee", "L_Ankle", "L_Toe", "R_Hip", "R_Knee", "R_Ankle", "R_Toe",
                   "Torso", "Spine", "Chest", "Neck", "Head", "L_Thorax", "L_Shoulder", "L_Elbow", "L_Wrist",
                   "L_Hand", "R_Thorax", "R_Shoulder", "R_Elbow", "R_Wrist", "R_Hand"]
SMPL_PARENTS = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
# rough bone offsets (metres) of an average SMPL body in its parent frame; only used to make clips plausible
_SMPL_OFFSETS = [
    (0, 0, 0), (0.0, 0.07, -0.09), (0.0, 0.03, -0.38), (0.0, -0.01, -0.40), (0.12, 0.02, -0.06),
    (0.0, -0.07, -0.09), (0.0, -0.03, -0.38), (0.0, 0.01, -0.40), (0.12, -0.02, -0.06),
    (-0.02, 0.0, 0.12), (0.0, 0.0, 0.14), (0.0, 0.0, 0.06), (-0.03, 0.0, 0.21), (0.05, 0.0, 0.09),
    (-0.02, 0.08, 0.11), (-0.01, 0.09, 0.03), (-0.02, 0.26, -0.01), (0.0, 0.25, 0.01), (-0.01, 0.08, -0.01),
    (-0.02, -0.08, 0.11), (-0.01, -0.09, 0.03), (-0.02, -0.26, -0.01), (0.0, -0.25, 0.01), (-0.01, -0.08, -0.01),
]
# env_im.yaml: key_bodies / reset_bodies, resolved against SMPL_BODY_NAMES
SMPL_KEY_BODIES = [7, 3, 22, 17]
SMPL_RESET_BODIES = [0, 1, 2, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23]
# has_dof_subset: drop L_Toe, R_Toe, L_Hand, R_Hand joints from the AMP observation (humanoid.py:388-413)
SMPL_DOF_SUBSET = [i for j in range(23) if SMPL_BODY_NAMES[j + 1] not in ("L_Toe", "R_Toe", "L_Hand", "R_Hand")
                   for i in (3 * j, 3 * j + 1, 3 * j + 2)]


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz), dim=-1)


def _qconj(q):
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def _qrot(q, v):
    u, w = q[..., :3], q[..., 3:4]
    t = 2.0 * torch.linalg.cross(u, v, dim=-1)
    return v + w * t + torch.linalg.cross(u, t, dim=-1)


def _qexp(e):
    """exp map (axis * angle) -> unit quaternion xyzw."""
    ang = e.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    k = torch.where(ang > 1e-8, torch.sin(half) / ang.clamp_min(1e-8), 0.5 * torch.ones_like(ang))
    return torch.cat((e * k, torch.cos(half)), dim=-1)


def _qlog(q):
    """unit quaternion -> exp map with angle in (-pi, pi]."""
    q = torch.where(q[..., 3:4] < 0, -q, q)
    s = q[..., :3].norm(dim=-1, keepdim=True)
    ang = 2.0 * torch.atan2(s, q[..., 3:4])
    k = torch.where(s > 1e-8, ang / s.clamp_min(1e-8), 2.0 * torch.ones_like(s))
    return q[..., :3] * k


def _unit(q):
    return q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)


@dataclass
class MotionData:
    """Flat frame tables exactly as MotionLibBase holds them (motion_lib_base.py:300-307), fp32 / int64."""
    gts: torch.Tensor            # [F, J, 3] global body translation
    grs: torch.Tensor            # [F, J, 4] global body rotation (xyzw)
    lrs: torch.Tensor            # [F, J, 4] local (parent-relative) rotation
    gvs: torch.Tensor            # [F, J, 3] global linear velocity
    gavs: torch.Tensor           # [F, J, 3] global angular velocity
    dvs: torch.Tensor            # [F, J-1, 3] joint (dof) velocity
    lengths: torch.Tensor        # [M] clip length in seconds = dt * (num_frames - 1)
    num_frames: torch.Tensor     # [M] int64
    dts: torch.Tensor            # [M] seconds per frame
    length_starts: torch.Tensor  # [M] int64 first row of each clip in the tables

    def to(self, device):
        return MotionData(*[getattr(self, f).to(device) for f in self.__dataclass_fields__])

    @property
    def num_motions(self):
        return int(self.lengths.shape[0])

    @property
    def num_bodies(self):
        return int(self.gts.shape[1])


def make_motions(num_motions: int, seed: int = 0, num_bodies: int = 24, min_frames: int = 60,
                 max_frames: int = 300, fps: float = 30.0) -> MotionData:
    """Random-walk clips through FK.  num_bodies != 24 builds a chain-like tree of that size (H1/G1/SMPL-X shapes)."""
    g = torch.Generator().manual_seed(seed)
    J = num_bodies
    if J == 24:
        parents, offs = SMPL_PARENTS, torch.tensor(_SMPL_OFFSETS, dtype=torch.float32)
    else:
        parents = [-1] + [max(0, j - 1 - (j % 3 == 0) * 2) for j in range(1, J)]
        offs = torch.randn(J, 3, generator=g) * 0.15
        offs[0] = 0
    nfr = torch.randint(min_frames, max_frames + 1, (num_motions,), generator=g)
    starts = torch.cumsum(nfr, 0) - nfr
    F = int(nfr.sum())
    dt = 1.0 / fps

    # one long random walk, cut into clips (clips only need to be smooth inside themselves)
    step = torch.randn(F, J, 3, generator=g) * 0.04
    clip_id = torch.repeat_interleave(torch.arange(num_motions), nfr)
    first = torch.zeros(F, dtype=torch.bool)
    first[starts] = True
    step[first] = torch.randn(num_motions, J, 3, generator=g) * 0.5      # initial pose of each clip
    # segmented cumulative sum
    csum = torch.cumsum(step, 0)
    base = csum[starts] - step[starts]
    walk = csum - base[clip_id]
    walk[:, 0, :2] *= 0.3                                                  # keep the root roughly upright
    lrs = _unit(_qexp(walk))
    root_step = torch.randn(F, 3, generator=g) * 0.02
    root_step[first] = torch.randn(num_motions, 3, generator=g) * torch.tensor([1.0, 1.0, 0.05]) + torch.tensor([0, 0, 0.9])
    rc = torch.cumsum(root_step, 0)
    root = rc - (rc[starts] - root_step[starts])[clip_id]

    grs = torch.empty(F, J, 4)
    gts = torch.empty(F, J, 3)
    for j in range(J):
        p = parents[j]
        if p < 0:
            grs[:, j] = lrs[:, j]
            gts[:, j] = root
        else:
            grs[:, j] = _unit(_qmul(grs[:, p], lrs[:, j]))
            gts[:, j] = gts[:, p] + _qrot(grs[:, p], offs[j].expand(F, 3))

    # forward differences inside each clip (last frame repeats the previous velocity)
    nxt = torch.arange(F) + 1
    last = torch.zeros(F, dtype=torch.bool)
    last[starts + nfr - 1] = True
    nxt = torch.where(last, torch.arange(F), nxt)
    prv = torch.where(last, torch.arange(F) - 1, torch.arange(F))
    gvs = (gts[nxt] - gts[prv]) / dt
    gavs = _qlog(_qmul(grs[nxt], _qconj(grs[prv]))) / dt
    dvs = _qlog(_qmul(_qconj(lrs[prv, 1:]), lrs[nxt, 1:])) / dt
    return MotionData(gts=gts.contiguous(), grs=grs.contiguous(), lrs=lrs.contiguous(), gvs=gvs.contiguous(),
                      gavs=gavs.contiguous(), dvs=dvs.contiguous(),
                      lengths=(dt * (nfr - 1)).float(), num_frames=nfr.long(),
                      dts=torch.full((num_motions,), dt, dtype=torch.float32), length_starts=starts.long())


@dataclass
class EnvState:
    """Per-step simulator-side inputs of the env hot path (layout contract: phc/env/tasks/humanoid.py:179-247)."""
    body_state: torch.Tensor      # [N, J, 13] pos(0:3) rot(3:7) vel(7:10) ang_vel(10:13)
    dof_state: torch.Tensor       # [N, D, 2]  (pos, vel) interleaved
    dof_force: torch.Tensor       # [N, D]
    progress: torch.Tensor        # [N] int64 (already incremented for this step)
    motion_ids: torch.Tensor      # [N] int64
    start_times: torch.Tensor     # [N] fp32  _motion_start_times
    start_offsets: torch.Tensor   # [N] fp32  _motion_start_times_offset
    global_offset: torch.Tensor   # [N, 3]
    amp_hist: torch.Tensor        # [N, S, A] newest-first AMP window before this step

    def to(self, device):
        return EnvState(*[getattr(self, f).to(device) for f in self.__dataclass_fields__])


def _lerp_state(m: MotionData, ids, t):
    """Cheap reference pose (nearest lower frame; good enough to centre the synthetic simulator state)."""
    ph = torch.clip(t / m.lengths[ids], 0, 1)
    f = (ph * (m.num_frames[ids] - 1)).long() + m.length_starts[ids]
    return m.gts[f], m.grs[f], m.gvs[f], m.gavs[f], m.lrs[f], m.dvs[f]


def make_env_state(m: MotionData, num_envs: int, seed: int = 0, amp_dim: int = 196, amp_steps: int = 10,
                   dt: float = 1.0 / 30.0, max_progress: int = 299, with_offset: bool = False,
                   blend_jitter: bool = False) -> EnvState:
    """Seeded simulator state around the reference pose at the current motion time."""
    g = torch.Generator().manual_seed(1000 + seed)
    N, J = num_envs, m.num_bodies
    ids = (torch.arange(N) % m.num_motions).long()
    ln = m.lengths[ids]
    # sample_time_interval: start snapped to the 1/30 grid (motion_lib_base.py:414-423)
    start = ((torch.rand(N, generator=g) * ln) / (1 / 30)).long() * (1 / 30)
    start = start.float()
    if blend_jitter:                                   # leave the frame grid so blend is a generic value
        start = start + torch.rand(N, generator=g) * (1 / 30)
    progress = torch.randint(0, max_progress + 1, (N,), generator=g).long()
    progress[: max(1, N // 16)] = torch.arange(max(1, N // 16)) % 5          # exercises progress<=3 / <=1 masks
    off = torch.zeros(N)
    goff = torch.zeros(N, 3)
    if with_offset:
        goff[:, :2] = torch.randn(N, 2, generator=g)
    t_now = progress * dt + start + off
    p, q, v, w, lq, dv = _lerp_state(m, ids, t_now)
    p = p + goff[:, None, :]

    pos = p + torch.randn(N, J, 3, generator=g) * 0.05
    small = _unit(torch.cat((torch.randn(N, J, 3, generator=g) * 0.05, torch.ones(N, J, 1)), dim=-1))
    rot = _unit(_qmul(q, small))
    u = torch.rand(N, J, generator=g)
    rot = torch.where((u < 0.05)[..., None], q, rot)
    rnd = _unit(torch.randn(N, J, 4, generator=g))
    rot = torch.where((u > 0.95)[..., None], rnd, rot)
    vel = v + torch.randn(N, J, 3, generator=g) * 0.5
    ang = w + torch.randn(N, J, 3, generator=g) * 0.5
    # a few envs far from the reference so early termination fires
    far = torch.rand(N, generator=g) < 0.1
    pos = pos + far[:, None, None] * torch.randn(N, J, 3, generator=g) * 0.3
    body = torch.cat((pos, rot, vel, ang), dim=-1).contiguous()

    dof_pos = _qlog(lq[:, 1:]).reshape(N, -1) + torch.randn(N, (J - 1) * 3, generator=g) * 0.05
    dof_pos[0, :3] = 0.0                                                    # exact-zero exp map (default-axis branch)
    dof_vel = dv.reshape(N, -1) + torch.randn(N, (J - 1) * 3, generator=g) * 0.5
    dof_state = torch.stack((dof_pos, dof_vel), dim=-1).contiguous()
    dof_force = torch.randn(N, (J - 1) * 3, generator=g) * 50.0
    amp_hist = torch.randn(N, amp_steps, amp_dim, generator=g)
    return EnvState(body_state=body.float(), dof_state=dof_state.float(), dof_force=dof_force.float(),
                    progress=progress, motion_ids=ids, start_times=start, start_offsets=off,
                    global_offset=goff, amp_hist=amp_hist.float())


def make_rollout(num_envs: int, horizon: int, seed: int = 0):
    """Synthetic PPO rollout scalars for the GAE / advantage kernels: time-major [T, N, 1]."""
    g = torch.Generator().manual_seed(2000 + seed)
    T, N = horizon, num_envs
    rewards = torch.rand(T, N, 1, generator=g)
    values = torch.randn(T, N, 1, generator=g)
    next_values = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, generator=g) < 0.03)
    return dones.float(), values, rewards, next_values


# ----------------------------------------------------------------------------------------------------------------------
# Hinge-joint robots (Unitree H1: phc/data/cfg/robot/unitree_h1.yaml:27-67): 20 bodies, 19 one-dof joints, 3 "extend"
# bodies (hands, head) rigidly attached to a parent; reference motion in the layout of phc/utils/motion_lib_real.py
# ----------------------------------------------------------------------------------------------------------------------
H1_NUM_BODIES, H1_NUM_DOFS = 20, 19
H1_EXT_PARENTS = [15, 19, 0]                                   # left_elbow_link, right_elbow_link, pelvis
H1_EXT_POS = [[0.3, 0.0, 0.0], [0.3, 0.0, 0.0], [0.0, 0.0, 0.6]]
H1_KEY_BODIES = [5, 10, 15, 19]                                # ankles and elbows
# Unitree G1 (phc/data/cfg/robot/unitree_g1.yaml): 38 bodies, 37 hinge dofs, one extend body above the pelvis
G1_NUM_BODIES, G1_NUM_DOFS = 38, 37
G1_EXT_PARENTS = [0]
G1_EXT_POS = [[0.0, 0.0, 0.4]]
G1_KEY_BODIES = [6, 12, 18, 30]                                # ankle roll links and elbow roll links
SMPLX_KEY_BODIES = [8, 4, 41, 26]                              # four end-effector-like bodies of a 52-body tree


@dataclass
class RobotMotionData:
    gts_t: torch.Tensor          # [F, J+E, 3]  bodies then extend bodies (the *_t tables of motion_lib_real)
    grs_t: torch.Tensor          # [F, J+E, 4]
    gvs_t: torch.Tensor          # [F, J+E, 3]
    gavs_t: torch.Tensor         # [F, J+E, 3]
    dof_pos: torch.Tensor        # [F, D]
    dvs: torch.Tensor            # [F, D] dof velocity
    lengths: torch.Tensor
    num_frames: torch.Tensor
    dts: torch.Tensor
    length_starts: torch.Tensor
    num_bodies: int = H1_NUM_BODIES

    @property
    def num_motions(self):
        return int(self.lengths.shape[0])

    @property
    def num_ext(self):
        return int(self.gts_t.shape[1]) - self.num_bodies

    @property
    def num_dofs(self):
        return int(self.dof_pos.shape[1])


def make_robot_motions(num_motions: int, seed: int = 0, num_bodies: int = H1_NUM_BODIES, num_dofs: int = H1_NUM_DOFS,
                       ext_parents: Sequence[int] = tuple(H1_EXT_PARENTS), ext_pos=H1_EXT_POS, min_frames: int = 60,
                       max_frames: int = 300) -> RobotMotionData:
    m = make_motions(num_motions, seed=seed, num_bodies=num_bodies, min_frames=min_frames, max_frames=max_frames)
    F, J, E = m.gts.shape[0], num_bodies, len(ext_parents)
    par = torch.tensor(list(ext_parents))
    off = torch.tensor(ext_pos, dtype=torch.float32)
    ep = m.gts[:, par] + _qrot(m.grs[:, par], off[None].expand(F, E, 3))
    eq = m.grs[:, par]
    dt = float(m.dts[0])
    clip_id = torch.repeat_interleave(torch.arange(num_motions), m.num_frames)
    last = torch.zeros(F, dtype=torch.bool)
    last[m.length_starts + m.num_frames - 1] = True
    nxt = torch.where(last, torch.arange(F), torch.arange(F) + 1)
    prv = torch.where(last, torch.arange(F) - 1, torch.arange(F))
    ev = (ep[nxt] - ep[prv]) / dt
    ew = m.gavs[:, par]
    g = torch.Generator().manual_seed(7000 + seed)
    step = torch.randn(F, num_dofs, generator=g) * 0.03
    first = torch.zeros(F, dtype=torch.bool)
    first[m.length_starts] = True
    step[first] = torch.randn(num_motions, num_dofs, generator=g) * 0.4
    cs = torch.cumsum(step, 0)
    dof_pos = cs - (cs[m.length_starts] - step[m.length_starts])[clip_id]
    dvs = (dof_pos[nxt] - dof_pos[prv]) / dt
    cat = lambda a, b: torch.cat((a, b), dim=1).contiguous()
    return RobotMotionData(gts_t=cat(m.gts, ep), grs_t=cat(m.grs, eq), gvs_t=cat(m.gvs, ev), gavs_t=cat(m.gavs, ew),
                           dof_pos=dof_pos.float().contiguous(), dvs=dvs.float().contiguous(), lengths=m.lengths,
                           num_frames=m.num_frames, dts=m.dts, length_starts=m.length_starts, num_bodies=J)


def make_robot_env_state(m: RobotMotionData, num_envs: int, seed: int = 0, amp_dim: int = 63, amp_steps: int = 10,
                         dt: float = 1.0 / 30.0, max_progress: int = 299, with_offset: bool = False,
                         blend_jitter: bool = False) -> EnvState:
    """Simulator state of a hinge-joint robot around its reference pose: [N, J, 13] bodies, [N, D, 2] dofs."""
    g = torch.Generator().manual_seed(3000 + seed)
    N, J, D = num_envs, m.num_bodies, m.num_dofs
    ids = (torch.arange(N) % m.num_motions).long()
    ln = m.lengths[ids]
    start = (((torch.rand(N, generator=g) * ln) / (1 / 30)).long() * (1 / 30)).float()
    if blend_jitter:
        start = start + torch.rand(N, generator=g) * (1 / 30)
    progress = torch.randint(0, max_progress + 1, (N,), generator=g).long()
    progress[: max(1, N // 16)] = torch.arange(max(1, N // 16)) % 5
    off, goff = torch.zeros(N), torch.zeros(N, 3)
    if with_offset:
        goff[:, :2] = torch.randn(N, 2, generator=g)
    t_now = progress * dt + start + off
    f = (torch.clip(t_now / ln, 0, 1) * (m.num_frames[ids] - 1)).long() + m.length_starts[ids]
    p, q, v, w = m.gts_t[f, :J] + goff[:, None, :], m.grs_t[f, :J], m.gvs_t[f, :J], m.gavs_t[f, :J]
    pos = p + torch.randn(N, J, 3, generator=g) * 0.05
    small = _unit(torch.cat((torch.randn(N, J, 3, generator=g) * 0.05, torch.ones(N, J, 1)), dim=-1))
    rot = _unit(_qmul(q, small))
    u = torch.rand(N, J, generator=g)
    rot = torch.where((u < 0.05)[..., None], q, rot)
    rot = torch.where((u > 0.95)[..., None], _unit(torch.randn(N, J, 4, generator=g)), rot)
    vel = v + torch.randn(N, J, 3, generator=g) * 0.5
    ang = w + torch.randn(N, J, 3, generator=g) * 0.5
    far = torch.rand(N, generator=g) < 0.1
    pos = pos + far[:, None, None] * torch.randn(N, J, 3, generator=g) * 0.3
    body = torch.cat((pos, rot, vel, ang), dim=-1).contiguous()
    dof_pos = m.dof_pos[f] + torch.randn(N, D, generator=g) * 0.05
    dof_vel = m.dvs[f] + torch.randn(N, D, generator=g) * 0.5
    dof_state = torch.stack((dof_pos, dof_vel), dim=-1).contiguous()
    dof_force = torch.randn(N, D, generator=g) * 50.0
    amp_hist = torch.randn(N, amp_steps, amp_dim, generator=g)
    return EnvState(body_state=body.float(), dof_state=dof_state.float(), dof_force=dof_force.float(), progress=progress,
                    motion_ids=ids, start_times=start, start_offsets=off, global_offset=goff, amp_hist=amp_hist.float())
