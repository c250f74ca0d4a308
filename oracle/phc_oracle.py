"""CPU oracle for the PHC hot path (TEST INFRASTRUCTURE -- never imported by the product path).

A plain fp32 torch-CPU restatement of the reference algorithm for every row of SURVEY.md section 8(a).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; `phc_b200/` must never do so (tests/test_no_oracle_in_product.py enforces it).

Pinned: every function below is checked against outputs of the UNMODIFIED reference
(ZhengyiLuo/PHC @ /root/reference, imported through tests/golden/ref_shim.py) on seeded inputs;
the vectors are committed under tests/golden/*.npz (generator: tests/golden/make_golden.py) and
replayed by tests/test_oracle_golden.py.  The rl_games-provided pieces (Gaussian neglogp,
policy_kl, Adam/clip wiring) have no in-tree reference: they are restated from rl_games==1.1.4
semantics and self-pinned against torch.distributions / torch.optim (header of each function).

Conventions (reference: phc/utils/isaacgym_torch_utils.py): quaternions are xyzw, fp32,
operation ORDER follows the reference expression by expression because `2*acos(w)` is
ill-conditioned near identity (SURVEY.md section 7 "Hard parts").
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# quaternion library  (reference: phc/utils/isaacgym_torch_utils.py, phc/utils/torch_utils.py)
# ----------------------------------------------------------------------------------------------
def qmul(a: Tensor, b: Tensor) -> Tensor:
    """Hamilton product, xyzw, 9-multiply factored form -- isaacgym_torch_utils.py:25-45."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    ww = (az + ax) * (bx + by)
    yy = (aw - ay) * (bw + bz)
    zz = (aw + ay) * (bw - bz)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (az - ax) * (bx - by))
    w = qq - ww + (az - ay) * (by - bz)
    x = qq - xx + (ax + aw) * (bx + bw)
    y = qq - yy + (aw - ax) * (by + bz)
    z = qq - zz + (az + ay) * (bw - bx)
    return torch.stack((x, y, z, w), dim=-1)


def qconj(q: Tensor) -> Tensor:
    """isaacgym_torch_utils.py:93-96."""
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def qrot(q: Tensor, v: Tensor) -> Tensor:
    """my_quat_rotate -- torch_utils.py:46-55:  v(2w^2-1) + 2w(qv x v) + 2 qv (qv.v)."""
    w = q[..., 3:4]
    u = q[..., :3]
    a = v * (2.0 * w * w - 1.0)
    b = torch.linalg.cross(u, v, dim=-1) * w * 2.0
    c = u * (u * v).sum(-1, keepdim=True) * 2.0
    return a + b + c


def _unit(x: Tensor, eps: float = 1e-9) -> Tensor:
    """normalize -- isaacgym_torch_utils.py:49-50."""
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def q_from_angle_axis(angle: Tensor, axis: Tensor) -> Tensor:
    """isaacgym_torch_utils.py:104-108."""
    th = (angle / 2).unsqueeze(-1)
    return _unit(torch.cat((_unit(axis) * th.sin(), th.cos()), dim=-1))


def wrap_angle(x: Tensor) -> Tensor:
    """normalize_angle -- isaacgym_torch_utils.py:111-112."""
    return torch.atan2(torch.sin(x), torch.cos(x))


def heading_angle(q: Tensor) -> Tensor:
    """calc_heading -- torch_utils.py:200-212 (x axis rotated by q, projected on xy)."""
    ex = torch.zeros_like(q[..., :3])
    ex[..., 0] = 1
    d = qrot(q, ex)
    return torch.atan2(d[..., 1], d[..., 0])


def heading_q(q: Tensor, inverse: bool) -> Tensor:
    """calc_heading_quat / calc_heading_quat_inv -- torch_utils.py:215-240."""
    h = heading_angle(q)
    ez = torch.zeros_like(q[..., :3])
    ez[..., 2] = 1
    return q_from_angle_axis(-h if inverse else h, ez)


def tan_norm(q: Tensor) -> Tensor:
    """quat_to_tan_norm -- torch_utils.py:101-113: [q*(1,0,0), q*(0,0,1)]."""
    ex = torch.zeros_like(q[..., :3])
    ex[..., 0] = 1
    ez = torch.zeros_like(q[..., :3])
    ez[..., 2] = 1
    return torch.cat((qrot(q, ex), qrot(q, ez)), dim=-1)


def q_to_angle_axis(q: Tensor) -> Tuple[Tensor, Tensor]:
    """quat_to_angle_axis -- torch_utils.py:58-78."""
    w = q[..., 3]
    s = torch.sqrt(1 - w * w)
    ang = wrap_angle(2 * torch.acos(w))
    axis = q[..., :3] / s.unsqueeze(-1)
    keep = s.abs() > 1e-5
    ez = torch.zeros_like(axis)
    ez[..., 2] = 1
    ang = torch.where(keep, ang, torch.zeros_like(ang))
    axis = torch.where(keep.unsqueeze(-1), axis, ez)
    return ang, axis


def q_to_exp_map(q: Tensor) -> Tensor:
    """quat_to_exp_map -- torch_utils.py:91-98."""
    ang, axis = q_to_angle_axis(q)
    return ang.unsqueeze(-1) * axis


def exp_map_to_q(e: Tensor) -> Tensor:
    """exp_map_to_quat via exp_map_to_angle_axis -- torch_utils.py:147-173."""
    ang = torch.norm(e, dim=-1)
    axis = e / ang.unsqueeze(-1)
    ang = wrap_angle(ang)
    keep = ang.abs() > 1e-5
    ez = torch.zeros_like(e)
    ez[..., 2] = 1
    ang = torch.where(keep, ang, torch.zeros_like(ang))
    axis = torch.where(keep.unsqueeze(-1), axis, ez)
    return q_from_angle_axis(ang, axis)


def slerp(q0: Tensor, q1: Tensor, t: Tensor) -> Tensor:
    """torch_utils.py:176-197.  t broadcasts against [..., 1]."""
    c = (q0 * q1).sum(-1, keepdim=True)
    q1 = torch.where(c < 0, -q1, q1)
    c = c.abs()
    half = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    ra = torch.sin((1 - t) * half) / s
    rb = torch.sin(t * half) / s
    out = ra * q0 + rb * q1
    out = torch.where(s.abs() < 0.001, 0.5 * q0 + 0.5 * q1, out)
    out = torch.where(c.abs() >= 1, q0, out)
    return out


_BASE_ROT_CONJ = (-0.5, -0.5, -0.5, 0.5)


def strip_base_rot(q: Tensor) -> Tensor:
    """remove_base_rot -- phc/env/tasks/humanoid.py:1935-1939 (used when not upright)."""
    b = torch.tensor(_BASE_ROT_CONJ, dtype=q.dtype).expand_as(q)
    return qmul(q, b)


# ----------------------------------------------------------------------------------------------
# K1  self observation   (reference: phc/env/tasks/humanoid.py:1994-2050)
# ----------------------------------------------------------------------------------------------
def self_obs(body_pos: Tensor, body_rot: Tensor, body_vel: Tensor, body_ang_vel: Tensor,
             local_root_obs: bool = True, root_height_obs: bool = True, upright: bool = True,
             shape_params: Optional[Tensor] = None, limb_weights: Optional[Tensor] = None) -> Tensor:
    """compute_humanoid_observations_smpl_max (humanoid.py:1994-2050).  [N,J,*] -> [N, 1+15J-3 (+ shape columns + limb-weight columns)];
    shape_params / limb_weights are the has_smpl_params / has_limb_weight_params tails (:2043-2047)."""
    N, J, _ = body_pos.shape
    root_pos = body_pos[:, 0]
    root_rot = body_rot[:, 0]
    if not upright:
        root_rot = strip_base_rot(root_rot)
    hinv = heading_q(root_rot, inverse=True).unsqueeze(1)            # [N,1,4] broadcast over bodies
    lp = qrot(hinv, body_pos - root_pos.unsqueeze(1)).reshape(N, J * 3)[:, 3:]
    lr = tan_norm(qmul(hinv.expand(N, J, 4), body_rot)).reshape(N, J * 6)
    if not local_root_obs:
        lr = lr.clone()
        lr[:, 0:6] = tan_norm(root_rot)
    lv = qrot(hinv, body_vel).reshape(N, J * 3)
    lw = qrot(hinv, body_ang_vel).reshape(N, J * 3)
    cols = []
    if root_height_obs:
        cols.append(root_pos[:, 2:3])
    cols += [lp, lr, lv, lw]
    if shape_params is not None:
        cols.append(shape_params)
    if limb_weights is not None:
        cols.append(limb_weights)
    return torch.cat(cols, dim=-1)


# ----------------------------------------------------------------------------------------------
# K3  imitation (task) observation v6   (reference: phc/env/tasks/humanoid_im.py:1308-1358)
# ----------------------------------------------------------------------------------------------
def task_obs_v6(root_pos: Tensor, root_rot: Tensor, body_pos: Tensor, body_rot: Tensor, body_vel: Tensor,
                body_ang_vel: Tensor, ref_pos: Tensor, ref_rot: Tensor, ref_vel: Tensor, ref_ang_vel: Tensor,
                time_steps: int = 1, upright: bool = True) -> Tensor:
    """sim [N,J,*]; ref [N*T,J,*] (env-major, T samples per env) -> [N, T*J*24]."""
    N, J, _ = body_pos.shape
    T = time_steps
    if not upright:
        root_rot = strip_base_rot(root_rot)
    hinv = heading_q(root_rot, inverse=True).view(N, 1, 1, 4)
    h = heading_q(root_rot, inverse=False).view(N, 1, 1, 4)
    rp, rr = ref_pos.view(N, T, J, 3), ref_rot.view(N, T, J, 4)
    rv, rw = ref_vel.view(N, T, J, 3), ref_ang_vel.view(N, T, J, 3)
    bp, br = body_pos.view(N, 1, J, 3), body_rot.view(N, 1, J, 4)
    bv, bw = body_vel.view(N, 1, J, 3), body_ang_vel.view(N, 1, J, 3)
    hinv_e, h_e = hinv.expand(N, T, J, 4), h.expand(N, T, J, 4)

    d_pos = qrot(hinv, rp - bp)
    d_rot = tan_norm(qmul(qmul(hinv_e, qmul(rr, qconj(br.expand(N, T, J, 4)))), h_e))
    d_vel = qrot(hinv, rv - bv)
    d_ang = qrot(hinv, rw - bw)
    l_pos = qrot(hinv, rp - root_pos.view(N, 1, 1, 3))
    l_rot = tan_norm(qmul(hinv_e, rr))
    parts = [x.reshape(N, T, -1) for x in (d_pos, d_rot, d_vel, d_ang, l_pos, l_rot)]
    return torch.cat(parts, dim=-1).reshape(N, -1)


# ----------------------------------------------------------------------------------------------
# K4  tracking reward   (reference: humanoid_im.py:1523-1554 + power term :939-946)
# ----------------------------------------------------------------------------------------------
DEFAULT_RWD = dict(k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1,
                   w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1)   # humanoid_im.py:57


def imitation_reward(body_pos, body_rot, body_vel, body_ang_vel, ref_pos, ref_rot, ref_vel, ref_ang_vel,
                     spec: Dict[str, float] = DEFAULT_RWD) -> Tuple[Tensor, Tensor]:
    e_pos = ((ref_pos - body_pos) ** 2).mean(-1).mean(-1)
    ang = q_to_angle_axis(qmul(ref_rot, qconj(body_rot)))[0]
    e_rot = (ang ** 2).mean(-1)
    e_vel = ((ref_vel - body_vel) ** 2).mean(-1).mean(-1)
    e_ang = ((ref_ang_vel - body_ang_vel) ** 2).mean(-1).mean(-1)
    r = torch.stack((torch.exp(-spec["k_pos"] * e_pos), torch.exp(-spec["k_rot"] * e_rot),
                     torch.exp(-spec["k_vel"] * e_vel), torch.exp(-spec["k_ang_vel"] * e_ang)), dim=-1)
    rew = spec["w_pos"] * r[:, 0] + spec["w_rot"] * r[:, 1] + spec["w_vel"] * r[:, 2] + spec["w_ang_vel"] * r[:, 3]
    return rew, r


def power_reward(dof_force: Tensor, dof_vel: Tensor, progress: Tensor, coef: float) -> Tensor:
    """humanoid_im.py:939-946."""
    p = -coef * (dof_force * dof_vel).abs().sum(-1)
    return torch.where(progress <= 3, torch.zeros_like(p), p)


# ----------------------------------------------------------------------------------------------
# K5  reset / termination   (reference: humanoid_im.py:1580-1608, caller :1117-1190)
# ----------------------------------------------------------------------------------------------
def im_reset(progress: Tensor, body_pos_sub: Tensor, ref_pos_sub: Tensor, pass_time: Tensor,
             term_dist_sub: Tensor, early_term: bool = True, no_collision: bool = False,
             use_mean: bool = False) -> Tuple[Tensor, Tensor]:
    terminated = torch.zeros_like(progress)
    if early_term:
        d = torch.norm(body_pos_sub - ref_pos_sub, dim=-1)
        if use_mean:
            fallen = torch.any(d.mean(-1, keepdim=True) > term_dist_sub[0], dim=-1)
        else:
            fallen = torch.any(d > term_dist_sub, dim=-1)
        fallen = fallen & (progress > 1)
        if no_collision:
            fallen = torch.zeros_like(fallen)
        terminated = torch.where(fallen, torch.ones_like(progress), terminated)
    reset = torch.where(pass_time, torch.ones_like(progress), terminated)
    return reset, terminated


# ----------------------------------------------------------------------------------------------
# K6  AMP observation   (reference: phc/env/tasks/humanoid_amp.py:966-1011, humanoid.py:1755-1765)
# ----------------------------------------------------------------------------------------------
def amp_obs(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_pos,
            dof_subset: Optional[Tensor], local_root_obs: bool = True, root_height_obs: bool = True,
            upright: bool = True) -> Tensor:
    N = root_pos.shape[0]
    if not upright:
        root_rot = strip_base_rot(root_rot)
    hinv = heading_q(root_rot, inverse=True)
    rr = tan_norm(qmul(hinv, root_rot) if local_root_obs else root_rot)
    lv = qrot(hinv, root_vel)
    lw = qrot(hinv, root_ang_vel)
    lk = qrot(hinv.unsqueeze(1), key_pos - root_pos.unsqueeze(1)).reshape(N, -1)
    if dof_subset is not None:
        dof_pos = dof_pos[:, dof_subset]
        dof_vel = dof_vel[:, dof_subset]
    jo = tan_norm(exp_map_to_q(dof_pos.reshape(-1, 3))).reshape(N, -1)
    cols = [root_pos[:, 2:3]] if root_height_obs else []
    cols += [rr, lv, lw, jo, dof_vel, lk]
    return torch.cat(cols, dim=-1)


# ----------------------------------------------------------------------------------------------
# K2  motion library frame blend + state   (reference: phc/utils/motion_lib_base.py:437-567)
# ----------------------------------------------------------------------------------------------
class MotionTables:
    """The flat frame tables MotionLibBase builds at load time (motion_lib_base.py:300-307)."""

    def __init__(self, gts, grs, lrs, gvs, gavs, dvs, lengths, num_frames, dts, length_starts):
        self.gts, self.grs, self.lrs, self.gvs, self.gavs, self.dvs = gts, grs, lrs, gvs, gavs, dvs
        self.lengths, self.num_frames, self.dts, self.length_starts = lengths, num_frames, dts, length_starts


def frame_blend(time: Tensor, length: Tensor, num_frames: Tensor, dt: Tensor):
    """_calc_frame_blend -- motion_lib_base.py:549-559 (idx int64, fp32 op order preserved)."""
    phase = torch.clip(time / length, 0.0, 1.0)
    time = torch.where(time < 0, torch.zeros_like(time), time)
    i0 = (phase * (num_frames - 1)).long()
    i1 = torch.min(i0 + 1, num_frames - 1)
    blend = torch.clip((time - i0 * dt) / dt, 0.0, 1.0)
    return i0, i1, blend


def motion_state(tab: MotionTables, ids: Tensor, times: Tensor, offset: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """get_motion_state -- motion_lib_base.py:437-520 (SMPL variant: dof_pos from slerped local rotations)."""
    i0, i1, blend = frame_blend(times, tab.lengths[ids], tab.num_frames[ids], tab.dts[ids])
    f0 = i0 + tab.length_starts[ids]
    f1 = i1 + tab.length_starts[ids]
    b = blend.view(-1, 1, 1)
    lerp = lambda t: (1.0 - b) * t[f0] + b * t[f1]
    pos = lerp(tab.gts)
    if offset is not None:
        pos = pos + offset[:, None, :]
    vel, ang, dvel = lerp(tab.gvs), lerp(tab.gavs), lerp(tab.dvs)
    lrot = slerp(tab.lrs[f0], tab.lrs[f1], b)
    rot = slerp(tab.grs[f0], tab.grs[f1], b)
    dof_pos = q_to_exp_map(lrot[:, 1:]).reshape(len(ids), -1)
    return dict(root_pos=pos[:, 0].clone(), root_rot=rot[:, 0].clone(), dof_pos=dof_pos,
                root_vel=vel[:, 0].clone(), root_ang_vel=ang[:, 0].clone(),
                dof_vel=dvel.reshape(len(ids), -1), rg_pos=pos, rb_rot=rot, body_vel=vel, body_ang_vel=ang)


# ----------------------------------------------------------------------------------------------
# Whole env step after physics   (reference: humanoid.py:1634-1665 post_physics_step ->
#   humanoid_im.py:873-948 _compute_reward, :1117-1190 _compute_reset, :694-871 _compute_observations,
#   humanoid_amp.py:194-210, :662-707)
# ----------------------------------------------------------------------------------------------
class StepConfig:
    def __init__(self, dt=1.0 / 30.0, upright=True, local_root_obs=True, root_height_obs=True,
                 rwd=DEFAULT_RWD, power_reward=True, power_coef=0.0005, early_term=True, no_collision=False,
                 use_mean=False, key_bodies=(7, 3, 22, 17), reset_bodies=None, term_dist=0.25,
                 dof_subset=None, time_steps=1, traj_dt=0.0, num_amp_steps=10, track_bodies=None, full_body_reward=True):
        # env.trackBodies / env.full_body_reward (humanoid_im.py:64-66, :926-935): body ids of the tracked subset (None = all)
        self.track_bodies = None if track_bodies is None else list(track_bodies)
        self.full_body_reward = full_body_reward
        self.dt, self.upright, self.local_root_obs, self.root_height_obs = dt, upright, local_root_obs, root_height_obs
        self.rwd, self.power_reward, self.power_coef = dict(rwd), power_reward, power_coef
        self.early_term, self.no_collision, self.use_mean = early_term, no_collision, use_mean
        self.key_bodies, self.reset_bodies, self.term_dist = list(key_bodies), reset_bodies, term_dist
        self.dof_subset, self.time_steps, self.traj_dt, self.num_amp_steps = dof_subset, time_steps, traj_dt, num_amp_steps


def env_step(tab: MotionTables, cfg: StepConfig, body_state: Tensor, dof_state: Tensor, dof_force: Tensor,
             progress: Tensor, motion_ids: Tensor, start_times: Tensor, start_offsets: Tensor,
             global_offset: Tensor, amp_hist: Tensor, occlusion: Optional[Tensor] = None, shape_params: Optional[Tensor] = None,
             limb_weights: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """One post-physics env step on [N,J,13] rigid-body state (progress already incremented).
    amp_hist: [N,S,A] newest-first window BEFORE this step; returned 'amp_obs_buf' is the window after.
    occlusion [N, K] bool: random_occlu_idx of _occl_training (humanoid_im.py:797-804); shape_params / limb_weights: self-obs tails."""
    N, J, _ = body_state.shape
    bp, br, bv, bw = body_state[..., 0:3], body_state[..., 3:7], body_state[..., 7:10], body_state[..., 10:13]
    dof_pos, dof_vel = dof_state[..., 0], dof_state[..., 1]
    out: Dict[str, Tensor] = {}

    # reward + reset at the CURRENT motion time (humanoid_im.py:879, :1118)
    t_now = progress * cfg.dt + start_times + start_offsets
    ref = motion_state(tab, motion_ids, t_now, global_offset)
    tb = list(range(J)) if cfg.track_bodies is None else cfg.track_bodies
    rs = list(range(J)) if cfg.full_body_reward else tb            # humanoid_im.py:912-935
    rew, raw = imitation_reward(bp[:, rs], br[:, rs], bv[:, rs], bw[:, rs], ref["rg_pos"][:, rs], ref["rb_rot"][:, rs], ref["body_vel"][:, rs],
                                ref["body_ang_vel"][:, rs], cfg.rwd)
    if cfg.power_reward:
        pw = power_reward(dof_force, dof_vel, progress, cfg.power_coef)
        rew = rew + pw
        raw = torch.cat((raw, pw[:, None]), dim=-1)
    out["rew"], out["reward_raw"] = rew, raw
    # flags.im_eval extras of HumanoidIm.post_physics_step (humanoid_im.py:674-680)
    out["mpjpe"] = (bp - ref["rg_pos"]).norm(dim=-1).mean(dim=-1)
    out["body_pos_gt"] = ref["rg_pos"]
    rb = list(range(J)) if cfg.reset_bodies is None else list(cfg.reset_bodies)
    td = torch.full((J,), cfg.term_dist) if not torch.is_tensor(cfg.term_dist) else cfg.term_dist
    pass_time = t_now >= tab.lengths[motion_ids]
    ref_reset = ref["rg_pos"][:, rb].clone()
    if occlusion is not None:        # humanoid_im.py:1180-1181: an occluded body cannot fail the distance test (indexed by BODY id there)
        oc_rb = occlusion.bool()[:, rb]
        ref_reset[oc_rb] = bp[:, rb][oc_rb]
    out["reset"], out["terminate"] = im_reset(progress, bp[:, rb], ref_reset, pass_time, td[rb],
                                              cfg.early_term, cfg.no_collision, cfg.use_mean)

    # observation for the NEXT step (humanoid_im.py:744-754)
    T = cfg.time_steps
    t_next = ((progress[:, None] + 1) * cfg.dt + torch.arange(T)[None, :] * cfg.traj_dt
              + start_times[:, None] + start_offsets[:, None]).flatten()
    refn = motion_state(tab, motion_ids.repeat_interleave(T), t_next, global_offset.repeat_interleave(T, dim=0))
    so = self_obs(bp, br, bv, bw, cfg.local_root_obs, cfg.root_height_obs, cfg.upright, shape_params, limb_weights)
    # tracked subset of the simulated and the reference bodies (humanoid_im.py:762-770), then the occlusion overwrite (:797-804)
    sub = lambda x: x[:, tb]
    r_pos, r_rot, r_vel, r_ang = (sub(refn[k]).clone() for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel"))
    if occlusion is not None:
        assert T == 1
        oc = occlusion.bool()
        r_pos[oc], r_rot[oc], r_vel[oc], r_ang[oc] = sub(bp)[oc], sub(br)[oc], sub(bv)[oc], sub(bw)[oc]
    to = task_obs_v6(bp[:, 0], br[:, 0], sub(bp), sub(br), sub(bv), sub(bw), r_pos, r_rot, r_vel, r_ang, T, cfg.upright)
    out["obs"] = torch.cat((so, to), dim=-1)
    out["ref_body_pos"] = refn["rg_pos"].view(N, T, J, 3)[:, 0]
    out["ref_body_rot"] = refn["rb_rot"].view(N, T, J, 4)[:, 0]
    out["ref_body_vel"] = refn["body_vel"].view(N, T, J, 3)[:, 0]
    out["ref_body_ang_vel"] = refn["body_ang_vel"].view(N, T, J, 3)[:, 0]
    out["ref_dof_pos"] = refn["dof_pos"].view(N, T, -1)[:, 0]

    # AMP observation: shift history by one, newest in slot 0 (humanoid_amp.py:662-670, :672-707)
    cur = amp_obs(bp[:, 0], br[:, 0], bv[:, 0], bw[:, 0], dof_pos, dof_vel, bp[:, cfg.key_bodies],
                  cfg.dof_subset, cfg.local_root_obs, cfg.root_height_obs, cfg.upright)
    out["amp_obs"] = cur
    out["amp_obs_buf"] = torch.cat((cur[:, None], amp_hist[:, :-1]), dim=1)
    return out


def env_step_getup(tab: MotionTables, cfg: StepConfig, body_state: Tensor, dof_state: Tensor, dof_force: Tensor,
                   progress: Tensor, motion_ids: Tensor, start_times: Tensor, start_offsets: Tensor, global_offset: Tensor,
                   amp_hist: Tensor, point_goal: Tensor, cycle_counter: Tensor, cycle_phase: Tensor, zero_out_far: bool = True,
                   cycle_motion: bool = True, close_distance: float = 0.25, far_distance: float = 3.0,
                   max_episode_length: int = 300) -> Dict[str, Tensor]:
    """The env step of env_im_getup_mcp.yaml (`zero_out_far: True`, `zero_out_far_train: False`, `cycle_motion: True`) -- the
    configuration HumanoidImMCP trains in:
      * pre_physics `_update_cycle_count` (humanoid_im.py:1076-1079): `cycle_counter` is the value BEFORE the decrement;
      * `_compute_reward` with the point-goal mix (:890-905, compute_point_goal_reward :1557-1562);
      * `_compute_reset` with clip wrap-around (:1120-1146): an env whose motion time passed the clip length gets a new start
        time (`cycle_phase` = the uniform numbers sample_time_interval draws, motion_lib_base.py:414-423), offset -progress*dt,
        a global offset that puts the clip's root under the humanoid, cycle_counter 60; pass_time = progress >= max_len - 1;
      * `_compute_task_obs` overwrites for far references (:783-796) and `_point_goal` update (:792).
    Returns env_step's dict plus the re-based bookkeeping tensors."""
    N, J, _ = body_state.shape
    bp, br, bv, bw = body_state[..., 0:3], body_state[..., 3:7], body_state[..., 7:10], body_state[..., 10:13]
    dof_pos, dof_vel = dof_state[..., 0], dof_state[..., 1]
    out: Dict[str, Tensor] = {}
    cc = torch.clamp_min(cycle_counter - 1, 0)
    start_times, start_offsets, global_offset = start_times.clone(), start_offsets.clone(), global_offset.clone()

    t_now = progress * cfg.dt + start_times + start_offsets
    ref = motion_state(tab, motion_ids, t_now, global_offset)
    im_rew, im_raw = imitation_reward(bp, br, bv, bw, ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], cfg.rwd)
    if zero_out_far:
        distance = torch.norm(bp[:, 0] - ref["root_pos"], dim=-1)
        far = distance > 0.25                                                # transition_distance (a literal in the reference)
        pg = torch.clamp(point_goal - distance, max=1 / 3) * 9
        raw = torch.zeros(N, 4)
        raw[:, 0] = pg
        rew = torch.where(far, pg, pg + im_rew * 0.5)
        raw = torch.where(far[:, None], raw, raw + im_raw * 0.5)
    else:
        rew, raw = im_rew, im_raw
    if cfg.power_reward:
        pw = power_reward(dof_force, dof_vel, progress, cfg.power_coef)
        rew = rew + pw
        raw = torch.cat((raw, pw[:, None]), dim=-1)
    out["rew"], out["reward_raw"] = rew, raw

    pass_len = t_now >= tab.lengths[motion_ids]
    if cycle_motion:
        pass_time = progress >= max_episode_length - 1
        if bool(pass_len.any()):
            w = pass_len
            start_offsets[w] = -progress[w] * cfg.dt
            curr_fps = 1 / 30
            start_times[w] = ((cycle_phase[w] * tab.lengths[motion_ids[w]]) / curr_fps).long() * curr_fps
            # get_root_pos_smpl (motion_lib_base.py:522-547): root position of the clip at the new start time, no offset
            i0, i1, bl = frame_blend(start_times[w], tab.lengths[motion_ids[w]], tab.num_frames[motion_ids[w]], tab.dts[motion_ids[w]])
            f0, f1 = i0 + tab.length_starts[motion_ids[w]], i1 + tab.length_starts[motion_ids[w]]
            b = bl.view(-1, 1, 1)
            root = ((1.0 - b) * tab.gts[f0] + b * tab.gts[f1])[:, 0]
            global_offset[w, :2] = bp[w, 0, :2] - root[:, :2]
            cc[w] = 60
            t_now = progress * cfg.dt + start_times + start_offsets
            ref = motion_state(tab, motion_ids, t_now, global_offset)
    else:
        pass_time = pass_len
    rb = list(range(J)) if cfg.reset_bodies is None else list(cfg.reset_bodies)
    td = torch.full((J,), cfg.term_dist) if not torch.is_tensor(cfg.term_dist) else cfg.term_dist
    reset, term = im_reset(progress, bp[:, rb], ref["rg_pos"][:, rb], pass_time, td[rb], cfg.early_term, cfg.no_collision, cfg.use_mean)
    rec = (~pass_time) & (cc > 0)
    out["reset"] = torch.where(rec, torch.zeros_like(reset), reset)
    out["terminate"] = torch.where(rec, torch.zeros_like(term), term)

    t_next = (progress + 1) * cfg.dt + start_times + start_offsets
    refn = motion_state(tab, motion_ids, t_next, global_offset)
    rp, rr, rv, rw = refn["rg_pos"].clone(), refn["rb_rot"].clone(), refn["body_vel"].clone(), refn["body_ang_vel"].clone()
    new_goal = point_goal.clone()
    if zero_out_far:
        distance = torch.norm(bp[:, 0] - rp[:, 0], dim=-1)
        z = distance > close_distance
        rp[z, 1:], rr[z, 1:] = bp[z, 1:], br[z, 1:]
        rv[z], rw[z] = bv[z], bw[z]
        new_goal = distance
        vz = distance > far_distance
        rp[vz, 0] = ((rp[vz, 0] - bp[vz, 0]) / distance[vz, None] * far_distance) + bp[vz, 0]
    so = self_obs(bp, br, bv, bw, cfg.local_root_obs, cfg.root_height_obs, cfg.upright)
    to = task_obs_v6(bp[:, 0], br[:, 0], bp, br, bv, bw, rp, rr, rv, rw, 1, cfg.upright)
    out["obs"] = torch.cat((so, to), dim=-1)
    out["ref_body_pos"], out["ref_body_rot"], out["ref_body_vel"] = refn["rg_pos"], refn["rb_rot"], refn["body_vel"]
    out["ref_body_ang_vel"] = refn["body_ang_vel"]
    cur = amp_obs(bp[:, 0], br[:, 0], bv[:, 0], bw[:, 0], dof_pos, dof_vel, bp[:, cfg.key_bodies],
                  cfg.dof_subset, cfg.local_root_obs, cfg.root_height_obs, cfg.upright)
    out["amp_obs"] = cur
    out["amp_obs_buf"] = torch.cat((cur[:, None], amp_hist[:, :-1]), dim=1)
    out.update(start_times=start_times, start_offsets=start_offsets, global_offset=global_offset, cycle_counter=cc,
               point_goal=new_goal)
    return out


def amp_obs_demo(tab: MotionTables, cfg: StepConfig, motion_ids: Tensor, times0: Tensor,
                 first_step: int = 0, num_steps: Optional[int] = None) -> Tensor:
    """build_amp_obs_demo (humanoid_amp.py:253-284) / _init_amp_obs_ref (:575-603, first_step=1): AMP
    observations of the reference motion at t0 - k*dt, k = first_step .. first_step+num_steps-1.  -> [n, S, A]."""
    S = cfg.num_amp_steps if num_steps is None else num_steps
    n = len(motion_ids)
    k = -cfg.dt * (torch.arange(0, S) + first_step)
    times = (times0[:, None] + k[None, :]).flatten()
    st = motion_state(tab, motion_ids.repeat_interleave(S), times)
    a = amp_obs(st["root_pos"], st["root_rot"], st["root_vel"], st["root_ang_vel"], st["dof_pos"], st["dof_vel"],
                st["rg_pos"][:, cfg.key_bodies], cfg.dof_subset, cfg.local_root_obs, cfg.root_height_obs, cfg.upright)
    return a.view(n, S, -1)


# ----------------------------------------------------------------------------------------------
# Hinge-joint robots (H1 / G1): phc/utils/motion_lib_real.py:236-361, humanoid_im.py:74-82 and :916-923 (extend bodies in
# the reward), humanoid_amp.py:1062-1104 (build_amp_observations_robot)
# ----------------------------------------------------------------------------------------------
class RobotTables:
    """motion_lib_real's tables: gts/grs/gvs/gavs over the J simulated bodies, the *_t tables over J + E (bodies then
    extend bodies), dof_pos / dvs [F, D]."""

    def __init__(self, gts_t, grs_t, gvs_t, gavs_t, dof_pos, dvs, lengths, num_frames, dts, length_starts, num_bodies):
        self.gts_t, self.grs_t, self.gvs_t, self.gavs_t, self.dof_pos, self.dvs = gts_t, grs_t, gvs_t, gavs_t, dof_pos, dvs
        self.lengths, self.num_frames, self.dts, self.length_starts, self.num_bodies = lengths, num_frames, dts, length_starts, num_bodies


def motion_state_robot(tab: RobotTables, ids: Tensor, times: Tensor, offset: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """MotionLibReal.get_motion_state (motion_lib_real.py:236-361): dof_pos is interpolated linearly like dof_vel."""
    J = tab.num_bodies
    i0, i1, blend = frame_blend(times, tab.lengths[ids], tab.num_frames[ids], tab.dts[ids])
    f0 = i0 + tab.length_starts[ids]
    f1 = i1 + tab.length_starts[ids]
    b1 = blend.view(-1, 1)
    b = blend.view(-1, 1, 1)
    lerp = lambda t: (1.0 - b) * t[f0] + b * t[f1]
    pos_t = lerp(tab.gts_t)
    if offset is not None:
        pos_t = pos_t + offset[:, None, :]
    vel_t, ang_t = lerp(tab.gvs_t), lerp(tab.gavs_t)
    rot_t = slerp(tab.grs_t[f0], tab.grs_t[f1], b)
    dof_vel = (1.0 - b1) * tab.dvs[f0] + b1 * tab.dvs[f1]
    dof_pos = (1.0 - b1) * tab.dof_pos[f0] + b1 * tab.dof_pos[f1]
    pos, rot, vel, ang = pos_t[:, :J], rot_t[:, :J], vel_t[:, :J], ang_t[:, :J]
    return dict(root_pos=pos[:, 0].clone(), root_rot=rot[:, 0].clone(), dof_pos=dof_pos, root_vel=vel[:, 0].clone(),
                root_ang_vel=ang[:, 0].clone(), dof_vel=dof_vel, rg_pos=pos, rb_rot=rot, body_vel=vel, body_ang_vel=ang,
                rg_pos_t=pos_t, rg_rot_t=rot_t, body_vel_t=vel_t, body_ang_vel_t=ang_t)


def amp_obs_robot(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_pos, local_root_obs=True,
                  root_height_obs=True, upright=True) -> Tensor:
    """build_amp_observations_robot (humanoid_amp.py:1062-1104)."""
    if not upright:
        root_rot = strip_base_rot(root_rot)
    hinv = heading_q(root_rot, inverse=True)
    rr = tan_norm(qmul(hinv, root_rot) if local_root_obs else root_rot)
    lv, lw = qrot(hinv, root_vel), qrot(hinv, root_ang_vel)
    lk = key_pos - root_pos[:, None, :]
    K = lk.shape[1]
    lk = qrot(hinv[:, None, :].expand(-1, K, -1).reshape(-1, 4), lk.reshape(-1, 3)).view(lk.shape[0], K * 3)
    parts = ([root_pos[:, 2:3]] if root_height_obs else []) + [rr, lv, lw, dof_pos, dof_vel, lk]
    return torch.cat(parts, dim=-1)


def env_step_robot(tab: RobotTables, cfg: "StepConfig", ext_parents, ext_pos, body_state: Tensor, dof_state: Tensor,
                   dof_force: Tensor, progress: Tensor, motion_ids: Tensor, start_times: Tensor, start_offsets: Tensor,
                   global_offset: Tensor, amp_hist: Tensor) -> Dict[str, Tensor]:
    """env_step for humanoid_type h1 / g1: the tracking reward also sees the extend bodies (humanoid_im.py:916-923); AMP
    observation from build_amp_observations_robot; observations and reset as for SMPL on the J simulated bodies."""
    N, J, _ = body_state.shape
    bp, br, bv, bw = body_state[..., 0:3], body_state[..., 3:7], body_state[..., 7:10], body_state[..., 10:13]
    dof_pos, dof_vel = dof_state[..., 0], dof_state[..., 1]
    par = torch.as_tensor(list(ext_parents))
    off = torch.as_tensor(ext_pos, dtype=torch.float32)
    E = len(par)
    out: Dict[str, Tensor] = {}
    t_now = progress * cfg.dt + start_times + start_offsets
    ref = motion_state_robot(tab, motion_ids, t_now, global_offset)
    ext_cur = qrot(br[:, par].reshape(-1, 4), off[None].expand(N, E, 3).reshape(-1, 3)).view(N, E, 3) + bp[:, par]
    bp_e = torch.cat((bp, ext_cur), dim=1)
    br_e = torch.cat((br, br[:, par]), dim=1)
    rp_e = torch.cat((ref["rg_pos"], ref["rg_pos_t"][:, J:]), dim=1)
    rr_e = torch.cat((ref["rb_rot"], ref["rg_rot_t"][:, J:]), dim=1)
    rew, raw = imitation_reward(bp_e, br_e, bv, bw, rp_e, rr_e, ref["body_vel"], ref["body_ang_vel"], cfg.rwd)
    if cfg.power_reward:
        pw = power_reward(dof_force, dof_vel, progress, cfg.power_coef)
        rew = rew + pw
        raw = torch.cat((raw, pw[:, None]), dim=-1)
    out["rew"], out["reward_raw"] = rew, raw
    rb = list(range(J)) if cfg.reset_bodies is None else list(cfg.reset_bodies)
    td = torch.full((J,), cfg.term_dist) if not torch.is_tensor(cfg.term_dist) else cfg.term_dist
    pass_time = t_now >= tab.lengths[motion_ids]
    out["reset"], out["terminate"] = im_reset(progress, bp[:, rb], ref["rg_pos"][:, rb], pass_time, td[rb],
                                              cfg.early_term, cfg.no_collision, cfg.use_mean)
    t_next = (progress + 1) * cfg.dt + start_times + start_offsets
    refn = motion_state_robot(tab, motion_ids, t_next, global_offset)
    so = self_obs(bp, br, bv, bw, cfg.local_root_obs, cfg.root_height_obs, cfg.upright)
    to = task_obs_v6(bp[:, 0], br[:, 0], bp, br, bv, bw, refn["rg_pos"], refn["rb_rot"], refn["body_vel"], refn["body_ang_vel"],
                     1, cfg.upright)
    out["obs"] = torch.cat((so, to), dim=-1)
    out["ref_body_pos"], out["ref_body_rot"], out["ref_body_vel"] = refn["rg_pos"], refn["rb_rot"], refn["body_vel"]
    out["ref_pose_t"] = torch.cat((refn["rg_pos_t"], refn["rg_rot_t"], refn["body_vel_t"], refn["body_ang_vel_t"]), dim=-1)
    cur = amp_obs_robot(bp[:, 0], br[:, 0], bv[:, 0], bw[:, 0], dof_pos, dof_vel, bp[:, cfg.key_bodies], cfg.local_root_obs,
                        cfg.root_height_obs, cfg.upright)
    out["amp_obs"] = cur
    out["amp_obs_buf"] = torch.cat((cur[:, None], amp_hist[:, :-1]), dim=1)
    return out


def amp_obs_demo_robot(tab: RobotTables, cfg: "StepConfig", motion_ids: Tensor, times0: Tensor, first_step: int = 0,
                       num_steps: Optional[int] = None) -> Tensor:
    S = cfg.num_amp_steps if num_steps is None else num_steps
    n = len(motion_ids)
    k = -cfg.dt * (torch.arange(0, S) + first_step)
    times = (times0[:, None] + k[None, :]).flatten()
    st = motion_state_robot(tab, motion_ids.repeat_interleave(S), times)
    a = amp_obs_robot(st["root_pos"], st["root_rot"], st["root_vel"], st["root_ang_vel"], st["dof_pos"], st["dof_vel"],
                      st["rg_pos"][:, cfg.key_bodies], cfg.local_root_obs, cfg.root_height_obs, cfg.upright)
    return a.view(n, S, -1)


# ----------------------------------------------------------------------------------------------
# K8  running mean/std   (reference: phc/utils/running_mean_std.py:56-109)
# ----------------------------------------------------------------------------------------------
def rms_normalize(x: Tensor, mean64: Tensor, var64: Tensor, eps: float = 1e-5) -> Tensor:
    return torch.clamp((x - mean64.float()) / torch.sqrt(var64.float() + eps), -5.0, 5.0)


def rms_unnormalize(y: Tensor, mean64: Tensor, var64: Tensor, eps: float = 1e-5) -> Tensor:
    return torch.sqrt(var64.float() + eps) * torch.clamp(y, -5.0, 5.0) + mean64.float()


def rms_update(mean64: Tensor, var64: Tensor, count64: Tensor, x: Tensor):
    """Parallel-variance merge of the batch moments (unbiased batch var) into fp64 running stats."""
    bm, bv, bc = x.mean(0), x.var(0), x.shape[0]
    delta = bm - mean64
    tot = count64 + bc
    new_mean = mean64 + delta * bc / tot
    m2 = var64 * count64 + bv * bc + delta ** 2 * count64 * bc / tot
    return new_mean, m2 / tot, tot


# ----------------------------------------------------------------------------------------------
# K10/K11  GAE and advantage normalisation   (reference: phc/learning/common_agent.py:493-505, :589-599)
# ----------------------------------------------------------------------------------------------
def gae(fdones: Tensor, values: Tensor, rewards: Tensor, next_values: Tensor, gamma: float, tau: float) -> Tensor:
    """Time-major [T,N,1] inputs -> advantages [T,N,1]."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    for t in reversed(range(T)):
        nd = (1.0 - fdones[t]).unsqueeze(1)
        delta = rewards[t] + gamma * next_values[t] - values[t]
        last = delta + gamma * tau * nd * last
        adv[t] = last
    return adv


def normalize_advantages(returns: Tensor, values: Tensor, normalize: bool = True) -> Tensor:
    a = (returns - values).sum(dim=1)
    if normalize:
        a = (a - a.mean()) / (a.std() + 1e-8)
    return a


# ----------------------------------------------------------------------------------------------
# K9/K12  networks and losses
#   MLP build: phc/learning/network_builder.py:105-124; actor/critic/disc eval: amp_network_builder.py:58-216
#   rl_games==1.1.4 (not in tree) ModelA2CContinuousLogStd: sigma=exp(logstd), neglogp, entropy -- restated,
#   self-pinned against torch.distributions.Normal in tests/test_oracle_learning.py.
# ----------------------------------------------------------------------------------------------
def act_fn(name: str):
    return {"relu": torch.relu, "silu": torch.nn.functional.silu, "elu": torch.nn.functional.elu,
            "None": (lambda x: x)}[name]


def mlp_forward(x: Tensor, weights, biases, act: str = "relu", final_linear: bool = True) -> Tensor:
    """Linear->act for every hidden layer; last (W,b) is the output head without activation."""
    f = act_fn(act)
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = torch.nn.functional.linear(x, w, b)
        if not (final_linear and i == n - 1):
            x = f(x)
    return x


def gaussian_neglogp(a: Tensor, mu: Tensor, sigma: Tensor, logstd: Tensor) -> Tensor:
    return (0.5 * (((a - mu) / sigma) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * a.shape[-1] + logstd.sum(-1))


def policy_kl(mu, sigma, mu_old, sigma_old) -> Tensor:
    """rl_games torch_ext.policy_kl(p0=new, p1=old, reduce=True) as called at amp_agent.py:660:
    KL(new || old) with rl_games' +1e-5 regularisers, summed over actions, mean over batch."""
    c1 = torch.log(sigma_old / sigma + 1e-5)
    c2 = (sigma ** 2 + (mu_old - mu) ** 2) / (2.0 * (sigma_old ** 2 + 1e-5))
    return (c1 + c2 - 0.5).sum(-1).mean()


def actor_loss(old_neglogp, neglogp, adv, e_clip) -> Tensor:
    """common_agent.py:564-574."""
    ratio = torch.exp(old_neglogp - neglogp)
    return torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip))


def critic_loss(values, returns) -> Tensor:
    """common_agent.py:576-587 with clip_value False."""
    return (returns - values) ** 2


def bound_loss(mu, soft_bound: float = 1.0) -> Tensor:
    """common_agent.py:512-520."""
    return (torch.clamp_max(mu + soft_bound, 0.0) ** 2 + torch.clamp_min(mu - soft_bound, 0.0) ** 2).sum(-1)


def disc_reward(logits: Tensor, scale: float = 2.0) -> Tensor:
    """amp_agent.py:864-878."""
    prob = 1 / (1 + torch.exp(-logits))
    return -torch.log(torch.maximum(1 - prob, torch.tensor(0.0001))) * scale


def disc_loss(agent_logit, demo_logit, demo_obs, logit_w, all_disc_w, logit_reg=0.01, grad_penalty=5.0,
              weight_decay=1e-4) -> Dict[str, Tensor]:
    """amp_agent.py:732-789.  demo_obs must require grad and demo_logit be computed from it."""
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    loss = 0.5 * (bce(agent_logit, torch.zeros_like(agent_logit)) + bce(demo_logit, torch.ones_like(demo_logit)))
    logit_l = torch.sum(torch.square(logit_w))
    loss = loss + logit_reg * logit_l
    g = torch.autograd.grad(demo_logit, demo_obs, grad_outputs=torch.ones_like(demo_logit), create_graph=True,
                            retain_graph=True, only_inputs=True)[0]
    gp = torch.mean(torch.sum(torch.square(g), dim=-1))
    loss = loss + grad_penalty * gp
    if weight_decay != 0:
        loss = loss + weight_decay * torch.sum(torch.square(torch.cat([w.flatten() for w in all_disc_w])))
    return dict(disc_loss=loss, disc_grad_penalty=gp.detach(), disc_logit_loss=logit_l.detach(),
                disc_agent_acc=(agent_logit < 0).float().mean(), disc_demo_acc=(demo_logit > 0).float().mean())
