"""TEST INFRASTRUCTURE ONLY -- CPU (numpy fp64) restatement of the reference's motion LOADER: the math between the on-disk
clip format and the tables the hot path reads.  Imported by tests/ only; the product path
(`phc_b200.motion_loader` -> `phc_motion_load` in csrc/motion_load.cu) never touches it.

Pinned: tests/test_oracle_golden.py replays tests/golden/load.npz, produced by the UNMODIFIED
`MotionLibSMPL.load_motion_with_skeleton` (tests/golden/make_golden.py:gen_load).

Reference chain restated here, per clip (all in float64 like the reference, cast to float32 at the end as
motion_lib_base.py:300-307 does with `.float()`):
  1. heading randomisation            phc/utils/motion_lib_smpl.py:141-149   (scipy Rotation about z, composed on the left)
  2. local rotations from global ones poselib/skeleton/skeleton3d.py:444-461  quat_mul_norm(conj(parent), child)
  3. forward kinematics               skeleton3d.py:390-408, core/rotation3d.py:318-327 (transform_mul, chain re-composed
                                      from the normalised local rotations)
  4. linear velocity                  skeleton3d.py:1100-1107  np.gradient / dt, gaussian_filter1d(sigma=2, mode="nearest")
  5. angular velocity                 skeleton3d.py:1110-1121  angle-axis of quat_mul_norm(r[t+1], conj(r[t])) / dt, same filter
  6. dof velocities                   phc/utils/motion_lib_base.py:47-70     phc quat_to_angle_axis of conj(l[t]) * l[t+1]
Quaternions are xyzw.
"""
from __future__ import annotations

import numpy as np

GAUSS_SIGMA = 2.0
GAUSS_RADIUS = int(4.0 * GAUSS_SIGMA + 0.5)        # scipy.ndimage.gaussian_filter1d: truncate = 4.0 -> radius 8, 17 taps


def gauss_weights() -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma=2, order=0, radius=8)."""
    x = np.arange(-GAUSS_RADIUS, GAUSS_RADIUS + 1, dtype=np.float64)
    w = np.exp(-0.5 / (GAUSS_SIGMA * GAUSS_SIGMA) * x * x)
    return w / w.sum()


def qmul(a, b):
    """Hamilton product (core/rotation3d.py:15-27; phc/utils/isaacgym_torch_utils.py:25-45 gives the same value)."""
    x1, y1, z1, w1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    x2, y2, z2, w2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                     w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], -1)


def qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def qnormalize_pos(q):
    """core/rotation3d.py:93-99 quat_normalize = unit(quat_pos(q)): real part made non-negative, then unit length."""
    q = np.where(q[..., 3:] < 0, -q, q)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def qmul_norm(a, b):
    return qnormalize_pos(qmul(a, b))


def qrotate(q, v):
    """core/rotation3d.py:206-212: imaginary part of q * (v, 0) * conj(q)."""
    vq = np.concatenate([v, np.zeros_like(v[..., :1])], -1)
    return qmul(qmul(q, vq), qconj(q))[..., :3]


def poselib_angle_axis(q):
    """core/rotation3d.py:231-241: angle = acos(clamp(2w^2-1)), axis = xyz / max(|xyz|, 1e-9)."""
    s = np.clip(2.0 * q[..., 3] ** 2 - 1.0, -1.0, 1.0)
    ang = np.arccos(s)
    n = np.maximum(np.linalg.norm(q[..., :3], axis=-1, keepdims=True), 1e-9)
    return ang, q[..., :3] / n


def phc_qmul_f32(a, b):
    """phc/utils/isaacgym_torch_utils.py:25-45 quat_mul, evaluated in float32 in the reference's operation order."""
    a, b = a.astype(np.float32), b.astype(np.float32)
    x1, y1, z1, w1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    x2, y2, z2, w2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = np.float32(0.5) * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return np.stack([x, y, z, w], -1)


def phc_angle_axis_f32(q):
    """phc/utils/torch_utils.py:58-78 quat_to_angle_axis in float32 (the dof velocities go through it in float32: poselib's
    SkeletonState.local_rotation is assembled into a float32 `quat_identity_like` buffer, skeleton3d.py:449-459)."""
    q = q.astype(np.float32)
    w = q[..., 3]
    one = np.float32(1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        sin_t = np.sqrt(one - w * w)
        ang = np.float32(2.0) * np.arccos(w)
        ang = np.arctan2(np.sin(ang), np.cos(ang))           # normalize_angle
        axis = q[..., :3] / sin_t[..., None]
    mask = np.abs(sin_t) > np.float32(1e-5)
    ang = np.where(mask, ang, np.float32(0.0))
    axis = np.where(mask[..., None], axis, np.array([0.0, 0.0, 1.0], np.float32))
    return ang.astype(np.float32), axis.astype(np.float32)


def filter_time(x):
    """gaussian_filter1d(x, 2, axis=0, mode="nearest") over one clip [F, ...]."""
    w = gauss_weights()
    F = x.shape[0]
    out = np.zeros_like(x)
    for k in range(-GAUSS_RADIUS, GAUSS_RADIUS + 1):
        idx = np.clip(np.arange(F) + k, 0, F - 1)
        out += w[k + GAUSS_RADIUS] * x[idx]
    return out


def load_clip(pose_quat_global, root_trans, fps, parents, offsets, heading=None):
    """One clip: pose_quat_global [F,J,4], root_trans [F,3] (float64), offsets [J,3] = skeleton_tree.local_translation,
    heading = angle about z or None (flags.im_eval / flags.test: no randomisation).  Returns float32 tables."""
    g = np.asarray(pose_quat_global, np.float64)
    t = np.asarray(root_trans, np.float64)
    off = np.asarray(offsets, np.float64)
    F, J, _ = g.shape
    if heading is not None:
        hq = np.array([0.0, 0.0, np.sin(0.5 * heading), np.cos(0.5 * heading)])
        g = g / np.linalg.norm(g, axis=-1, keepdims=True)                  # Rotation.from_quat normalises its input
        g = qmul(np.broadcast_to(hq, g.shape), g)                          # random_heading_rot * R(q), as_quat(): no sign fix
        g = g / np.linalg.norm(g, axis=-1, keepdims=True)
        c, s = np.cos(heading), np.sin(heading)
        t = np.stack([c * t[:, 0] - s * t[:, 1], s * t[:, 0] + c * t[:, 1], t[:, 2]], -1)   # trans @ R.T
    lr = np.empty_like(g)
    for j in range(J):
        p = int(parents[j])
        lr[:, j] = g[:, j] if p < 0 else qmul_norm(qconj(g[:, p]), g[:, j])
    # the reference writes the local rotations into a float32 buffer (quat_identity_like -> torch.ones default dtype,
    # skeleton3d.py:449-459): everything downstream (FK chain, dof velocities) sees the float32-rounded values
    lr32 = lr.astype(np.float32)
    lr = lr32.astype(np.float64)
    chain = np.empty_like(g)
    pos = np.empty((F, J, 3))
    for j in range(J):
        p = int(parents[j])
        if p < 0:
            chain[:, j], pos[:, j] = lr[:, j], t
        else:
            chain[:, j] = qmul_norm(chain[:, p], lr[:, j])
            pos[:, j] = qrotate(chain[:, p], np.broadcast_to(off[j], (F, 3))) + pos[:, p]
    dt = 1 / fps
    vel = filter_time(np.gradient(pos, axis=0) / dt)
    dq = np.zeros_like(g)
    dq[..., 3] = 1.0
    dq[:-1] = qmul_norm(g[1:], qconj(g[:-1]))
    ang, ax = poselib_angle_axis(dq)
    angvel = filter_time(ax * ang[..., None] / dt)
    dang, dax = phc_angle_axis_f32(phc_qmul_f32(lr32[:-1] * np.array([-1, -1, -1, 1], np.float32), lr32[1:]))
    dv = (dax * dang[..., None] / np.float32(1.0 / fps))[:, 1:]          # float32 tensor / python float
    dv = np.concatenate([dv, dv[-1:]], 0)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(gts=f32(pos), grs=f32(g), lrs=lr32, gvs=f32(vel), gavs=f32(angvel), dvs=f32(dv))


def load_clips(pose_quat_global, root_trans, num_frames, fps, parents, offsets, heading=None):
    """Concatenated clips -> concatenated tables (the torch.cat of motion_lib_base.py:300-307)."""
    outs, s = [], 0
    for i, F in enumerate(int(n) for n in num_frames):
        outs.append(load_clip(pose_quat_global[s:s + F], root_trans[s:s + F], float(fps[i]), parents, offsets[i],
                              None if heading is None else float(heading[i])))
        s += F
    return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
