"""GPU parity tests of the env-side hot path: CUDA (through the C ABI) vs the oracle and the committed goldens.
Tolerance: rtol 1e-5 / atol 1e-6 fp32 on observations and rewards (north_star); integers bit-exact."""
import pytest
import torch

from oracle import phc_oracle as O
from phc_b200 import ops, synthetic as syn
from tests.helpers import close, env_state_from, load, motion_data_from, oracle_tables, smpl_step_config

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def smpl_cfg(**kw):
    base = dict(key_bodies=syn.SMPL_KEY_BODIES, reset_bodies=syn.SMPL_RESET_BODIES, dof_subset=syn.SMPL_DOF_SUBSET)
    base.update(kw)
    return ops.EnvStepConfig(**base)


def pack(m: syn.MotionData):
    d = m.to(DEV)
    return ops.pack_motion_lib(d.gts, d.grs, d.gvs, d.gavs, d.lrs, d.dvs, d.lengths, d.num_frames, d.dts, d.length_starts)


def run_cuda_step(m, st, cfg, **plan_kw):
    mlib = pack(m)
    s = st.to(DEV)
    plan = ops.EnvStepPlan(cfg, mlib, s.body_state, s.dof_state, s.dof_force, s.progress, s.motion_ids, s.start_times,
                           s.start_offsets, s.global_offset, amp_obs_buf=s.amp_hist.clone(), with_ref_buffers=True, **plan_kw)
    plan.run()
    torch.cuda.synchronize()
    return plan


def check_against(plan, exp, tag=""):
    # atol 2e-6: velocity columns are differences of O(10) operands, so one fp32 ulp of an operand (9.5e-7 at 8..16)
    # survives cancellation as an absolute error; everything else holds at rtol 1e-5 / atol 1e-6.
    close(plan.obs.cpu(), exp["obs"], atol=2e-6, what=f"{tag} obs")
    close(plan.rew.cpu(), exp["rew"], what=f"{tag} rew")
    close(plan.reward_raw.cpu(), exp["reward_raw"], what=f"{tag} reward_raw")
    close(plan.reset.cpu(), exp["reset"], what=f"{tag} reset")
    close(plan.terminate.cpu(), exp["terminate"], what=f"{tag} terminate")
    close(plan.amp_obs_buf.cpu(), exp["amp_obs_buf"], what=f"{tag} amp_obs_buf")
    close(plan.ref_body_pos.cpu(), exp["ref_body_pos"], what=f"{tag} ref_body_pos")
    close(plan.ref_body_rot.cpu(), exp["ref_body_rot"], what=f"{tag} ref_body_rot")
    close(plan.ref_body_vel.cpu(), exp["ref_body_vel"], what=f"{tag} ref_body_vel")


@pytest.mark.parametrize("tag,in_tag,kw", [("A", "A", {}), ("B", "B", {}),
                                            ("C", "A", dict(upright=False, local_root_obs=False)),
                                            ("D", "B", dict(term_use_mean=True))])
def test_env_step_vs_reference_golden(tag, in_tag, kw):
    """CUDA vs outputs of the UNMODIFIED reference classes (tests/golden/envstep.npz)."""
    g = load("envstep.npz")
    plan = run_cuda_step(motion_data_from(g), env_state_from(g, in_tag), smpl_cfg(**kw))
    exp = {k: g[f"{tag}_out_{k}"] for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf",
                                             "ref_body_pos", "ref_body_rot", "ref_body_vel")}
    check_against(plan, exp, tag)


@pytest.mark.parametrize("n,seed,jitter", [(1, 0, False), (5, 1, True), (257, 2, False), (1024, 3, True)])
def test_env_step_vs_oracle(n, seed, jitter):
    """Ragged sizes (1 env, non-multiple of the CTA tile, > one wave of warps) against the oracle."""
    m = syn.make_motions(n, seed=seed, min_frames=30, max_frames=90)
    st = syn.make_env_state(m, n, seed=seed, max_progress=80, with_offset=jitter, blend_jitter=jitter)
    plan = run_cuda_step(m, st, smpl_cfg())
    exp = O.env_step(oracle_tables(m), smpl_step_config(), st.body_state, st.dof_state, st.dof_force, st.progress,
                     st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    check_against(plan, exp, f"n={n}")
    assert exp["reset"].sum() > 0 or n < 8


def test_env_step_shared_clips_and_padded_bodies():
    """Fewer clips than envs (ids wrap) and bodies_per_env > J with a non-multiple-of-4 row (plain-load path)."""
    n = 64
    m = syn.make_motions(7, seed=5, min_frames=100, max_frames=120)
    st = syn.make_env_state(m, n, seed=5, max_progress=60)
    exp = O.env_step(oracle_tables(m), smpl_step_config(), st.body_state, st.dof_state, st.dof_force, st.progress,
                     st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    check_against(run_cuda_step(m, st, smpl_cfg()), exp, "shared")
    padded = torch.cat((st.body_state, torch.full((n, 3, 13), 77.0)), dim=1).contiguous()   # 27 bodies per env
    st2 = syn.EnvState(**{**{k: getattr(st, k) for k in st.__dataclass_fields__}, "body_state": padded})
    check_against(run_cuda_step(m, st2, smpl_cfg()), exp, "padded")


def test_env_step_future_tracks():
    """T = 3 future reference samples (fut_tracks): observation layout [B, T, J*24]."""
    n = 96
    m = syn.make_motions(n, seed=9, min_frames=40, max_frames=80)
    st = syn.make_env_state(m, n, seed=9, max_progress=30, blend_jitter=True)
    cfg_o = smpl_step_config(time_steps=3, traj_dt=3 / 30.0)
    exp = O.env_step(oracle_tables(m), cfg_o, st.body_state, st.dof_state, st.dof_force, st.progress, st.motion_ids,
                     st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    plan = run_cuda_step(m, st, smpl_cfg(time_steps=3, traj_dt=3 / 30.0))
    assert plan.obs.shape[1] == 358 + 3 * 576
    check_against(plan, exp, "fut")


def test_env_step_other_body_count():
    """J = 20 (H1-sized tree), all joints in the AMP obs, no power reward, out-of-place AMP window."""
    n, J = 130, 20
    m = syn.make_motions(n, seed=4, num_bodies=J, min_frames=30, max_frames=60)
    A = 1 + 12 + 9 * (J - 1) + 3 * 2
    st = syn.make_env_state(m, n, seed=4, amp_dim=A, max_progress=40)
    kw = dict(key_bodies=[5, 9], reset_bodies=None, dof_subset=None, power_reward=False)
    exp = O.env_step(oracle_tables(m), O.StepConfig(**kw), st.body_state, st.dof_state, st.dof_force, st.progress,
                     st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    mlib = pack(m)
    s = st.to(DEV)
    hist_in = s.amp_hist.clone()
    plan = ops.EnvStepPlan(ops.EnvStepConfig(**kw), mlib, s.body_state, s.dof_state, None, s.progress, s.motion_ids,
                           s.start_times, s.start_offsets, s.global_offset, amp_hist_in=hist_in, with_ref_buffers=True)
    plan.run()
    torch.cuda.synchronize()
    assert plan.reward_raw.shape[1] == 4
    check_against(plan, exp, "J20")
    assert torch.equal(hist_in.cpu(), st.amp_hist)          # out-of-place: the input window is untouched


def test_env_step_is_idempotent_and_deterministic():
    """Size-independent property at the bench size: same inputs -> bit-identical outputs, run twice."""
    n = 4096
    m = syn.make_motions(n, seed=11, min_frames=60, max_frames=120)
    st = syn.make_env_state(m, n, seed=11)
    p1 = run_cuda_step(m, st, smpl_cfg())
    p2 = run_cuda_step(m, st, smpl_cfg())
    for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf"):
        assert torch.equal(getattr(p1, k), getattr(p2, k)), k
    assert torch.isfinite(p1.obs).all() and torch.isfinite(p1.rew).all()
    # rotating the whole world about z leaves the (heading-local) observation and the reward unchanged
    th = 0.7
    qz = torch.tensor([0.0, 0.0, torch.sin(torch.tensor(th / 2)), torch.cos(torch.tensor(th / 2))])
    from phc_b200.synthetic import _qmul, _qrot
    def rot_tab(t, is_q):
        return _qmul(qz.expand_as(t), t) if is_q else _qrot(qz.expand(*t.shape[:-1], 4), t)
    m2 = syn.MotionData(gts=rot_tab(m.gts, False), grs=rot_tab(m.grs, True), lrs=m.lrs, gvs=rot_tab(m.gvs, False),
                        gavs=rot_tab(m.gavs, False), dvs=m.dvs, lengths=m.lengths, num_frames=m.num_frames, dts=m.dts,
                        length_starts=m.length_starts)
    b = st.body_state
    b2 = torch.cat((rot_tab(b[..., 0:3], False), rot_tab(b[..., 3:7], True), rot_tab(b[..., 7:10], False),
                    rot_tab(b[..., 10:13], False)), dim=-1).contiguous()
    st2 = syn.EnvState(**{**{k: getattr(st, k) for k in st.__dataclass_fields__}, "body_state": b2})
    p3 = run_cuda_step(m2, st2, smpl_cfg())
    close(p3.obs.cpu(), p1.obs.cpu(), rtol=1e-3, atol=2e-4, what="z-rotation invariance of obs")
    close(p3.rew.cpu(), p1.rew.cpu(), rtol=1e-4, atol=1e-5, what="z-rotation invariance of reward")
    assert torch.equal(p3.reset, p1.reset)


def test_motion_state_and_amp_demo_vs_golden():
    g = load("motion.npz")
    mlib = pack(motion_data_from(g))
    res = ops.motion_state(mlib, g["ids"].to(DEV), g["times"].to(DEV), g["offset"].to(DEV))
    for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_vel", "rg_pos", "rb_rot", "body_vel", "body_ang_vel"):
        close(res[k].cpu(), g["out_" + k], what=k)
    close(res["dof_pos"].cpu(), g["out_dof_pos"], rtol=1e-4, atol=2e-5, what="dof_pos (acos near identity)")
    res_no = ops.motion_state(mlib, g["ids"].to(DEV), g["times"].to(DEV), None, want_dof=False)
    close(res_no["rg_pos"].cpu(), g["out_noffset_rg_pos"], what="rg_pos no offset")

    e = load("envstep.npz")
    mlib2 = pack(motion_data_from(e))
    demo = ops.amp_obs_demo(mlib2, smpl_cfg(), e["demo_ids"].to(DEV), e["demo_t0"].to(DEV))
    close(demo.cpu(), e["demo_out"], rtol=1e-4, atol=2e-5, what="amp_obs_demo")
    hist = ops.amp_obs_demo(mlib2, smpl_cfg(), e["demo_ids"].to(DEV), e["demo_t0"].to(DEV), first_step=1, num_steps=9)
    exp = O.amp_obs_demo(oracle_tables(motion_data_from(e)), smpl_step_config(), e["demo_ids"], e["demo_t0"], 1, 9)
    close(hist.cpu(), exp, rtol=1e-4, atol=2e-5, what="amp history init")


def test_amp_ring_equals_reference_window_shift():
    """The AMP ring (one slot written per step + export) reproduces the reference's per-step window shift."""
    n = 70
    m = syn.make_motions(n, seed=21, min_frames=60, max_frames=90)
    st = syn.make_env_state(m, n, seed=21, max_progress=10)
    mlib = pack(m)
    s = st.to(DEV)
    common = (mlib, s.body_state, s.dof_state, s.dof_force, s.progress, s.motion_ids, s.start_times, s.start_offsets, s.global_offset)
    shift = ops.EnvStepPlan(smpl_cfg(), *common, amp_obs_buf=s.amp_hist.clone())
    ring = ops.EnvStepPlan(smpl_cfg(), *common, amp_obs_buf=s.amp_hist.clone(), amp_ring=True)
    window = torch.zeros_like(ring.amp_obs_buf)
    for step in range(13):                                 # more than one full revolution of the 10-slot ring
        s.body_state.add_(0.01 * torch.randn_like(s.body_state))
        s.body_state[..., 3:7] = torch.nn.functional.normalize(s.body_state[..., 3:7], dim=-1)
        s.progress.add_(1)
        shift.run()
        ring.advance_ring()
        ring.run()
        ops.amp_window_export(ring.amp_obs_buf, ring.ring_head, window)
        torch.cuda.synchronize()
        assert torch.equal(window, shift.amp_obs_buf), f"step {step}"
        assert torch.equal(ring.obs, shift.obs) and torch.equal(ring.rew, shift.rew)


def test_env_step_future_tracks_vs_reference_golden():
    """fut_tracks (3 future samples, 0.1 s apart): CUDA vs outputs of the unmodified reference (tests/golden/fut.npz).  The ref_body_*
    side buffers are compared for env 0 only: under fut_tracks the reference stores env 0's first sample into every env's row
    (humanoid_im.py:857-861 indexes the flat [B*T, J, 3] tensor), a quirk that is documented, not mirrored."""
    g = load("fut.npz")
    st = syn.EnvState(**{k: g[f"in_{k}"] for k in syn.EnvState.__dataclass_fields__})
    plan = run_cuda_step(motion_data_from(g), st, smpl_cfg(time_steps=3, traj_dt=1 / 10))
    assert plan.obs.shape[1] == 358 + 3 * 576
    close(plan.obs.cpu(), g["out_obs"], atol=3e-6, what="fut obs")
    close(plan.rew.cpu(), g["out_rew"], what="fut rew")
    close(plan.reward_raw.cpu(), g["out_reward_raw"], what="fut reward_raw")
    close(plan.reset.cpu(), g["out_reset"], what="fut reset")
    close(plan.terminate.cpu(), g["out_terminate"], what="fut terminate")
    close(plan.amp_obs_buf.cpu(), g["out_amp_obs_buf"], what="fut amp_obs_buf")
    close(plan.ref_body_pos.cpu()[0], g["out_ref_body_pos"][0], atol=2e-6, what="fut ref_body_pos (env 0)")


def test_reset_path_pieces_vs_reference_golden():
    """History initialisation of freshly reset envs (HumanoidAMP._init_amp_obs_ref) and the start-time draw
    (MotionLibBase.sample_time_interval) against outputs of the unmodified reference (tests/golden/reset.npz)."""
    import ctypes as C
    from phc_b200 import _lib
    g = load("reset.npz")
    mlib = pack(motion_data_from(g))
    ids, t0 = g["motion_ids"].to(DEV), g["t0"].to(DEV)
    hist = ops.amp_obs_demo(mlib, smpl_cfg(), ids, t0, first_step=1, num_steps=9)
    close(hist.cpu(), g["hist_after"][g["env_ids"]], rtol=1e-4, atol=2e-5, what="_init_amp_obs_ref")
    # phc_reset_bookkeeping: start_times = trunc(phase * len / (1/30)) * (1/30) for the masked envs, counters cleared
    lib = _lib.load()
    n = int(ids.shape[0])
    em = torch.zeros(n, 4, dtype=torch.int32, device=DEV)
    _lib.check(lib.phc_env_motion_gather(C.byref(mlib.c), ids.data_ptr(), n, em.data_ptr(), torch.cuda.current_stream().cuda_stream))
    mask = torch.tensor([1, 1, 0, 1, 1], dtype=torch.int64, device=DEV)
    phase = g["phase"].to(DEV).contiguous()
    start = torch.full((n,), -1.0, device=DEV)
    off, goff = torch.full((n,), 7.0, device=DEV), torch.full((n, 3), 7.0, device=DEV)
    cc = torch.full((n,), 5, dtype=torch.int32, device=DEV)
    prog = torch.full((n,), 9, dtype=torch.int64, device=DEV)
    rst, term = torch.ones(n, dtype=torch.int64, device=DEV), torch.ones(n, dtype=torch.int64, device=DEV)
    _lib.check(lib.phc_reset_bookkeeping(mask.data_ptr(), phase.data_ptr(), em.data_ptr(), n, start.data_ptr(), off.data_ptr(), goff.data_ptr(),
                                         cc.data_ptr(), prog.data_ptr(), rst.data_ptr(), term.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    m = mask.bool().cpu()
    close(start.cpu()[m], g["sampled_times"][m], rtol=0, atol=1e-6, what="sample_time_interval")
    assert float(start.cpu()[~m][0]) == -1.0
    assert bool((off.cpu()[m] == 0).all()) and bool((goff.cpu()[m] == 0).all()) and bool((cc.cpu()[m] == 0).all())
    assert bool((prog.cpu()[m] == 0).all()) and bool((rst.cpu()[m] == 0).all()) and bool((term.cpu()[m] == 0).all())
    assert float(off.cpu()[~m][0]) == 7.0 and int(prog.cpu()[~m][0]) == 9


@pytest.mark.parametrize("tag", ["E", "F", "G"])
def test_tracked_subset_occlusion_and_shape_columns_vs_reference_golden(tag):
    """vr.npz from the unmodified reference (make_golden.gen_vr): env_vr.yaml's trackBodies = reset_bodies = Head + both hands with the
    subset reward and the shape / limb-weight columns of smpl_humanoid_shape.yaml (E), the subset with the full-body reward (F),
    occlusion training on the full body in observation and reset test (G)."""
    g = load("vr.npz")
    track = g["track"].tolist()
    subset = tag in ("E", "F")
    cfg = smpl_cfg(track_bodies=track if subset else None, reset_bodies=track if subset else syn.SMPL_RESET_BODIES, full_body_reward=tag != "E")
    st = syn.EnvState(**{k: g[f"in_{k}"] for k in syn.EnvState.__dataclass_fields__})
    kw = {}
    if tag == "E":
        kw = dict(shape_params=g["E_shape"][:, :-6].contiguous().to(DEV), limb_weights=g["E_limb"].to(DEV))
    if tag == "G":
        kw = dict(occlusion=g["G_occlusion"].to(DEV))
    plan = run_cuda_step(motion_data_from(g), st, cfg, **kw)
    assert plan.obs.shape[1] == g[f"{tag}_out_obs"].shape[1]
    check_against(plan, {k: g[f"{tag}_out_{k}"] for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos",
                                                          "ref_body_rot", "ref_body_vel")}, tag)


def test_humanoid_im_reads_track_bodies_and_occlusion_from_the_env_config():
    """env_vr.yaml through the task class: trackBodies / reset_bodies by NAME, full_body_reward False; and occlusion_training on the full
    body (the pattern _update_occl_training leaves behind) -- the step agrees with the oracle fed the same options."""
    from oracle import phc_oracle as O
    from phc_b200.env.humanoid_im import HumanoidIm
    from tests.helpers import oracle_tables, smpl_step_config
    n = 96
    m = syn.make_motions(n, seed=12, min_frames=40, max_frames=80)
    vr = ["Head", "L_Hand", "R_Hand"]
    ids = [syn.SMPL_BODY_NAMES.index(b) for b in vr]
    task = HumanoidIm({"env": {"num_envs": n, "trackBodies": vr, "reset_bodies": vr, "full_body_reward": False}, "motion_data": m, "seed": 1})
    assert task.get_task_obs_size() == 3 * 24 and task.get_obs_size() == 358 + 72 and task._reset_bodies_id == ids
    task.reset()
    for _ in range(2):
        hist = task._amp_obs_buf.cpu().clone()
        task.step(None)
        torch.cuda.synchronize()
        exp = O.env_step(oracle_tables(m), smpl_step_config(track_bodies=ids, reset_bodies=ids, full_body_reward=False), task._rigid_body_state_reshaped.cpu(),
                         task._dof_state.cpu(), task.dof_force_tensor.cpu(), task.progress_buf.cpu(), task._sampled_motion_ids.cpu(),
                         task._motion_start_times.cpu(), torch.zeros(n), torch.zeros(n, 3), hist)
        close(task.obs_buf.cpu(), exp["obs"], atol=2e-6, what="vr obs")
        close(task.rew_buf.cpu(), exp["rew"], what="vr rew")
        assert torch.equal(task.reset_buf.cpu(), exp["reset"]) and torch.equal(task._terminate_buf.cpu(), exp["terminate"])
    occ = HumanoidIm({"env": {"num_envs": n, "occlusion_training": True}, "motion_data": m, "seed": 1})
    occ.reset()
    hist = occ._amp_obs_buf.cpu().clone()
    occ.step(None)
    torch.cuda.synchronize()
    pattern = occ.random_occlu_idx.cpu()
    assert pattern[:, :9].all() and not pattern[:, 9:].any()          # what the reference's _update_occl_training ends with
    exp = O.env_step(oracle_tables(m), smpl_step_config(), occ._rigid_body_state_reshaped.cpu(), occ._dof_state.cpu(), occ.dof_force_tensor.cpu(),
                     occ.progress_buf.cpu(), occ._sampled_motion_ids.cpu(), occ._motion_start_times.cpu(), torch.zeros(n), torch.zeros(n, 3), hist,
                     occlusion=pattern)
    close(occ.obs_buf.cpu(), exp["obs"], atol=2e-6, what="occlusion obs")
    assert torch.equal(occ.reset_buf.cpu(), exp["reset"])
