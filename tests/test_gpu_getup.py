"""GPU parity of the env_im_getup_mcp.yaml extras of the fused step (zero_out_far + cycle_motion, the configuration
HumanoidImMCP trains in): CUDA through the C ABI vs the golden from the UNMODIFIED reference (tests/golden/getup.npz) and vs
the oracle at the bench size.  Tolerances as test_gpu_env_step.py; integers bit-exact."""
import pytest
import torch

from oracle import phc_oracle as O
from phc_b200 import ops, synthetic as syn
from tests.helpers import close, load, motion_data_from, oracle_tables, smpl_step_config
from tests.test_gpu_env_step import check_against, pack, smpl_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_getup(m, st, point_goal, cycle_counter, cycle_phase, max_len, **plan_kw):
    mlib = pack(m)
    s = st.to(DEV)
    extra = dict(point_goal=point_goal.to(DEV).clone(), cycle_counter=cycle_counter.to(DEV).to(torch.int32).clone(),
                 cycle_phase=cycle_phase.to(DEV).clone())
    cfg = smpl_cfg(zero_out_far=True, cycle_motion=True, max_episode_length=max_len)
    plan = ops.EnvStepPlan(cfg, mlib, s.body_state, s.dof_state, s.dof_force, s.progress, s.motion_ids, s.start_times,
                           s.start_offsets, s.global_offset, amp_obs_buf=s.amp_hist.clone(), with_ref_buffers=True, **extra, **plan_kw)
    plan.run()
    torch.cuda.synchronize()
    return plan, s, extra


def check_getup(plan, s, extra, exp, tag):
    check_against(plan, exp, tag)
    close(s.start_times.cpu(), exp["start_times"], what=f"{tag} start_times")
    close(s.start_offsets.cpu(), exp["start_offsets"], what=f"{tag} start_offsets")
    close(s.global_offset.cpu(), exp["global_offset"], what=f"{tag} global_offset")
    close(extra["point_goal"].cpu(), exp["point_goal"], what=f"{tag} point_goal")
    assert torch.equal(extra["cycle_counter"].cpu().long(), exp["cycle_counter"].long()), f"{tag} cycle_counter"


def test_getup_step_vs_reference_golden():
    g = load("getup.npz")
    st = syn.EnvState(**{k: g[f"in_{k}"] for k in syn.EnvState.__dataclass_fields__})
    plan, s, extra = run_getup(motion_data_from(g), st, g["in_point_goal"], g["in_cycle_counter"], g["in_cycle_phase"], 15)
    exp = {k[4:]: v for k, v in g.items() if k.startswith("out_")}
    check_getup(plan, s, extra, exp, "getup golden")


def _far_state(n, seed):
    m = syn.make_motions(n, seed=seed, min_frames=30, max_frames=90)
    st = syn.make_env_state(m, n, seed=seed, max_progress=120, with_offset=True)     # progress beyond many clip ends -> wraps
    g = torch.Generator().manual_seed(seed)
    k = n // 4
    st.global_offset[:k, :2] += torch.randn(k, 2, generator=g) * 4.0
    st.global_offset[k:2 * k, :2] += torch.randn(k, 2, generator=g) * 0.8
    st.global_offset[2 * k:2 * k + k // 2, :2] += torch.randn(k // 2, 2, generator=g) * 0.15
    cc = torch.tensor([0, 0, 0, 1, 2, 7], dtype=torch.int32)[torch.randint(0, 6, (n,), generator=g)]
    return m, st, torch.rand(n, generator=g) * 6, cc, torch.rand(n, generator=g)


@pytest.mark.parametrize("n,seed", [(3, 1), (257, 2), (4096, 3)])
def test_getup_step_vs_oracle(n, seed):
    m, st, pg, cc, ph = _far_state(n, seed)
    exp = O.env_step_getup(oracle_tables(m), smpl_step_config(), st.body_state, st.dof_state, st.dof_force, st.progress, st.motion_ids,
                           st.start_times, st.start_offsets, st.global_offset, st.amp_hist, pg, cc, ph, max_episode_length=100)
    plan, s, extra = run_getup(m, st, pg, cc, ph, 100)
    check_getup(plan, s, extra, exp, f"getup n={n}")
    if n >= 257:
        wrapped = (exp["cycle_counter"] == 60).sum()
        assert wrapped > 0 and (exp["reward_raw"][:, 1] == 0).sum() > 0 and (exp["reward_raw"][:, 1] != 0).sum() > 0


def test_getup_instantiation_without_events_is_the_plain_step():
    """cycle_motion on, but no clip wraps and no counter is running: the GETUP instantiation must reproduce the plain kernel
    (a separate template instantiation, so equality is asserted at the parity tolerance, integers exactly)."""
    n = 512
    m = syn.make_motions(n, seed=8, min_frames=200, max_frames=260)
    st = syn.make_env_state(m, n, seed=8, max_progress=20)
    st.start_times.mul_(0.5)                       # no env reaches the end of its clip
    mlib = pack(m)
    s1, s2 = st.to(DEV), st.to(DEV)
    plain = ops.EnvStepPlan(smpl_cfg(), mlib, s1.body_state, s1.dof_state, s1.dof_force, s1.progress, s1.motion_ids, s1.start_times,
                            s1.start_offsets, s1.global_offset, amp_obs_buf=s1.amp_hist.clone(), with_ref_buffers=True)
    plain.run()
    cyc = ops.EnvStepPlan(smpl_cfg(cycle_motion=True, max_episode_length=300), mlib, s2.body_state, s2.dof_state, s2.dof_force, s2.progress,
                          s2.motion_ids, s2.start_times, s2.start_offsets, s2.global_offset, amp_obs_buf=s2.amp_hist.clone(),
                          with_ref_buffers=True, cycle_counter=torch.zeros(n, dtype=torch.int32, device=DEV), cycle_phase=torch.rand(n, device=DEV))
    cyc.run()
    torch.cuda.synchronize()
    exp = {k: getattr(plain, k).cpu() for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos", "ref_body_rot", "ref_body_vel")}
    check_against(cyc, exp, "getup-idle")
    assert torch.equal(s2.start_times, s1.start_times) and torch.equal(s2.global_offset, s1.global_offset)


def test_getup_env_runs_with_and_without_pose_cache():
    """HumanoidIm in the getup configuration for 90 steps (clips of 1-2 s wrap several times): the pose-cache path and the
    re-interpolating path stay bit-identical through wrap-arounds and resets, and the motion time never runs past the clip."""
    from phc_b200.env.humanoid_im import HumanoidIm
    n = 256
    outs = []
    for cache in (True, False):
        torch.manual_seed(0)
        m = syn.make_motions(n, seed=12, min_frames=30, max_frames=60)
        env = HumanoidIm({"env": {"num_envs": n, "cycle_motion": True, "zero_out_far": True, "zero_out_far_train": False,
                                  "episode_length": 70, "enableEarlyTermination": False}, "motion_data": m, "seed": 3, "ref_pose_cache": cache},
                         device_type="cuda", device_id=0)
        env.reset()
        trace = []
        for step in range(90):
            env.step(None)
            t_now = env.progress_buf * env.dt + env._motion_start_times + env._motion_start_times_offset
            assert bool((t_now <= env._motion_lib.lengths[env._sampled_motion_ids] + 2 * env.dt).all()), f"step {step}: clip overrun"
            trace.append((env.obs_buf.clone(), env.rew_buf.clone(), env.reset_buf.clone(), env._cycle_counter.clone(), env._point_goal.clone()))
            done = env.reset_buf.nonzero().flatten()
            if len(done):
                env.reset(done)
        outs.append(trace)
        assert int(max(t[3].max() for t in trace)) == 60          # a clip wrapped
    for a, b in zip(*outs):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
