"""GPU parity of the hinge-joint robot path (Unitree H1: 20 bodies + 3 extend bodies in the reward, 19 one-dof joints,
obs 778, AMP 63 -- config 5 of BASELINE.json) against outputs of the real reference (tests/golden/h1.npz: MotionLibReal,
HumanoidIm h1 branches, build_amp_observations_robot) and against the oracle at 4096 envs.
Tolerances as in test_gpu_env_step.py: rtol 1e-5 / atol 1e-6 (2e-6 on observations), integers bit-exact."""
import pytest
import torch

from oracle import phc_oracle as O
from phc_b200 import ops, synthetic as syn
from tests.helpers import close, env_state_from, h1_step_config, load, robot_motion_data_from, robot_tables_from

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def h1_cfg(**kw):
    base = dict(key_bodies=syn.H1_KEY_BODIES, reset_bodies=None, dof_subset=None, ext_parents=syn.H1_EXT_PARENTS, ext_pos=syn.H1_EXT_POS)
    base.update(kw)
    return ops.EnvStepConfig(**base)


def pack(m: syn.RobotMotionData):
    t = lambda x: x.to(DEV)
    return ops.pack_robot_motion_lib(t(m.gts_t), t(m.grs_t), t(m.gvs_t), t(m.gavs_t), t(m.dof_pos), t(m.dvs), m.num_bodies, t(m.lengths),
                                     t(m.num_frames), t(m.dts), t(m.length_starts))


def tables(m: syn.RobotMotionData):
    return O.RobotTables(m.gts_t, m.grs_t, m.gvs_t, m.gavs_t, m.dof_pos, m.dvs, m.lengths, m.num_frames, m.dts, m.length_starts, m.num_bodies)


def run_cuda_step(mlib, st, cfg, **kw):
    s = st.to(DEV)
    plan = ops.EnvStepPlan(cfg, mlib, s.body_state, s.dof_state, s.dof_force, s.progress, s.motion_ids, s.start_times, s.start_offsets,
                           s.global_offset, amp_obs_buf=s.amp_hist.clone(), with_ref_buffers=True, **kw)
    plan.run()
    torch.cuda.synchronize()
    return plan


def check(plan, exp, tag):
    close(plan.obs.cpu(), exp["obs"], atol=2e-6, what=f"{tag} obs")
    for k in ("rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos", "ref_body_rot", "ref_body_vel"):
        close(getattr(plan, k).cpu(), exp[k], what=f"{tag} {k}")


def test_h1_dims():
    mlib = pack(syn.make_robot_motions(2, seed=0, min_frames=10, max_frames=12))
    assert (mlib.num_bodies, mlib.num_ext_bodies, mlib.num_dofs, mlib.dofs) == (20, 3, 19, 19)
    assert mlib.frames_body.shape[1] == 300 and mlib.frames_joint.shape[1] == 40


def test_h1_motion_state_vs_motion_lib_real_golden():
    g = load("h1.npz")
    mlib = pack(robot_motion_data_from(g))
    out = ops.motion_state(mlib, g["ms_ids"].to(DEV), g["ms_times"].to(DEV), g["ms_offset"].to(DEV))
    torch.cuda.synchronize()
    for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel", "dof_pos", "dof_vel", "rg_pos_t", "rg_rot_t", "body_vel_t", "body_ang_vel_t"):
        close(out[k].cpu(), g["ms_out_" + k], what="h1 motion_state " + k)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_h1_env_step_vs_reference_golden(tag):
    g = load("h1.npz")
    plan = run_cuda_step(pack(robot_motion_data_from(g)), env_state_from(g, tag), h1_cfg())
    assert plan.obs_dim == 778 and plan.amp_dim == 63
    check(plan, {k: g[f"{tag}_out_{k}"] for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos",
                                                   "ref_body_rot", "ref_body_vel")}, "h1 " + tag)


def test_h1_amp_obs_demo_vs_reference_golden():
    g = load("h1.npz")
    out = ops.amp_obs_demo(pack(robot_motion_data_from(g)), h1_cfg(), g["demo_ids"].to(DEV), g["demo_t0"].to(DEV))
    torch.cuda.synchronize()
    close(out.cpu(), g["demo_out"], what="h1 amp_obs_demo")


def test_h1_env_step_4096_vs_oracle_and_pose_cache():
    """BASELINE config 5 size: 4096 envs; then the pose-cache launch (reward from the cached J + E records) gives the same bits."""
    n = 4096
    m = syn.make_robot_motions(64, seed=5, min_frames=40, max_frames=80)
    st = syn.make_robot_env_state(m, n, seed=6, with_offset=True, blend_jitter=True)
    mlib = pack(m)
    plan = run_cuda_step(mlib, st, h1_cfg())
    exp = O.env_step_robot(tables(m), h1_step_config(), syn.H1_EXT_PARENTS, syn.H1_EXT_POS, st.body_state, st.dof_state, st.dof_force,
                           st.progress, st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    check(plan, exp, "h1 4096")
    assert int(plan.terminate.sum()) > 0
    # cached reward reference: run the previous step's observation launch (progress - 1) to fill the cache, then the step
    s = st.to(DEV)
    cache = torch.zeros(n, mlib.frames_body.shape[1], device=DEV)
    prev = ops.EnvStepPlan(h1_cfg(), mlib, s.body_state, s.dof_state, s.dof_force, s.progress - 1, s.motion_ids, s.start_times,
                           s.start_offsets, s.global_offset, obs_only=True, with_amp=False, ref_cache=cache)
    prev.run()
    ms = O.motion_state_robot(tables(m), st.motion_ids, st.progress * (1.0 / 30.0) + st.start_times + st.start_offsets, st.global_offset)
    exp_cache = torch.zeros(n, 300)
    exp_cache[:, :299] = torch.cat([ms[k] for k in ("rg_pos_t", "rg_rot_t", "body_vel_t", "body_ang_vel_t")], dim=-1).reshape(n, -1)
    torch.cuda.synchronize()
    close(cache.cpu(), exp_cache, what="pose cache rows (J + E records)")
    cached = ops.EnvStepPlan(h1_cfg(), mlib, s.body_state, s.dof_state, s.dof_force, s.progress, s.motion_ids, s.start_times,
                             s.start_offsets, s.global_offset, amp_obs_buf=s.amp_hist.clone(), ref_cache=cache, reward_from_cache=True)
    cached.run()
    torch.cuda.synchronize()
    for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf"):
        assert torch.equal(getattr(cached, k), getattr(plan, k)), f"cached vs re-interpolated: {k} differs"


def test_h1_task_rollout_and_agent_epoch():
    """HumanoidIm with robot tables (humanoid_type h1): reset / step against the oracle, then one AMPAgent epoch."""
    from phc_b200.env.humanoid_im import HumanoidIm, RLGPUEnv
    from phc_b200.learning.amp_agent import AMPAgent
    n = 64
    m = syn.make_robot_motions(n, seed=8, min_frames=40, max_frames=90)
    task = HumanoidIm({"env": {"num_envs": n}, "motion_data": m, "seed": 8})
    assert task.humanoid_type == "h1" and task.get_obs_size() == 778 and task.get_action_size() == 19 and task.get_num_amp_obs() == 630
    task.reset()
    hist = task._amp_obs_buf.cpu().clone()
    task.step(None)
    torch.cuda.synchronize()
    exp = O.env_step_robot(tables(m), h1_step_config(), syn.H1_EXT_PARENTS, syn.H1_EXT_POS, task._rigid_body_state_reshaped.cpu(),
                           task._dof_state.cpu(), task.dof_force_tensor.cpu(), task.progress_buf.cpu(), task._sampled_motion_ids.cpu(),
                           task._motion_start_times.cpu(), torch.zeros(n), torch.zeros(n, 3), hist)
    close(task.obs_buf.cpu(), exp["obs"], atol=2e-6, what="task obs")
    close(task.rew_buf.cpu(), exp["rew"], what="task rew")
    close(task.reset_buf.cpu(), exp["reset"], what="task reset")
    close(task._amp_obs_buf.cpu(), exp["amp_obs_buf"], what="task amp window")
    agent = AMPAgent("t", {"vec_env": RLGPUEnv(task), "horizon_length": 8, "minibatch_size": 256, "amp_minibatch_size": 64,
                           "mini_epochs": 2, "amp_obs_demo_buffer_size": 2048, "amp_replay_buffer_size": 2048, "amp_batch_size": 128,
                           "network": {"mlp": {"units": [128, 64], "activation": "relu"}, "disc": {"units": [128, 64], "activation": "relu"}}})
    agent.obs = agent.env_reset()
    agent._init_amp_demo_buf()
    p0 = agent.model.params.clone()
    agent.train_epoch()
    torch.cuda.synchronize()
    assert torch.isfinite(agent.model.params).all() and not torch.equal(agent.model.params, p0)
