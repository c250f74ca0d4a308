"""GPU tests of the host mirrors: HumanoidIm (reset path with masks, step) against the oracle, and AMPAgent end to end
(rollout + update on small sizes: finite, parameters move, buffers consistent)."""
import pytest
import torch

from oracle import phc_oracle as O
from phc_b200 import synthetic as syn
from phc_b200.env.humanoid_im import HumanoidIm, RLGPUEnv
from phc_b200.learning.amp_agent import AMPAgent
from tests.helpers import close, oracle_tables, smpl_step_config

pytestmark = pytest.mark.gpu


def make_task(n, seed=0):
    m = syn.make_motions(n, seed=seed, min_frames=40, max_frames=90)
    return m, HumanoidIm({"env": {"num_envs": n}, "motion_data": m, "seed": seed})


def test_reset_then_step_match_oracle():
    n = 200
    m, task = make_task(n, seed=1)
    tab, cfg = oracle_tables(m), smpl_step_config()
    obs = task.reset()                               # first launch is the (smaller-smem) obs-only plan: regression for the
    torch.cuda.synchronize()                         # cudaFuncSetAttribute ordering bug found by bench.py
    ids, t0 = task._sampled_motion_ids.cpu(), task._motion_start_times.cpu()
    assert int(task.progress_buf.abs().sum()) == 0
    # reference-state init: the simulator tensors hold the reference pose at the sampled start time
    st = O.motion_state(tab, ids, t0, torch.zeros(n, 3))
    body = task._rigid_body_state_reshaped.cpu()
    close(body[..., 0:3], st["rg_pos"], what="reset pos")
    close(body[..., 3:7], st["rb_rot"], what="reset rot")
    close(task._dof_state.cpu()[..., 0], st["dof_pos"], rtol=1e-4, atol=2e-5, what="reset dof_pos")
    # AMP window re-initialised from the reference motion (current + 9 history slots)
    close(task._amp_obs_buf.cpu(), O.amp_obs_demo(tab, cfg, ids, t0), rtol=1e-4, atol=2e-5, what="amp window after reset")
    # observation of the reset envs: self obs + task obs at t0 + dt
    nxt = O.motion_state(tab, ids, (torch.zeros(n) + 1) * cfg.dt + t0, torch.zeros(n, 3))
    bp, br, bv, bw = body[..., 0:3], body[..., 3:7], body[..., 7:10], body[..., 10:13]
    exp_obs = torch.cat((O.self_obs(bp, br, bv, bw), O.task_obs_v6(bp[:, 0], br[:, 0], bp, br, bv, bw, nxt["rg_pos"], nxt["rb_rot"],
                                                                     nxt["body_vel"], nxt["body_ang_vel"])), dim=-1)
    # 5e-6: env_step.cu is compiled with FMA contraction on, the oracle rounds every product; the angular-velocity differences (values
    # of order 1-10, rotated into the heading frame) carry a few ulps of that (seen: 1 of 186800 entries at 3.1e-6)
    close(obs.cpu(), exp_obs, atol=5e-6, what="obs after reset")

    # one env step on a fresh simulator snapshot
    hist = task._amp_obs_buf.cpu().clone()
    task.step(None)
    torch.cuda.synchronize()
    exp = O.env_step(tab, cfg, task._rigid_body_state_reshaped.cpu(), task._dof_state.cpu(), task.dof_force_tensor.cpu(),
                     task.progress_buf.cpu(), ids, t0, torch.zeros(n), torch.zeros(n, 3), hist)
    close(task.obs_buf.cpu(), exp["obs"], atol=5e-6, what="obs")
    close(task.rew_buf.cpu(), exp["rew"], what="rew")
    close(task.reset_buf.cpu(), exp["reset"], what="reset")
    close(task._amp_obs_buf.cpu(), exp["amp_obs_buf"], what="amp window")

    # masked reset: only flagged envs change
    mask = (torch.arange(n, device=task.device) % 3 == 0).long()
    before_obs, before_prog, before_start = task.obs_buf.clone(), task.progress_buf.clone(), task._motion_start_times.clone()
    task.reset(mask)
    torch.cuda.synchronize()
    keep = mask == 0
    assert torch.equal(task.obs_buf[keep], before_obs[keep]) and torch.equal(task.progress_buf[keep], before_prog[keep])
    assert torch.equal(task._motion_start_times[keep], before_start[keep])
    assert int(task.progress_buf[mask == 1].abs().sum()) == 0 and not torch.equal(task.obs_buf[mask == 1], before_obs[mask == 1])
    # reference-style index list selects the same envs
    task.reset(torch.nonzero(mask).flatten())
    assert torch.equal(task._reset_mask, mask)


def test_agent_epoch_small():
    n = 64
    _, task = make_task(n, seed=2)
    agent = AMPAgent("t", {"vec_env": RLGPUEnv(task), "horizon_length": 8, "minibatch_size": 256, "amp_minibatch_size": 64,
                           "mini_epochs": 2, "amp_obs_demo_buffer_size": 2048, "amp_replay_buffer_size": 2048,
                           "amp_batch_size": 128, "network": {"mlp": {"units": [128, 64], "activation": "relu"},
                                                              "disc": {"units": [128, 64], "activation": "relu"}}})
    agent.obs = agent.env_reset()
    agent._init_amp_demo_buf()
    p0 = agent.model.params.clone()
    mean0 = agent.running_mean_std.running_mean.clone()
    info = None
    for _ in range(2):
        info = agent.train_epoch()
    torch.cuda.synchronize()
    assert torch.isfinite(agent.model.params).all() and not torch.equal(agent.model.params, p0)
    assert agent.opt_step == 2 * 2 * (8 * n // 256)
    assert not torch.equal(agent.running_mean_std.running_mean, mean0) and float(agent.running_mean_std.count) > 1
    assert torch.isfinite(info["returns"]).all() and torch.isfinite(info["disc_rewards"]).all()
    r = agent.train_result_dict()
    assert all(v == v for v in r.values())          # no NaN
    # GAE bookkeeping: returns = advantages + values on the flattened rollout
    eb = agent.experience_buffer
    adv = O.gae(eb["dones"].cpu(), eb["values"].cpu(), (0.5 * eb["rewards"] + 0.5 * info["disc_rewards"].view(n, 8, 1).transpose(0, 1)).cpu(),
                eb["next_values"].cpu(), 0.99, 0.95)
    close(info["returns"].view(n, 8, 1).transpose(0, 1).cpu(), adv + eb["values"].cpu(), rtol=1e-4, atol=1e-4, what="returns")
    # checkpoint round trip with the reference's keys
    sd = agent.get_full_state_weights()
    assert {"model", "running_mean_std", "reward_mean_std", "amp_input_mean_std"} <= set(sd)
    assert "a2c_network.actor_mlp.0.weight" in sd["model"] and sd["model"]["a2c_network.actor_mlp.0.weight"].shape == (128, 934)
    agent.model.params.zero_()
    agent.set_full_state_weights(sd)
    assert torch.equal(agent.model.state_dict()["a2c_network._disc_logits.weight"], sd["model"]["a2c_network._disc_logits.weight"])


@pytest.mark.parametrize("n", [4096, 16384])
def test_specialised_step_kernel_matches_oracle_at_bench_sizes(n):
    """The kernel bench.py's roofline object times -- env_step_kernel<1, 24, false, FAST=true> at 4096 envs (and the 16384-env
    point) -- against the pinned oracle over three consecutive steps (pose cache, AMP ring with the head on the device,
    per-env motion records): observations, reward terms, reset / terminate flags (bit-exact) and the AMP window."""
    from phc_b200 import _lib
    lib = _lib.load()
    m = syn.make_motions(min(n, 4096), seed=17, min_frames=40, max_frames=90)       # 16384 envs share 4096 clips
    task = HumanoidIm({"env": {"num_envs": n}, "motion_data": m, "seed": 17})
    tab, cfg = oracle_tables(m), smpl_step_config()
    torch.manual_seed(3)
    task.reset()
    torch.cuda.synchronize()
    ids, t0 = task._sampled_motion_ids.cpu(), task._motion_start_times.cpu()
    f0 = lib.phc_env_step_fast_launches()
    for step in range(3):
        hist = task._amp_obs_buf.cpu().clone()
        task.step(None)
        torch.cuda.synchronize()
        exp = O.env_step(tab, cfg, task._rigid_body_state_reshaped.cpu(), task._dof_state.cpu(), task.dof_force_tensor.cpu(),
                         task.progress_buf.cpu(), ids, t0, torch.zeros(n), torch.zeros(n, 3), hist)
        # 15 M observation entries per step at 16384 envs: the velocity-difference columns subtract O(30 m/s) operands, one ulp of
        # which (3.8e-6) survives the cancellation in a handful of entries -- atol 5e-6 here, 2e-6 in the small-batch tests
        close(task.obs_buf.cpu(), exp["obs"], atol=5e-6, what=f"obs (step {step}, {n} envs)")
        close(task.rew_buf.cpu(), exp["rew"], what=f"rew (step {step})")
        close(task.reward_raw.cpu(), exp["reward_raw"], what=f"reward_raw (step {step})")
        assert torch.equal(task.reset_buf.cpu(), exp["reset"]) and torch.equal(task._terminate_buf.cpu(), exp["terminate"]), f"step {step}: reset / terminate"
        close(task._amp_obs_buf.cpu(), exp["amp_obs_buf"], what=f"amp window (step {step})")
    assert lib.phc_env_step_fast_launches() - f0 == 3, "the steady-state launches must take the specialised instantiation"


def test_rollout_dataset_and_minibatch_gather_path_vs_oracle():
    """what happens BETWEEN the rollout and the gradient -- discount_values, the combined reward,
    prepare_dataset (advantage / value normalisation, swap_and_flatten01), the index-composed minibatch (rows addressed through
    idx, replay / demo rows through composed indices, gathered inside the normalisation kernels) -- replayed on the CPU with
    plain torch indexing from the SAME experience buffer and fed to the pinned oracle update; losses and every parameter
    gradient must agree with what the agent's pipelined minibatch produced."""
    from oracle import ppo_oracle as PO
    n, T, mb, Bd = 64, 8, 256, 64
    m, task = make_task(n, seed=9)
    net_cfg = {"mlp": {"units": [128, 64], "activation": "relu"}, "disc": {"units": [128, 64], "activation": "relu"}}
    agent = AMPAgent("t", {"vec_env": RLGPUEnv(task), "horizon_length": T, "minibatch_size": mb, "amp_minibatch_size": Bd, "mini_epochs": 1,
                           "amp_obs_demo_buffer_size": 1024, "amp_replay_buffer_size": 1024, "amp_batch_size": 128, "network": net_cfg,
                           "graph_rollout": False})
    agent.obs = agent.env_reset()
    agent._init_amp_demo_buf()
    agent.train_epoch()                                   # statistics, replay buffer and weights leave their initial state
    cpu = lambda t: t.detach().cpu().clone()
    bd = agent.play_steps()
    torch.cuda.synchronize()
    eb = {k: cpu(v) for k, v in agent.experience_buffer.items()}
    # --- rollout side on the CPU: rewards, GAE, returns
    rew = agent._task_reward_w * eb["rewards"] + agent._disc_reward_w * cpu(bd["disc_rewards"]).view(n, T, 1).transpose(0, 1)
    adv = O.gae(eb["dones"], eb["values"], rew, eb["next_values"], agent.gamma, agent.tau)
    flat = lambda t: t.transpose(0, 1).reshape(n * T, *t.shape[2:])
    close(cpu(bd["mb_advs"]), flat(adv), rtol=1e-4, atol=1e-5, what="GAE advantages of the rollout")
    returns, values = flat(adv + eb["values"]), flat(eb["values"])
    # --- the epoch's bookkeeping exactly as train_epoch does it, then ONE pipelined minibatch
    agent._update_amp_demos()
    N = bd["amp_obs"].shape[0]
    bd["amp_obs_demo_idx"] = agent._amp_obs_demo_buffer.sample_indices(N)
    agent._amp_replay_src = agent._amp_replay_buffer.data
    bd["amp_obs_replay_idx"] = agent._amp_replay_buffer.sample_indices(N)
    vms = agent.value_mean_std
    v_mean, v_var, v_cnt = cpu(vms.running_mean), cpu(vms.running_var), cpu(vms.count)
    agent.set_train()
    agent.prepare_dataset(bd)
    ds = agent.dataset
    exp_adv = O.normalize_advantages(returns, values)
    close(cpu(ds["advantages"]), exp_adv, rtol=1e-4, atol=1e-5, what="normalised advantages")
    close(cpu(ds["old_values"]), O.rms_normalize(values, v_mean, v_var), rtol=1e-5, atol=1e-6, what="normalised old values")
    v_mean2, v_var2, _ = O.rms_update(v_mean, v_var, v_cnt, values.double())
    exp_ret = O.rms_normalize(returns, v_mean2, v_var2)   # the second call sees the statistics the first one updated
    close(cpu(ds["returns"]), exp_ret, rtol=1e-4, atol=1e-5, what="normalised returns")
    idx = agent._idx_buf[:mb]
    obs_mean, obs_var = cpu(agent.running_mean_std_temp.running_mean), cpu(agent.running_mean_std_temp.running_var)
    a_mean, a_var, a_cnt = cpu(agent._amp_input_mean_std.running_mean), cpu(agent._amp_input_mean_std.running_var), cpu(agent._amp_input_mean_std.count)
    sd = {k: cpu(v) for k, v in agent.model.state_dict().items()}
    replay_rows, demo_rows = cpu(agent._amp_replay_buffer.data), cpu(agent._amp_obs_demo_buffer.data)
    prepared = agent._prepare_minibatch(idx)
    agent._compute_gradients(idx, prepared)
    torch.cuda.synchronize()
    # --- the same minibatch with plain indexing
    i = idx.cpu()
    ia = i[:Bd]
    amp_flat = flat(eb["amp_obs"])
    blocks = [amp_flat[ia], replay_rows[cpu(bd["amp_obs_replay_idx"])[ia]], demo_rows[cpu(bd["amp_obs_demo_idx"])[ia]]]
    normed = []
    for blk in blocks:                                     # three RunningMeanStd.forward calls in train mode: normalise, then update
        normed.append(O.rms_normalize(blk, a_mean, a_var))
        a_mean, a_var, a_cnt = O.rms_update(a_mean, a_var, a_cnt, blk.double())
    batch = dict(obs_n=O.rms_normalize(flat(eb["obses"])[i], obs_mean, obs_var), actions=flat(eb["actions"])[i], old_neglogp=flat(eb["neglogpacs"])[i],
                 advantages=exp_adv[i], old_mu=flat(eb["mus"])[i], old_sigma=flat(eb["sigmas"])[i], returns=exp_ret[i],
                 amp_agent=normed[0], amp_replay=normed[1], amp_demo=normed[2])
    close(cpu(prepared[0])[:, :agent.obs_dim], batch["obs_n"], rtol=1e-5, atol=1e-6, what="gathered + normalised observations")
    close(cpu(prepared[1])[:, :agent.amp_obs_dim], torch.cat(normed), rtol=1e-5, atol=2e-6, what="gathered + normalised AMP rows [agent | replay | demo]")
    cfg = dict(e_clip=agent.e_clip, critic_coef=agent.critic_coef, entropy_coef=agent.entropy_coef, bounds_loss_coef=agent.bounds_loss_coef,
               disc_coef=agent._disc_coef, disc_logit_reg=agent._disc_logit_reg, disc_grad_penalty=agent._disc_grad_penalty,
               disc_weight_decay=agent._disc_weight_decay, grad_norm=agent.grad_norm, learning_rate=agent.last_lr, truncate_grads=True)
    exp = PO.minibatch_update(sd, batch, cfg, n_hidden=2, dtype=torch.float64, mu_override=cpu(agent._ws_actor["out"])[:, :agent.actions_num].double())
    r = agent.train_result_dict()
    close(torch.tensor(r["critic_loss"]), exp["c_loss"].float(), rtol=2e-4, atol=1e-5, what="critic loss")
    close(torch.tensor(r["actor_loss"]), exp["a_loss"].float(), rtol=1e-3, atol=1e-4, what="actor loss")
    close(torch.tensor(r["disc_grad_penalty"]), exp["disc"]["disc_grad_penalty"].float(), rtol=1e-3, atol=1e-6, what="gradient penalty")
    net = agent.model
    for l in net.all_layers():
        for kind, got in (("weight", net.weight(l, True)[:, :l.in_dim]), ("bias", net.bias(l, True))):
            ge = exp["grads"][f"a2c_network.{l.name}.{kind}"]
            err, sc = float((cpu(got).double() - ge).abs().max()), float(ge.abs().max()) + 1e-30
            assert err <= 2e-3 * sc, f"gradient of {l.name}.{kind}: {err:.3e} vs scale {sc:.3e}"   # a mis-routed row would be O(1)


def test_graph_replayed_rollout_equals_eager_rollout():
    """AMPAgent.play_steps as ONE CUDA graph (second rollout captured, later ones replayed) against the eager loop: same seeds,
    same generator state -> the same experience buffer, bit for bit, epoch after epoch (resets, AMP ring head on the device,
    carried episode statistics and the random draws inside the graph included)."""
    n = 96
    cfgs = {"horizon_length": 8, "minibatch_size": 256, "amp_minibatch_size": 64, "mini_epochs": 1, "amp_obs_demo_buffer_size": 1024,
            "amp_replay_buffer_size": 1024, "amp_batch_size": 128,
            "network": {"mlp": {"units": [64, 32], "activation": "relu"}, "disc": {"units": [64, 32], "activation": "relu"}}}
    agents = []
    for graphed in (False, True):
        _, task = make_task(n, seed=5)
        torch.manual_seed(1234)
        ag = AMPAgent("t", dict(cfgs, vec_env=RLGPUEnv(task), graph_rollout=graphed))
        ag.obs = ag.env_reset()
        ag._init_amp_demo_buf()
        agents.append(ag)
    assert agents[1]._graph_rollout and not agents[0]._graph_rollout
    for epoch in range(4):
        outs = []
        for ag in agents:
            torch.manual_seed(77 + epoch)
            ag.set_eval()
            bd = ag.play_steps()
            torch.cuda.synchronize()
            outs.append({k: v.clone() for k, v in ag.experience_buffer.items()} | {"returns": bd["returns"].clone(), "cur_rew": ag.current_rewards.clone()})
        for k in outs[0]:
            assert torch.equal(outs[0][k], outs[1][k]), f"epoch {epoch}: {k} differs between the eager and the graph-replayed rollout"
    assert agents[1]._rollout_graph is not None and agents[1]._rollout_graph_launches > 8 * 10


def test_deferred_critic_rollout_equals_per_step_critic():
    """values / next_values evaluated after the loop on the stored obses / next_obses (chunks of the minibatch size, obs and next-obs
    chunks in the same grouped launches) against the reference's order (critic inside every step, amp_agent.py:329-372): row results
    of the GEMM do not depend on the batching -> the same experience buffer, returns and advantages bit for bit; the ragged last
    chunk (8 x 96 = 768 rows in chunks of 256 / 3 x 96 = 288 rows in chunks of 256) included."""
    n = 96
    for horizon in (8, 3):
        cfgs = {"horizon_length": horizon, "minibatch_size": 256 if horizon == 8 else 96, "amp_minibatch_size": 64, "mini_epochs": 1,
                "amp_obs_demo_buffer_size": 1024, "amp_replay_buffer_size": 1024, "amp_batch_size": 128, "graph_rollout": False,
                "network": {"mlp": {"units": [64, 32], "activation": "relu"}, "disc": {"units": [64, 32], "activation": "relu"}}}
        agents = []
        for deferred in (False, True):
            _, task = make_task(n, seed=5)
            torch.manual_seed(1234)
            ag = AMPAgent("t", dict(cfgs, vec_env=RLGPUEnv(task), deferred_critic=deferred))
            if horizon == 3:
                ag.minibatch_size = 256                 # 288 rows in chunks of 256: the workspaces below are re-made for it
                ag._x_mb = torch.zeros(256, ag.obs_pad, device=ag.device)
                ag._ws_critic = ag.engine.workspace("critic", ag.model.critic, 256)
                if deferred:
                    ag._x_mb2 = torch.zeros(256, ag.obs_pad, device=ag.device)
                    ag._ws_critic2 = ag.engine.workspace("critic2", ag.model.critic, 256)
            ag.obs = ag.env_reset()
            ag._init_amp_demo_buf()
            agents.append(ag)
        assert agents[1]._defer_critic and not agents[0]._defer_critic
        for epoch in range(2):
            outs = []
            for ag in agents:
                torch.manual_seed(77 + epoch)
                ag.set_eval()
                bd = ag.play_steps()
                torch.cuda.synchronize()
                outs.append({k: v.clone() for k, v in ag.experience_buffer.items()} | {"returns": bd["returns"].clone(), "advs": bd["mb_advs"].clone()})
            assert float(outs[0]["next_values"].abs().sum()) > 0
            for k in outs[0]:
                assert torch.equal(outs[0][k], outs[1][k]), f"horizon {horizon} epoch {epoch}: {k} differs between the per-step and the deferred critic"


def test_ref_pose_cache_is_bit_identical_to_reinterpolation():
    """PHC_FLAG_REWARD_FROM_CACHE (the pose interpolated for the observation of step s is the reward-time pose of step
    s+1) against re-interpolating every step: identical bits over a rollout with resets in between."""
    n = 160
    m = syn.make_motions(n, seed=3, min_frames=20, max_frames=40)          # short clips: some envs run past the end
    # both through the SAME (generic) kernel instantiation: the property is the cache, not the code generation of two templates
    tasks = [HumanoidIm({"env": {"num_envs": n}, "motion_data": m, "seed": 3, "ref_pose_cache": c, "specialised_step": False}) for c in (True, False)]
    assert tasks[0]._use_ref_cache and not tasks[1]._use_ref_cache
    for t in tasks:
        torch.manual_seed(99)                    # start times are drawn from the global generator: same draws for both
        t.reset()
    for step in range(12):
        for t in tasks:
            t.step(None)
        torch.cuda.synchronize()
        a, b = tasks
        for name in ("obs_buf", "rew_buf", "reward_raw", "reset_buf", "_terminate_buf", "progress_buf"):
            assert torch.equal(getattr(a, name), getattr(b, name)), f"step {step}: {name} differs"
        assert torch.equal(a._amp_obs_buf, b._amp_obs_buf), f"step {step}: AMP window differs"
        assert torch.equal(a.ref_body_pos, b.ref_body_pos) and torch.equal(a.ref_body_rot, b.ref_body_rot)
        assert torch.equal(a.ref_body_vel, b.ref_body_vel)
        if step % 3 == 2:            # the agent's env_reset(done_indices) (here: done envs plus every 5th env)
            mask = ((a.reset_buf != 0) | (torch.arange(n, device=a.device) % 5 == step % 5)).long()
            for t in tasks:
                torch.manual_seed(100 + step)
                t.reset(mask.clone())
            assert torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a._motion_start_times, b._motion_start_times)
    assert int((tasks[0].reset_buf != 0).sum()) >= 0


def test_shipped_config_takes_the_specialised_step_kernel():
    """HumanoidIm's default (shipped im.yaml / env_im.yaml) steady-state launch must qualify for the compile-time specialised
    kernel (phc_env_step_fast_launches), and that kernel must agree with the generic instantiation on the same inputs."""
    from phc_b200 import _lib
    from phc_b200.env.humanoid_im import HumanoidIm
    lib = _lib.load()
    n = 300
    m = syn.make_motions(n, seed=31, min_frames=40, max_frames=80)
    env = HumanoidIm({"env": {"num_envs": n}, "motion_data": m, "seed": 5}, device_type="cuda", device_id=0)
    env.reset()
    before = lib.phc_env_step_fast_launches()
    cache_in = env._ref_cache.clone()               # the launch consumes the cached pose and replaces it
    env.step(None)
    torch.cuda.synchronize()
    assert lib.phc_env_step_fast_launches() == before + 1
    fast = {k: getattr(env, k).clone() for k in ("obs_buf", "rew_buf", "reward_raw", "reset_buf", "_terminate_buf")}
    fast["amp"] = env._amp_obs_buf.clone()
    fast["cache"] = env._ref_cache.clone()
    # the same step through the generic instantiation: an all-ones env mask disqualifies the launch from the fast variant
    p = env._plan
    ones = torch.ones(n, dtype=torch.int64, device=env.device)
    env._ref_cache.copy_(cache_in)
    p.args.only_where = ones.data_ptr()
    p.run()
    torch.cuda.synchronize()
    p.args.only_where = None
    assert lib.phc_env_step_fast_launches() == before + 1
    for k in ("obs_buf", "rew_buf", "reward_raw"):
        close(getattr(env, k).cpu(), fast[k].cpu(), atol=2e-6, what=f"fast vs generic {k}")
    assert torch.equal(env.reset_buf, fast["reset_buf"]) and torch.equal(env._terminate_buf, fast["_terminate_buf"])
    close(env._amp_obs_buf.cpu(), fast["amp"].cpu(), what="fast vs generic amp")
    close(env._ref_cache.cpu(), fast["cache"].cpu(), what="fast vs generic pose cache")


def test_arrival_ordered_step_kernel_matches_the_fast_instantiation(tmp_path):
    """env_step_fast.cu (what phc_env_step launches for the shipped steady state) against env_step_kernel<1, 24, false, true> on the
    device: the same seeded three steps in two fresh processes (PHC_ENV_FASTK is read once per process).  The two kernels are the same
    source expressions (bit-identical in the -ffp-contract=off CPU emulation, tests/test_env_step_emu_cpu.py); on the device nvcc
    contracts mul + add pairs per kernel, so floats agree to rounding (the fast-vs-generic tolerance), integers exactly."""
    import os
    import subprocess
    import sys
    script = (
        "import sys, torch\n"
        "from phc_b200 import synthetic as syn\n"
        "from phc_b200.env.humanoid_im import HumanoidIm\n"
        "n = 1000\n"
        "m = syn.make_motions(n, seed=23, min_frames=30, max_frames=70)\n"
        "task = HumanoidIm({'env': {'num_envs': n}, 'motion_data': m, 'seed': 23})\n"
        "torch.manual_seed(5); task.reset(); out = []\n"
        "for step in range(3):\n"
        "    task.step(None); torch.cuda.synchronize()\n"
        "    out.append({k: getattr(task, k).cpu().clone() for k in ('obs_buf', 'rew_buf', 'reward_raw', 'reset_buf', '_terminate_buf', '_amp_obs_buf', '_ref_cache')})\n"
        "torch.save(out, sys.argv[1])\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("1", "0"):
        f = str(tmp_path / f"fastk{flag}.pt")
        r = subprocess.run([sys.executable, "-c", script, f], env=dict(os.environ, PHC_ENV_FASTK=flag, PYTHONPATH=root), cwd=root,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(f))
    for step, (a, b) in enumerate(zip(*outs)):
        for k in a:
            if a[k].dtype.is_floating_point:
                close(a[k], b[k], atol=2e-6, what=f"step {step}: {k}, env_step_fast_kernel vs the FAST instantiation")
            else:
                assert torch.equal(a[k], b[k]), f"step {step}: {k} differs between env_step_fast_kernel and the FAST instantiation"


def test_im_eval_extras_match_oracle():
    """flags.im_eval: extras['mpjpe'] / 'body_pos_gt' (humanoid_im.py:674-680) and the mean-distance termination, against the oracle."""
    n = 130
    m = syn.make_motions(n, seed=17, min_frames=40, max_frames=90)
    task = HumanoidIm({"env": {"num_envs": n}, "motion_data": m, "seed": 2, "im_eval": True})
    task.reset()
    for _ in range(3):
        task.step(None)
    torch.cuda.synchronize()
    st_in = dict(body=task._rigid_body_state_reshaped[:, :24].cpu().clone(), dof=task._dof_state.cpu().clone(), force=task.dof_force_tensor.cpu().clone(),
                 prog=task.progress_buf.cpu().clone() + 1, ids=task._sampled_motion_ids.cpu(), st=task._motion_start_times.cpu().clone(),
                 off=task._motion_start_times_offset.cpu().clone(), goff=task._global_offset.cpu().clone(), hist=task._amp_obs_buf.cpu().clone())
    task.sim.simulate(None)
    st_in["body"] = task._rigid_body_state_reshaped[:, :24].cpu().clone()
    st_in["dof"], st_in["force"] = task._dof_state.cpu().clone(), task.dof_force_tensor.cpu().clone()
    task.post_physics_step()
    torch.cuda.synchronize()
    exp = O.env_step(oracle_tables(m), smpl_step_config(use_mean=True), st_in["body"], st_in["dof"], st_in["force"], st_in["prog"], st_in["ids"],
                     st_in["st"], st_in["off"], st_in["goff"], st_in["hist"])
    close(task.extras["mpjpe"].cpu(), exp["mpjpe"], what="mpjpe")
    close(task.extras["body_pos_gt"].cpu(), exp["body_pos_gt"], what="body_pos_gt")
    close(task.rew_buf.cpu(), exp["rew"], what="rew (im_eval)")
    assert torch.equal(task.reset_buf.cpu(), exp["reset"]) and torch.equal(task._terminate_buf.cpu(), exp["terminate"])
    assert torch.equal(task.extras["body_pos"], task._rigid_body_pos)
