"""Pin oracle/phc_oracle.py against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.  Tolerance: rtol 1e-5 / atol 1e-6 fp32 (north_star), except where a
comment says otherwise (ill-conditioned acos near identity -- SURVEY.md section 7)."""
import os

import numpy as np
import torch

from oracle import phc_oracle as O
import pytest

from phc_b200 import synthetic as syn
from tests.helpers import close, env_state_from, load, smpl_step_config, tables_from


def test_quaternion_library():
    g = load("quat.npz")
    a, b, v, t, e = g["a"], g["b"], g["v"], g["t"], g["e"]
    close(O.qmul(a, b), g["quat_mul"], what="quat_mul")
    close(O.qconj(a), g["quat_conjugate"], what="quat_conjugate")
    close(O.qrot(a, v), g["my_quat_rotate"], what="my_quat_rotate")
    close(O.tan_norm(a), g["quat_to_tan_norm"], what="tan_norm")
    ang, axis = O.q_to_angle_axis(a)
    close(ang, g["angle"], what="angle")
    close(axis, g["axis"], what="axis")
    ang_s, axis_s = O.q_to_angle_axis(g["small"])
    close(ang_s, g["angle_small"], what="angle_small")
    close(axis_s, g["axis_small"], rtol=1e-4, what="axis_small")
    close(O.q_to_exp_map(a), g["quat_to_exp_map"], what="exp_map")
    close(O.exp_map_to_q(e), g["exp_map_to_quat"], what="exp_map_to_quat")
    close(O.slerp(a, b, t), g["slerp"], what="slerp")
    close(O.heading_angle(a), g["calc_heading"], what="heading")
    close(O.heading_q(a, False), g["calc_heading_quat"], what="heading_q")
    close(O.heading_q(a, True), g["calc_heading_quat_inv"], what="heading_q_inv")
    close(O.strip_base_rot(a), g["remove_base_rot"], what="remove_base_rot")


def test_motion_state():
    g = load("motion.npz")
    tab = tables_from(g)
    ids, times = g["ids"], g["times"]
    i0, i1, bl = O.frame_blend(times, tab.lengths[ids], tab.num_frames[ids], tab.dts[ids])
    close(i0, g["idx0"], what="idx0")
    close(i1, g["idx1"], what="idx1")
    close(bl, g["blend"], what="blend")
    res = O.motion_state(tab, ids, times, g["offset"])
    for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_vel", "rg_pos", "rb_rot", "body_vel", "body_ang_vel"):
        close(res[k], g["out_" + k], what=k)
    # dof_pos = 2*acos(w)*axis of a slerped quaternion: near-identity joints are ill-conditioned
    close(res["dof_pos"], g["out_dof_pos"], rtol=1e-4, atol=1e-5, what="dof_pos")
    close(O.motion_state(tab, ids, times)["rg_pos"], g["out_noffset_rg_pos"], what="rg_pos (no offset)")


def _check_step(g, tag, in_tag, cfg):
    tab = tables_from(g)
    st = env_state_from(g, in_tag)
    out = O.env_step(tab, cfg, st.body_state, st.dof_state, st.dof_force, st.progress, st.motion_ids,
                     st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    close(out["obs"], g[f"{tag}_out_obs"], what=f"{tag} obs")
    close(out["obs"][:, :g[f"{tag}_out_self_obs"].shape[1]], g[f"{tag}_out_self_obs"], what=f"{tag} self_obs")
    close(out["rew"], g[f"{tag}_out_rew"], what=f"{tag} rew")
    close(out["reward_raw"], g[f"{tag}_out_reward_raw"], what=f"{tag} reward_raw")
    close(out["reset"], g[f"{tag}_out_reset"], what=f"{tag} reset")
    close(out["terminate"], g[f"{tag}_out_terminate"], what=f"{tag} terminate")
    close(out["amp_obs_buf"], g[f"{tag}_out_amp_obs_buf"], what=f"{tag} amp_obs_buf")
    close(out["ref_body_pos"], g[f"{tag}_out_ref_body_pos"], what=f"{tag} ref_body_pos")
    close(out["ref_body_rot"], g[f"{tag}_out_ref_body_rot"], what=f"{tag} ref_body_rot")
    close(out["ref_body_vel"], g[f"{tag}_out_ref_body_vel"], what=f"{tag} ref_body_vel")
    close(out["ref_dof_pos"], g[f"{tag}_out_ref_dof_pos"], rtol=1e-4, atol=1e-5, what=f"{tag} ref_dof_pos")
    return out


def test_env_step_shipped_config():
    g = load("envstep.npz")
    out = _check_step(g, "A", "A", smpl_step_config())
    assert out["terminate"].sum() > 0 and out["reset"].sum() >= out["terminate"].sum()   # the case exercises early termination
    assert (out["reward_raw"][:, 4] == 0).any() and (out["reward_raw"][:, 4] != 0).any()  # power reward mask both ways


def test_env_step_generic_blend_and_offset():
    _check_step(load("envstep.npz"), "B", "B", smpl_step_config())


def test_env_step_not_upright_global_root():
    _check_step(load("envstep.npz"), "C", "A", smpl_step_config(upright=False, local_root_obs=False))


def test_env_step_im_eval_mean_termination():
    _check_step(load("envstep.npz"), "D", "B", smpl_step_config(use_mean=True))


def test_amp_obs_demo():
    g = load("envstep.npz")
    out = O.amp_obs_demo(tables_from(g), smpl_step_config(), g["demo_ids"], g["demo_t0"])
    # joint exp-maps of the reference pose go through acos near identity -> looser on those columns
    close(out, g["demo_out"], rtol=1e-4, atol=2e-5, what="amp_obs_demo")


def test_learning_pieces():
    g = load("learn.npz")
    adv = O.gae(g["gae_fdones"], g["gae_values"], g["gae_rewards"], g["gae_next_values"], 0.99, 0.95)
    close(adv, g["gae_adv"], what="gae")
    T, N = adv.shape[0], adv.shape[1]
    flat = lambda x: x.transpose(0, 1).reshape(T * N, -1)
    close(O.normalize_advantages(flat(adv + g["gae_values"]), flat(g["gae_values"])), g["adv_norm"], what="adv_norm")
    close(O.actor_loss(g["al_old"], g["al_new"], g["al_adv"], 0.2), g["al_out"], what="actor_loss")
    close(O.critic_loss(g["cl_v"], g["cl_r"]), g["cl_out"], what="critic_loss")
    close(O.bound_loss(g["bl_mu"]), g["bl_out"], what="bound_loss")

    ws = [g["d_w1"], g["d_w2"], g["d_w3"]]
    bs = [g["d_b1"], g["d_b2"], g["d_b3"]]
    params = [p.clone().requires_grad_(True) for p in (ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    W, B = params[0::2], params[1::2]
    x_demo = g["d_x_demo"].clone().requires_grad_(True)
    la = O.mlp_forward(g["d_x_agent"], W, B)
    ld = O.mlp_forward(x_demo, W, B)
    info = O.disc_loss(la, ld, x_demo, W[2], W)
    close(info["disc_loss"], g["d_loss"], what="disc_loss")
    close(info["disc_grad_penalty"], g["d_gp"], what="disc_gp")
    close(info["disc_logit_loss"], g["d_logit_loss"], what="disc_logit_loss")
    close(info["disc_agent_acc"], g["d_agent_acc"], what="acc_a")
    close(info["disc_demo_acc"], g["d_demo_acc"], what="acc_d")
    grads = torch.autograd.grad(info["disc_loss"], params)
    for i, gr in enumerate(grads):
        close(gr, g[f"d_grad{i}"], rtol=1e-4, atol=1e-6, what=f"disc grad {i}")
    with torch.no_grad():
        dr = O.disc_reward(O.mlp_forward(g["d_x_agent"], ws, bs))
    close(dr, g["d_reward"], what="disc_reward")
    close(0.5 * torch.ones(48, 1) * 0.7 + 0.5 * dr, g["d_combined"], what="combined")

    mean, var, cnt = torch.zeros(12, dtype=torch.float64), torch.ones(12, dtype=torch.float64), torch.ones((), dtype=torch.float64)
    for i in range(3):
        x = g[f"rms_x{i}"]
        close(O.rms_normalize(x, mean, var), g[f"rms_y{i}"], what=f"rms_y{i}")
        mean, var, cnt = O.rms_update(mean, var, cnt, x)
    close(mean, g["rms_mean"], rtol=1e-12, atol=1e-12, what="rms_mean")
    close(var, g["rms_var"], rtol=1e-12, atol=1e-12, what="rms_var")
    close(cnt, g["rms_count"], what="rms_count")
    close(O.rms_unnormalize(g["rms_x0"] * 0.1, mean, var), g["rms_unnorm"], what="rms_unnorm")


def test_gaussian_pieces_self_pinned():
    """rl_games==1.1.4 pieces (absent from /root/reference): pinned against torch.distributions."""
    gen = torch.Generator().manual_seed(0)
    mu, a = torch.randn(16, 69, generator=gen), torch.randn(16, 69, generator=gen)
    logstd = torch.full((69,), -2.9).expand(16, 69)
    sigma = torch.exp(logstd)
    ref = -torch.distributions.Normal(mu, sigma).log_prob(a).sum(-1)
    close(O.gaussian_neglogp(a, mu, sigma, logstd), ref, rtol=1e-5, atol=1e-3, what="neglogp")
    # KL(new || old); sigma ~ 0.5 so rl_games' +1e-5 regularisers are negligible against the closed form
    s_new, s_old = torch.full((16, 69), 0.5), torch.full((16, 69), 0.55)
    mu_old = mu + 0.05
    kl = torch.distributions.kl_divergence(torch.distributions.Normal(mu, s_new), torch.distributions.Normal(mu_old, s_old)).sum(-1).mean()
    close(O.policy_kl(mu, s_new, mu_old, s_old), kl, rtol=1e-3, atol=1e-3, what="policy_kl")


# ------------------------------------------------------------------------------------------------------------------
# PNN / MCP (row a19): oracle/mcp_oracle.py against the real PNN / load_pnn / load_mcp_mlp / HumanoidImMCP.step
# ------------------------------------------------------------------------------------------------------------------
def _mcp():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mcp.npz"))
    g = {k: z[k] for k in z.files}
    sd = {k[len("model/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("model/")}
    return g, sd


def test_pnn_forward_matches_reference():
    from oracle import mcp_oracle as mo
    g, sd = _mcp()
    K = int(g["num_prim"])
    x = torch.from_numpy(g["x"])
    for k in range(K):
        torch.testing.assert_close(mo.pnn_forward(sd, x, K, idx=k), torch.from_numpy(g[f"col{k}"])[0], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(torch.stack(mo.pnn_forward(sd, x, K), 0), torch.from_numpy(g["all"]), rtol=1e-6, atol=1e-6)


def test_pnn_load_actor_and_freeze_match_reference():
    from oracle import mcp_oracle as mo
    g, sd = _mcp()
    single = {k[len("single/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("single/")}
    for k, v in mo.pnn_load_actor(single, n_hidden=len(g["units"]), idx=1).items():
        assert torch.equal(sd[k], v), k
    names = [str(n) for n in g["param_names"]]
    assert mo.pnn_trainable(names, 1) == [bool(v) for v in g["trainable_after_freeze1"]]


def test_mcp_step_actions_match_reference():
    from oracle import mcp_oracle as mo
    g, sd = _mcp()
    K = int(g["num_prim"])
    args = (torch.from_numpy(g["obs_buf"]), torch.from_numpy(g["rms_mean"]), torch.from_numpy(g["rms_var"]), sd, torch.from_numpy(g["weights"]), K)
    assert (np.abs(g["obs_buf"]) > 5).any()
    torch.testing.assert_close(mo.mcp_step_actions(*args), torch.from_numpy(g["actions"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(mo.mcp_step_actions(*args, discrete=True), torch.from_numpy(g["actions_discrete"]), rtol=1e-6, atol=1e-6)


def test_mcp_composer_keeps_final_relu():
    from oracle import mcp_oracle as mo
    g, _ = _mcp()
    comp = {k[len("composer/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("composer/")}
    out = mo.mlp_forward(comp, "a2c_network.composer.", torch.from_numpy(g["x"]), ending_act=True)
    torch.testing.assert_close(out, torch.from_numpy(g["composer_out"]), rtol=1e-6, atol=1e-6)
    assert (g["composer_out"] == 0).any() and (g["composer_out"] >= 0).all()
    out = mo.mlp_forward(comp, "a2c_network.composer.", torch.from_numpy(g["x"]), ending_act=True, act="silu")
    torch.testing.assert_close(out, torch.from_numpy(g["composer_out_silu"]), rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------------------------
# Hinge-joint robot (H1, config 5 of BASELINE.json): oracle against the real MotionLibReal / HumanoidIm h1 branches
# ------------------------------------------------------------------------------------------------------------------
def test_h1_motion_state_matches_motion_lib_real():
    from tests.helpers import robot_tables_from
    g = load("h1.npz")
    out = O.motion_state_robot(robot_tables_from(g), g["ms_ids"], g["ms_times"], g["ms_offset"])
    for k in ("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "rg_pos", "rb_rot", "body_vel", "body_ang_vel",
              "rg_pos_t", "rg_rot_t", "body_vel_t", "body_ang_vel_t"):
        close(out[k], g["ms_out_" + k], what="h1 motion_state " + k)


def test_h1_env_step_matches_reference():
    from tests.helpers import robot_tables_from, h1_step_config
    g = load("h1.npz")
    tab = robot_tables_from(g)
    for tag in ("A", "B"):
        st = env_state_from(g, tag)
        out = O.env_step_robot(tab, h1_step_config(), g["ext_parents"].tolist(), g["ext_pos"], st.body_state, st.dof_state, st.dof_force,
                               st.progress, st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
        assert out["obs"].shape[1] == 778 and out["amp_obs"].shape[1] == 63          # H1 row of SURVEY.md section 0
        for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos", "ref_body_rot", "ref_body_vel"):
            close(out[k], g[f"{tag}_out_{k}"], what=f"h1 {tag} {k}")
    assert g["A_out_terminate"].sum() > 0


def test_h1_amp_obs_demo_matches_reference():
    from tests.helpers import robot_tables_from, h1_step_config
    g = load("h1.npz")
    out = O.amp_obs_demo_robot(robot_tables_from(g), h1_step_config(), g["demo_ids"], g["demo_t0"])
    close(out, g["demo_out"], what="h1 amp_obs_demo")


# ----------------------------------------------------------------------------------------------------------------
# motion LOADER: oracle/motion_load_oracle.py against the real MotionLibSMPL.load_motion_with_skeleton
# ----------------------------------------------------------------------------------------------------------------
def test_motion_loader_matches_reference():
    import numpy as np
    from oracle import motion_load_oracle as ML
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "load.npz"))
    out = ML.load_clips(z["pose_quat_global"], z["root_trans"], z["num_frames"], z["fps"], z["parents"], z["offsets"], z["heading"])
    for k in ("gts", "grs", "lrs", "gvs", "gavs"):          # float64 pipeline: identical after the float32 cast
        close(out[k], z[k], rtol=1e-6, atol=1e-7, what=f"load {k}")
    close(out["dvs"], z["dvs"], rtol=1e-5, atol=2e-5, what="load dvs")   # float32 2*acos(w) of frame-to-frame rotations


def test_motion_loader_filter_taps_are_scipys():
    import numpy as np
    from scipy.ndimage import gaussian_filter1d
    from oracle import motion_load_oracle as ML
    x = np.random.default_rng(0).standard_normal((23, 3, 2))
    np.testing.assert_allclose(ML.filter_time(x), gaussian_filter1d(x, 2, axis=0, mode="nearest"), rtol=1e-13, atol=1e-14)


# ----------------------------------------------------------------------------------------------------------------
# env_im_getup_mcp.yaml: zero_out_far + cycle_motion (the configuration HumanoidImMCP trains in)
# ----------------------------------------------------------------------------------------------------------------
def getup_oracle_step(g):
    st = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    return O.env_step_getup(tables_from(g), smpl_step_config(), st["body_state"], st["dof_state"], st["dof_force"], st["progress"],
                            st["motion_ids"], st["start_times"], st["start_offsets"], st["global_offset"], st["amp_hist"],
                            st["point_goal"], st["cycle_counter"], st["cycle_phase"], max_episode_length=15)


def test_env_step_getup_zero_out_far_cycle_motion():
    g = load("getup.npz")
    out = getup_oracle_step(g)
    assert int(g["in_wrap"].sum()) >= 8 and 8 <= int((g["out_reward_raw"][:, 1] == 0).sum()) <= 40      # both branches exercised
    for k in ("start_times", "start_offsets", "global_offset", "point_goal", "rew", "reward_raw", "ref_body_pos", "ref_body_rot",
              "ref_body_vel", "amp_obs_buf"):
        close(out[k], g["out_" + k], what=f"getup {k}")
    close(out["obs"], g["out_obs"], atol=2e-6, what="getup obs")
    for k in ("reset", "terminate"):
        close(out[k], g["out_" + k], what=f"getup {k}")
    close(out["cycle_counter"].long(), g["out_cycle_counter"].long(), what="getup cycle_counter")


# ----------------------------------------------------------------------------------------------------------------
# Shapes with more than 32 bodies (Unitree G1: 38 + 1 extend body; SMPL-X: 52): the oracle is pinned ahead of the fused step
# kernel, which returns PHC_ERR_UNSUPPORTED for them today (one body per lane)
# ----------------------------------------------------------------------------------------------------------------
def _g1_tables(g):
    f = lambda k: g["tab_" + k]
    return O.RobotTables(f("gts_t"), f("grs_t"), f("gvs_t"), f("gavs_t"), f("dof_pos"), f("dvs"), f("lengths"), f("num_frames"), f("dts"),
                         f("length_starts"), 38)


def test_g1_shapes_match_reference():
    from phc_b200 import synthetic as syn
    g = load("g1.npz")
    tab = _g1_tables(g)
    out = O.motion_state_robot(tab, g["ms_ids"], g["ms_times"], g["ms_offset"])
    for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel", "dof_pos", "dof_vel", "rg_pos_t", "rg_rot_t"):
        close(out[k], g["ms_out_" + k], what="g1 motion_state " + k)
    cfg = O.StepConfig(key_bodies=syn.G1_KEY_BODIES, reset_bodies=None, dof_subset=None)
    for tag in ("A", "B"):
        st = env_state_from(g, tag)
        o = O.env_step_robot(tab, cfg, g["ext_parents"].tolist(), g["ext_pos"], st.body_state, st.dof_state, st.dof_force, st.progress,
                             st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
        assert o["obs"].shape[1] == 1 + 15 * 38 - 3 + 24 * 38 and o["amp_obs"].shape[1] == 13 + 2 * 37 + 12
        for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos", "ref_body_rot", "ref_body_vel"):
            close(o[k], g[f"{tag}_out_{k}"], atol=2e-6, what=f"g1 {tag} {k}")
    close(O.amp_obs_demo_robot(tab, cfg, g["demo_ids"], g["demo_t0"]), g["demo_out"], what="g1 amp_obs_demo")


def test_smplx_shapes_match_reference():
    from phc_b200 import synthetic as syn
    g = load("smplx.npz")
    st = syn.EnvState(**{k: g[f"in_{k}"] for k in syn.EnvState.__dataclass_fields__})
    cfg = O.StepConfig(key_bodies=syn.SMPLX_KEY_BODIES, reset_bodies=None, dof_subset=None)
    o = O.env_step(tables_from(g), cfg, st.body_state, st.dof_state, st.dof_force, st.progress, st.motion_ids, st.start_times,
                   st.start_offsets, st.global_offset, st.amp_hist)
    assert o["obs"].shape[1] == 1 + 15 * 52 - 3 + 24 * 52
    for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf", "ref_body_pos", "ref_body_rot", "ref_body_vel"):
        close(o[k], g[f"out_{k}"], atol=2e-6, what=f"smplx {k}")


def test_env_step_future_tracks_matches_reference():
    """fut_tracks with 3 future samples 0.1 s apart (tests/golden/fut.npz from the real _compute_task_obs / v6)."""
    from phc_b200 import synthetic as syn
    g = load("fut.npz")
    st = syn.EnvState(**{k: g[f"in_{k}"] for k in syn.EnvState.__dataclass_fields__})
    o = O.env_step(tables_from(g), smpl_step_config(time_steps=3, traj_dt=1 / 10), st.body_state, st.dof_state, st.dof_force, st.progress,
                   st.motion_ids, st.start_times, st.start_offsets, st.global_offset, st.amp_hist)
    assert o["obs"].shape[1] == 358 + 3 * 576
    for k in ("obs", "rew", "reward_raw", "reset", "terminate", "amp_obs_buf"):
        close(o[k], g[f"out_{k}"], atol=2e-6, what=f"fut {k}")
    # Reference quirk, not mirrored: under fut_tracks `self.ref_body_pos[env_ids] = ref_rb_pos[..., 0, :, :]` (humanoid_im.py:857-861)
    # indexes the FLAT [B*T, J, 3] tensor, so every env receives env 0's first sample.  Oracle and kernel keep each env's own
    # first sample (what the non-fut branch stores); env 0 is where the two agree.
    for k in ("ref_body_pos", "ref_body_rot", "ref_body_vel"):
        ref = g[f"out_{k}"]
        assert float((ref - ref[0:1]).abs().max()) == 0.0
        close(o[k][0], ref[0], atol=2e-6, what=f"fut {k} (env 0)")


def test_reset_path_pieces_match_reference():
    """_init_amp_obs_ref (history slots of freshly reset envs) and sample_time_interval, tests/golden/reset.npz."""
    g = load("reset.npz")
    hist = O.amp_obs_demo(tables_from(g), smpl_step_config(), g["motion_ids"], g["t0"], first_step=1, num_steps=9)
    close(hist, g["hist_after"][g["env_ids"]], rtol=1e-4, atol=2e-5, what="_init_amp_obs_ref")
    untouched = torch.ones(g["hist_after"].shape[0], dtype=torch.bool)
    untouched[g["env_ids"]] = False
    assert torch.equal(g["hist_after"][untouched], g["hist_before"][untouched])
    ln = g["tab_lengths"][g["motion_ids"]]
    t = ((g["phase"] * ln) / (1 / 30)).long() * (1 / 30)          # the arithmetic phc_reset_bookkeeping implements
    assert torch.equal(t, g["sampled_times"])


def test_env_step_getup_smplx_shapes():
    """The getup configuration at SMPL-X shapes (env_im_x_getup_mcp.yaml), tests/golden/getup_smplx.npz."""
    from phc_b200 import synthetic as syn
    g = load("getup_smplx.npz")
    st = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    cfg = O.StepConfig(key_bodies=syn.SMPLX_KEY_BODIES, reset_bodies=None, dof_subset=None)
    out = O.env_step_getup(tables_from(g), cfg, st["body_state"], st["dof_state"], st["dof_force"], st["progress"], st["motion_ids"],
                           st["start_times"], st["start_offsets"], st["global_offset"], st["amp_hist"], st["point_goal"], st["cycle_counter"],
                           st["cycle_phase"], max_episode_length=15)
    assert int(g["in_wrap"].sum()) >= 5 and int((g["out_reward_raw"][:, 1] != 0).sum()) >= 2
    for k in ("start_times", "start_offsets", "global_offset", "point_goal", "rew", "reward_raw", "ref_body_pos", "ref_body_rot", "ref_body_vel",
              "amp_obs_buf", "reset", "terminate"):
        close(out[k], g["out_" + k], what=f"getup smplx {k}")
    close(out["obs"], g["out_obs"], atol=2e-6, what="getup smplx obs")


def _vr_case(g, tag):
    """vr.npz (make_golden.gen_vr): E = Head + hands subset, subset reward, shape + limb-weight columns; F = the subset with the
    full-body reward; G = every body tracked, occlusion training."""
    track = g["track"].tolist()
    subset = tag in ("E", "F")
    cfg = smpl_step_config(track_bodies=track if subset else None, reset_bodies=track if subset else syn.SMPL_RESET_BODIES,
                           full_body_reward=tag != "E")
    kw = {}
    if tag == "E":
        kw = dict(shape_params=g["E_shape"][:, :-6], limb_weights=g["E_limb"])
    if tag == "G":
        kw = dict(occlusion=g["G_occlusion"])
    return cfg, kw


@pytest.mark.parametrize("tag", ["E", "F", "G"])
def test_env_step_tracked_subset_occlusion_shape_columns(tag):
    g = load("vr.npz")
    cfg, kw = _vr_case(g, tag)
    st = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    out = O.env_step(tables_from(g), cfg, st["body_state"], st["dof_state"], st["dof_force"], st["progress"], st["motion_ids"], st["start_times"],
                     st["start_offsets"], st["global_offset"], st["amp_hist"], **kw)
    close(out["obs"], g[f"{tag}_out_obs"], atol=2e-6, what=f"{tag} obs")
    close(out["rew"], g[f"{tag}_out_rew"], what=f"{tag} rew")
    close(out["reward_raw"], g[f"{tag}_out_reward_raw"], what=f"{tag} reward_raw")
    assert torch.equal(out["reset"], g[f"{tag}_out_reset"]) and torch.equal(out["terminate"], g[f"{tag}_out_terminate"])
    close(out["ref_body_pos"], g[f"{tag}_out_ref_body_pos"], what=f"{tag} ref_body_pos")
    if tag == "G":
        assert g["G_occlusion"].any()
