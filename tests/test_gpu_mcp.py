"""GPU parity of the PNN / MCP row: frozen-primitive evaluation, action mixing, the PNN training
column and the ReLU-ended composer, against the golden vectors of the real reference (tests/golden/mcp.npz) and the oracle.

Tolerance: the MLPs run as 3xTF32 tensor-core GEMMs with fp32 accumulation; per layer the error is <= ~4e-6*|x||W| (see
tests/test_gpu_learner.py), so outputs of the 3-layer stacks are compared at 1e-5 relative to the output scale.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mcp_oracle as mo
from phc_b200 import _lib, synthetic as syn
from phc_b200.learning.networks import AMPNetwork, MLPEngine, round4

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def golden():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mcp.npz"))
    g = {k: z[k] for k in z.files}
    sd = {k[len("model/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("model/")}
    return g, sd


def close_scale(got, ref, tol=1e-5, what=""):
    ref = ref.to(torch.float64)
    err = (got.detach().cpu().to(torch.float64) - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-3)
    assert err <= tol * scale, f"{what}: max err {err:.3e} > {tol:g} * {scale:.3e}"


def checkpoint(g, sd):
    return {"model": sd, "running_mean_std": {"running_mean": torch.from_numpy(g["rms_mean"]), "running_var": torch.from_numpy(g["rms_var"])}}


@pytest.mark.parametrize("backend", ["tc5", "mma"])
def test_frozen_pnn_all_columns_match_reference(backend):
    from phc_b200.learning.network_loader import load_pnn
    g, sd = golden()
    K = int(g["num_prim"])
    pnn = load_pnn(checkpoint(g, sd), num_prim=K, device=DEV, backend=backend)
    x_ref = torch.from_numpy(g["x"])
    x = pnn.input_buffer(x_ref.shape[0])
    x[:, :x_ref.shape[1]] = x_ref.to(DEV)
    out = pnn.forward_all(x)
    torch.cuda.synchronize()
    close_scale(out[:, :, :pnn.action_dim], torch.from_numpy(g["all"]), what="PNN.forward(idx=-1)")


@pytest.mark.parametrize("discrete", [False, True])
def test_mcp_compose_actions_match_reference(discrete):
    """HumanoidImMCP.compose_actions on an instance assembled by hand (as the golden generator does for the reference)."""
    from phc_b200.env.humanoid_im_mcp import HumanoidImMCP
    from phc_b200.learning.amp_agent import RunningMeanStd
    from phc_b200.learning.network_loader import load_pnn
    g, sd = golden()
    K, N, D, A = int(g["num_prim"]), g["obs_buf"].shape[0], int(g["obs_dim"]), int(g["act_dim"])
    env = object.__new__(HumanoidImMCP)
    env.num_envs, env.num_prim, env.num_dof, env.discrete_mcp, env.device = N, K, A, discrete, torch.device(DEV)
    env.pnn = load_pnn(checkpoint(g, sd), num_prim=K, device=DEV)
    env._pnn_rms = RunningMeanStd(D, DEV, epsilon=1e-5)
    env._pnn_rms.running_mean.copy_(torch.from_numpy(g["rms_mean"]))
    env._pnn_rms.running_var.copy_(torch.from_numpy(g["rms_var"]))
    env._pnn_rms.freeze()
    env.obs_buf = torch.from_numpy(g["obs_buf"]).to(DEV)
    env._mixed = torch.zeros(N, A, device=DEV)
    env._lib = _lib.load()
    act = env.compose_actions(torch.from_numpy(g["weights"]).to(DEV))
    torch.cuda.synchronize()
    close_scale(act, torch.from_numpy(g["actions_discrete" if discrete else "actions"]), what="mixed action")


def test_mcp_combine_kernel_exact():
    """phc_mcp_combine alone: same fp32 operation order as sum(weights[:, :, None] * x_all, dim=1) -> bit exact."""
    lib = _lib.load()
    gen = torch.Generator().manual_seed(3)
    for (N, K, A) in [(1, 1, 1), (7, 3, 69), (1000, 8, 69), (0, 3, 5)]:
        w = torch.randn(N, K, generator=gen)
        prim = torch.randn(K, N, round4(A) + 4, generator=gen)
        out = torch.zeros(N, A, device=DEV)
        wd, pd = w.to(DEV), prim.to(DEV)
        _lib.check(lib.phc_mcp_combine(wd.data_ptr(), max(K, 1), pd.data_ptr(), pd.stride(1), pd.stride(0), N, K, A, 0, out.data_ptr(), A, None))
        ref = torch.zeros(N, A)
        for k in range(K):                                   # torch.sum over dim=1 adds k = 0, 1, ... in order for small K
            ref = ref + w[:, k:k + 1] * prim[k, :, :A]
        assert torch.equal(out.cpu(), ref), (N, K, A)
        if N:
            _lib.check(lib.phc_mcp_combine(wd.data_ptr(), K, pd.data_ptr(), pd.stride(1), pd.stride(0), N, K, A, 1, out.data_ptr(), A, None))
            best = torch.argmax(w, dim=1)
            assert torch.equal(out.cpu(), prim[best, torch.arange(N), :A] + 0.0)
    assert lib.phc_mcp_combine(None, 1, None, 1, 1, 1, 1, 1, 0, None, 1, None) != 0


def _load_columns(net, sd):
    for l in net.all_layers():
        k = f"a2c_network.{l.name}.weight"
        if k in sd:
            net.set_layer(l, sd[k], sd[k[:-6] + "bias"])
    net.refresh_split()


@pytest.mark.parametrize("backend", ["tc5", "mma"])
def test_pnn_training_column_forward_backward(backend):
    """amp_pnn network: forward = column `training_prim` only (amp_network_pnn_builder.py:65); the backward fills that
    column's gradients and leaves every other column's at zero (frozen, pnn.py:45-51)."""
    g, sd = golden()
    K, D, A, units = int(g["num_prim"]), int(g["obs_dim"]), int(g["act_dim"]), [int(u) for u in g["units"]]
    tp = 1
    net = AMPNetwork(D, A, 8, units=units, disc_units=(8,), device=DEV, kind="amp_pnn", num_prim=K, training_prim=tp)
    _load_columns(net, sd)
    assert set(k for k in net.state_dict() if ".pnn." in k) == set(k for k in sd if ".pnn." in k)
    eng = MLPEngine(net, backend)
    x_ref = torch.from_numpy(g["x"])
    B = x_ref.shape[0]
    x = torch.zeros(B, round4(D), device=DEV)
    x[:, :D] = x_ref.to(DEV)
    ws = eng.workspace("a", net.actor, B)
    mu = eng.forward(net.actor, x, ws)
    close_scale(mu[:, :A], torch.from_numpy(g[f"col{tp}"])[0], what="training column forward")
    dout = torch.randn(B, A, generator=torch.Generator().manual_seed(5))
    ws["dout"].zero_()
    ws["dout"][:, :A] = dout.to(DEV)
    net.grads.zero_()
    eng.backward(net.actor, x, ws)
    torch.cuda.synchronize()
    # autograd on the oracle (fp64)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items() if ".pnn." in k}
    (mo.pnn_forward(sd64, x_ref.double(), K, idx=tp) * dout.double()).sum().backward()
    for k, col in enumerate(net.pnn_actors):
        for l in col.layers:
            gw, gb = net.weight(l, True)[:, :l.in_dim], net.bias(l, True)
            if k != tp:
                assert float(gw.abs().sum()) == 0.0 and float(gb.abs().sum()) == 0.0, f"column {k} must stay frozen"
            else:
                close_scale(gw, sd64[f"a2c_network.{l.name}.weight"].grad, tol=2e-5, what=f"{l.name}.weight grad")
                close_scale(gb, sd64[f"a2c_network.{l.name}.bias"].grad, tol=2e-5, what=f"{l.name}.bias grad")


@pytest.mark.parametrize("activation", ["relu", "silu"])
@pytest.mark.parametrize("backend", ["tc5", "mma"])
def test_mcp_composer_forward_backward(backend, activation):
    """amp_mcp network: composer keeps the activation after its last Linear (ending_act) in forward and backward
    (ReLU: im_mcp.yaml, SiLU: im_mcp_big.yaml)."""
    g, _ = golden()
    comp = {k[len("composer/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("composer/")}
    K, D, units = int(g["num_prim"]), int(g["obs_dim"]), [int(u) for u in g["units"]]
    net = AMPNetwork(D, K, 8, units=units, disc_units=(8,), activation=activation, device=DEV, kind="amp_mcp", num_prim=K)
    _load_columns(net, comp)
    assert set(k for k in net.state_dict() if "composer" in k) == set(comp)
    eng = MLPEngine(net, backend)
    x_ref = torch.from_numpy(g["x"])
    B = x_ref.shape[0]
    x = torch.zeros(B, round4(D), device=DEV)
    x[:, :D] = x_ref.to(DEV)
    ws = eng.workspace("a", net.actor, B)
    out = eng.forward(net.actor, x, ws)
    ref = torch.from_numpy(g["composer_out" if activation == "relu" else "composer_out_silu"])
    close_scale(out[:, :K], ref, what="composer forward")
    if activation == "relu":
        live = ref.abs() > 1e-4                                        # away from the ReLU kink the zero pattern must agree
        assert torch.equal((out[:, :K].cpu() > 0)[live], (ref > 0)[live])
    dout = torch.randn(B, K, generator=torch.Generator().manual_seed(6))
    ws["dout"].zero_()
    ws["dout"][:, :K] = dout.to(DEV)
    net.grads.zero_()
    eng.backward(net.actor, x, ws)
    torch.cuda.synchronize()
    c64 = {k: v.double().requires_grad_(True) for k, v in comp.items()}
    (mo.mlp_forward(c64, "a2c_network.composer.", x_ref.double(), ending_act=True, act=activation) * dout.double()).sum().backward()
    for l in net.actor.layers:
        close_scale(net.weight(l, True)[:, :l.in_dim], c64[f"a2c_network.{l.name}.weight"].grad, tol=2e-5, what=f"{l.name}.weight grad")
        close_scale(net.bias(l, True), c64[f"a2c_network.{l.name}.bias"].grad, tol=2e-5, what=f"{l.name}.bias grad")


def _pnn_checkpoint(obs_dim, act_dim, K, units=(64, 32), seed=0):
    gen = torch.Generator().manual_seed(seed)
    sd, d = {}, obs_dim
    for k in range(K):
        d = obs_dim
        for i, u in enumerate(list(units) + [act_dim]):
            sd[f"a2c_network.pnn.actors.{k}.{2 * i}.weight"] = torch.randn(u, d, generator=gen) / d ** 0.5
            sd[f"a2c_network.pnn.actors.{k}.{2 * i}.bias"] = torch.randn(u, generator=gen) * 0.1
            d = u
    sd["a2c_network.mu.bias"] = torch.zeros(act_dim)
    rms = {"running_mean": torch.randn(obs_dim, generator=gen, dtype=torch.float64) * 0.1,
           "running_var": torch.rand(obs_dim, generator=gen, dtype=torch.float64) + 0.5}
    return {"model": sd, "running_mean_std": rms}


def test_humanoid_im_mcp_task_and_agent_epoch():
    """The whole MCP configuration (config 4 of BASELINE.json): HumanoidImMCP env + amp_mcp composer agent.
    The env's mixed action equals the oracle's on the env's own observation; one training epoch moves only the composer,
    critic and discriminator (the primitives live in the env and stay frozen)."""
    from phc_b200.env.humanoid_im import RLGPUEnv
    from phc_b200.env.humanoid_im_mcp import HumanoidImMCP
    from phc_b200.learning.amp_agent import AMPAgent
    n, K = 64, 3
    m = syn.make_motions(n, seed=4, min_frames=40, max_frames=90)
    probe = HumanoidImMCP.__mro__[1]({"env": {"num_envs": 4}, "motion_data": m, "seed": 0})
    ck = _pnn_checkpoint(probe.get_obs_size(), probe.num_dof, K)
    task = HumanoidImMCP({"env": {"num_envs": n, "num_prim": K, "has_pnn": True, "has_lateral": False}, "motion_data": m, "seed": 4},
                         pnn_checkpoint=ck)
    assert task.get_action_size() == K and task.get_task_obs_size_detail()["num_prim"] == K
    task.reset()
    w = torch.relu(torch.randn(n, K, generator=torch.Generator().manual_seed(1))).to(DEV)
    obs_before = task.obs_buf.clone()
    task.step(w)
    torch.cuda.synchronize()
    exp = mo.mcp_step_actions(obs_before.cpu(), ck["running_mean_std"]["running_mean"], ck["running_mean_std"]["running_var"], ck["model"],
                              w.cpu(), K, dtype=torch.float64)
    close_scale(task.actions, exp, what="HumanoidImMCP mixed action")
    frozen = task.pnn.net.params.clone()

    agent = AMPAgent("t", {"vec_env": RLGPUEnv(task), "horizon_length": 8, "minibatch_size": 256, "amp_minibatch_size": 64,
                           "mini_epochs": 2, "amp_obs_demo_buffer_size": 2048, "amp_replay_buffer_size": 2048, "amp_batch_size": 128,
                           "network": {"name": "amp_mcp", "has_softmax": False, "ending_act": True,     # im_mcp.yaml:15-16
                                       "mlp": {"units": [128, 64], "activation": "relu"},
                                       "disc": {"units": [128, 64], "activation": "relu"}}})
    assert agent.model.kind == "amp_mcp" and agent.model.action_dim == K and agent.model.actor.head_relu
    assert "a2c_network.composer.4.weight" in agent.model.state_dict()
    agent.obs = agent.env_reset()
    agent._init_amp_demo_buf()
    p0 = agent.model.params.clone()
    agent.train_epoch()
    torch.cuda.synchronize()
    assert torch.isfinite(agent.model.params).all() and not torch.equal(agent.model.params, p0)
    assert torch.equal(task.pnn.net.params, frozen)


def test_amp_pnn_agent_trains_only_its_column():
    from phc_b200.env.humanoid_im import HumanoidIm, RLGPUEnv
    from phc_b200.learning.amp_agent import AMPAgent
    n, K, tp = 64, 3, 1
    m = syn.make_motions(n, seed=5, min_frames=40, max_frames=90)
    task = HumanoidIm({"env": {"num_envs": n, "num_prim": K, "training_prim": tp}, "motion_data": m, "seed": 5})
    agent = AMPAgent("t", {"vec_env": RLGPUEnv(task), "horizon_length": 8, "minibatch_size": 256, "amp_minibatch_size": 64,
                           "mini_epochs": 2, "amp_obs_demo_buffer_size": 2048, "amp_replay_buffer_size": 2048, "amp_batch_size": 128,
                           "network": {"name": "amp_pnn", "mlp": {"units": [128, 64], "activation": "relu"},
                                       "disc": {"units": [128, 64], "activation": "relu"}}})
    net = agent.model
    assert net.kind == "amp_pnn" and net.num_prim == K and net.training_prim == tp and net.actor is net.pnn_actors[tp]
    agent.obs = agent.env_reset()
    agent._init_amp_demo_buf()
    before = {l.name: net.weight(l).clone() for l in net.all_layers()}
    agent.train_epoch()
    torch.cuda.synchronize()
    for k, col in enumerate(net.pnn_actors):
        for l in col.layers:
            same = torch.equal(net.weight(l), before[l.name])
            assert same == (k != tp), f"column {k} layer {l.name}: {'unchanged' if same else 'changed'}"
    assert not torch.equal(net.weight(net.critic.head), before[net.critic.head.name])
