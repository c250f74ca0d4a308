"""Drives the CUDA learner kernels for one minibatch the way AMPAgent.calc_gradients does, from pre-normalised inputs
(test infrastructure shared by tests/test_gpu_learner.py)."""
import torch

from phc_b200 import _lib
from phc_b200.learning.networks import MLPEngine, round4, _splits


def run_cuda_minibatch(net, batch, cfg, backend=None):
    from phc_b200.learning.amp_agent import AMPAgent
    dev = net.device
    lib = _lib.load()
    eng = MLPEngine(net, backend=backend)
    B, Bd, A = batch["obs_n"].shape[0], batch["amp_agent"].shape[0], net.action_dim
    z = lambda *s: torch.zeros(*s, device=dev)
    x = z(B, round4(net.obs_dim)); x[:, :net.obs_dim] = batch["obs_n"].to(dev)
    xa = z(3 * Bd, round4(net.amp_dim))
    xa[:Bd, :net.amp_dim] = batch["amp_agent"].to(dev)
    xa[Bd:2 * Bd, :net.amp_dim] = batch["amp_replay"].to(dev)
    xa[2 * Bd:, :net.amp_dim] = batch["amp_demo"].to(dev)
    # a minimal object with the attributes the agent's update code reads
    ag = AMPAgent.__new__(AMPAgent)
    ag._lib, ag.model, ag.engine, ag.device = lib, net, eng, dev
    ag._disc_coef, ag._disc_grad_penalty = cfg["disc_coef"], cfg["disc_grad_penalty"]
    ag._stats = z(16)
    du = net.disc.hidden
    ag._gp_u = [z(Bd, round4(l.out_dim)) for l in du]
    ag._gp_e = [z(Bd, round4(l.out_dim)) for l in du]
    ag._gp_g = z(Bd, round4(net.amp_dim))
    wa, wc, wd = eng.workspace("a", net.actor, B), eng.workspace("c", net.critic, B), eng.workspace("d", net.disc, 3 * Bd)
    net.grads.zero_()
    t = {k: batch[k].to(dev).contiguous() for k in ("actions", "old_neglogp", "advantages", "old_mu", "old_sigma")}
    rets = batch["returns"].to(dev).reshape(-1).contiguous()
    if eng.backend == "tc5s":          # the grouped path of AMPAgent._update_grouped on the same pre-normalised inputs
        from phc_b200.learning.amp_agent import PhaseTimer
        ag._ws_actor, ag._ws_critic, ag._ws_disc, ag.timer = wa, wc, wd, PhaseTimer(False)
        ag._disc_logit_reg, ag._disc_weight_decay = cfg["disc_logit_reg"], cfg["disc_weight_decay"]

        def losses():
            mu, val, logits = wa["out"], wc["out"], wd["out"]
            _lib.check(lib.phc_ppo_actor_grad(mu.data_ptr(), mu.stride(0), net.sigma.data_ptr(), t["actions"].data_ptr(), t["old_neglogp"].data_ptr(),
                                              t["advantages"].data_ptr(), t["old_mu"].data_ptr(), t["old_sigma"].data_ptr(), B, A, cfg["e_clip"],
                                              cfg["bounds_loss_coef"], 1.0 / B, wa["dout"].data_ptr(), wa["dout"].stride(0), ag._stats.data_ptr(), None))
            _lib.check(lib.phc_ppo_critic_grad(val.data_ptr(), val.stride(0), rets.data_ptr(), B, cfg["critic_coef"], 1.0 / B,
                                               wc["dout"].data_ptr(), wc["dout"].stride(0), ag._stats.data_ptr(), None))
            _lib.check(lib.phc_disc_logit_grad(logits.data_ptr(), logits.stride(0), 2 * Bd, Bd, cfg["disc_coef"], wd["dout"].data_ptr(),
                                               wd["dout"].stride(0), ag._stats.data_ptr(), None))
        ag._grouped_core(x, xa, Bd, None, losses)
        mu, val = wa["out"], wc["out"]
        return _finish(ag, net, lib, cfg, mu, val, A, B, Bd, dev)
    mu = eng.forward(net.actor, x, wa)
    val = eng.forward(net.critic, x, wc)
    _lib.check(lib.phc_ppo_actor_grad(mu.data_ptr(), mu.stride(0), net.sigma.data_ptr(), t["actions"].data_ptr(), t["old_neglogp"].data_ptr(),
                                      t["advantages"].data_ptr(), t["old_mu"].data_ptr(), t["old_sigma"].data_ptr(), B, A, cfg["e_clip"],
                                      cfg["bounds_loss_coef"], 1.0 / B, wa["dout"].data_ptr(), wa["dout"].stride(0), ag._stats.data_ptr(), None))
    _lib.check(lib.phc_ppo_critic_grad(val.data_ptr(), val.stride(0), rets.data_ptr(), B, cfg["critic_coef"], 1.0 / B,
                                       wc["dout"].data_ptr(), wc["dout"].stride(0), ag._stats.data_ptr(), None))
    eng.backward(net.actor, x, wa)
    eng.backward(net.critic, x, wc)
    logits = eng.forward(net.disc, xa, wd)
    _lib.check(lib.phc_disc_logit_grad(logits.data_ptr(), logits.stride(0), 2 * Bd, Bd, cfg["disc_coef"], wd["dout"].data_ptr(),
                                       wd["dout"].stride(0), ag._stats.data_ptr(), None))
    eng.backward(net.disc, xa, wd)
    ag._disc_grad_penalty_backward(xa[2 * Bd:], [h[2 * Bd:] for h in wd["h"]], Bd)
    head = net.disc.head
    _lib.check(lib.phc_axpy2d(net.weight(head).data_ptr(), head.in_pad, net.weight(head, True).data_ptr(), head.in_pad, 1, head.in_dim,
                              2.0 * cfg["disc_coef"] * cfg["disc_logit_reg"], ag._stats[11:].data_ptr(), None))
    for l in net.disc.layers:
        _lib.check(lib.phc_axpy2d(net.weight(l).data_ptr(), l.in_pad, net.weight(l, True).data_ptr(), l.in_pad, l.out_dim, l.in_dim,
                                  2.0 * cfg["disc_coef"] * cfg["disc_weight_decay"], ag._stats[12:].data_ptr(), None))
    return _finish(ag, net, lib, cfg, mu, val, A, B, Bd, dev)


def _finish(ag, net, lib, cfg, mu, val, A, B, Bd, dev):
    torch.cuda.synchronize()
    grads = {}
    for l in net.all_layers():
        grads[f"a2c_network.{l.name}.weight"] = net.weight(l, True)[:, :l.in_dim].clone()
        grads[f"a2c_network.{l.name}.bias"] = net.bias(l, True).clone()
        assert float(net.weight(l, True)[:, l.in_dim:].abs().sum()) == 0.0, "pad columns of the gradient must stay zero"
    gs = torch.zeros(1, dtype=torch.float64, device=dev)
    m, v = torch.zeros_like(net.params), torch.zeros_like(net.params)
    _lib.check(lib.phc_grad_sumsq(net.grads.data_ptr(), net.num_floats, gs.data_ptr(), None))
    _lib.check(lib.phc_adam_step(net.params.data_ptr(), net.grads.data_ptr(), m.data_ptr(), v.data_ptr(), net.num_floats, gs.data_ptr(),
                                 1.0, cfg["grad_norm"], cfg["learning_rate"], 0.9, 0.999, 1e-8, 1, None))
    torch.cuda.synchronize()
    ag._last_B, ag._last_Bd = B, Bd
    return dict(mu=mu[:, :A].clone(), values=val[:, :1].clone(), stats=ag.train_result_dict(), grads=grads,
                total_norm=float(gs.sqrt()), new_params=net.state_dict())
