"""CPU oracle of one PPO + AMP minibatch update (TEST INFRASTRUCTURE, see oracle/phc_oracle.py header).

Follows AMPAgent.calc_gradients (phc/learning/amp_agent.py:554-688) line by line with torch autograd on the CPU:
normalise -> actor / critic / 3x disc forward -> losses (common_agent.py:512-587, amp_agent.py:732-789) ->
loss = a + critic_coef*c - entropy_coef*ent + bounds_loss_coef*b + disc_coef*disc -> backward ->
clip_grad_norm_(grad_norm) -> Adam(lr, eps=1e-8).  The loss pieces are the functions of oracle/phc_oracle.py that are
pinned against the reference's own methods (tests/golden/learn.npz); rl_games' Gaussian head / Adam wiring is
restated (rl_games==1.1.4 is not in the reference tree) -- parity for those pieces is self-pinned.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import phc_oracle as O


def stack_params(sd: Dict[str, torch.Tensor], prefix: str, head: str, n_hidden: int):
    names = [f"{prefix}.{2 * i}" for i in range(n_hidden)] + [head]
    return [sd[f"a2c_network.{n}.weight"] for n in names], [sd[f"a2c_network.{n}.bias"] for n in names]


def minibatch_update(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], cfg: Dict, n_hidden: int = 2,
                     adam_state=None, step: int = 1, dtype=torch.float32, mu_override=None, mlp_act: str = "relu",
                     n_hidden_disc=None) -> Dict:
    """sd: reference-keyed state dict (fp32, un-padded).  batch: obs_n [B,obs] (already normalised), actions, old_neglogp,
    advantages, old_mu, old_sigma, returns [B,1], amp_agent / amp_replay / amp_demo [Bd, amp] (already normalised)."""
    p = {k: v.clone().to(dtype).requires_grad_(k != "a2c_network.sigma") for k, v in sd.items()}
    batch = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in batch.items()}
    aw, ab = stack_params(p, "actor_mlp", "mu", n_hidden)
    cw, cb = stack_params(p, "critic_mlp", "value", n_hidden)
    dw, db = stack_params(p, "_disc_mlp", "_disc_logits", n_hidden if n_hidden_disc is None else n_hidden_disc)
    logstd = p["a2c_network.sigma"]

    mu = O.mlp_forward(batch["obs_n"], aw, ab, act=mlp_act)      # mlp.activation of the yaml; the discriminator is relu
    mu_exact = mu.detach().clone()
    if mu_override is not None:
        # evaluate the loss AT the given policy mean (same gradient path): with sigma = exp(-2.9) an fp32-level difference in
        # mu moves neglogp by ~150x that, so loss/backward arithmetic can only be compared tightly at identical mu
        mu = mu + (mu_override.to(dtype) - mu).detach()
    values = O.mlp_forward(batch["obs_n"], cw, cb, act=mlp_act)
    sigma = torch.exp(mu * 0.0 + logstd)
    neglogp = O.gaussian_neglogp(batch["actions"], mu, sigma, (mu * 0.0 + logstd))
    a_loss = O.actor_loss(batch["old_neglogp"], neglogp, batch["advantages"], cfg["e_clip"]).mean()
    c_loss = O.critic_loss(values, batch["returns"]).mean()
    b_loss = O.bound_loss(mu).mean()
    entropy = torch.distributions.Normal(mu, sigma).entropy().sum(-1).mean()

    demo = batch["amp_demo"].clone().requires_grad_(True)
    la = O.mlp_forward(batch["amp_agent"], dw, db)
    lr_ = O.mlp_forward(batch["amp_replay"], dw, db)
    ld = O.mlp_forward(demo, dw, db)
    dinfo = O.disc_loss(torch.cat([la, lr_], dim=0), ld, demo, dw[-1], dw, cfg["disc_logit_reg"], cfg["disc_grad_penalty"],
                        cfg["disc_weight_decay"])
    loss = (a_loss + cfg["critic_coef"] * c_loss - cfg["entropy_coef"] * entropy + cfg["bounds_loss_coef"] * b_loss
            + cfg["disc_coef"] * dinfo["disc_loss"])
    params = [v for k, v in p.items() if v.requires_grad]
    names = [k for k, v in p.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, params)
    gdict = dict(zip(names, grads))
    with torch.no_grad():
        kl = O.policy_kl(mu, sigma, batch["old_mu"], batch["old_sigma"])

    # clip + Adam (torch's own implementations are the restatement of rl_games' optimiser wiring)
    leaves = [torch.nn.Parameter(p[k].detach().clone()) for k in names]
    for leaf, g in zip(leaves, grads):
        leaf.grad = g.clone()
    total_norm = torch.nn.utils.clip_grad_norm_(leaves, cfg["grad_norm"]) if cfg.get("truncate_grads", True) else None
    opt = torch.optim.Adam(leaves, lr=cfg["learning_rate"], eps=1e-8)
    if adam_state is not None:
        opt.load_state_dict(adam_state)
    opt.step()
    new_sd = {k: l.detach() for k, l in zip(names, leaves)}
    return dict(loss=loss.detach(), a_loss=a_loss.detach(), c_loss=c_loss.detach(), b_loss=b_loss.detach(), kl=kl,
                entropy=entropy.detach(), disc=dinfo, grads=gdict, new_params=new_sd, total_norm=total_norm, mu=mu_exact,
                values=values.detach(), logits_agent=la.detach(), logits_demo=ld.detach())
