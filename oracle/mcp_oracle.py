"""CPU restatement of the PNN / MCP pieces of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(phc_b200/) never does.  Pinned by tests/golden/mcp.npz, which tests/golden/make_golden.py:gen_mcp produced by
running the unmodified reference classes (PNN, load_pnn, load_mcp_mlp, HumanoidImMCP.step).
"""
from typing import Dict, List

import torch


def _layer_ids(sd: Dict[str, torch.Tensor], prefix: str) -> List[int]:
    return sorted({int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix) and k.endswith(".weight")})


def mlp_forward(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, ending_act: bool = False, act: str = "relu") -> torch.Tensor:
    """nn.Sequential(Linear, ReLU, ..., Linear[, ReLU]) stored under `prefix` + '<2i>.weight/bias'.
    PNN column: pnn.py:22-31 (no activation after the last Linear); composer: amp_network_mcp_builder.py:57-63 and
    network_loader.py:38-40 (activation kept after the last Linear)."""
    ids = _layer_ids(sd, prefix)
    f = {"relu": torch.relu, "silu": torch.nn.functional.silu}[act]
    h = x
    for n, i in enumerate(ids):
        h = h @ sd[f"{prefix}{i}.weight"].to(h.dtype).T + sd[f"{prefix}{i}.bias"].to(h.dtype)
        if n < len(ids) - 1 or ending_act:
            h = f(h)
    return h


def pnn_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_prim: int, idx: int = -1, prefix: str = "a2c_network.pnn.actors."):
    """PNN.forward without laterals (pnn.py:99-107): idx != -1 -> that column's output, else the list of all of them."""
    if idx != -1:
        return mlp_forward(sd, f"{prefix}{idx}.", x)
    return [mlp_forward(sd, f"{prefix}{k}.", x) for k in range(num_prim)]


def pnn_load_actor(single: Dict[str, torch.Tensor], n_hidden: int, idx: int, prefix: str = "a2c_network.pnn.actors.") -> Dict[str, torch.Tensor]:
    """PNN.load_actor (pnn.py:53-60): actor_mlp.<2i> -> actors[idx].<2i>, mu -> actors[idx].<2*n_hidden>."""
    out = {}
    for i in range(n_hidden):
        for p in ("weight", "bias"):
            out[f"{prefix}{idx}.{2 * i}.{p}"] = single[f"a2c_network.actor_mlp.{2 * i}.{p}"]
    for p in ("weight", "bias"):
        out[f"{prefix}{idx}.{2 * n_hidden}.{p}"] = single[f"a2c_network.mu.{p}"]
    return out


def pnn_trainable(names: List[str], idx: int) -> List[bool]:
    """PNN.freeze_pnn(idx) (pnn.py:45-51): columns < idx are frozen, the rest keep requires_grad."""
    return [int(n.split(".")[1]) >= idx for n in names]


def mcp_step_actions(obs_buf: torch.Tensor, running_mean: torch.Tensor, running_var: torch.Tensor, sd: Dict[str, torch.Tensor],
                     weights: torch.Tensor, num_prim: int, discrete: bool = False, dtype=torch.float32) -> torch.Tensor:
    """HumanoidImMCP.step up to pre_physics_step (humanoid_im_mcp.py:64-82)."""
    mean = running_mean.float().to(dtype)
    var = running_var.float().to(dtype)
    cur = (obs_buf.to(dtype) - mean) / torch.sqrt(var + 1e-05)
    cur = torch.clamp(cur, min=-5.0, max=5.0)
    if discrete:
        weights = torch.nn.functional.one_hot(torch.argmax(weights, dim=1), num_classes=num_prim).to(dtype)
    x_all = torch.stack(pnn_forward(sd, cur, num_prim), dim=1)
    return torch.sum(weights.to(dtype)[:, :, None] * x_all, dim=1)
