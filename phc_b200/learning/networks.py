"""Actor / critic / discriminator networks of the AMP agent on flat fp32 buckets, computed by libphc_b200.so.

Mirrors the module tree the reference builds with rl_games' builders so that checkpoints interchange:
  AMPBuilder.Network (phc/learning/amp_network_builder.py:13-249) on top of A2CBuilder.Network
  (phc/learning/network_builder.py:130-330): `actor_mlp` / `critic_mlp` / `_disc_mlp` are nn.Sequential(Linear, act, ...)
  so the Linear layers sit at even indices (`actor_mlp.0`, `actor_mlp.2`), heads are `mu`, `value`, `_disc_logits`,
  `sigma` is a fixed (requires_grad False) log-std initialised to -2.9 (im.yaml:22-27).
State-dict keys are `a2c_network.<name>.{weight,bias}` exactly as the reference's `model.state_dict()`.

Storage: every trainable tensor lives in ONE flat parameter bucket (and one flat gradient bucket of the same layout)
so that the per-minibatch all-reduce, the global-norm clip and Adam are single passes.  Weight rows are padded to a
multiple of 4 floats (934 -> 936) because the GEMM loads 16-byte chunks; pad columns stay exactly zero.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib
from ..ops import _ptr, _stream


def round4(x: int) -> int:
    return (x + 3) & ~3


@dataclass
class LinearSpec:
    name: str          # e.g. "actor_mlp.0", "mu"
    in_dim: int
    out_dim: int
    w_off: int = 0     # offsets into the flat bucket (floats)
    b_off: int = 0

    @property
    def in_pad(self) -> int:
        return round4(self.in_dim)


class MLPStack:
    """One MLP = hidden Linear+act layers followed by a linear head."""

    def __init__(self, prefix: str, head: str, in_dim: int, units: Sequence[int], out_dim: int):
        self.layers: List[LinearSpec] = []
        d = in_dim
        for i, u in enumerate(units):
            self.layers.append(LinearSpec(f"{prefix}.{2 * i}", d, u))
            d = u
        self.layers.append(LinearSpec(head, d, out_dim))
        self.in_dim, self.out_dim = in_dim, out_dim

    @property
    def hidden(self) -> List[LinearSpec]:
        return self.layers[:-1]

    @property
    def head(self) -> LinearSpec:
        return self.layers[-1]


class AMPNetwork:
    """Parameter container (flat buckets + named views).  Compute lives in MLPEngine."""

    def __init__(self, obs_dim: int, action_dim: int, amp_dim: int, units: Sequence[int] = (1024, 512),
                 disc_units: Sequence[int] = (1024, 512), activation: str = "relu", sigma_init: float = -2.9,
                 device="cuda:0", seed: int = 0):
        if activation != "relu":
            raise NotImplementedError("only the relu MLPs of im.yaml are built so far (silu of im_big.yaml: next)")
        self.device = torch.device(device)
        self.obs_dim, self.action_dim, self.amp_dim = obs_dim, action_dim, amp_dim
        self.actor = MLPStack("actor_mlp", "mu", obs_dim, units, action_dim)
        self.critic = MLPStack("critic_mlp", "value", obs_dim, units, 1)
        self.disc = MLPStack("_disc_mlp", "_disc_logits", amp_dim, disc_units, 1)
        off = 0
        for st in (self.actor, self.critic, self.disc):
            for l in st.layers:
                l.w_off = off
                off += l.out_dim * l.in_pad
                l.b_off = off
                off += round4(l.out_dim)
        self.num_floats = off
        self.params = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.sigma = torch.full((action_dim,), float(sigma_init), dtype=torch.float32, device=self.device)
        self._init_default(seed)

    # ---- views -------------------------------------------------------------------------------------------
    def weight(self, l: LinearSpec, grad: bool = False) -> torch.Tensor:
        buf = self.grads if grad else self.params
        return buf[l.w_off:l.w_off + l.out_dim * l.in_pad].view(l.out_dim, l.in_pad)

    def bias(self, l: LinearSpec, grad: bool = False) -> torch.Tensor:
        buf = self.grads if grad else self.params
        return buf[l.b_off:l.b_off + l.out_dim]

    def all_layers(self) -> List[LinearSpec]:
        return self.actor.layers + self.critic.layers + self.disc.layers

    # ---- init: PyTorch's default nn.Linear init (`initializer: default`), disc biases zero, logits U(-1, 1) --------
    def _init_default(self, seed: int) -> None:
        g = torch.Generator().manual_seed(seed)
        for st in (self.actor, self.critic, self.disc):
            for l in st.layers:
                bound = 1.0 / math.sqrt(l.in_dim)
                w = (torch.rand(l.out_dim, l.in_dim, generator=g) * 2 - 1) * bound     # kaiming_uniform(a=sqrt(5))
                b = (torch.rand(l.out_dim, generator=g) * 2 - 1) * bound
                if st is self.disc:
                    b.zero_()                                                          # amp_network_builder.py:240-244
                    if l is st.head:
                        w = torch.rand(l.out_dim, l.in_dim, generator=g) * 2 - 1       # DISC_LOGIT_INIT_SCALE = 1 (:246)
                self.set_layer(l, w, b)

    def set_layer(self, l: LinearSpec, w: torch.Tensor, b: torch.Tensor) -> None:
        W = self.weight(l)
        W.zero_()
        W[:, :l.in_dim] = w.to(self.device, torch.float32)
        self.bias(l).copy_(b.to(self.device, torch.float32))

    # ---- checkpoint interchange with the reference ----------------------------------------------------------------
    def state_dict(self, prefix: str = "a2c_network.") -> Dict[str, torch.Tensor]:
        sd = {prefix + "sigma": self.sigma.clone()}
        for l in self.all_layers():
            sd[f"{prefix}{l.name}.weight"] = self.weight(l)[:, :l.in_dim].clone()
            sd[f"{prefix}{l.name}.bias"] = self.bias(l).clone()
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "a2c_network.") -> None:
        for l in self.all_layers():
            self.set_layer(l, sd[f"{prefix}{l.name}.weight"], sd[f"{prefix}{l.name}.bias"])
        if prefix + "sigma" in sd:
            self.sigma.copy_(sd[prefix + "sigma"].to(self.device))

    def get_disc_logit_weights(self) -> torch.Tensor:
        return self.weight(self.disc.head)[:, :self.disc.head.in_dim].flatten()

    def get_disc_weights(self) -> List[torch.Tensor]:
        return [self.weight(l)[:, :l.in_dim].flatten() for l in self.disc.layers]


def _splits(tiles: int, K: int) -> int:
    """split-K factor for the weight-gradient GEMMs (M, N are layer widths, K is the batch)."""
    want = max(1, (2 * 148 + tiles - 1) // tiles)
    return int(max(1, min(want, K // 512, 64)))


class MLPEngine:
    """Forward / backward of the MLP stacks through phc_gemm, with per-batch-size activation workspaces."""

    def __init__(self, net: AMPNetwork):
        self.net = net
        self.lib = _lib.load()
        self.dev = net.device
        self._ws: Dict[Tuple[str, int], Dict[str, torch.Tensor]] = {}

    # -- raw GEMM ------------------------------------------------------------------------------------------------
    def gemm(self, A, a_k, B, b_k, C, M, N, K, alpha=1.0, bias=None, relu=False, mask=None, accumulate=False, k_splits=1):
        lda = A.stride(0)
        ldb = B.stride(0)
        rc = self.lib.phc_gemm(A.data_ptr(), lda, 1 if a_k else 0, B.data_ptr(), ldb, 1 if b_k else 0, C.data_ptr(),
                               C.stride(0), M, N, K, alpha, _ptr(bias), 1 if relu else 0, _ptr(mask),
                               mask.stride(0) if mask is not None else 0, 1 if accumulate else 0, k_splits, _stream())
        if rc:
            _lib.check(rc, "phc_gemm")

    def colsum(self, X, M, N, out, alpha=1.0, accumulate=True):
        rc = self.lib.phc_colsum(X.data_ptr(), X.stride(0), M, N, alpha, out.data_ptr(), 1 if accumulate else 0, _stream())
        if rc:
            _lib.check(rc, "phc_colsum")

    # -- workspaces ----------------------------------------------------------------------------------------------
    def workspace(self, tag: str, st: MLPStack, batch: int) -> Dict[str, torch.Tensor]:
        key = (tag, batch)
        ws = self._ws.get(key)
        if ws is None:
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.dev)
            ws = {"h": [z(batch, round4(l.out_dim)) for l in st.hidden], "out": z(batch, round4(st.out_dim)),
                  "dh": [z(batch, round4(l.out_dim)) for l in st.hidden], "dout": z(batch, round4(st.out_dim))}
            self._ws[key] = ws
        return ws

    # -- forward: x is [B, in_pad] (zero padded) -------------------------------------------------------------------
    def forward(self, st: MLPStack, x: torch.Tensor, ws: Dict[str, torch.Tensor]) -> torch.Tensor:
        net, B = self.net, x.shape[0]
        cur = x
        for l, h in zip(st.hidden, ws["h"]):
            self.gemm(cur, True, net.weight(l), True, h, B, l.out_dim, l.in_dim, bias=net.bias(l), relu=True)
            cur = h
        l = st.head
        self.gemm(cur, True, net.weight(l), True, ws["out"], B, l.out_dim, l.in_dim, bias=net.bias(l))
        return ws["out"]

    # -- backward: ws["dout"] holds d(loss)/d(out) [B, round4(out)]; accumulates into net.grads --------------------
    def backward(self, st: MLPStack, x: torch.Tensor, ws: Dict[str, torch.Tensor], dx: Optional[torch.Tensor] = None) -> None:
        net, B = self.net, x.shape[0]
        acts = [x] + ws["h"]
        dcur = ws["dout"]
        for li in range(len(st.layers) - 1, -1, -1):
            l = st.layers[li]
            a_in = acts[li]
            tiles = ((l.out_dim + 127) // 128) * ((l.in_dim + 127) // 128)
            # dW[out, in] += dY^T X
            self.gemm(dcur, False, a_in, False, net.weight(l, grad=True), l.out_dim, l.in_dim, B, accumulate=True,
                      k_splits=_splits(tiles, B))
            self.colsum(dcur, B, l.out_dim, net.bias(l, grad=True))
            if li > 0:
                # dX = dY W, masked by the ReLU of the layer below
                self.gemm(dcur, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim, mask=acts[li])
                dcur = ws["dh"][li - 1]
            elif dx is not None:
                self.gemm(dcur, True, net.weight(l), False, dx, B, l.in_dim, l.out_dim)
