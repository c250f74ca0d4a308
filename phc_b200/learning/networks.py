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
import os
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

    def __init__(self, prefix: str, head: str, in_dim: int, units: Sequence[int], out_dim: int, head_relu: bool = False,
                 activation: str = "relu"):
        self.layers: List[LinearSpec] = []
        d = in_dim
        for i, u in enumerate(units):
            self.layers.append(LinearSpec(f"{prefix}.{2 * i}", d, u))
            d = u
        self.layers.append(LinearSpec(head, d, out_dim))
        self.in_dim, self.out_dim = in_dim, out_dim
        self.head_relu = head_relu          # MCP composer: the last Linear is followed by the activation too (ending_act;
                                            # the name is historical: it is the stack's own activation, relu or silu)
        assert activation in ("relu", "silu")
        self.activation = activation        # hidden activation: nn.ReLU (im.yaml) or nn.SiLU (im_big / im_pnn_big / im_mcp_big)

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
                 device="cuda:0", seed: int = 0, kind: str = "amp", num_prim: int = 4, training_prim: int = 0):
        """kind: 'amp' (AMPBuilder, actor_mlp + mu), 'amp_pnn' (AMPPNNBuilder: `num_prim` independent actor columns
        `pnn.actors.K`, column `training_prim` is the one evaluated / trained -- pnn.py:11-131, amp_network_pnn_builder.py:23-87)
        or 'amp_mcp' (AMPMCPBuilder: `composer` MLP whose `action_dim` = num_prim outputs keep the final ReLU --
        amp_network_mcp_builder.py:23-91)."""
        if activation not in ("relu", "silu"):
            raise NotImplementedError(f"activation {activation!r}: the shipped configs use relu and silu")
        self.activation = activation        # mlp.activation; the discriminator is relu in every shipped config
        assert kind in ("amp", "amp_pnn", "amp_mcp")
        self.device = torch.device(device)
        self.kind, self.num_prim, self.training_prim = kind, num_prim, training_prim
        self.obs_dim, self.action_dim, self.amp_dim = obs_dim, action_dim, amp_dim
        n_h = len(units)
        if kind == "amp_pnn":
            self.pnn_actors = [MLPStack(f"pnn.actors.{k}", f"pnn.actors.{k}.{2 * n_h}", obs_dim, units, action_dim, activation=activation)
                               for k in range(num_prim)]
            self.actor = self.pnn_actors[training_prim]
            actor_stacks = self.pnn_actors
        elif kind == "amp_mcp":
            st = MLPStack("composer", f"composer.{2 * n_h}", obs_dim, units, action_dim, head_relu=True, activation=activation)
            self.actor, actor_stacks = st, [st]
        else:
            self.actor = MLPStack("actor_mlp", "mu", obs_dim, units, action_dim, activation=activation)
            actor_stacks = [self.actor]
        self.actor_stacks = actor_stacks
        self.critic = MLPStack("critic_mlp", "value", obs_dim, units, 1, activation=activation)
        self.disc = MLPStack("_disc_mlp", "_disc_logits", amp_dim, disc_units, 1)
        off = 0
        for st in (*actor_stacks, self.critic, self.disc):
            for l in st.layers:
                l.w_off = off
                off += l.out_dim * l.in_pad
                l.b_off = off
                off += round4(l.out_dim)
        self.num_floats = off
        self.params = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=self.device)
        # 3xTF32 operand split of the weights for the tcgen05 GEMM (refreshed after every optimiser step)
        self.params_hi = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.params_lo = torch.zeros(off, dtype=torch.float32, device=self.device)
        # tc5s: the low TF32 term of the weights (lo = rna_tf32(W - trunc_tf32(W))), the operand the GEMM's splitter warps would
        # otherwise recompute on every tile visit; refreshed lazily (torch-side writes bump params._version) and after Adam
        self.params_lo_trunc = torch.zeros(off, dtype=torch.float32, device=self.device)
        self._lo_version = -1
        self.sigma = torch.full((action_dim,), float(sigma_init), dtype=torch.float32, device=self.device)
        self._init_default(seed)

    # ---- views -------------------------------------------------------------------------------------------
    def weight(self, l: LinearSpec, grad: bool = False, part: Optional[str] = None) -> torch.Tensor:
        buf = self.grads if grad else {None: self.params, "hi": self.params_hi, "lo": self.params_lo}[part]
        return buf[l.w_off:l.w_off + l.out_dim * l.in_pad].view(l.out_dim, l.in_pad)

    def refresh_split(self) -> None:
        """hi = rna_tf32(W), lo = rna_tf32(W - hi) over the whole bucket (one streaming pass, 5.5 M floats)."""
        n = self.num_floats
        _lib.check(_lib.load().phc_split_tf32(self.params.data_ptr(), n, 1, n, self.params_hi.data_ptr(), self.params_lo.data_ptr(),
                                              n, _stream()), "phc_split_tf32")

    def refresh_split_lo(self) -> None:
        """params_lo_trunc <- split_lo(params) (one streaming pass over the bucket); call after anything that writes `params`
        behind torch's back (the Adam kernel, a collective)."""
        _lib.check(_lib.load().phc_split_lo(self.params.data_ptr(), self.params_lo_trunc.data_ptr(), self.num_floats, _stream()), "phc_split_lo")
        self._lo_version = self.params._version

    def weight_lo(self, l: LinearSpec) -> torch.Tensor:
        if self._lo_version != self.params._version:
            self.refresh_split_lo()
        return self.params_lo_trunc[l.w_off:l.w_off + l.out_dim * l.in_pad].view(l.out_dim, l.in_pad)

    def bias(self, l: LinearSpec, grad: bool = False) -> torch.Tensor:
        buf = self.grads if grad else self.params
        return buf[l.b_off:l.b_off + l.out_dim]

    def all_layers(self) -> List[LinearSpec]:
        return [l for st in self.actor_stacks for l in st.layers] + self.critic.layers + self.disc.layers

    def load_actor_column(self, checkpoint_model: Dict[str, torch.Tensor], idx: int = 0) -> None:
        """PNN.load_actor (pnn.py:53-60): copy a single-policy checkpoint (actor_mlp.* / mu.*) into primitive column idx."""
        col = self.pnn_actors[idx]
        names = [f"a2c_network.actor_mlp.{2 * i}" for i in range(len(col.hidden))] + ["a2c_network.mu"]
        for l, n in zip(col.layers, names):
            self.set_layer(l, checkpoint_model[n + ".weight"], checkpoint_model[n + ".bias"])

    # ---- init: PyTorch's default nn.Linear init (`initializer: default`), disc biases zero, logits U(-1, 1) --------
    def _init_default(self, seed: int) -> None:
        g = torch.Generator().manual_seed(seed)
        for st in (*self.actor_stacks, self.critic, self.disc):
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
    want = max(1, (4 * 148 + tiles - 1) // tiles)      # ~4 tiles per SM: measured 176 vs 211 us (1024x934, K=16384) against 2 per SM
    return int(max(1, min(want, K // 512, 64)))


def group_splits(K: int) -> int:
    """split-K factor of a weight-gradient problem inside a grouped launch (K = batch rows): about 32 k-blocks of 32 rows per
    split, so a dW tile costs about as much as a dX / forward tile of the same launch (the static tile striding then balances)."""
    env = os.environ.get("PHC_DW_KB_PER_SPLIT")
    per = int(env) if env else 32
    return int(max(1, min(64, round(((K + 31) // 32) / per))))


class MLPEngine:
    """Forward / backward of the MLP stacks through phc_gemm, with per-batch-size activation workspaces."""

    _mode_set = None       # the library's precision switch is process wide: remember what was set last

    def __init__(self, net: AMPNetwork, backend: Optional[str] = None, precision: str = "fp32"):
        """precision: "fp32" = 3xTF32, fp32-equivalent (default, the parity path); "tf32" = one tensor-core pass per product
        (PHC_GEMM_TF32_SINGLE_PASS, ~1e-3 relative, tc5s back end only): the reduced-precision mode of BASELINE.json configs[3]."""
        assert precision in ("fp32", "tf32")
        self.precision = precision
        self.net = net
        self.lib = _lib.load()
        self.dev = net.device
        self._ws: Dict[Tuple[str, int], Dict[str, torch.Tensor]] = {}
        # "tc5s" (default): grouped tcgen05 / TMEM / TMA 3xTF32 with the operand split in shared memory (gemm_tc5s.cu);
        # "tc5": the round-1 kernel with operands pre-split in global memory (gemm_tc5.cu); "mma": warp-level mma.sync 3xTF32
        # (gemm.cu).  The latter two stay as cross-check implementations (PHC_GEMM=tc5 | mma)
        self.backend = backend or os.environ.get("PHC_GEMM", "tc5s")
        assert self.backend in ("mma", "tc5", "tc5s")
        if precision == "tf32" and self.backend != "tc5s":
            raise ValueError("precision='tf32' (single tensor-core pass) exists on the tc5s back end only")
        self._companions: Dict[Tuple[int, Tuple[int, ...], Tuple[int, ...]], Tuple[torch.Tensor, torch.Tensor]] = {}
        # pre-split weight operand (PhcGemmDesc.B_lo): only the -DPHC_TC5S_BLO_IN_STAGE build of the library uses it (measured, not adopted)
        self.presplit = os.environ.get("PHC_TC5S_PRESPLIT", "0") == "1"
        self.gemm_flops = 0.0          # algorithmic fp32 FLOPs (2 M N K) of every grouped launch so far (bench.py reads it)
        if self.backend == "tc5":
            net.refresh_split()

    # -- operand split bookkeeping for the tcgen05 path --------------------------------------------------------------
    def companions(self, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(hi, lo) buffers that shadow activation tensor `t` (same shape / strides)."""
        key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()))
        c = self._companions.get(key)
        if c is None:
            base_rows, ld = t.shape[0], t.stride(0)
            c = (torch.zeros(base_rows, ld, device=self.dev)[:, :t.shape[1]], torch.zeros(base_rows, ld, device=self.dev)[:, :t.shape[1]])
            self._companions[key] = c
        return c

    def split(self, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        hi, lo = self.companions(t)
        rc = self.lib.phc_split_tf32(t.data_ptr(), t.stride(0), t.shape[0], t.shape[1], hi.data_ptr(), lo.data_ptr(), hi.stride(0), _stream())
        if rc:
            _lib.check(rc, "phc_split_tf32")
        return hi, lo

    def _weight_parts(self, W: torch.Tensor) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """If W is a view into the parameter bucket, the matching views of the pre-split buckets."""
        p = self.net.params
        off = (W.data_ptr() - p.data_ptr()) // 4
        if 0 <= off < p.numel() and W.data_ptr() >= p.data_ptr():
            n = W.shape[0] * W.stride(0)
            return (self.net.params_hi[off:off + n].view(W.shape[0], W.stride(0))[:, :W.shape[1]],
                    self.net.params_lo[off:off + n].view(W.shape[0], W.stride(0))[:, :W.shape[1]])
        return None

    # -- raw GEMM ------------------------------------------------------------------------------------------------
    def gemm(self, A, a_k, B, b_k, C, M, N, K, alpha=1.0, bias=None, relu=False, mask=None, accumulate=False, k_splits=1,
             a_split=None, b_split=None, split_out: bool = False, act: Optional[int] = None):
        """act: PHC_ACT_* code (default: RELU if relu else NONE); mask: the epilogue's `aux` matrix (see include/phc_b200.h).
        a_split / b_split: already up-to-date (hi, lo) companions of the operand (skips the split pass);
        split_out: also produce C's companions in the epilogue (tc5 only)."""
        lda = A.stride(0)
        ldb = B.stride(0)
        if act is None:
            act = _lib.PHC_ACT_RELU if relu else _lib.PHC_ACT_NONE
        if self.backend == "tc5":
            Ah, Al = a_split or self._weight_parts(A) or self.split(A)
            Bh, Bl = b_split or self._weight_parts(B) or self.split(B)
            Ch = Cl = None
            if split_out and not accumulate:
                Ch, Cl = self.companions(C)
            rc = self.lib.phc_gemm_tc5(Ah.data_ptr(), Al.data_ptr(), lda, 1 if a_k else 0, Bh.data_ptr(), Bl.data_ptr(), ldb,
                                       1 if b_k else 0, C.data_ptr(), _ptr(Ch), _ptr(Cl), C.stride(0), M, N, K, alpha, _ptr(bias),
                                       act, _ptr(mask), mask.stride(0) if mask is not None else 0,
                                       1 if accumulate else 0, k_splits, _stream())
            if rc:
                _lib.check(rc, "phc_gemm_tc5")
            return (Ch, Cl) if Ch is not None else None
        if self.backend == "tc5s":
            self._set_mode()
        fn = self.lib.phc_gemm_tc5s if self.backend == "tc5s" else self.lib.phc_gemm
        rc = fn(A.data_ptr(), lda, 1 if a_k else 0, B.data_ptr(), ldb, 1 if b_k else 0, C.data_ptr(),
                               C.stride(0), M, N, K, alpha, _ptr(bias), act, _ptr(mask),
                               mask.stride(0) if mask is not None else 0, 1 if accumulate else 0, k_splits, _stream())
        if rc:
            _lib.check(rc, "phc_gemm")

    # -- grouped launches (tc5s only): one persistent kernel over the tiles of up to PHC_GEMM_GROUP_MAX independent GEMMs ------
    def gdesc(self, A, a_k, B, b_k, C, M, N, K, alpha=1.0, bias=None, act=0, aux=None, accumulate=False, k_splits=1, b_lo=None):
        """b_lo: the pre-split low part of B (same shape / stride), see wlo()."""
        return _lib.PhcGemmDesc(A.data_ptr(), A.stride(0), 1 if a_k else 0, B.data_ptr(), B.stride(0), 1 if b_k else 0, C.data_ptr(),
                                C.stride(0), M, N, K, alpha, _ptr(bias), act, _ptr(aux), aux.stride(0) if aux is not None else 0,
                                1 if accumulate else 0, k_splits, _ptr(b_lo))

    def wlo(self, l: LinearSpec) -> Optional[torch.Tensor]:
        """The pre-split lo view of layer l's weight for a B operand (None: single-pass mode or PHC_TC5S_PRESPLIT=0)."""
        return self.net.weight_lo(l) if self.presplit and self.precision != "tf32" else None

    def fwd_desc(self, st: MLPStack, li: int, x: torch.Tensor, ws: Dict[str, torch.Tensor]):
        """Layer li of the forward pass of stack st on batch x / workspace ws (the same epilogues as forward())."""
        net, l, B = self.net, st.layers[li], x.shape[0]
        inp = x if li == 0 else ws["h"][li - 1]
        silu = st.activation == "silu"
        if li < len(st.hidden):
            out = ws["h"][li]
            act = _lib.PHC_ACT_SILU if silu else (_lib.PHC_ACT_RELU_BITS if "hbits" in ws else _lib.PHC_ACT_RELU)
            aux = ws["z"][li] if silu else (ws["hbits"][li] if "hbits" in ws else None)
        else:
            out = ws["out"]
            act = _lib.PHC_ACT_NONE if not st.head_relu else (_lib.PHC_ACT_SILU if silu else (_lib.PHC_ACT_RELU_BITS if "obits" in ws else _lib.PHC_ACT_RELU))
            aux = None if not st.head_relu else (ws["z_out"] if silu else ws.get("obits"))
        return self.gdesc(inp, True, net.weight(l), True, out, B, l.out_dim, l.in_dim, bias=net.bias(l), act=act, aux=aux, b_lo=self.wlo(l))

    def bwd_descs(self, st: MLPStack, li: int, x: torch.Tensor, ws: Dict[str, torch.Tensor], dx: Optional[torch.Tensor] = None):
        """(dW, dX) problems of layer li: dW[out, in] += dY^T X (split-K, reduce-add into the gradient bucket) and
        dX = (dY W) * act'(layer below).  dX is None for the first layer unless `dx` is given."""
        net, l, B = self.net, st.layers[li], x.shape[0]
        dY = ws["dout"] if li == len(st.layers) - 1 else ws["dh"][li]
        inp = x if li == 0 else ws["h"][li - 1]
        dw = self.gdesc(dY, False, inp, False, net.weight(l, grad=True), l.out_dim, l.in_dim, B, accumulate=True, k_splits=group_splits(B))
        dxd = None
        if li > 0:
            if st.activation == "silu":
                dxd = self.gdesc(dY, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim, act=_lib.PHC_ACT_SILU_BWD, aux=ws["z"][li - 1], b_lo=self.wlo(l))
            elif "hbits" in ws:
                dxd = self.gdesc(dY, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim, act=_lib.PHC_ACT_MASK_BITS, aux=ws["hbits"][li - 1], b_lo=self.wlo(l))
            else:
                dxd = self.gdesc(dY, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim, aux=ws["h"][li - 1], b_lo=self.wlo(l))
        elif dx is not None:
            dxd = self.gdesc(dY, True, net.weight(l), False, dx, B, l.in_dim, l.out_dim, b_lo=self.wlo(l))
        return dw, dxd

    def _set_mode(self) -> None:
        mode = _lib.PHC_GEMM_TF32_SINGLE_PASS if self.precision == "tf32" else _lib.PHC_GEMM_FP32_3XTF32
        if MLPEngine._mode_set != mode:
            _lib.check(self.lib.phc_gemm_set_precision(mode), "phc_gemm_set_precision")
            MLPEngine._mode_set = mode

    def run_group(self, descs) -> None:
        self._set_mode()
        descs = [d for d in descs if d is not None]
        for i in range(0, len(descs), _lib.PHC_GEMM_GROUP_MAX):
            part = descs[i:i + _lib.PHC_GEMM_GROUP_MAX]
            self.gemm_flops += sum(2.0 * d.M * d.N * d.K for d in part)
            arr = (_lib.PhcGemmDesc * len(part))(*part)
            rc = self.lib.phc_gemm_group(arr, len(part), _stream())
            if rc:
                _lib.check(rc, "phc_gemm_group")

    def forward_group(self, items) -> None:
        """items: [(stack, x, ws)]: the stacks advance layer by layer together, one grouped launch per layer index."""
        depth = max(len(st.layers) for st, _, _ in items)
        for li in range(depth):
            self.run_group([self.fwd_desc(st, li, x, ws) for st, x, ws in items if li < len(st.layers)])

    def colsum(self, X, M, N, out, alpha=1.0, accumulate=True):
        rc = self.lib.phc_colsum(X.data_ptr(), X.stride(0), M, N, alpha, out.data_ptr(), 1 if accumulate else 0, _stream())
        if rc:
            _lib.check(rc, "phc_colsum")

    def colsum_group(self, items) -> None:
        """items: [(X, M, N, out)] -- out[n] += sum_m X[m, n] for all of them in one launch (phc_colsum_group)."""
        for i in range(0, len(items), _lib.PHC_GEMM_GROUP_MAX):
            part = items[i:i + _lib.PHC_GEMM_GROUP_MAX]
            arr = (_lib.PhcColsumDesc * len(part))(*[_lib.PhcColsumDesc(X.data_ptr(), X.stride(0), M, N, 1.0, out.data_ptr()) for X, M, N, out in part])
            rc = self.lib.phc_colsum_group(arr, len(part), _stream())
            if rc:
                _lib.check(rc, "phc_colsum_group")

    # -- workspaces ----------------------------------------------------------------------------------------------
    def workspace(self, tag: str, st: MLPStack, batch: int) -> Dict[str, torch.Tensor]:
        key = (tag, batch)
        ws = self._ws.get(key)
        if ws is None:
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.dev)
            ws = {"h": [z(batch, round4(l.out_dim)) for l in st.hidden], "out": z(batch, round4(st.out_dim)),
                  "dh": [z(batch, round4(l.out_dim)) for l in st.hidden], "dout": z(batch, round4(st.out_dim))}
            if st.activation == "relu" and self.backend == "tc5s":     # ReLU backward from 1 bit per element (PHC_ACT_RELU_BITS)
                ws["hbits"] = [torch.zeros(batch, (l.out_dim + 31) // 32, dtype=torch.int32, device=self.dev) for l in st.hidden]
                if st.head_relu:
                    ws["obits"] = torch.zeros(batch, (st.out_dim + 31) // 32, dtype=torch.int32, device=self.dev)
            if st.activation == "silu":          # SiLU backward needs the pre-activations
                ws["z"] = [z(batch, round4(l.out_dim)) for l in st.hidden]
                if st.head_relu:
                    ws["z_out"] = z(batch, round4(st.out_dim))
            self._ws[key] = ws
        return ws

    # -- forward: x is [B, in_pad] (zero padded) -------------------------------------------------------------------
    def forward(self, st: MLPStack, x: torch.Tensor, ws: Dict[str, torch.Tensor]) -> torch.Tensor:
        net, B = self.net, x.shape[0]
        tc5 = self.backend == "tc5"
        cur, cur_split = x, (self.split(x) if tc5 else None)
        ws["x_split"] = cur_split
        ws["h_split"] = []
        silu = st.activation == "silu"
        bits = ws.get("hbits")
        for i, (l, h) in enumerate(zip(st.hidden, ws["h"])):
            cur_split = self.gemm(cur, True, net.weight(l), True, h, B, l.out_dim, l.in_dim, bias=net.bias(l),
                                  act=_lib.PHC_ACT_SILU if silu else (_lib.PHC_ACT_RELU_BITS if bits else _lib.PHC_ACT_RELU),
                                  mask=ws["z"][i] if silu else (bits[i] if bits else None), a_split=cur_split, split_out=True)
            ws["h_split"].append(cur_split)
            cur = h
        l = st.head
        head_act = _lib.PHC_ACT_NONE if not st.head_relu else (_lib.PHC_ACT_SILU if silu else _lib.PHC_ACT_RELU)
        self.gemm(cur, True, net.weight(l), True, ws["out"], B, l.out_dim, l.in_dim, bias=net.bias(l), a_split=cur_split,
                  act=head_act, mask=ws["z_out"] if (st.head_relu and silu) else None)
        return ws["out"]

    # -- backward: ws["dout"] holds d(loss)/d(out) [B, round4(out)]; accumulates into net.grads --------------------
    def backward(self, st: MLPStack, x: torch.Tensor, ws: Dict[str, torch.Tensor], dx: Optional[torch.Tensor] = None) -> None:
        net, B = self.net, x.shape[0]
        tc5 = self.backend == "tc5"
        acts = [x] + ws["h"]
        act_splits = ([ws.get("x_split")] + list(ws.get("h_split", []))) if tc5 else [None] * len(acts)
        dcur = ws["dout"]
        if st.head_relu:                                    # MCP composer: activation after the head (ending_act)
            silu = st.activation == "silu"
            aux = ws["z_out"] if silu else ws["out"]
            rc = self.lib.phc_act_backward(dcur.data_ptr(), dcur.stride(0), aux.data_ptr(), aux.stride(0), B, st.out_dim,
                                           _lib.PHC_ACT_SILU if silu else _lib.PHC_ACT_RELU, _stream())
            if rc:
                _lib.check(rc, "phc_act_backward")
        dsplit = self.split(dcur) if tc5 else None          # the loss kernels wrote dout: split it once for both GEMMs
        for li in range(len(st.layers) - 1, -1, -1):
            l = st.layers[li]
            a_in = acts[li]
            tiles = ((l.out_dim + 127) // 128) * ((l.in_dim + 127) // 128)
            # dW[out, in] += dY^T X
            self.gemm(dcur, False, a_in, False, net.weight(l, grad=True), l.out_dim, l.in_dim, B, accumulate=True,
                      k_splits=_splits(tiles, B), a_split=dsplit, b_split=act_splits[li])
            self.colsum(dcur, B, l.out_dim, net.bias(l, grad=True))
            if li > 0:
                # dX = dY W, times the derivative of the activation of the layer below (ReLU: its output > 0; SiLU: at z)
                if st.activation == "silu":
                    nsplit = self.gemm(dcur, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim,
                                       mask=ws["z"][li - 1], act=_lib.PHC_ACT_SILU_BWD, a_split=dsplit, split_out=True)
                elif ws.get("hbits"):
                    nsplit = self.gemm(dcur, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim,
                                       mask=ws["hbits"][li - 1], act=_lib.PHC_ACT_MASK_BITS)
                else:
                    nsplit = self.gemm(dcur, True, net.weight(l), False, ws["dh"][li - 1], B, l.in_dim, l.out_dim, mask=acts[li],
                                       a_split=dsplit, split_out=True)
                dcur, dsplit = ws["dh"][li - 1], nsplit
            elif dx is not None:
                self.gemm(dcur, True, net.weight(l), False, dx, B, l.in_dim, l.out_dim, a_split=dsplit)
