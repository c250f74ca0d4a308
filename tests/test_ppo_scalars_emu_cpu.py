"""The GAE scan and the advantage normalisation (phc_b200/csrc/ppo_scalars.cu, verbatim) on the CPU: block-level emulation
(tests/emu/: every thread of a 1024- / 256-thread block is a std::thread, __syncthreads = a block barrier, `__shared__` =
function-local statics) against CommonAgent.discount_values / _calc_advs of the unmodified reference (tests/golden/learn.npz)."""
import ctypes as C
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

from tests.helpers import close, load      # noqa: E402

P = C.c_void_p


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import build_emu
    lib = C.CDLL(build_emu.build_scalars(str(tmp_path_factory.mktemp("semu"))))
    lib.emu_gae.argtypes = [P, P, P, P, C.c_int32, C.c_int64, C.c_float, C.c_float, P, P]
    lib.emu_adv_norm.argtypes = [P, P, C.c_int64, C.c_int32, P, P]
    return lib


def test_gae_and_adv_norm_vs_reference_golden(emu):
    g = load("learn.npz")
    fd, v, r, nv = (g[k].float().contiguous() for k in ("gae_fdones", "gae_values", "gae_rewards", "gae_next_values"))
    T, N = int(fd.shape[0]), int(fd.shape[1])
    adv, ret = torch.zeros(T, N), torch.zeros(T, N)
    assert emu.emu_gae(fd.data_ptr(), v.data_ptr(), r.data_ptr(), nv.data_ptr(), T, N, 0.99, 0.95, adv.data_ptr(), ret.data_ptr()) == 0
    close(adv.view_as(g["gae_adv"]), g["gae_adv"], rtol=1e-5, atol=5e-6, what="discount_values")
    close(ret, adv + v.view(T, N), what="returns = advs + values")
    flat = lambda t: t.reshape(T, N).transpose(0, 1).reshape(-1).contiguous()
    rets, vals = flat(ret), flat(v)
    out = torch.zeros(T * N)
    ws = torch.zeros(2 * 296, dtype=torch.float64)
    emu.emu_adv_norm(rets.data_ptr(), vals.data_ptr(), T * N, 1, out.data_ptr(), ws.data_ptr())
    close(out, g["adv_norm"].reshape(-1), rtol=1e-4, atol=1e-5, what="_calc_advs")


def test_gae_long_horizon_chunks(emu):
    """T > 32 walks the time axis in chunks of 32 with a carried A_{t+1}; N not a multiple of the 32-env tile."""
    from oracle import phc_oracle as O
    g = torch.Generator().manual_seed(0)
    T, N = 75, 45
    fd = (torch.rand(T, N, generator=g) < 0.05).float()
    v, r, nv = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
    adv = torch.zeros(T, N)
    emu.emu_gae(fd.data_ptr(), v.data_ptr(), r.data_ptr(), nv.data_ptr(), T, N, 0.99, 0.95, adv.data_ptr(), None)
    u = lambda t: t.unsqueeze(-1)
    close(adv, O.gae(fd, u(v), u(r), u(nv), 0.99, 0.95).view(T, N), rtol=1e-5, atol=5e-6, what="gae T=75")


# ---- ppo_update.cu: RunningMeanStd, discriminator reward, Gaussian head ----------------------------------------------
@pytest.fixture(scope="module")
def upd(tmp_path_factory):
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import build_emu
    lib = C.CDLL(build_emu.build_update(str(tmp_path_factory.mktemp("uemu"))))
    lib.emu_rms_apply.argtypes = [P, C.c_int64, C.c_int64, C.c_int32, P, P, C.c_float, C.c_int32, P, C.c_int64, P]
    lib.emu_rms_update.argtypes = [P, C.c_int64, C.c_int64, C.c_int32, P, P, P, P, P]
    lib.emu_disc_reward.argtypes = [P, C.c_int64, P, C.c_int64, C.c_float, C.c_float, C.c_float, P, P]
    lib.emu_gaussian_sample.argtypes = [P, C.c_int64, P, P, C.c_int64, C.c_int32, P, P, P, P]
    return lib


def test_running_mean_std_kernels_vs_reference_golden(upd):
    """RunningMeanStd.forward in train mode over three batches (normalise with the current stats, then fold the batch in), the
    final fp64 statistics and the un-normalise direction: phc/utils/running_mean_std.py through tests/golden/learn.npz."""
    g = load("learn.npz")
    d = 12
    mean, var, cnt = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64), torch.ones((), dtype=torch.float64)
    acc = torch.zeros(2 * d, dtype=torch.float64)
    for i in range(3):
        x = g[f"rms_x{i}"].float().contiguous()
        y = torch.zeros_like(x)
        upd.emu_rms_apply(x.data_ptr(), d, x.shape[0], d, mean.data_ptr(), var.data_ptr(), 1e-5, 0, y.data_ptr(), d, None)
        close(y, g[f"rms_y{i}"], what=f"rms_y{i}")
        upd.emu_rms_update(x.data_ptr(), d, x.shape[0], d, mean.data_ptr(), var.data_ptr(), cnt.data_ptr(), acc.data_ptr(), None)
    # the reference takes the batch mean / var in float32 (input.mean / input.var) before the float64 merge; the kernel
    # accumulates the moments in float64 throughout, so the statistics agree to float32 rounding of the batch moments
    close(mean, g["rms_mean"], rtol=1e-6, atol=1e-7, what="running_mean")
    close(var, g["rms_var"], rtol=1e-6, atol=1e-7, what="running_var")
    close(cnt, g["rms_count"], what="count")
    x = (g["rms_x0"] * 0.1).float().contiguous()
    y = torch.zeros_like(x)
    upd.emu_rms_apply(x.data_ptr(), d, x.shape[0], d, mean.data_ptr(), var.data_ptr(), 1e-5, 1, y.data_ptr(), d, None)
    close(y, g["rms_unnorm"], what="unnorm")
    # row-gathered form (index-composed minibatches): rows 5, 0, 31 of batch 1
    idx = torch.tensor([5, 0, 31], dtype=torch.int64)
    x1 = g["rms_x1"].float().contiguous()
    yg = torch.zeros(3, d)
    upd.emu_rms_apply(x1.data_ptr(), d, 3, d, mean.data_ptr(), var.data_ptr(), 1e-5, 0, yg.data_ptr(), d, idx.data_ptr())
    yf = torch.zeros_like(x1)
    upd.emu_rms_apply(x1.data_ptr(), d, x1.shape[0], d, mean.data_ptr(), var.data_ptr(), 1e-5, 0, yf.data_ptr(), d, None)
    assert torch.equal(yg, yf[idx])


def test_disc_reward_kernel_vs_reference_golden(upd):
    """AMPAgent._calc_disc_rewards + _combine_rewards (amp_agent.py:848-878) on the golden discriminator logits."""
    from oracle import phc_oracle as O
    g = load("learn.npz")
    logits = O.mlp_forward(g["d_x_agent"], [g["d_w1"], g["d_w2"], g["d_w3"]], [g["d_b1"], g["d_b2"], g["d_b3"]]).float().contiguous()
    n = logits.shape[0]
    task = torch.full((n,), 0.7)
    dr, comb = torch.zeros(n), torch.zeros(n)
    upd.emu_disc_reward(logits.data_ptr(), 1, task.data_ptr(), n, 2.0, 0.5, 0.5, dr.data_ptr(), comb.data_ptr())
    close(dr.view(n, 1), g["d_reward"], what="disc reward")
    close(comb.view(n, 1), g["d_combined"], what="combined reward")


def test_gaussian_head_kernel_self_pinned(upd):
    """rl_games' ModelA2CContinuousLogStd (absent from the reference tree): action = mu + sigma * eps, neglogp; against
    torch.distributions.Normal like the oracle's own pin."""
    gen = torch.Generator().manual_seed(2)
    n, A = 37, 69
    mu, noise = torch.randn(n, A, generator=gen), torch.randn(n, A, generator=gen)
    logstd = torch.full((A,), -2.9)
    act, nlp, mus, sig = torch.zeros(n, A), torch.zeros(n), torch.zeros(n, A), torch.zeros(n, A)
    upd.emu_gaussian_sample(mu.data_ptr(), A, logstd.data_ptr(), noise.data_ptr(), n, A, act.data_ptr(), nlp.data_ptr(), mus.data_ptr(), sig.data_ptr())
    close(act, mu + logstd.exp() * noise, what="actions")
    ref = -torch.distributions.Normal(mu, logstd.exp().expand_as(mu)).log_prob(act).sum(-1)
    close(nlp, ref, rtol=1e-5, atol=1e-4, what="neglogp")
    assert torch.equal(mus, mu) and torch.allclose(sig, logstd.exp().expand_as(mu))
