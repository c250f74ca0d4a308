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
