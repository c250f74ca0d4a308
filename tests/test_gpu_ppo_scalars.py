"""GPU parity of the PPO scalar kernels (GAE warp scan, advantage normalisation) vs golden + oracle."""
import pytest
import torch

from oracle import phc_oracle as O
from phc_b200 import ops, synthetic as syn
from tests.helpers import close, load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_gae_vs_reference_golden():
    g = load("learn.npz")
    adv, ret = ops.gae(g["gae_fdones"].to(DEV), g["gae_values"].to(DEV), g["gae_rewards"].to(DEV), g["gae_next_values"].to(DEV), 0.99, 0.95)
    close(adv.cpu(), g["gae_adv"], rtol=1e-5, atol=2e-6, what="gae")
    close(ret.cpu(), g["gae_adv"] + g["gae_values"], rtol=1e-5, atol=2e-6, what="returns")
    T, N = adv.shape[0], adv.shape[1]
    flat = lambda x: x.transpose(0, 1).reshape(T * N, -1).contiguous()
    close(ops.adv_norm(flat(ret), flat(g["gae_values"].to(DEV))).cpu(), g["adv_norm"], rtol=1e-5, atol=2e-6, what="adv_norm")


@pytest.mark.parametrize("T,N", [(1, 1), (32, 4096), (7, 33), (45, 70), (64, 31), (100, 5)])
def test_gae_vs_oracle(T, N):
    fd, v, r, nv = syn.make_rollout(N, T, seed=T + N)
    fd[T // 2] = 1.0                                   # a full row of episode ends: scan segments restart everywhere
    exp = O.gae(fd, v, r, nv, 0.99, 0.95)
    adv, ret = ops.gae(fd.to(DEV), v.to(DEV), r.to(DEV), nv.to(DEV), 0.99, 0.95)
    close(adv.cpu(), exp, rtol=1e-5, atol=5e-6, what="gae")
    close(ret.cpu(), exp + v, rtol=1e-5, atol=5e-6, what="returns")
    if T * N >= 2:
        e2 = O.normalize_advantages((exp + v).reshape(-1, 1), v.reshape(-1, 1))
        close(ops.adv_norm(ret.reshape(-1, 1), v.to(DEV).reshape(-1, 1)).cpu(), e2, rtol=2e-5, atol=5e-6, what="adv_norm")


def test_gae_linearity_full_size():
    """Size-independent property at the bench size: GAE is linear in (rewards, values, next_values) for fixed dones."""
    T, N = 32, 16384
    fd, v, r, nv = [x.to(DEV) for x in syn.make_rollout(N, T, seed=1)]
    _, v2, r2, nv2 = [x.to(DEV) for x in syn.make_rollout(N, T, seed=2)]
    a1 = ops.gae(fd, v, r, nv, 0.99, 0.95, want_returns=False)
    a2 = ops.gae(fd, v2, r2, nv2, 0.99, 0.95, want_returns=False)
    a12 = ops.gae(fd, v + 2 * v2, r + 2 * r2, nv + 2 * nv2, 0.99, 0.95, want_returns=False)
    close(a12.cpu(), (a1 + 2 * a2).cpu(), rtol=1e-4, atol=1e-4, what="linearity")
    adv = ops.adv_norm((a1 + v).reshape(-1), v.reshape(-1))
    assert abs(float(adv.mean())) < 1e-4 and abs(float(adv.std()) - 1.0) < 1e-4


def test_pd_targets_bit_exact_incl_reduce_action_and_frozen_dofs():
    """phc_pd_targets = Humanoid._action_to_pd_targets (humanoid.py:1711-1713) as pre_physics_step uses it (:1540-1556): the
    affine map, the reduce_action scatter (actions_full[:, action_idx] = actions) and the frozen hand / toe dofs; bit-exact
    against the torch expressions of the reference evaluated on the CPU."""
    from phc_b200 import synthetic as syn
    from phc_b200.env.humanoid_im import HumanoidIm
    n = 300
    task = HumanoidIm({"env": {"num_envs": n}, "motion_data": syn.make_motions(8, seed=0, min_frames=20, max_frames=30), "seed": 0})
    D = task.num_dof
    g = torch.Generator().manual_seed(4)
    off, sc = torch.randn(D, generator=g), torch.rand(D, generator=g) * 3 + 0.1
    act = torch.randn(n, D, generator=g)
    task.set_pd_action_map(off, sc)
    got = task._action_to_pd_targets(act.to(task.device))
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), off + sc * act)
    # reduce_action + frozen toes / hands
    idx = torch.randperm(D, generator=g)[:40].sort().values
    frozen = [3, 4, 5, 60, 61, 62]
    task.set_pd_action_map(off, sc, action_idx=idx.tolist(), zero_dofs=frozen)
    a2 = torch.randn(n, 40, generator=g)
    got = task._action_to_pd_targets(a2.to(task.device))
    torch.cuda.synchronize()
    full = torch.zeros(n, D)
    full[:, idx] = a2
    exp = off + sc * full
    exp[:, frozen] = 0
    assert torch.equal(got.cpu(), exp)
    # step() hands the targets (not the raw action) to the backend
    seen = {}
    orig = task.sim.simulate
    task.sim.simulate = lambda a: (seen.__setitem__("a", a.clone()), orig(a))[1]
    task.reset()
    task.step(a2.to(task.device))
    torch.cuda.synchronize()
    assert torch.equal(seen["a"].cpu(), exp)
