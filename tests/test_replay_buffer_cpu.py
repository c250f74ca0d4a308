"""ReplayBuffer (row a18) against the reference's own class, run side by side on CPU with the same seed: circular store,
the pre-fill `% head` rule, the permutation refresh.  phc/learning/replay_buffer.py only needs torch, so it is imported as is
(skipped on the GPU box, which has no reference checkout)."""
import importlib.util
import os

import pytest
import torch

REF = "/root/reference/phc/learning/replay_buffer.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(REF), reason="reference checkout not present (GPU box)")


def _ref_class():
    spec = importlib.util.spec_from_file_location("ref_replay_buffer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ReplayBuffer


def test_store_and_sample_match_the_reference():
    from phc_b200.learning.amp_agent import ReplayBuffer
    size, width = 50, 7

    class Stream:                       # each buffer draws its permutations from its own copy of the same generator state
        def __init__(self, seed):
            torch.manual_seed(seed)
            self.state = torch.get_rng_state()

        def __enter__(self):
            torch.set_rng_state(self.state)

        def __exit__(self, *a):
            self.state = torch.get_rng_state()

    s_ref, s_ours = Stream(5), Stream(5)
    with s_ref:
        ref = _ref_class()(size, "cpu")
    with s_ours:
        ours = ReplayBuffer(size, width, "cpu")
    g = torch.Generator().manual_seed(1)
    for step, n in enumerate([8, 8, 8, 20, 13, 50, 3]):          # partial fill, wrap-around, a full-size store
        rows = torch.randn(n, width, generator=g)
        ref.store({"amp_obs": rows})
        ours.store(rows)
        assert ours.get_total_count() == ref.get_total_count() and ours.get_buffer_size() == ref.get_buffer_size()
        assert torch.equal(ours.data, ref._data_buf["amp_obs"]), f"store {step}"
        for k in (5, 17, 30):                                        # crosses the permutation refresh several times
            with s_ref:
                want = ref.sample(k)["amp_obs"]
            with s_ours:
                idx = ours.sample_indices(k)
            assert torch.equal(ours.data[idx], want), f"sample after store {step}"
