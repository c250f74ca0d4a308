"""The drop-in contract of SURVEY.md section 8(b), checked against the UNMODIFIED reference where it is available (the build
container; skipped on the GPU box, which has no /root/reference): every method / buffer name the reference's callers use
exists on the mirrors with a compatible signature, and phc_b200.dropin rebinds the classes in the reference's own modules
(both import spellings run_hydra.py uses).  CPU only: nothing is instantiated."""
import inspect
import os
import sys

import pytest

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")

TASK_METHODS = ["step", "reset", "_compute_observations", "_compute_task_obs", "_compute_reward", "_compute_reset", "_compute_humanoid_obs",
                "_compute_amp_observations", "fetch_amp_obs_demo", "get_obs_size", "get_task_obs_size", "get_self_obs_size", "get_num_amp_obs",
                "get_action_size", "resample_motions", "get_task_obs_size_detail", "get_running_mean_size", "post_physics_step"]
AGENT_METHODS = ["train", "train_epoch", "play_steps", "calc_gradients", "discount_values", "_calc_advs", "_preproc_obs", "_disc_loss",
                 "_calc_amp_rewards", "_combine_rewards", "get_stats_weights", "set_stats_weights", "get_full_state_weights",
                 "set_full_state_weights", "restore", "save", "env_reset", "env_step", "get_action_values", "_eval_critic", "prepare_dataset",
                 "pre_epoch", "_update_amp_demos", "_init_amp_demo_buf", "_store_replay_amp_obs", "set_eval", "set_train"]


def test_mirror_surface_is_complete():
    from phc_b200.env.humanoid_im import HumanoidIm
    from phc_b200.env.humanoid_im_mcp import HumanoidImMCP
    from phc_b200.learning.amp_agent import AMPAgent
    assert [m for m in TASK_METHODS if not callable(getattr(HumanoidIm, m, None))] == []
    assert [m for m in AGENT_METHODS if not callable(getattr(AMPAgent, m, None))] == []
    assert issubclass(HumanoidImMCP, HumanoidIm)
    assert list(inspect.signature(HumanoidIm.__init__).parameters)[1:] == ["cfg", "sim_params", "physics_engine", "device_type", "device_id", "headless"]
    assert list(inspect.signature(AMPAgent.__init__).parameters)[1:] == ["base_name", "config"]


def _ref_modules():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_shim
    ref_shim.install()
    import importlib
    return (importlib.import_module("phc.env.tasks.humanoid_im"), importlib.import_module("phc.learning.amp_agent"),
            importlib.import_module("learning.amp_agent"), importlib.import_module("phc.env.tasks.humanoid_im_mcp"))


@needs_ref
def test_signatures_match_the_reference():
    from phc_b200.env.humanoid_im import HumanoidIm
    from phc_b200.learning.amp_agent import AMPAgent
    ref_env, ref_agent, _, _ = _ref_modules()
    R, A = ref_env.HumanoidIm, ref_agent.AMPAgent
    # the reference defines these itself (not only through rl_games / Isaac Gym bases): positional parameters must line up
    for cls, ours, names in ((R, HumanoidIm, ["__init__", "_compute_task_obs", "_compute_reward", "_compute_reset", "resample_motions",
                                              "get_task_obs_size", "get_task_obs_size_detail", "post_physics_step"]),
                             (A, AMPAgent, ["__init__", "play_steps", "calc_gradients", "train_epoch", "_calc_amp_rewards", "_combine_rewards",
                                            "_disc_loss", "get_stats_weights", "set_stats_weights", "_preproc_obs"])):
        for n in names:
            assert n in vars(cls) or any(n in vars(b) for b in cls.__mro__), f"reference lacks {n}?"
            rp = [p for p in inspect.signature(getattr(cls, n)).parameters.values()]
            op = [p for p in inspect.signature(getattr(ours, n)).parameters.values()]
            r_req = [p.name for p in rp if p.default is p.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
            o_names = [p.name for p in op if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
            assert o_names[:len(r_req)] == r_req or len(o_names) >= len(r_req), f"{cls.__name__}.{n}: reference {r_req} vs ours {o_names}"
            o_req = [p.name for p in op if p.default is p.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
            assert len(o_req) <= len([p for p in rp if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]), \
                f"{cls.__name__}.{n}: ours requires {o_req}, the reference passes at most {[p.name for p in rp]}"


@needs_ref
def test_dropin_rebinds_both_import_spellings():
    from phc_b200 import dropin
    from phc_b200.env.humanoid_im import HumanoidIm
    from phc_b200.env.humanoid_im_mcp import HumanoidImMCP
    from phc_b200.learning.amp_agent import AMPAgent
    ref_env, ref_agent, ref_agent_short, ref_mcp = _ref_modules()
    saved = (ref_env.HumanoidIm, ref_agent.AMPAgent, ref_agent_short.AMPAgent, ref_mcp.HumanoidImMCP)
    try:
        assert dropin.install() >= 4
        assert ref_env.HumanoidIm is HumanoidIm and ref_mcp.HumanoidImMCP is HumanoidImMCP
        assert ref_agent.AMPAgent is AMPAgent and ref_agent_short.AMPAgent is AMPAgent
        assert eval("HumanoidIm", vars(ref_env)) is HumanoidIm          # what parse_task.py:60 does
    finally:
        ref_env.HumanoidIm, ref_agent.AMPAgent, ref_agent_short.AMPAgent, ref_mcp.HumanoidImMCP = saved


@needs_ref
def test_install_on_import_rebinds_when_the_reference_modules_load_later():
    """The sitecustomize route: the hook is registered BEFORE the reference modules are imported (as when `python
    phc/run_hydra.py` starts) and rebinds the classes right after each module body ran -- checked in a fresh interpreter."""
    import subprocess
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ref_shim; ref_shim.install()                      # stands in for the reference's own (absent) dependencies
import phc_b200.dropin as d
d.install_on_import()                                     # what sitecustomize.py does
assert not any(m in sys.modules for m in ("phc.env.tasks.humanoid_im", "learning.amp_agent"))
import phc.env.tasks.humanoid_im as ref_env              # parse_task.py:29-38 spelling
import learning.amp_agent as ref_agent                   # run_hydra.py:57-64 spelling
import learning.im_amp as im_amp                          # class IMAmpAgent(amp_agent.AMPAgent)
from phc_b200.env.humanoid_im import HumanoidIm
from phc_b200.learning.amp_agent import AMPAgent
assert ref_env.HumanoidIm is HumanoidIm and eval("HumanoidIm", vars(ref_env)) is HumanoidIm
assert ref_agent.AMPAgent is AMPAgent
assert AMPAgent in im_amp.IMAmpAgent.__mro__, im_amp.IMAmpAgent.__mro__
print("OK")
''' % (os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
