"""N-rank NCCL check of the data-parallel learner (run under torchrun): every rank trains 2 epochs on its own envs; afterwards the
parameter buckets, Adam moments and normaliser statistics must be IDENTICAL on all ranks (same reduced gradient, same clip norm, same
step), and different from a rank that trained alone on rank 0's data (the all-reduce really mixed the gradients).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_ddp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
device = torch.device("cuda", local)
torch.cuda.set_device(device)
dist.init_process_group("nccl", device_id=device)

from phc_b200 import synthetic as syn
from phc_b200.env.humanoid_im import HumanoidIm, RLGPUEnv
from phc_b200.learning.amp_agent import AMPAgent


def make(multi, seed_rank, n=256):
    task = HumanoidIm({"env": {"num_envs": n}, "motion_data": syn.make_motions(n, seed=seed_rank), "seed": seed_rank}, device_type="cuda", device_id=local)
    cfg = {"multi_gpu": multi, "seed": 0, "device": str(device), "vec_env": RLGPUEnv(task), "horizon_length": 8, "minibatch_size": 512,
           "amp_minibatch_size": 128, "mini_epochs": 2, "amp_obs_demo_buffer_size": 4096, "amp_replay_buffer_size": 4096, "amp_batch_size": 256,
           "network": {"mlp": {"units": [256, 128], "activation": "relu"}, "disc": {"units": [256, 128], "activation": "relu"}}}
    ag = AMPAgent("ddp", cfg)
    ag.obs = ag.env_reset()
    ag._init_amp_demo_buf()
    return ag


torch.manual_seed(100 + rank)
ag = make(True, rank)
for _ in range(2):
    ag.train_epoch()
torch.cuda.synchronize()
for name, t in (("params", ag.model.params), ("exp_avg", ag.exp_avg), ("exp_avg_sq", ag.exp_avg_sq), ("sigma", ag.model.sigma)):
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, t), f"rank {rank}: {name} differs from rank 0 after 2 data-parallel epochs"
assert torch.isfinite(ag.model.params).all()
if rank == 0:
    torch.manual_seed(100)
    solo = make(False, 0)
    for _ in range(2):
        solo.train_epoch()
    torch.cuda.synchronize()
    d = float((solo.model.params - ag.model.params).abs().max())
    assert d > 0.0, "the data-parallel parameters equal a single-rank run: the all-reduce did nothing"
    print(f"check_ddp: {world} ranks identical after 2 epochs; max |param - solo| = {d:.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
