"""CUDA-event timing of the fused env-step kernel alone (the roofline object of bench.py), for quick A/B checks.
    python tools/time_env.py [num_envs] [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from phc_b200 import synthetic as syn
from phc_b200.env.humanoid_im import HumanoidIm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
task = HumanoidIm({"env": {"num_envs": n}, "motion_data": syn.make_motions(n, seed=0), "seed": 0})
task.reset()
peak, src = bench.measured_peak_gbs()
r = bench.env_kernel_roofline(task, peak, src, iters=iters)
print(json.dumps({k: r[k] for k in ("kernel_us", "kernel_us_event_pair", "kernel_us_differential_raw", "achieved", "frac", "frac_event_pair")} | {"num_envs": n}))
