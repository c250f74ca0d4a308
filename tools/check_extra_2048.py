"""The 8-rank extra configuration of bench.py (2048 envs per rank: BASELINE configs[2], 16384 envs over 8 GPUs) exercised on ONE GPU:
the same run_extra_config call the 8-rank run makes, so that a shape problem shows up here and not inside the scaling run."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
peak, src = bench.measured_peak_gbs()
bench.WORKLOADS["smpl_2048"] = dict(bench.WORKLOADS["smpl"], envs=2048, desc=bench.WORKLOADS["smpl"]["desc"] + " -- 16384 envs sharded over 8 GPUs (BASELINE configs[2])")
out = bench.run_extra_config("smpl_2048", dev, 0, 1, peak, src)
print(json.dumps(out)[:1500])
assert "error" not in out, out
