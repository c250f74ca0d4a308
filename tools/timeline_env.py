"""Where the time of ONE launch of the fused env-step kernel goes (diagnostic; needs the PHC_EXP_TIMELINE build):
    python -c "from phc_b200 import build as b; b.build_variant('tl', ['-DPHC_EXP_TIMELINE'])"
    PHC_LIB_PATH=$PWD/phc_b200/lib/alt_tl/libphc_b200.so python tools/timeline_env.py [num_envs]
Every env's warp stamps %globaltimer at the kernel's phase boundaries (entry, loads issued, state + pose cache landed, end of
phase A, observation frames landed, rows staged, bulk stores read) plus %smid.  L2 is flushed before the launch; a one-thread
kernel-free reference is the CUDA-event time of the same launch.  Prints percentiles per stamp relative to the earliest entry."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from phc_b200 import _lib
from phc_b200 import synthetic as syn
from phc_b200.env.humanoid_im import HumanoidIm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
task = HumanoidIm({"env": {"num_envs": n}, "motion_data": syn.make_motions(n, seed=0), "seed": 0})
task.reset()
lib = _lib.load()
for fn in (lib.phc_exp_set_timeline, lib.phc_exp_set_timeline_fast):      # one buffer pointer per kernel source file
    fn.argtypes = [C.c_void_p]
    fn.restype = C.c_int
tl = torch.zeros(n, 8, dtype=torch.int64, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for i in range(6):
    task.sim.simulate(None)
    flush.fill_(float(i))
    task._plan.run()
_lib.check(lib.phc_exp_set_timeline(tl.data_ptr()))
_lib.check(lib.phc_exp_set_timeline_fast(tl.data_ptr()))
runs = []
for i in range(5):
    task.sim.simulate(None)
    flush.fill_(float(i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    task._plan.run()
    e1.record()
    torch.cuda.synchronize()
    t = tl.cpu().numpy().astype(np.int64)
    runs.append((e0.elapsed_time(e1) * 1e3, t.copy()))
names = ["entry", "loads issued", "state+cache landed", "phase A done", "obs frames landed", "rows staged", "stores read (exit)"]
ev_us, t = runs[-1]
t0 = t[:, 0].min()
rel = (t[:, :7] - t0) / 1e3          # us
out = {"num_envs": n, "event_pair_us": ev_us, "span_us": float(rel[:, 6].max()), "stamps_us": {}}
print(f"num_envs {n}: CUDA-event pair {ev_us:.2f} us; first entry -> last exit {rel[:, 6].max():.2f} us")
print(f"{'stamp':24s} {'min':>7s} {'p10':>7s} {'p50':>7s} {'p90':>7s} {'max':>7s}")
for k, nm in enumerate(names):
    q = np.percentile(rel[:, k], [0, 10, 50, 90, 100])
    out["stamps_us"][nm] = [float(x) for x in q]
    print(f"{nm:24s} " + " ".join(f"{x:7.2f}" for x in q))
d = np.diff(rel, axis=1)
print("per-warp phase durations (us):")
for k in range(6):
    q = np.percentile(d[:, k], [0, 10, 50, 90, 100])
    print(f"  {names[k]:>20s} -> {names[k + 1]:20s} " + " ".join(f"{x:7.2f}" for x in q))
sm = t[:, 7]
per_sm_first = np.array([rel[sm == s, 0].min() for s in np.unique(sm)])
per_sm_last = np.array([rel[sm == s, 6].max() for s in np.unique(sm)])
per_sm_cnt = np.array([(sm == s).sum() for s in np.unique(sm)])
print(f"SMs used {len(per_sm_cnt)}, warps per SM min/max {per_sm_cnt.min()}/{per_sm_cnt.max()}")
print("per-SM first entry  : " + " ".join(f"{x:7.2f}" for x in np.percentile(per_sm_first, [0, 10, 50, 90, 100])))
print("per-SM last exit    : " + " ".join(f"{x:7.2f}" for x in np.percentile(per_sm_last, [0, 10, 50, 90, 100])))
# entry order inside one SM: the k-th CTA of an SM
order = np.zeros(n)
for s in np.unique(sm):
    idx = np.where(sm == s)[0]
    order[idx[np.argsort(rel[idx, 0], kind="stable")]] = np.arange(len(idx))
for k in (0, 4, 8, 12, 16, 20, 24, 27):
    m = order == k
    if m.any():
        print(f"  warp #{k:2d} of its SM: entry {rel[m, 0].mean():6.2f}  landed {rel[m, 2].mean():6.2f}  A done {rel[m, 3].mean():6.2f}  "
              f"frames {rel[m, 4].mean():6.2f}  staged {rel[m, 5].mean():6.2f}  exit {rel[m, 6].mean():6.2f}")
out["runs_event_us"] = [r[0] for r in runs]
out["runs_span_us"] = [float(((r[1][:, 6].max() - r[1][:, 0].min()) / 1e3)) for r in runs]
print(json.dumps(out))
