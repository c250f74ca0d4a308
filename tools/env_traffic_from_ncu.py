"""DRAM bytes per launch of the fused env-step kernel from one `ncu --set full` capture -> profiles/env_step_traffic.json (what
bench.py reports as roofline.traffic):  python tools/env_traffic_from_ncu.py gpurun_out/env_step_r2.ncu-rep [num_envs] [out...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
outs = sys.argv[3:] or [os.path.join(ROOT, "profiles", "env_step_traffic.json")]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, row = raw[0], raw[1], raw[2]


def nbytes(name):
    v, u = float(row[hdr.index(name)].replace(",", "")), units[hdr.index(name)].lower()
    return int(round(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]))


rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
d = {"kernel": row[hdr.index("Kernel Name")].split("(")[0], "num_envs": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
     "dram_bytes_per_launch": rd + wr,
     "source": f"{os.path.basename(rep)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch, L2 flushed); summary in profiles/env_step_r2_ncu.md"}
for o in outs:
    json.dump(d, open(o, "w"), indent=1)
print(json.dumps(d))
