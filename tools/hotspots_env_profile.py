"""Join the SASS source page of an ncu capture of the fused env-step kernel (`Instructions Executed` per SASS instruction) with
the line table of the SAME kernel in the local cubin (nvdisasm -g), and print the executed warp instructions per env by
source function.  Works because the captured binary and the local object are the same build (the instruction streams are
compared opcode by opcode first).

    python tools/hotspots_env_profile.py gpurun_out/env_step_r1c.ncu-rep Li1ELi24ELb0ELb1 [num_envs] > profiles/env_step_r1c_hotspots.md
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, mangled = sys.argv[1], sys.argv[2]
n_env = int(sys.argv[3]) if len(sys.argv) > 3 else 4096

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "phc_b200", "lib", "obj", "env_step.o")], cwd=td, capture_output=True)
    txt = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(os.path.join(td, "*.cubin"))[0]], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith("\t.section\t.text.") and mangled in l][0]
end = next((i for i, l in enumerate(txt) if i > start and l.startswith("\t.section")), len(txt))
cur, ins = None, []
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(1).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(ins) == len(data), f"local cubin has {len(ins)} instructions, the capture {len(data)}: not the same build"
opc = lambda t: (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
assert all(opc(a[0]) == opc(r[1].strip()) for a, r in zip(ins, data)), "instruction streams differ: not the same build"

src = {f: open(os.path.join(ROOT, "phc_b200", "csrc", f)).read().split("\n") for f in ("env_step.cu", "phc_math.cuh", "phc_common.cuh")}


def owner(f, n):
    if f == "env_step.cu":
        return "env_step.cu: kernel body (indexing, staging, stores, control)"
    if f not in src:
        return {"sm_30_intrinsics.hpp": "CUDA headers: __shfl_sync / __syncwarp / vote"}.get(f, f)
    for i in range(min(n, len(src[f])) - 1, -1, -1):
        line = src[f][i]
        m = re.match(r"(?:PHC_HD|__device__ __forceinline__)\s+[\w:<> ]+?\b(\w+)\s*\(", line)
        if m:
            return f"{f}: {m.group(1)}"
    return f


agg, samp = collections.Counter(), collections.Counter()
for (_, cur), r in zip(ins, data):
    k = owner(*cur) if cur else "?"
    agg[k] += int(r[ie])
    samp[k] += int(r[ss])
tot = sum(agg.values())
print(f"# Executed warp instructions per env by source function -- {mangled} ({os.path.basename(rep)})\n")
print(f"`python tools/hotspots_env_profile.py {rep} {mangled} {n_env}`: {tot} executed warp instructions = {tot / n_env:.0f} per env; attribution is by the "
      "INNERMOST inlined function of each SASS instruction (libm bodies count towards the phc_math.cuh function that calls them).\n")
print("| function | warp instr / env | share | pc samples |\n|---|---:|---:|---:|")
for k, v in agg.most_common(28):
    print(f"| `{k}` | {v / n_env:.1f} | {100 * v / tot:.1f}% | {samp[k]} |")

body, bsamp = collections.Counter(), collections.Counter()
for (_, cur), r in zip(ins, data):
    if cur and cur[0] == "env_step.cu":
        body[cur[1]] += int(r[ie])
        bsamp[cur[1]] += int(r[ss])
print("\n## The kernel body by source line (env_step.cu, top 24)\n\n| line | warp instr / env | pc samples | source |\n|---:|---:|---:|---|")
for ln, v in body.most_common(24):
    print(f"| {ln} | {v / n_env:.1f} | {bsamp[ln]} | `{src['env_step.cu'][ln - 1].strip()[:110].replace('|', '/')}` |")
