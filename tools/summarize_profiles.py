"""Turn the scratch ncu artefacts under gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r1
writes profiles/launches_<round>.csv.gz (raw ncu launch list), profiles/launches_<round>_summary.md (per-kernel totals and
shares) and profiles/env_step_<round>_ncu.md (key metrics of the fused env-step kernel from the --set full capture)."""
import collections
import csv
import gzip
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

launches = os.path.join(src, f"launches_{tag}.csv")
if os.path.exists(launches):
    with open(launches, "rb") as f, gzip.open(os.path.join(dst, f"launches_{tag}.csv.gz"), "wb") as g:
        shutil.copyfileobj(f, g)
    lines = [l for l in open(launches) if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rd:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])[:90]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        u = r[ci["Metric Unit"]]
        v *= {"usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6}.get(u, 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(dst, f"launches_{tag}_summary.md"), "w") as f:
        f.write(f"# ncu launch list summary ({tag})\n\nCommand: `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv python bench.py --steps 1 --warmup 0 "
                f"--no-cpu-baseline --no-e2e` (cold-cache, serialised: compare SHARES).  {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms of kernel time.\n\n"
                "| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.1f} |\n")

rep = os.path.join(src, f"prof_env_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
            "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
            "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_not_selected",
            "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_selected"]
    with open(os.path.join(dst, f"env_step_{tag}_ncu.md"), "w") as f:
        f.write(f"# phc::env_step_kernel<1> -- ncu --set full ({tag})\n\nCommand: `ncu --set full --clock-control none --import-source on -k regex:env_step_kernel "
                "-s 6 -c 1 python tools/profile_env.py 4096 12` (4096 envs, one clip per env, L2 flushed before each launch).\n\n| metric | value | unit |\n|---|---:|---|\n")
        for r in rows[2:3]:
            for w in want:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
print("profiles written:", sorted(os.listdir(dst)))
