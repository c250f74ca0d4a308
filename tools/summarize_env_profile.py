"""Summarise one `ncu --set full --import-source on` capture of the fused env-step kernel into profiles/env_step_<tag>_ncu.md:
key counters (raw page) plus the executed-instruction composition by opcode class (SASS source page: `Instructions Executed`
per SASS instruction) -- the evidence behind the "issue-bound, not memory-bound" reading in profiles/README.md.

    python tools/summarize_env_profile.py gpurun_out/env_step_r1c.ncu-rep r1c [num_envs]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
n_env = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
kregex = sys.argv[4] if len(sys.argv) > 4 else "env_step_kernel"


def page(name, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


raw = page("raw")
hdr, units, row = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]
kernel = row[hdr.index("Kernel Name")]

sass = page("source", "--print-source", "sass")
sh = sass[1]
ie = sh.index("Instructions Executed")
GROUPS = {"fp32 arithmetic (FFMA FMUL FADD FSETP FSEL MUFU FCHK I2F ...)": "FFMA FADD FMUL FSETP FSEL FMNMX MUFU FCHK F2I I2F I2FP F2F HFMA2 FRND F2FP".split(),
          "integer / predicate / address (IADD3 IMAD LOP3 SHF LEA ISETP SEL MOV PLOP3 ...)": "IADD3 IMAD LOP3 SHF LEA ISETP SEL MOV PLOP3 VIADD IABS PRMT R2P P2R IMNMX VIMNMX POPC FLO BREV VIADDMNMX".split(),
          "control (BRA BSSY BSYNC WARPSYNC CALL EXIT NOP ...)": "BRA BSSY BSYNC WARPSYNC ENDCOLLECTIVE CALL RET EXIT NOP BAR YIELD BREAK".split(),
          "memory (LDS STS LDG STG LDC LDCU SYNCS UBLKCP fences ...)": "LDS STS LDG STG LDC LDCU LDL STL ULDC SYNCS UBLKCP FENCE MEMBAR ERRBAR CCTL UTMACMDFLUSH LDGDEPBAR DEPBAR UTMALDG UTMASTG ATOMS".split(),
          "warp shuffle / vote": "SHFL VOTE VOTEU MATCH REDUX".split(),
          "uniform datapath (R2UR UMOV UISETP ...)": "R2UR UMOV UISETP ULOP3 UIADD3 USHF UIMAD ULEA S2R S2UR CS2R UPLOP3 USEL UFLO UPOPC ELECT UP2UR UR2UP UPRMT".split()}
per_op = collections.Counter()
for r in sass[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    txt = r[1].strip().split()
    op = (txt[1] if txt[0].startswith("@") else txt[0]).split(".")[0].rstrip(";")
    per_op[op] += int(r[ie])
tot = sum(per_op.values())
by_group = collections.Counter()
for op, v in per_op.items():
    by_group[next((g for g, l in GROUPS.items() if op in l), "other")] += v

out = os.path.join(ROOT, "profiles", f"env_step_{tag}_ncu.md")
with open(out, "w") as f:
    f.write(f"# {kernel.split('(')[0]} -- ncu --set full ({tag})\n\nCommand: `ncu --set full --clock-control none --import-source on -k regex:{kregex} "
            f"-s 6 -c 1 python tools/profile_env.py {n_env} 12` ({n_env} envs, one clip per env, L2 flushed before each launch); summary by "
            "`tools/summarize_env_profile.py`.\n\n| metric | value | unit |\n|---|---:|---|\n")
    for w in want:
        if w in hdr:
            f.write(f"| {w} | {row[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
    f.write(f"\n## Executed warp instructions by class (SASS source page; {tot} total = {tot / n_env:.0f} per env)\n\n| class | per env | share |\n|---|---:|---:|\n")
    for g, v in by_group.most_common():
        f.write(f"| {g} | {v / n_env:.0f} | {100 * v / tot:.1f}% |\n")
    f.write("\nTop opcodes per env: " + ", ".join(f"{op} {v / n_env:.0f}" for op, v in per_op.most_common(16)) + "\n")
print("wrote", out)
