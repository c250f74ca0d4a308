"""Summarise `ncu --set full` captures of the MLP GEMM kernels into profiles/<out>.md: per captured launch the tensor-pipe
activity, issue / LSU / L2 / DRAM utilisation, bytes, and the warp-stall mix (sampling) -- the counters VERDICT r1 asked for.

    python tools/summarize_gemm_profile.py profiles/gemm_tc5s_r2_ncu.md "label=report.ncu-rep" ["label2=report2.ncu-rep" ...]
The three launches of each report are the PPO shapes tools/profile_gemm.py runs: forward obs -> 1024 (M 16384, N 1024, K 934,
bias + ReLU), input gradient 512 -> 1024 with the ReLU mask (M 16384, N 1024, K 512), weight gradient 1024 x 934 over the batch
(K 16384, split-K 9)."""
import csv
import os
import subprocess
import sys

out_path = sys.argv[1]
SHAPES = ["fwd obs->1024 (M 16384, N 1024, K 934)", "dX 512->1024 + mask (M 16384, N 1024, K 512)", "dW 1024x934 (K 16384, split-K 9)"]
FLOPS = [3 * 2.0 * 16384 * 1024 * 934, 3 * 2.0 * 16384 * 1024 * 512, 3 * 2.0 * 1024 * 934 * 16384]
WANT = [("gpu__time_duration.sum", "duration us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU smem wavefronts %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"), ("launch__registers_per_thread", "registers"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("gpc__cycles_elapsed.avg.per_second", "SM clock GHz")]
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "selected", "not_selected", "barrier", "math_pipe_throttle", "mio_throttle", "branch_resolving",
          "no_instructions", "sleeping", "membar", "lg_throttle", "tex_throttle", "dispatch_stall", "drain"]
with open(out_path, "w") as f:
    f.write("# MLP GEMM kernels -- `ncu --set full --clock-control none --import-source on` (round 2)\n\n"
            "Command per variant: `ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 3 -c 3 -o gpurun_out/<name> "
            "python tools/profile_gemm.py <variant>` (one warm launch of each shape, then the profiled one); summarised by "
            "`tools/summarize_gemm_profile.py`.  ncu replays each launch ~40 times with cold caches and its own clock behaviour: read the SHARES, "
            "the durations of record are the CUDA-event numbers of `gemm_microbench_r2*.log` and of bench.py's `roofline_gemm`.\n"
            "`tensor pipe active %` is per SM cycle; a 3xTF32 product is three tensor-core instructions, so 100 % would be the dense TF32 peak.\n\n")
    for spec in sys.argv[2:]:
        label, rep = spec.split("=", 1)
        raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
        hdr, rows = raw[0], raw[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        f.write(f"## {label}\n\nKernel: `{rows[0][ix['Kernel Name']].split('(')[0]}`\n\n| metric | " + " | ".join(SHAPES[:len(rows)]) + " |\n|---|" + "---:|" * len(rows) + "\n")
        for key, name in WANT:
            if key in ix:
                f.write(f"| {name} | " + " | ".join(r[ix[key]] for r in rows) + " |\n")
        f.write("| tensor TFLOP/s under ncu (3 x 2MNK / duration) | " + " | ".join(f"{FLOPS[i] / float(r[ix['gpu__time_duration.sum']]) * 1e-6:.0f}" for i, r in enumerate(rows)) + " |\n")
        f.write("\nWarp-stall samples (all warps: 4 epilogue, 1 TMA, 1 MMA-issue, 8 splitter; most warps of a warp-specialised kernel WAIT by design):\n\n| stall | " +
                " | ".join(s.split(" (")[0][:14] for s in SHAPES[:len(rows)]) + " |\n|---|" + "---:|" * len(rows) + "\n")
        for st in STALLS:
            key = f"smsp__pcsamp_warps_issue_stalled_{st}"
            if key in ix and any(r[ix[key]] not in ("", "0") for r in rows):
                f.write(f"| {st} | " + " | ".join(r[ix[key]] for r in rows) + " |\n")
        f.write("\n")
print("wrote", out_path)
