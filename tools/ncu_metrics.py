#!/usr/bin/env python
"""Key metrics of `ncu --set full` reports as a markdown table (one column per report).

    python tools/ncu_metrics.py gpurun_out/prof_r01b_*.ncu-rep > profiles/r01b_ncu_full_summary.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTA/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (smem)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("sm__icc_request_hit_rate.pct", "instruction cache hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
]


def load(path):
    """one (name, grid, block, metrics) tuple per kernel of the report"""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
        grid = d.get("Grid Size", ("", ""))[0]; block = d.get("Block Size", ("", ""))[0]
        res.append((name, grid, block, d))
    return res


def main(paths):
    reps = [r for p in paths for r in load(p)]
    print("ncu --set full --clock-control none, one launch per kernel (command: profiles/README.md)")
    print()
    print("| metric | " + " | ".join(r[0] for r in reps) + " |")
    print("|---|" + "---:|" * len(reps))
    print("| grid / block | " + " | ".join(f"{r[1]} / {r[2]}" for r in reps) + " |")
    for key, label in WANT:
        cells = []
        for r in reps:
            v, u = r[3].get(key, ("", ""))
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        print(f"| {label} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
