#!/usr/bin/env python
"""Per-launch DRAM traffic (and FP64 instruction counts) of `ncu --set full` reports as JSON,
read by bench.py for the `roofline.traffic` field.

    python tools/ncu_traffic.py <grid N> [--merge profiles/ncu_traffic.json] gpurun_out/prof_<tag>_*.ncu-rep > new.json

A report may hold several kernels; `--merge old.json` keeps the entries of kernels that were not
re-captured (each entry names the kernel and report it came from).
"""
import csv
import io
import json
import subprocess
import sys

CLASS = {"k_fz": None, "k_fy": "k_fft_y", "k_fx": "k_x_green", "k_iz": "k_inv_z", "k_iz_pipe": "k_inv_z", "k_update_mm10": "k_update_mm10",
         "k_pk1_tangent": "k_pk1_tangent", "k_cg_update": "k_cg_update", "k_update_mm01": "k_update_mm01"}


def val(d, key):
    v, u = d.get(key, ("0", ""))
    f = float(v.replace(",", "")) if v else 0.0
    return f * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)


def main(grid, paths, merge=None):
    out = {"grid": grid, "source": "ncu --set full --clock-control none (see tools/ncu_full.sh, profiles/README.md)", "kernels": {}}
    if merge:
        with open(merge) as f:
            out["kernels"] = json.load(f)["kernels"]
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        for row in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], row)}
            name = d["Kernel Name"][0]
            base = name.replace("void ", "").split("<")[0].split("(")[0]
            cls = CLASS.get(base, base)
            if base.startswith("k_update_mm10"):      # k_update_mm10[_lf][_u], _taylor, _mts: one profiler class
                cls = "k_update_mm10"
            if base == "k_fz":
                cls = "k_fwd_z_K4" if any(("<%d, %d>" % (grid, m)) in name for m in (1, 2, 3)) else "k_fwd_z"
            ent = {"kernel": name, "report": p.split("/")[-1],
                   "dram_bytes_read": val(d, "dram__bytes_read.sum"), "dram_bytes_write": val(d, "dram__bytes_write.sum"),
                   "duration_us": float(d["gpu__time_duration.sum"][0].replace(",", "")) *
                   {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[d["gpu__time_duration.sum"][1]]}
            ent["dram_bytes"] = ent["dram_bytes_read"] + ent["dram_bytes_write"]
            cyc = d.get("sm__cycles_elapsed.max", ("0", ""))[0].replace(",", "")
            cyc = float(cyc) if cyc else 0.0
            for nm in ("dfma", "dmul", "dadd"):
                k = f"smsp__sass_thread_inst_executed_op_{nm}_pred_on.sum.per_cycle_elapsed"
                if k in d and d[k][0] and cyc:
                    ent[nm] = float(d[k][0].replace(",", "")) * cyc          # thread instructions per launch
            if "dfma" in ent:
                ent["fp64_flop"] = 2 * ent["dfma"] + ent.get("dmul", 0.0) + ent.get("dadd", 0.0)
                ent["fp64_pipe_active_pct"] = float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0])
            out["kernels"][cls] = ent
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    args = sys.argv[1:]
    mg = None
    if "--merge" in args:
        i = args.index("--merge"); mg = args[i + 1]; del args[i:i + 2]
    main(int(args[0]), args[1:], mg)
