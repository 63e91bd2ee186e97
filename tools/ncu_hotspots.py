#!/usr/bin/env python
"""Source-level hot spots of one kernel from an `ncu --set full --import-source on` report: the per-SASS-instruction
table of ncu's source page (executed instructions, warp-stall samples by reason) joined with the line info of the SAME
build of the object file (`nvdisasm -g`, needs `-lineinfo`), aggregated per source line.

    python tools/ncu_hotspots.py gpurun_out/prof_r02z_update.ncu-rep k_update_mm10_lf_u cpfft_b200/build/material.o \
        _Z18k_update_mm10_lf_u7UpdArgs [--top 40] [--opcodes]

How profiles/r02r_update_hotspots.md was made.  The object file must be the build that was profiled (the report holds
absolute addresses only; rows are matched by their offset from the kernel's first instruction)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

STALLS = ["stall_wait", "stall_no_inst", "stall_long_sb", "stall_selected", "stall_short_sb", "stall_branch_resolving",
          "stall_math", "stall_not_selected", "stall_dispatch", "stall_lg", "stall_mio"]


def line_info(obj, mangled):
    d = tempfile.mkdtemp()
    subprocess.check_call(f"cd {d} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null", shell=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(d, cub)], text=True).split("\n")
    out, cur, inside = {}, ("?", 0), False
    for l in txt:
        if l.startswith(mangled + ":"):
            inside = True
            continue
        if inside and l.startswith("//---"):
            break
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", l)
        if m:
            out[int(m.group(1), 16)] = cur
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep, kname, obj, mangled = args[:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    off2line = line_info(obj, mangled)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, data = rows[1], []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if r:
            data.append(r)
    col = {n: hdr.index(n) for n in ["Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"] + STALLS}
    base = int(data[0][col["Address"]], 16)
    agg = collections.defaultdict(collections.Counter)
    tot, ops = collections.Counter(), collections.Counter()
    for r in data:
        key = off2line.get(int(r[col["Address"]], 16) - base, ("?", 0))
        a = agg[key]
        vals = {"samp": int(r[col["# Samples"]]), "ex": int(r[col["Instructions Executed"]]), "thr": int(r[col["Thread Instructions Executed"]]), "n": 1}
        vals.update({s: int(r[col[s]]) for s in STALLS})
        a.update(vals); tot.update(vals)
        ops[re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip()).split()[0].split(".")[0]] += vals["ex"]
    print(f"{kname}: {tot['ex']:.4g} warp instructions, {tot['samp']} samples; stall reasons: " +
          ", ".join(f"{s[6:]} {100.0 * tot[s] / tot['samp']:.1f} %" for s in STALLS[:7]))
    if "--opcodes" in sys.argv:
        print("executed opcodes: " + ", ".join(f"{o} {100.0 * n / tot['ex']:.1f} %" for o, n in ops.most_common(14)))
    print()
    print("| source line | SASS instr | executed | samples | lanes | wait | no_inst | long_sb | short_sb |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
        print(f"| {f}:{ln} | {a['n']} | {100.0 * a['ex'] / tot['ex']:.2f} % | {100.0 * a['samp'] / tot['samp']:.2f} % | {a['thr'] / max(a['ex'], 1):.1f} | "
              f"{a['stall_wait']} | {a['stall_no_inst']} | {a['stall_long_sb']} | {a['stall_short_sb']} |")


if __name__ == "__main__":
    main()
