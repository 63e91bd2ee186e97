"""Run one of the reference's input decks end to end on the GPU:

    python -m cpfft_b200 examples/test_mm10.in [--outdir results] [--device 0]

Mirrors `compute` of the reference's main program (FFT_finite_3d.f:138-147): the initial
drive_eps_sig sweep, the FFT_nr3 step loop with the reference's log lines on stdout, and on the
steps selected by `output results steps ...` the flat-text result files of ouresult.f
(wes#####_text stresses, wee#####_text strains).  The CUDA library does all the work; this
module is host glue (deck reader, printing, file writing)."""
from __future__ import annotations

import argparse
import os
import sys
import time

from . import Solver, read_deck
from .results import write_step, write_model


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m cpfft_b200")
    ap.add_argument("deck")
    ap.add_argument("--outdir", default=".")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--steps", type=int, default=0, help="run only the first STEPS load steps")
    args = ap.parse_args(argv)
    os.makedirs(args.outdir, exist_ok=True)
    prob = read_deck(args.deck)
    nstep = args.steps or prob.nstep
    print(f" >> deck {args.deck}: grid {prob.N}^3, {len(prob.materials)} material(s), {nstep} load step(s)")
    if prob.model_file:                      # `output model "<file>"`: the mesh the nodal results refer to
        mf = write_model(os.path.join(args.outdir, os.path.basename(prob.model_file)), prob.N, prob.lengths, prob.name)
        print(f" >>>>>> Model description file: {mf}")
    s = Solver(prob, device=args.device)
    t0 = time.time()
    s.drive_eps_sig(1, 0)
    # the reference's three wall-time buckets (thyme.f:15-47): 1 pcg, 2 sig-eps, 3 patran output
    times = [[0.0, 0], [0.0, 0], [0.0, 0]]
    for step in range(1, nstep + 1):
        r = s.FFT_nr3(nstep=1, first=step - 1)
        sys.stdout.write(r["log"])
        times[0][0] += float(r["buckets"][0]); times[0][1] += len(r["cg_iters"][0])
        times[1][0] += float(r["buckets"][1]); times[1][1] += int(r["counters"][1])
        if step in prob.out_steps:
            t_out = time.time()
            write_step(args.outdir, step, s.download("URCS_N1", 1), s.download("EPS_N1", 1), prob.name, prob.N,
                       Fn1=s.download("FN1"), lengths=prob.lengths)
            times[2][0] += time.time() - t_out; times[2][1] += 1
            print(f"       results of step {step} written to {args.outdir}")
    fails = s.material_failures()[0]
    print(f"\n >> analysis done: {nstep} steps, {time.time() - t0:.2f} s, mm10 local failures {fails}")
    print_timings(times, time.time() - t0)
    return 0


def print_timings(times, total):
    """the exit summary of outime (outime.f:29-49, formats 9000-9020)"""
    print("\n\n\n\n >>>>>  solution timings   <<<<<")
    labels = ("pcg solution vector update:   ", "sig-eps & internal force:     ", "patran output:                ")
    for label, (secs, calls) in zip(labels, times):
        if calls < 0.01:
            continue
        print(f"\n\n  calculations for {label}")
        print(f"\n   wall time (secs): {secs:10.4f}{100.0 * secs / max(total, 1e-30):5.1f} (%) no. calls: {calls:8d}")


if __name__ == "__main__":
    sys.exit(main())
