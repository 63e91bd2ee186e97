#!/usr/bin/env python
"""bench.py -- voxel-iterations/s of the CPFFT hot path on B200.

A "step" is one load step of FFT_nr3 (FFT_nr3.f:51-185) on the synthetic Voronoi polycrystal of
BASELINE.json (1000 random-orientation fcc grains, mm10 / Voce) under the SURVEY.md 8d loading:
finite-strain uniaxial tension, F_xx driven at 0.1 % per step with P_yy = P_zz = 0 -- the
stress-BC loop (FFT_nr3.f:127-165) around the Newton loop, one drive_eps_sig sweep and one CG
solve per Newton iteration, tangent_homo (9 CG solves) per stress iteration.  `--strain-bc` is the
pure-strain variant for stage timing.  metric = voxels x G_K_dF applications / second (SURVEY.md
8d "VI/s": one iteration = one Green-operator application with the material update amortised).

Every timed step goes through the C ABI with HOST buffers: pinned F is uploaded, cpfft_FFT_nr3 runs
the load step, F and P are downloaded.  `value` sums the device time of the K steps with the inputs
already resident (event after the upload -> event before the download); `e2e` is the time of the
same K steps including the copies.  Both are the max over ranks.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # CPU oracle (port of the reference) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel-iterations/sec (mm10 update + FFT/Green step)"
UNIT = "voxel-iterations/s"
GRID_FOR_GPUS = {1: 256, 2: 320, 4: 400, 8: 512}   # ~16.8 M voxels per GPU (weak scaling)
CPU_SAMPLE_N = 64                                   # bounded CPU sample of the same workload: 64^3 is out of the host LLC
CPU_SAMPLE_APPLIES = 40                             # G_K_dF applications (CG iterations) per reference "step"
CPU_SWEEP_EVERY = 5                                 # one drive_eps_sig sweep per 5 x 40 = 200 applications: the GPU workload's mix
PARITY_FIXTURE = os.path.join(ROOT, "tests", "golden", "poly32_strain.npz")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """The reference's CPU path: no Fortran compiler / MKL exists in this image, so this is the
    C++/OpenMP oracle port (oracle/), all host threads, on a BOUNDED sample of the workload.

    The GPU workload at 256^3 costs the oracle hours per load step, so a reference "step" is a slice
    of one: CPU_SAMPLE_APPLIES iterations of a real CG solve (G_K_dF with the K4 contraction + the CG
    vector work, FFT_nr3.f:214-360) on the 64^3 polycrystal (same generator, grains scaled with the
    volume, out of the LLC) in the plastic regime, and one drive_eps_sig sweep every CPU_SWEEP_EVERY
    steps -- 1 sweep per 200 applications, the mix the stress-BC load steps of the GPU arm have."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ctypes as C
    from oracle import Oracle
    from cpfft_b200.polycrystal import polycrystal
    N = args.cpu_n
    ngr = max(8, int(round(args.grains * (N / 256.0) ** 3)))
    prob = polycrystal(N, ngrains=ngr, stress_bc=bool(args.stress_bc))
    cores = os.cpu_count() or 1
    o = Oracle(prob, threads=cores, polar="double")      # the literal double arithmetic of polar.f, as the reference runs it
    dp = C.POINTER(C.c_double)
    n3 = N ** 3
    # bring the sample into the plastic regime without paying for converged load steps: two 0.2 % increments
    # of the macroscopic stretch, each followed by the plastic sweep and a committed state (drive_eps_sig +
    # update, FFT_nr3.f:107-123, 173-179), then the 0.1 % increment of a benchmark load step
    o.drive_eps_sig(1, 0)
    Fbar = np.zeros((9, 1))
    step = 0
    for inc in (0.002, 0.002, 0.001):
        step += 1
        Fbar[0] += inc; Fbar[4] -= 0.3 * inc; Fbar[8] -= 0.3 * inc
        o.Fn1[:] = np.eye(3).reshape(9, 1) + Fbar
        if o.drive_eps_sig(step, 1):
            raise SystemExit("oracle sweep failed while preparing the CPU sample")
        if inc != 0.001:
            o.Fn[:] = o.Fn1
            o.update()
    b = -o.G_K_dF(o.Pn1, 0)                      # right-hand side of the Newton correction (FFT_nr3.f:110-111)
    x = np.zeros_like(b)
    W, K = args.warmup, args.steps
    cnt, sec = np.zeros(3, dtype=np.int64), np.zeros(2)

    def sample(k):
        t0 = time.perf_counter()
        if k % CPU_SWEEP_EVERY == 0:
            o.drive_eps_sig(step, 1)
        it, rr = C.c_int32(0), C.c_double(0)
        rc = o.L.orc_fftPcg_capped(o.h, b.ctypes.data_as(dp), x.ctypes.data_as(dp), 1e-14, CPU_SAMPLE_APPLIES, C.byref(it), C.byref(rr))
        assert rc == 0 and it.value == CPU_SAMPLE_APPLIES, (rc, it.value)
        return time.perf_counter() - t0, it.value

    for k in range(W):
        sample(k)
    secs, applies = 0.0, 0
    for k in range(K):
        dt, it = sample(W + k)
        secs += dt; applies += it
    value = n3 * applies / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {N}^3 Voronoi polycrystal ({ngr} random-orientation fcc grains, mm10/Voce), plastic regime: "
                               f"bounded CPU sample of the {GRID_FOR_GPUS.get(args.gpus, 256)}^3 GPU workload",
                   "grid": N, "applies": applies, "step": f"{CPU_SAMPLE_APPLIES} CG iterations (G_K_dF + CG vector work) "
                   f"+ one drive_eps_sig sweep every {CPU_SWEEP_EVERY} steps"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{K} x {CPU_SAMPLE_APPLIES} CG iterations + {len([k for k in range(W, W + K) if k % CPU_SWEEP_EVERY == 0])} "
                                   f"drive_eps_sig sweeps on a {N}^3 polycrystal at 0.5 % strain after {W} warm-up samples; "
                                   f"C++/OpenMP restatement (oracle/), not the ifort/MKL binary"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def ncu_profile_data():
    """per-launch DRAM traffic / FP64 counts measured by `ncu --set full` (tools/ncu_full.sh,
    tools/ncu_traffic.py), committed under profiles/; None when absent"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def algorithmic_bytes_per_voxel(cls, N, cg_frac=0.0, cg_fused_frac=0.0):
    """SURVEY.md 8d per-unit figures (FP64, half spectrum, Ghat and phases recomputed), plus the
    minimum traffic of the CG vector work that is fused into the z passes:
      forward z pass of CG iterations 2.. of a solve (share cg_fused_frac of its launches):
          p <- r + beta p : r read 72, p write 72;   x += alpha p : x read + write 144   (p is read anyway)
      inverse z pass inside CG (share cg_frac): p.Ap partial sums: p read 72 (Ap is in registers)."""
    half = 16.0 * (N // 2 + 1) / N          # complex half-spectrum bytes per voxel-component
    return {
        "k_fwd_z_K4": (81 + 9) * 8 + 9 * half + 288.0 * cg_fused_frac,
        "k_fwd_z": 9 * 8 + 9 * half,
        "k_fft_y": 2 * 9 * half,
        "k_x_green": 2 * 9 * half,
        "k_inv_z": 9 * half + 9 * 8 + 72.0 * cg_frac,
        "k_pk1_tangent": (18 + 6 + 36) * 8 + (9 + 81) * 8,
        "k_update_mm10": 8.0 * (18 + 9 + 6 + 30) + 8.0 * (9 + 6 + 9 + 36 + 36 + 80),
        "k_update_mm01": 8.0 * (18 + 11 + 9 + 6) + 8.0 * (9 + 6 + 9 + 11 + 36),
    }.get(cls)


def stage_rooflines(table, N, n3, cgits, nsolves, world, fp64_peak):
    """per-kernel-class table {ms, launches, share, algorithmic bytes, achieved GB/s, fraction of the
    measured HBM peak, ncu DRAM traffic / FP64 counts when profiles/ncu_traffic.json matches} and the
    `roofline` object of the kernel class that takes the most time.  table: {class: (ms, launches)}."""
    peak, which = measured_peaks()
    stages = {}
    tot_ms = sum(v[0] for v in table.values()) or 1.0
    for name, (kms, cnt) in table.items():
        if cnt == 0:
            continue
        # share of the launches of the z passes that carry fused CG work (counted, not assumed)
        cg_frac = min(1.0, cgits / cnt) if name == "k_inv_z" else 0.0
        cg_fused = min(1.0, max(0, cgits - nsolves) / cnt) if name == "k_fwd_z_K4" else 0.0
        b = algorithmic_bytes_per_voxel(name, N, cg_frac, cg_fused)
        ent = {"ms_total": kms, "launches": cnt, "share": kms / tot_ms, "ms_per_launch": kms / cnt}
        if b is not None:
            gbs = b * n3 / (kms / cnt * 1e-3) / 1e9
            ent.update({"alg_bytes_per_voxel": b, "achieved_gbs": gbs, "frac_of_hbm": gbs / peak})
        stages[name] = ent
    prof = ncu_profile_data()
    # ncu is a single-process tool: the captures are from the 256^3 1-GPU run.  The z passes, the CG vector kernels and the
    # material kernels touch rank-local data only, so their DRAM traffic per launch is that of the capture scaled by the
    # local voxel count (16.8 M per GPU at 256^3 x1 and 512^3 x8; 16.4 M / 16.0 M at 320^3 x2 / 400^3 x4).
    local_kernels = ("k_fwd_z_K4", "k_fwd_z", "k_inv_z", "k_update_mm10", "k_pk1_tangent")
    if prof and (prof.get("grid") == N or world > 1):
        scale = float(n3) / float(prof["grid"]) ** 3
        for name, ent in prof["kernels"].items():
            if name in stages:
                if "dram_bytes" not in ent:
                    pass
                elif world == 1:
                    stages[name]["ncu_dram_bytes_per_launch"] = ent["dram_bytes"]
                elif name in local_kernels:
                    stages[name]["ncu_dram_bytes_per_launch"] = ent["dram_bytes"] * scale
                    stages[name]["ncu_dram_bytes_source"] = f"1-GPU capture at {prof['grid']}^3 x local voxels / {prof['grid']}^3 (rank-local kernel)"
                if "fp64_flop" in ent and fp64_peak and world == 1:
                    # FP64 rate of the PROFILED launch: its own flop count over its own duration (the
                    # profiled k_update_mm10 launch is a plastic sweep; elastic iter-0 sweeps are a class of their own)
                    # (this run's launches of the class do other amounts of work -- the material sweeps differ in their
                    # Newton iteration counts -- so the capture's flop count is never divided by this run's times)
                    dur = ent.get("duration_us")
                    tf_ncu = ent["fp64_flop"] / (dur * 1e-6) / 1e12 if dur else None
                    stages[name].update({"fp64_flop_per_launch_ncu": ent["fp64_flop"], "ncu_duration_us": dur,
                                         "fp64_tflops_ncu_launch": tf_ncu, "fp64_tflops": tf_ncu,
                                         "fp64_peak_tflops_measured": fp64_peak,
                                         "frac_of_fp64_ncu_launch": tf_ncu / fp64_peak if tf_ncu else None,
                                         "frac_of_fp64": tf_ncu / fp64_peak if tf_ncu else None})
    cand = [k for k in stages if "achieved_gbs" in stages[k]]
    dom = max(cand, key=lambda k: stages[k]["ms_total"]) if cand else None
    roof = None
    if dom:
        d = stages[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": d["achieved_gbs"] / peak, "traffic": d.get("ncu_dram_bytes_per_launch"),
                "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)",
                "alg_bytes_per_launch": d["alg_bytes_per_voxel"] * n3, "ms_per_launch": d["ms_per_launch"]}
        if dom == "k_fwd_z_K4":
            # the same launch against SURVEY.md 8d's operator-only figure (K4 contraction + forward z pass, no fused CG work)
            op = algorithmic_bytes_per_voxel(dom, N)
            roof["frac_operator_only"] = op * n3 / (d["ms_per_launch"] * 1e-3) / 1e9 / peak
            roof["alg_bytes_per_voxel"] = d["alg_bytes_per_voxel"]; roof["alg_bytes_per_voxel_operator_only"] = op
        # one whole CG iteration (G_K_dF with K4 + the CG vector work): its kernels' mean times against the bytes it has to move
        names = ("k_fwd_z_K4", "k_fft_y", "k_x_green", "k_inv_z", "vector_ops", "exchange")
        if cgits > 0 and all(k in stages for k in names[:4]):
            it_ms = (stages["k_fwd_z_K4"]["ms_per_launch"] + 2 * stages["k_fft_y"]["ms_per_launch"] + stages["k_x_green"]["ms_per_launch"] +
                     stages["k_inv_z"]["ms_per_launch"] + stages.get("vector_ops", {}).get("ms_total", 0.0) / cgits +
                     stages.get("exchange", {}).get("ms_total", 0.0) / cgits)
            roof["cg_iteration"] = {"ms": it_ms, "alg_bytes_per_voxel_fused": 1872.0, "frac_fused": 1872.0 * n3 / (it_ms * 1e-3) / 1e9 / peak,
                                    "alg_bytes_per_voxel_survey_8d": 1224.0, "frac_survey_8d": 1224.0 * n3 / (it_ms * 1e-3) / 1e9 / peak}
    return stages, roof


def parity_run(Solver, polycrystal, world, rank, local_rank, nccl_id, dist, torch):
    """The 32^3, 64-grain benchmark polycrystal over 8 load steps (>= 6 plastic) on all ranks of this run against the
    oracle's frozen output (tests/golden/poly32_strain.npz, tools/make_golden_poly.py): Newton iterations per step
    identical, macroscopic stress <= 1e-10, F and P at 1024 sampled voxels <= 1e-9 (north_star tolerances)."""
    if not os.path.exists(PARITY_FIXTURE):
        return {"ok": None, "skipped": "fixture missing"}
    g = np.load(PARITY_FIXTURE)
    N, nstep = int(g["N"]), int(g["nstep"])
    if N % world:
        return {"ok": None, "skipped": f"{N} not divisible by {world} ranks"}
    nx = N // world
    p = polycrystal(N, ngrains=int(g["grains"]), stress_bc=bool(g["stress_bc"]), x_range=(rank * nx, (rank + 1) * nx))
    s = Solver(p, device=local_rank, rank=rank, world=world, nccl_id=nccl_id, local_slab=True)
    s.drive_eps_sig(1, 0)
    r = s.FFT_nr3(nstep=nstep)
    F, P = s.download("FN1"), s.download("PN1")
    lo, hi = rank * s.n3, (rank + 1) * s.n3
    idx = g["idx"]
    mine = (idx >= lo) & (idx < hi)
    eF = np.abs(F[:, idx[mine] - lo] - g["F"][:, mine]).max() if mine.any() else 0.0
    eP = np.abs(P[:, idx[mine] - lo] - g["P"][:, mine]).max() if mine.any() else 0.0
    errs = torch.tensor([eF, eP], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    s.close()
    scale_P, scale_F = float(np.abs(g["P"]).max()), float(np.abs(g["F"]).max())
    pbar_err = float(np.abs(r["Pbar"] - g["Pbar"]).max() / np.abs(g["Pbar"]).max())
    cg_ref = [[int(v) for v in row if v >= 0] for row in g["cg_iters"]]
    cg_dev = max((abs(a - b) for ra, rb in zip(r["cg_iters"], cg_ref) for a, b in zip(ra, rb)), default=0)
    out = {"case": "poly32_strain: 32^3, 64 grains, 8 load steps, oracle fixture", "ranks": world,
           "newton_iters": [int(v) for v in r["nr_iters"]], "newton_iters_equal": [int(v) for v in r["nr_iters"]] == [int(v) for v in g["nr_iters"]],
           "cg_solves_equal": [len(a) for a in r["cg_iters"]] == [len(b) for b in cg_ref], "cg_iters_max_abs_diff": int(cg_dev),
           "Pbar_rel_err": pbar_err, "F_rel_err_sampled": float(errs[0].item()) / scale_F, "P_rel_err_sampled": float(errs[1].item()) / scale_P,
           "tolerances": {"Pbar": 1e-10, "F": 1e-9, "P": 1e-9}}
    out["ok"] = bool(out["newton_iters_equal"] and out["cg_solves_equal"] and pbar_err <= 1e-10 and
                     out["F_rel_err_sampled"] <= 1e-9 and out["P_rel_err_sampled"] <= 1e-9)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cpfft_b200", choices=["cpfft_b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="override the grid edge N")
    ap.add_argument("--grains", type=int, default=1000)
    ap.add_argument("--cpu-n", type=int, default=CPU_SAMPLE_N)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="do not record per-kernel CUDA events inside the timed region")
    ap.add_argument("--variant", default="voce", choices=["voce", "mts", "taylor2", "taylor4", "bcc48"],
                    help="workload variant for kernel measurements (NOT the BASELINE.json metric unless 'voce'): "
                         "MTS hardening law, 2 / 4 crystals per material point (Taylor average), or 48-system bcc grains")
    ap.add_argument("--stress-bc", action="store_true",
                    help="time the K steps under the SURVEY.md 8d loading (F_xx driven, P_yy = P_zz = 0: stress-BC loop + "
                         "tangent_homo, ~4000 G_K_dF applications per load step) instead of its pure-strain variant")
    ap.add_argument("--strain-bc", action="store_true", help="(default for the K timed steps; kept for older command lines)")
    ap.add_argument("--stress-leg-steps", type=int, default=1,
                    help="load steps of the bounded stress-BC leg that precedes the K timed steps (0 = none)")
    ap.add_argument("--stress-leg-warmup", type=int, default=2, help="untimed stress-BC load steps before them (step 1 is elastic)")
    ap.add_argument("--no-parity", action="store_true", help="skip the 32^3 parity run against tests/golden/poly32_strain.npz")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from cpfft_b200 import Solver
    from cpfft_b200.polycrystal import polycrystal

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    from cpfft_b200.api import library_path
    if not os.path.exists(library_path()):       # tree without built artefacts: compile first (rank 0), nvcc is in the image
        if local_rank == 0:
            from cpfft_b200.build import build as build_cuda
            build_cuda()
        for _ in range(3000):                    # the other ranks wait for the file
            if os.path.exists(library_path()):
                break
            time.sleep(0.1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_nccl_id():
        """a fresh NCCL unique id for one Solver (one communicator), created on rank 0 and broadcast"""
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(Solver.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    N = args.grid or GRID_FOR_GPUS.get(world, 256)
    W, K = max(args.warmup, 0), max(args.steps, 1)
    stress_bc = bool(args.stress_bc)
    nx = N // world

    # ---- N-GPU correctness before anything is timed: the 32^3 polycrystal of tests/golden (8 load steps, oracle
    # output frozen by tools/make_golden_poly.py) on ALL ranks of this run, against the committed fixture ----
    parity = None
    if not args.no_parity:
        parity = parity_run(Solver, polycrystal, world, rank, local_rank, new_nccl_id(), dist if world > 1 else None, torch)

    Event = torch.cuda.Event
    nvox = float(N) ** 3

    # load increment: 1 % strain over 10 steps; the MTS variant's parameter set needs quarter-size steps -- with 0.1 % its
    # global Newton loop does not converge on grids >= 64^3, in the oracle either (profiles/r02z_mtsdiag.log)
    steps_per_percent = 40 if args.variant == "mts" else 10

    def make_solver(sbc, nsteps):
        prob = polycrystal(N, ngrains=args.grains, nstep=max(steps_per_percent, nsteps + 2), x_range=(rank * nx, (rank + 1) * nx), stress_bc=sbc)
        if args.variant != "voce":
            from cpfft_b200.polycrystal import workload_variant
            prob = workload_variant(prob, args.variant, args.grains)
        return Solver(prob, device=local_rank, rank=rank, world=world, nccl_id=new_nccl_id(), local_slab=True)

    def barrier(s):
        s.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(s, hF, hP, Wl, Kl, sample_clocks):
        """Wl untimed + Kl timed load steps of solver `s` through the C ABI with the host buffers hF / hP.
        Returns device seconds with resident inputs, seconds including the copies (both max over ranks) and counters."""
        stream = torch.cuda.ExternalStream(s.stream())
        s.drive_eps_sig(1, 0)                       # FFT_finite_3d.f:145
        step0 = 0
        for _ in range(Wl):                         # untimed warm-up load steps
            s.FFT_nr3(nstep=1, first=step0); step0 += 1
        s.download_ptr("FN1", hF.data_ptr())
        s.profile(not args.no_profile); s.profile_reset()
        launches0 = s.kernel_launches()
        clocks = ClockSampler(local_rank) if sample_clocks else None
        barrier(s)
        if clocks and rank == 0:
            clocks.start()
        ev0, ev1 = Event(enable_timing=True), Event(enable_timing=True)
        marks = []
        c = dict(applies=0, sweeps=0, cgits=0, nfail=0, nfail_final=0, nsolves=0, t_pcg=0.0, t_sig=0.0, nr_hist=[])
        ev0.record(stream)
        for _ in range(Kl):
            a, b = Event(enable_timing=True), Event(enable_timing=True)
            s.upload_ptr("FN1", hF.data_ptr())                 # host -> device: the step's deformation field
            a.record(stream)                                   # inputs resident in HBM: `value` starts here
            r = s.FFT_nr3(nstep=1, first=step0); step0 += 1    # one load step through the ABI
            b.record(stream)                                   # ... and stops here
            s.download_ptr("FN1", hF.data_ptr())               # device -> host: F and P of the step
            s.download_ptr("PN1", hP.data_ptr())
            marks.append((a, b))
            c["applies"] += int(r["counters"][0]); c["sweeps"] += int(r["counters"][1]); c["cgits"] += int(r["counters"][2])
            c["nfail"] += int(r["counters"][3]); c["nfail_final"] += int(r["counters"][4])
            c["t_pcg"] += float(r["buckets"][0]); c["t_sig"] += float(r["buckets"][1])
            c["nr_hist"].append(int(r["nr_iters"][0]))
            c["nsolves"] += sum(len(row) for row in r["cg_iters"])
        ev1.record(stream)
        barrier(s)
        c["clocks"] = clocks.stop() if (clocks and rank == 0) else None
        dev_ms = sum(a.elapsed_time(b) for a, b in marks)
        ms = torch.tensor([dev_ms, ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        c["secs"], c["secs_e2e"] = float(ms[0].item()) * 1e-3, float(ms[1].item()) * 1e-3
        c["launches"] = s.kernel_launches() - launches0
        c["table"] = s.profile_table()
        s.profile(False)
        return c

    n3 = nx * N * N
    hF = torch.empty(9 * n3, dtype=torch.float64).pin_memory()     # host buffers of the caller (pinned)
    hP = torch.empty(9 * n3, dtype=torch.float64).pin_memory()

    # ---- bounded leg under the SURVEY.md 8d loading (F_xx driven, P_yy = P_zz = 0): ~4000 G_K_dF applications per load
    # step -- 25 s per step at 256^3 -- so the K timed steps of the driver (K = 20, W = 5, 870 s per N in the scaling run)
    # cannot all be stress-BC steps; this leg measures the same metric on `--stress-leg-steps` of them, every run ----
    stress_leg = None
    if not stress_bc and args.stress_leg_steps > 0:
        sl = make_solver(True, args.stress_leg_warmup + args.stress_leg_steps)
        c = timed_steps(sl, hF, hP, args.stress_leg_warmup, args.stress_leg_steps, False)
        sl.close(); del sl
        torch.cuda.empty_cache()
        stress_leg = {"loading": "F_xx driven with P_yy = P_zz = 0 (stress-BC loop FFT_nr3.f:127-165 + tangent_homo), 0.1 % per load step",
                      "value": nvox * c["applies"] / c["secs"], "e2e_value": nvox * c["applies"] / c["secs_e2e"], "unit": UNIT,
                      "steps": args.stress_leg_steps, "warmup": args.stress_leg_warmup, "ms_per_step": 1e3 * c["secs"] / args.stress_leg_steps,
                      "newton_iters_per_step": c["nr_hist"], "G_K_dF_applies": c["applies"], "drive_eps_sig_sweeps": c["sweeps"],
                      "VG_per_s": nvox * c["applies"] / c["t_pcg"] if c["t_pcg"] > 0 else None,
                      "VU_per_s": nvox * c["sweeps"] / c["t_sig"] if c["t_sig"] > 0 else None,
                      "mm10_local_failures": {"all_sweeps": c["nfail"], "final_sweeps": c["nfail_final"]}}

    s = make_solver(stress_bc, W + K)
    assert s.n3 == n3
    c = timed_steps(s, hF, hP, W, K, True)
    applies, sweeps, cgits, nsolves = c["applies"], c["sweeps"], c["cgits"], c["nsolves"]
    nfail, nfail_final, t_pcg, t_sig, nr_hist = c["nfail"], c["nfail_final"], c["t_pcg"], c["t_sig"], c["nr_hist"]
    secs, secs_e2e, launches, table, clk = c["secs"], c["secs_e2e"], c["launches"], c["table"], c["clocks"]
    value = nvox * applies / secs
    fp64_peak = s.fp64_peak()
    e2e = None
    if not args.no_e2e:
        e2e = {"value": nvox * applies / secs_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(9 * n3 * 8) * world, "d2h_bytes_per_step": int(18 * n3 * 8) * world,   # all ranks
               "steps": K, "ms_per_step": 1e3 * secs_e2e / K,
               "what": "the same K load steps as `value`, each: cpfft_upload(FN1) from pinned host memory -> cpfft_FFT_nr3 -> "
                       "cpfft_download(FN1, PN1) to pinned host memory; `value` leaves the copies out",
               "check": float(hP.abs().max())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events on the launching stream) ----
    stages, roof = stage_rooflines(table, N, s.n3, cgits, nsolves, world, fp64_peak)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5",
                                  "--warmup", "1", "--cpu-n", str(args.cpu_n), "--grains", str(args.grains)] +
                                 (["--stress-bc"] if args.stress_bc else []), capture_output=True, text=True, timeout=900)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:  # reported, never hidden
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {N}^3 Voronoi polycrystal ({args.grains} random-orientation fcc grains, "
                               f"mm10/{ {'voce': 'Voce', 'mts': 'MTS (variant, not the BASELINE metric)', 'bcc48': 'Voce, bcc48 slip family (variant, not the BASELINE metric)'}.get(args.variant, 'Voce, ' + args.variant[-1] + ' crystals per point (variant, not the BASELINE metric)') }), "
                               "finite-strain uniaxial tension, " +
                               ("F_xx driven with P_yy = P_zz = 0 (stress-BC loop + tangent_homo)" if stress_bc else
                                "strain-controlled variant of SURVEY.md 8d for the K timed steps (F_yy = F_zz = -0.3 F_xx); the stress-BC loading of 8d "
                                "is measured by the bounded leg `stress_bc_leg` of this line") + f", {1.0 / steps_per_percent:g} % per load step",
                   "grid": N, "voxels": int(nvox), "voxels_per_gpu": int(s.n3), "parallelism": f"x-slabs x{world}",
                   "l2": "working set (>= 1.2 GB per field) far exceeds the 126 MB L2; no flush needed",
                   "step": "one FFT_nr3 load step", "newton_iters_per_step": nr_hist,
                   "G_K_dF_applies": applies, "drive_eps_sig_sweeps": sweeps, "cg_iterations": cgits, "cg_solves": nsolves,
                   "newton_normalised_voxel_updates_per_s": nvox * sweeps / secs,
                   # VG/s = voxels x G_K_dF applications / bucket 1 (pcg), VU/s = voxels x
                   # drive_eps_sig sweeps / bucket 2 (sig-eps), the reference's own thyme() buckets
                   "VG_per_s": nvox * applies / t_pcg if t_pcg > 0 else None,
                   "VU_per_s": nvox * sweeps / t_sig if t_sig > 0 else None,
                   "mm10_local_failures": {"all_sweeps": nfail, "final_sweeps": nfail_final},
                   "exchange_mode": s.exchange_mode(), "fp64_peak_tflops_measured": fp64_peak,
                   "weak_scaling_note": "256^3 / 320^3 / 400^3 / 512^3 on 1 / 2 / 4 / 8 GPUs (~16.8 M voxels and ~80 GB per GPU): "
                                        "512^3 needs >= 4 GPUs (4.8 KB of state per voxel), so strong scaling of it over 2 GPUs is not possible",
                   "even_N_convention": "Nyquist planes of Ghat zeroed (reference is only valid for odd N)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "stages": stages,
        "cpu_baseline": cpu, "parity": parity, "stress_bc_leg": stress_leg,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
