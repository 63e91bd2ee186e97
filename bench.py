#!/usr/bin/env python
"""bench.py -- voxel-iterations/s of the CPFFT hot path on B200.

A "step" is one load step of FFT_nr3 (FFT_nr3.f:51-185) on the synthetic Voronoi polycrystal of
BASELINE.json (1000 random-orientation fcc grains, mm10 / Voce): Newton loop, one
drive_eps_sig sweep per Newton iteration, one CG solve (tens of G_K_dF applications) per
Newton iteration.  metric = voxels x G_K_dF applications / second (SURVEY.md 8d "VI/s": one
iteration = one Green-operator application with the material update amortised).

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
CPU_SAMPLE_N = 32                                   # bounded CPU sample of the same workload


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
    C++/OpenMP oracle port (oracle/), all host threads, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import Oracle
    from cpfft_b200.polycrystal import polycrystal
    N = args.cpu_n
    prob = polycrystal(N, ngrains=max(8, int(1000 * (N / 256.0) ** 3)))
    cores = os.cpu_count() or 1
    o = Oracle(prob, threads=cores)
    o.drive_eps_sig(1, 0)
    W, K = args.warmup, args.steps
    o.FFT_nr3(nstep=W)        # warm-up load steps (also brings the sample into the plastic regime)
    applies, secs = 0, 0.0
    bc = prob.BC_all()
    for k in range(K):        # K further steps, state persists inside the oracle model
        step_bc = np.ascontiguousarray(bc[W + k:W + k + 1])
        res = _oracle_steps(o, step_bc, W + k + 1)
        applies += int(res["counters"][0]); secs += float(res["buckets"][2])
    value = N ** 3 * applies / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {N}^3 Voronoi polycrystal, fcc mm10/Voce, pure-strain uniaxial "
                               f"(bounded CPU sample of the {GRID_FOR_GPUS.get(args.gpus, 256)}^3 GPU workload)",
                   "grid": N, "applies": applies},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{K} plastic load steps of a {N}^3 polycrystal after {W} warm-up steps; "
                                   f"C++/OpenMP restatement (oracle/), not the ifort/MKL binary"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _oracle_steps(o, bc_rows, first_step):
    """run load steps with explicit BC rows, continuing the oracle's committed state."""
    import ctypes as C
    n = len(bc_rows)
    nbc = np.ascontiguousarray(o.prob.isNBC, dtype=np.int32)
    nr = np.zeros(n, dtype=np.int32); cg = np.full((n, 64), -1, dtype=np.int32)
    pb = np.zeros((n, 9)); bk = np.zeros(3); cnt = np.zeros(5, dtype=np.int64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    rc = o.L.orc_FFT_nr3_from(o.h, first_step, n, bc_rows.ctypes.data_as(dp), nbc.ctypes.data_as(ip),
                              nr.ctypes.data_as(ip), cg.ctypes.data_as(ip), 64, pb.ctypes.data_as(dp),
                              bk.ctypes.data_as(dp), cnt.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0, rc
    return {"nr_iters": nr, "buckets": bk, "counters": cnt, "Pbar": pb}


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
    if prof and prof.get("grid") == N and world == 1:
        for name, ent in prof["kernels"].items():
            if name in stages:
                stages[name]["ncu_dram_bytes_per_launch"] = ent["dram_bytes"]
                if "fp64_flop" in ent and fp64_peak:
                    tf = ent["fp64_flop"] / (stages[name]["ms_per_launch"] * 1e-3) / 1e12
                    stages[name].update({"fp64_flop_per_launch_ncu": ent["fp64_flop"], "fp64_tflops": tf,
                                         "fp64_peak_tflops_measured": fp64_peak, "frac_of_fp64": tf / fp64_peak})
    cand = [k for k in stages if "achieved_gbs" in stages[k]]
    dom = max(cand, key=lambda k: stages[k]["ms_total"]) if cand else None
    roof = None
    if dom:
        d = stages[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": d["achieved_gbs"] / peak, "traffic": d.get("ncu_dram_bytes_per_launch"),
                "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)",
                "alg_bytes_per_launch": d["alg_bytes_per_voxel"] * n3, "ms_per_launch": d["ms_per_launch"]}
    return stages, roof


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
    ap.add_argument("--variant", default="voce", choices=["voce", "mts", "taylor2", "taylor4"],
                    help="workload variant for kernel measurements (NOT the BASELINE.json metric unless 'voce'): "
                         "MTS hardening law, or 2 / 4 crystals per material point (Taylor average)")
    ap.add_argument("--stress-bc", action="store_true",
                    help="uniaxial tension with P_yy = P_zz = 0 (stress-BC loop, tangent_homo) instead of pure strain control")
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
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(Solver.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    N = args.grid or GRID_FOR_GPUS.get(world, 256)
    W, K = max(args.warmup, 0), max(args.steps, 1)
    nx = N // world
    prob = polycrystal(N, ngrains=args.grains, nstep=max(10, W + 2 * K + 2), x_range=(rank * nx, (rank + 1) * nx),
                       stress_bc=args.stress_bc)
    if args.variant != "voce":
        from cpfft_b200.polycrystal import workload_variant
        prob = workload_variant(prob, args.variant, args.grains)
    s = Solver(prob, device=local_rank, rank=rank, world=world, nccl_id=nccl_id, local_slab=True)
    stream = torch.cuda.ExternalStream(s.stream())

    def barrier():
        s.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s.drive_eps_sig(1, 0)                       # FFT_finite_3d.f:145
    step0 = 0
    for _ in range(W):                          # untimed warm-up load steps
        s.FFT_nr3(nstep=1, first=step0); step0 += 1
    s.profile(True); s.profile_reset()
    launches0 = s.kernel_launches()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    applies = sweeps = cgits = nfail = nfail_final = nsolves = 0
    t_pcg = t_sig = 0.0
    nr_hist = []
    for _ in range(K):
        r = s.FFT_nr3(nstep=1, first=step0); step0 += 1
        applies += int(r["counters"][0]); sweeps += int(r["counters"][1]); cgits += int(r["counters"][2])
        nfail += int(r["counters"][3]); nfail_final += int(r["counters"][4])
        t_pcg += float(r["buckets"][0]); t_sig += float(r["buckets"][1])
        nr_hist.append(int(r["nr_iters"][0]))
        nsolves += sum(len(row) for row in r["cg_iters"])
    ev1.record(stream)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    secs = float(ms.item()) * 1e-3
    launches = s.kernel_launches() - launches0
    table = s.profile_table()
    s.profile(False)
    nvox = float(N) ** 3
    value = nvox * applies / secs
    fp64_peak = s.fp64_peak()

    # ---- end-to-end through the public C ABI with HOST buffers (pinned) ----
    e2e = None
    if not args.no_e2e:
        n3 = s.n3
        hF = torch.empty(9 * n3, dtype=torch.float64).pin_memory()
        hP = torch.empty(9 * n3, dtype=torch.float64).pin_memory()
        s.download_ptr("FN1", hF.data_ptr())
        Ke = max(1, min(K, 2))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        app_e = 0
        for _ in range(Ke):
            s.upload_ptr("FN1", hF.data_ptr())                 # host -> device: current deformation field
            r = s.FFT_nr3(nstep=1, first=step0); step0 += 1    # one load step through the ABI
            s.download_ptr("FN1", hF.data_ptr())               # device -> host: F and P of the step
            s.download_ptr("PN1", hP.data_ptr())
            app_e += int(r["counters"][0])
        e1.record(stream)
        barrier()
        mse = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(mse, op=dist.ReduceOp.MAX)
        e2e = {"value": nvox * app_e / (float(mse.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(9 * n3 * 8) * world, "d2h_bytes_per_step": int(18 * n3 * 8) * world,   # all ranks
               "steps": Ke,
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
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2",
                                  "--warmup", "3", "--cpu-n", str(args.cpu_n)], capture_output=True, text=True, timeout=900)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:  # reported, never hidden
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {N}^3 Voronoi polycrystal ({args.grains} random-orientation fcc grains, "
                               f"mm10/{ {'voce': 'Voce', 'mts': 'MTS (variant, not the BASELINE metric)'}.get(args.variant, 'Voce, ' + args.variant[-1] + ' crystals per point (variant, not the BASELINE metric)') }), "
                               "finite-strain uniaxial tension, " +
                               ("F_xx driven with P_yy = P_zz = 0" if args.stress_bc else "strain-controlled") + ", 0.1 % per load step",
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
                   "even_N_convention": "Nyquist planes of Ghat zeroed (reference is only valid for odd N)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "stages": stages,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
