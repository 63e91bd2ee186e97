#!/usr/bin/env python
"""Multi-GPU self-check (run under torchrun on an N-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py [--grid 32]

Every rank owns an x-slab; the slab-decomposed run (all-to-all transposes of the half spectrum,
all-reduced norms) must reproduce the single-GPU run of the same problem: identical Newton and
CG iteration counts, fields equal to round-off.  Rank 0 prints one JSON line and exits non-zero
on a mismatch."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--grains", type=int, default=40)
    ap.add_argument("--variant", default="voce", choices=["voce", "mts", "taylor2", "taylor4"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cpfft_b200 import Solver
    from cpfft_b200.polycrystal import polycrystal, workload_variant
    from cpfft_b200.dist import broadcast_nccl_id, slab_range

    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nccl_id = broadcast_nccl_id(Solver, rank)
    N = args.grid
    x0, x1 = slab_range(N, rank, world)
    p_loc = polycrystal(N, ngrains=args.grains, x_range=(x0, x1))
    if args.variant != "voce":
        p_loc = workload_variant(p_loc, args.variant, args.grains)
    s = Solver(p_loc, device=lr, rank=rank, world=world, nccl_id=nccl_id, local_slab=True)
    s.drive_eps_sig(1, 0)
    r = s.FFT_nr3(nstep=args.steps)
    P = torch.from_numpy(s.download("PN1")).cuda()
    F = torch.from_numpy(s.download("FN1")).cuda()
    Pl = [torch.empty_like(P) for _ in range(world)]; Fl = [torch.empty_like(F) for _ in range(world)]
    dist.all_gather(Pl, P); dist.all_gather(Fl, F)
    ok, info = True, {}
    if rank == 0:
        Pm = torch.cat(Pl, dim=1).cpu().numpy(); Fm = torch.cat(Fl, dim=1).cpu().numpy()
        p = polycrystal(N, ngrains=args.grains)
        if args.variant != "voce":
            p = workload_variant(p, args.variant, args.grains)
        s1 = Solver(p, device=lr)
        s1.drive_eps_sig(1, 0)
        r1 = s1.FFT_nr3(nstep=args.steps)
        P1, F1 = s1.download("PN1"), s1.download("FN1")
        eP = float(np.abs(Pm - P1).max() / np.abs(P1).max()); eF = float(np.abs(Fm - F1).max() / np.abs(F1).max())
        ePb = float(np.abs(r["Pbar"] - r1["Pbar"]).max() / np.abs(r1["Pbar"]).max())
        same_nr = list(map(int, r["nr_iters"])) == list(map(int, r1["nr_iters"]))
        same_cg = [[int(v) for v in row] for row in r["cg_iters"]] == [[int(v) for v in row] for row in r1["cg_iters"]]
        ok = same_nr and same_cg and eP <= 1e-9 and eF <= 1e-9 and ePb <= 1e-10
        info = {"world": world, "grid": N, "exchange_mode": s.exchange_mode(), "nr_iters": list(map(int, r["nr_iters"])), "same_newton_counts": same_nr,
                "same_cg_counts": same_cg, "relerr_P": eP, "relerr_F": eF, "relerr_Pbar": ePb, "ok": ok,
                "applies": int(r["counters"][0])}
        print(json.dumps(info), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
