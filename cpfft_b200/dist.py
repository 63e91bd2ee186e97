"""Host-side plumbing of the slab-decomposed (multi-GPU) run: one process per GPU, launched by
torchrun; torch.distributed is used only for the rendezvous (broadcast of the NCCL unique id
that the C library needs) -- the data path (all-to-all transposes, all-reduced norms) lives in
libcpfft_b200.so.  The reference has no distributed path (mpi_code.f is all stubs)."""
from __future__ import annotations

import numpy as np


def slab_range(N: int, rank: int, world: int):
    """x-planes [x0, x1) owned by ``rank``: equal slabs, N must be divisible by world
    (cpfft_create enforces the same rule)."""
    if N % world != 0:
        raise ValueError(f"slab decomposition needs N ({N}) divisible by the number of ranks ({world})")
    nx = N // world
    return rank * nx, (rank + 1) * nx


def broadcast_nccl_id(Solver, rank: int, device=None) -> bytes:
    """rank 0 creates the 128-byte NCCL unique id; everybody receives it.  Works on any
    torch.distributed backend (gloo in the CPU tests, nccl on the GPU box)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = torch.tensor(list(Solver.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def gather_slabs(local: np.ndarray, axis: int = 1):
    """all-gather per-rank (ncomp, n3loc) slabs into the global field (test helper)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=axis).cpu().numpy()
