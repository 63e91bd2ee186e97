"""Host-side logic of the multi-GPU (slab-decomposed) run on CPU: two processes, gloo backend.
The data path itself (NCCL transposes inside libcpfft_b200.so) needs GPUs and is covered by
tools/multi_gpu_check.py on the B200 box; here: slab ownership, per-rank generation of the
synthetic polycrystal (each rank builds only its x-planes), the unique-id broadcast plumbing
and the gather helper."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeSolver:
    @staticmethod
    def nccl_unique_id():
        return bytes(range(128))


def _worker(rank, world, port, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpfft_b200.dist import slab_range, broadcast_nccl_id, gather_slabs
    from cpfft_b200.polycrystal import polycrystal
    x0, x1 = slab_range(N, rank, world)
    p = polycrystal(N, ngrains=30, x_range=(x0, x1))
    assert len(p.matlist) == (x1 - x0) * N * N
    ident = broadcast_nccl_id(_FakeSolver, rank)
    assert ident == bytes(range(128))
    ang = gather_slabs(np.ascontiguousarray(p.angles.T))          # (3, n3loc) -> (3, N^3)
    if rank == 0:
        np.save(os.path.join(out_dir, "angles.npy"), ang)
    # polycrystalline material points: the per-rank Taylor tables are slabs of the global ones
    from cpfft_b200.polycrystal import workload_variant
    pt = workload_variant(polycrystal(N, ngrains=30, x_range=(x0, x1)), "taylor2", 30)
    nc, ang_t, ids = pt.taylor_tables()
    assert nc == 2 and ang_t.shape == ((x1 - x0) * N * N, 2, 3) and ids is None and pt.taylor
    tay = gather_slabs(np.ascontiguousarray(ang_t.reshape(len(ang_t), 6).T))      # (6, n3loc) -> (6, N^3)
    if rank == 0:
        np.save(os.path.join(out_dir, "taylor.npy"), tay)
    # the reductions of the solver are sums over slabs: emulate P_bar
    local = torch.tensor([float(p.angles[:, 0].sum())], dtype=torch.float64)
    dist.all_reduce(local)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), local.numpy())
    dist.destroy_process_group()


def test_two_rank_slab_decomposition(tmp_path):
    N, world = 8, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    from cpfft_b200.polycrystal import polycrystal
    full = polycrystal(N, ngrains=30)
    ang = np.load(tmp_path / "angles.npy")
    assert np.array_equal(ang, full.angles.T)
    assert np.isclose(np.load(tmp_path / "sum.npy")[0], full.angles[:, 0].sum(), rtol=1e-14)
    from cpfft_b200.polycrystal import workload_variant
    full_t = workload_variant(polycrystal(N, ngrains=30), "taylor2", 30)
    assert np.array_equal(np.load(tmp_path / "taylor.npy"), np.asarray(full_t.angles).reshape(N ** 3, 6).T)


def test_slab_range_rules():
    from cpfft_b200.dist import slab_range
    assert slab_range(512, 3, 8) == (192, 256)
    assert [slab_range(256, r, 4) for r in range(4)] == [(0, 64), (64, 128), (128, 192), (192, 256)]
    with pytest.raises(ValueError):
        slab_range(255, 0, 2)
