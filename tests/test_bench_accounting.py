"""bench.py's roofline post-processing on the recorded per-kernel times of the round-1 run
(profiles/r01g_bench256_1gpu.json): runs on the CPU, no GPU work -- guards the code that turns the
kernel-class profile into the `stages` / `roofline` objects of the JSON line."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_stage_rooflines_on_recorded_profile():
    b = _bench()
    rec = json.loads(open(os.path.join(ROOT, "profiles", "r01g_bench256_1gpu.json")).read().strip().splitlines()[-1])
    table = {k: (v["ms_total"], v["launches"]) for k, v in rec["stages"].items()}
    cg, sweeps = rec["config"]["cg_iterations"], rec["config"]["drive_eps_sig_sweeps"]
    stages, roof = b.stage_rooflines(table, 256, 256 ** 3, cg, sweeps - 2, 1, rec["config"]["fp64_peak_tflops_measured"])
    assert roof["kernel"] == "k_fwd_z_K4" and roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    # fused CG work is counted where it is moved: 792.6 + 288 x (share of MODE 3 launches)
    assert 1070.0 < stages["k_fwd_z_K4"]["alg_bytes_per_voxel"] <= 1080.6
    assert 214.0 < stages["k_inv_z"]["alg_bytes_per_voxel"] <= 216.6
    assert 0.9 < roof["frac"] < 1.05 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert roof["traffic"] and 0.95 < roof["traffic"] / roof["alg_bytes_per_launch"] < 1.1     # ncu DRAM bytes ~ algorithmic
    assert 0.15 < stages["k_update_mm10"]["frac_of_fp64"] < 0.4
    # the FP64 rate of the profiled launch itself: its own flop count over its own duration (about 0.2 of the measured peak)
    assert 0.15 < stages["k_update_mm10"]["frac_of_fp64_ncu_launch"] < 0.25
    assert abs(sum(v["share"] for v in stages.values()) - 1.0) < 1e-9
    json.dumps({"stages": stages, "roofline": roof})          # serialisable


def test_algorithmic_bytes_table():
    b = _bench()
    assert abs(b.algorithmic_bytes_per_voxel("k_fwd_z_K4", 256) - (90 * 8 + 9 * 16 * 129 / 256)) < 1e-9
    assert b.algorithmic_bytes_per_voxel("k_fwd_z_K4", 256, 0.0, 1.0) - b.algorithmic_bytes_per_voxel("k_fwd_z_K4", 256) == 288.0
    assert b.algorithmic_bytes_per_voxel("k_inv_z", 256, 1.0) - b.algorithmic_bytes_per_voxel("k_inv_z", 256) == 72.0
    assert b.algorithmic_bytes_per_voxel("vector_ops", 256) is None


class _StubEvent:
    def __init__(self, enable_timing=True):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 250.0


class _StubSolver:
    """stands in for cpfft_b200.Solver in the dry run of bench.main(): hands back the recorded r01g
    kernel-class profile and plausible per-step counters; no device work"""
    def __init__(self, prob, **kw):
        self.prob, self.n3, self.step, self._launches = prob, len(prob.matlist), 0, 0
        rec = json.loads(open(os.path.join(ROOT, "profiles", "r01g_bench256_1gpu.json")).read().strip().splitlines()[-1])
        self._table = {k: (v["ms_total"], v["launches"]) for k, v in rec["stages"].items()}

    @staticmethod
    def nccl_unique_id():
        return bytes(128)

    def stream(self): return 0
    def synchronize(self): pass
    def drive_eps_sig(self, step, it): return 0
    def profile(self, on): pass
    def profile_reset(self): pass
    def kernel_launches(self): return self._launches
    def profile_table(self): return self._table
    def fp64_peak(self): return 36.0
    def exchange_mode(self): return "single"
    def download_ptr(self, name, ptr): pass
    def upload_ptr(self, name, ptr): pass
    def close(self): pass

    def FFT_nr3(self, nstep=1, first=0):
        import numpy as np
        assert first == self.step
        self.step += nstep
        self._launches += 1500
        return dict(rc=0, nr_iters=np.array([3]), cg_iters=[[17, 16, 12, 1]], Pbar=np.zeros((1, 9)),
                    buckets=np.array([1.0, 0.3]), counters=np.array([50, 5, 46, 0, 0]))


def test_bench_main_dry_run(monkeypatch, capsys):
    """bench.main() end to end on the CPU with the device layer stubbed out: every key of the JSON line the
    driver reads is produced (argument handling, counters, e2e block, stages / roofline, serialisation).
    The numbers are meaningless; tests/test_gpu_*.py and the round-end bench cover the real path."""
    import sys
    import torch
    import cpfft_b200
    b = _bench()
    monkeypatch.setattr(cpfft_b200, "Solver", _StubSolver)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda p: None)
    monkeypatch.setattr(torch.cuda, "Event", _StubEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    for extra in ([], ["--variant", "taylor2"], ["--stress-bc"]):
        monkeypatch.setattr(sys, "argv", ["bench.py", "--grid", "8", "--grains", "5", "--steps", "2", "--warmup", "3",
                                          "--no-cpu-baseline", "--no-parity"] + extra)
        b.main()
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "stages",
                    "cpu_baseline"):
            assert key in line, key
        assert line["metric"] == b.METRIC and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3
        assert line["value"] == 8 ** 3 * 100 / 0.5 and line["ms_per_step"] == 250.0     # 2 steps x 250 ms of device time
        assert line["e2e"]["value"] == 8 ** 3 * 100 / 0.25 and line["e2e"]["steps"] == 2     # stub events: 250 ms whatever the bracket
        assert (line["stress_bc_leg"] is None) == (extra == ["--stress-bc"])
        if line["stress_bc_leg"]:
            assert line["stress_bc_leg"]["steps"] == 1 and line["stress_bc_leg"]["value"] == 8 ** 3 * 50 / 0.25
        assert line["gpu_launches"] == 3000 and line["config"]["cg_solves"] == 8
        assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
        assert line["e2e"]["h2d_bytes_per_step"] == 9 * 512 * 8 and line["e2e"]["d2h_bytes_per_step"] == 18 * 512 * 8
        assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
        assert "workload" in line["config"] and "model" not in line["config"]
