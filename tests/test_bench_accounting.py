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
    assert 0.2 < stages["k_update_mm10"]["frac_of_fp64"] < 0.4
    assert abs(sum(v["share"] for v in stages.values()) - 1.0) < 1e-9
    json.dumps({"stages": stages, "roofline": roof})          # serialisable


def test_algorithmic_bytes_table():
    b = _bench()
    assert abs(b.algorithmic_bytes_per_voxel("k_fwd_z_K4", 256) - (90 * 8 + 9 * 16 * 129 / 256)) < 1e-9
    assert b.algorithmic_bytes_per_voxel("k_fwd_z_K4", 256, 0.0, 1.0) - b.algorithmic_bytes_per_voxel("k_fwd_z_K4", 256) == 288.0
    assert b.algorithmic_bytes_per_voxel("k_inv_z", 256, 1.0) - b.algorithmic_bytes_per_voxel("k_inv_z", 256) == 72.0
    assert b.algorithmic_bytes_per_voxel("vector_ops", 256) is None
