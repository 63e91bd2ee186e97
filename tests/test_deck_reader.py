"""The deck reader (cpfft_b200/deck.py) on the reference's shipped input decks
(examples/*.in copied verbatim as fixtures into tests/golden/decks): what inmat / incrystal /
inelem / inlodcase / inlod / indypm leave in the reference's modules."""
import numpy as np
import pytest

import os

from helpers import deck, stress_bc_variant, DECKS
from cpfft_b200.deck import _int_list, _tokens, _is_comment


def test_scanner_conventions():
    assert _is_comment("c this is a comment") and _is_comment("c") and not _is_comment("crystal 1")
    assert _tokens("harden_n 5,48") == ["harden_n", "5", ",", "48"]
    assert _int_list(["1-5"]) == [1, 2, 3, 4, 5]
    assert _int_list(["2-10", "by", "2"]) == [2, 4, 6, 8, 10]
    assert _int_list(["1", "2", "7-8", "material"]) == [1, 2, 7, 8]


def test_mm01_deck():
    p = deck("test_mm01.in")
    assert p.N == 7 and p.N3 == 343
    assert [m.type for m in p.materials] == [1, 1]
    a, b = p.materials
    # REAL*4 matprp slots (mod_fft.f:20, inmat.f:100-127)
    assert a.e == 12000.0 and b.e == 24000.0
    assert a.nu == float(np.float32(0.3)) and a.nu != 0.3
    assert (a.yld_pt, b.yld_pt, a.tan_e, a.beta) == (100.0, 200.0, 1000.0, 0.5)
    assert set(np.unique(p.matlist)) == {1, 2} and p.matlist.shape == (343,)
    assert p.nstep == 10 and np.allclose(p.mults, 0.1)
    assert np.allclose(p.FP_max, [0.3, 0, 0, 0, -0.1, 0, 0, 0, -0.1])
    assert not p.isNBC.any()
    assert (p.tolNR, p.tolPCG, p.maxIter, p.tstep) == (1e-5, 1e-10, 10, 10.0)
    bc = p.BC_all()
    assert bc.shape == (10, 9)
    # cumulative + identity on the strain-controlled diagonal (inlod.f:57-63)
    assert np.allclose(bc[-1], [1.3, 0, 0, 0, 0.9, 0, 0, 0, 0.9])
    assert np.allclose(bc[0], [1.03, 0, 0, 0, 0.99, 0, 0, 0, 0.99])


def test_mm10_deck():
    p = deck("test_mm10.in")
    assert p.N == 7 and len(p.materials) == 1 and p.materials[0].type == 10
    c = p.crystals[0]
    assert c.slip_type == 8 and c.elastic_type == 1 and c.h_type == 1 and c.alter_mode == 1
    # `harden_n 5,48`: the comma makes the reader fetch the next line, 48 is discarded
    # (incrystal.f:989-990)
    assert c.harden_n == 5.0
    assert (c.e, c.nu, c.theta_0, c.voche_m, c.tau_v, c.tau_y) == (200000.0, 0.3, 0.01, 1.0, 5000.0, 205.0)
    assert c.eps_dot_0_y == 4e-5
    # defaults of initialize_new_crystal (mod_crystals.f:392-403)
    assert (c.atol, c.rtol, c.atol1, c.rtol1, c.miter) == (1e-5, 5e-5, 1e-5, 1e-5, 30)
    # angle_bc.in: a bicrystal, 196 voxels at (0,0,0) and 147 at (45,0,0)
    vals, counts = np.unique(p.angles, axis=0, return_counts=True)
    assert vals.tolist() == [[0, 0, 0], [45, 0, 0]] and counts.tolist() == [196, 147]
    assert np.allclose(p.FP_max, [0.03, 0, 0, 0, -0.01, 0, 0, 0, -0.01]) and p.tstep == 10.0


def test_stress_bc_variant_table():
    p = stress_bc_variant(deck("test_mm01.in"))
    assert list(p.isNBC) == [0, 0, 0, 0, 1, 0, 0, 0, 1]
    bc = p.BC_all()
    # stress-controlled diagonals get no identity: the prescribed P is 0
    assert np.allclose(bc[:, 4], 0) and np.allclose(bc[:, 8], 0) and np.allclose(bc[-1, 0], 1.3)


def test_pod_layouts_match_header():
    """ctypes mirrors of cpfft_crystal / cpfft_material / cpfft_config (include/cpfft_b200.h)."""
    import ctypes as C
    from cpfft_b200.problem import CrystalPOD, MaterialPOD
    from cpfft_b200.api import Config
    assert C.sizeof(CrystalPOD) == 6 * 4 + (16 + 14) * 8      # Voce block + 14 MTS parameters
    assert C.sizeof(MaterialPOD) == 2 * 4 + 6 * 4
    assert C.sizeof(Config) == 6 * 4 + 3 * 8


def test_output_cards_and_project_name():
    p = deck("test_mm10.in")
    assert p.name == "test"
    assert p.out_steps == (2, 4, 6, 8, 10)           # `output results steps 2-10 by 2`


def test_flat_result_writer(tmp_path):
    """wes / wee flat text files of ouresult.f: header, 30e15.6 records, zero-filled tail."""
    from cpfft_b200.results import write_step, fortran_e, flat_name
    assert fortran_e(123.4567) == "   0.123457E+03" and fortran_e(-1.2345678e-4) == "  -0.123457E-03"
    assert fortran_e(0.0) == "   0.000000E+00" and fortran_e(1e-31) == "   0.000000E+00"
    assert fortran_e(9.9999996) == "   0.100000E+02"
    assert flat_name("stresses", 4) == "wes00004_text" and flat_name("strains", 12) == "wee00012_text"
    rng = np.random.default_rng(0)
    urcs, eps = rng.standard_normal((343, 9)) * 100, rng.standard_normal((343, 6)) * 1e-3
    write_step(str(tmp_path), 2, urcs, eps, "test", 7)
    lines = open(tmp_path / "wes00002_text").read().splitlines()
    assert lines[0] == "#" and lines[1].startswith("#  WARP3D element results: stresses")
    assert lines[3] == "#  Model nodes, elements:      512     343" and lines[5] == "#  Load(time) step:        2"
    rows = lines[7:]
    assert len(rows) == 343 and all(len(r) == 26 * 15 for r in rows)
    back = np.array([[float(r[15 * k:15 * k + 15]) for k in range(26)] for r in rows])
    assert np.abs(back[:, :6] - urcs[:, :6]).max() <= 1e-6 * np.abs(urcs).max() * 10 and (back[:, 6:] == 0).all()
    rows_e = open(tmp_path / "wee00002_text").read().splitlines()[7:]
    assert all(len(r) == 22 * 15 for r in rows_e)


def test_polycrystalline_points_from_a_crystal_file(tmp_path):
    """n_crystals > 1 with `crystal_input file` / `orientation_input file`: ncry consecutive lines
    per element, `elem psi theta phi crystal` (read_defs, mod_crystals.f:2233-2318)."""
    from cpfft_b200.deck import read_deck, read_crystal_file, DeckError
    p = read_deck(os.path.join(DECKS, "taylor_mm10.in"))
    assert p.N == 5 and p.ncmax == 2 and p.taylor
    assert p.materials[0].n_crystals == 2 and p.materials[0].crystal_input == 2
    assert p.angles.shape == (125, 2, 3) and p.crystal_ids.shape == (125, 2)
    rows = [l.split(",") for l in open(os.path.join(DECKS, "taylor_crystals.in")) if l.strip()]
    for r in (0, 1, 100, 249):
        e, c = int(rows[r][0]) - 1, r % 2
        assert np.allclose(p.angles[e, c], [float(v) for v in rows[r][1:4]])
        assert p.crystal_ids[e, c] == int(rows[r][4])
    assert set(np.unique(p.crystal_ids)) == {1, 2} and len(p.crystals) == 2
    assert p.crystals[0].slip_type == 8 and p.crystals[1].slip_type == 1
    # angles only / crystals only variants of the same file format
    f = tmp_path / "ang.in"
    f.write_text("1 10 20 30\n1 40 50 60\n2 1 2 3\n2 4 5 6\n")
    a, _ = read_crystal_file(str(f), 2, 2, True, False)
    assert np.allclose(a[0], [[10, 20, 30], [40, 50, 60]]) and np.allclose(a[1], [[1, 2, 3], [4, 5, 6]])
    f.write_text("1 2\n1 1\n2 1\n2 2\n")
    _, ids = read_crystal_file(str(f), 2, 2, False, True)
    assert ids.tolist() == [[2, 1], [1, 2]]
    f.write_text("1 10 20 30\n2 1 2 3\n2 4 5 6\n")            # element 1 has one line, two are needed
    with pytest.raises(DeckError):
        read_crystal_file(str(f), 2, 2, True, False)


def test_mts_crystal_keywords():
    """`hardening mts` and its property keywords (incrystal.f:165-236, 305-331)"""
    p = deck("mts_mm10.in")
    c = p.crystals[0]
    assert c.h_type == 2 and c.slip_type == 1
    assert (c.tau_a, c.tau_hat_y, c.g_0_y, c.tau_hat_v, c.g_0_v) == (20.0, 180.0, 0.4, 300.0, 1.2)
    assert (c.burgers, c.boltzman, c.eps_dot_0_y, c.eps_dot_0_v) == (2.5e-7, 1.3806e-20, 1.0e10, 1.0e10)
    assert (c.p_y, c.q_y, c.p_v, c.q_v, c.mu_0, c.D_0, c.T_0) == (0.5, 2.0, 0.5, 2.0, 80000.0, 3000.0, 200.0)
    assert c.theta_0 == 1500.0 and c.harden_n == 20.0


def test_cli_timing_summary_format(capsys):
    """the three thyme() buckets printed like outime (outime.f:29-49)"""
    from cpfft_b200.__main__ import print_timings
    print_timings([[1.2345, 12], [0.5, 30], [0.0, 0]], 2.0)
    out = capsys.readouterr().out
    assert ">>>>>  solution timings   <<<<<" in out
    assert "calculations for pcg solution vector update:" in out and "sig-eps & internal force:" in out
    assert "patran output" not in out                       # buckets without calls are skipped
    assert "wall time (secs):     1.2345 61.7 (%) no. calls:       12" in out


def test_single_crystal_points_with_a_crystal_file(tmp_path, oracle_built):
    """n_crystals 1 with `crystal_input file` and `orientation_input single`: the flat file carries
    `elem crystal` only (read_defs with angles = .false., mod_crystals.f:2262-2264); the oracle and the
    host build of the kernel source run it."""
    import shutil
    from cpfft_b200.deck import read_deck
    src = open(os.path.join(DECKS, "taylor_mm10.in")).read()
    src = src.replace("n_crystals 2", "n_crystals 1").replace("orientation_input file filename 'taylor_crystals.in'",
                                                                "orientation_input single angles 10.0 20.0 30.0 filename 'cry.in'")
    (tmp_path / "deck.in").write_text(src)
    (tmp_path / "cry.in").write_text("".join(f"{e} {1 + (e % 2)}\n" for e in range(1, 126)))
    p = read_deck(str(tmp_path / "deck.in"))
    assert p.ncmax == 1 and p.taylor and p.angles.shape == (125, 3) and p.crystal_ids.shape == (125, 1)
    assert np.all(p.angles == [10.0, 20.0, 30.0])
    assert p.crystal_ids[:4, 0].tolist() == [2, 1, 2, 1]
    from oracle import Oracle
    from host_kernels import HostKernels
    o, k = Oracle(p), HostKernels(p)
    assert o.H == k.H == 357                      # the bcc48 crystal forces the 48-system layout on every point
    o.drive_eps_sig(1, 0); k.drive_eps_sig(1, 0)
    assert np.abs(k.K4 - o.K4).max() <= 1e-9 * np.abs(o.K4).max()


def test_angle_type_radians(tmp_path):
    """`angle_type radians` (inmat.f:206-216): converted to the degrees the C ABI takes"""
    from cpfft_b200.deck import read_deck
    src = open(os.path.join(DECKS, "test_mm10.in")).read().replace("angle_type degrees", "angle_type radians")
    src = src.replace("filename 'angle_bc.in'", "filename 'ang.in'")
    (tmp_path / "deck.in").write_text(src)
    (tmp_path / "ang.in").write_text("".join(f"{e}, {0.25 * np.pi}, 0.0, 0.1\n" for e in range(1, 344)))
    p = read_deck(str(tmp_path / "deck.in"))
    assert np.allclose(p.angles, [45.0, 0.0, np.degrees(0.1)], rtol=0, atol=1e-12)


def test_cli_host_glue_with_the_oracle_standing_in(tmp_path, monkeypatch, capsys, oracle_built):
    """`python -m cpfft_b200 deck` end to end on the CPU: the CLI's host glue (deck reader, step loop,
    result files of the selected steps, timing summary) with the oracle standing in for the CUDA
    library behind the Solver interface (test infrastructure; the product Solver needs a GPU)."""
    import ctypes as C
    import cpfft_b200.__main__ as cli
    from oracle import Oracle

    class OracleSolver:
        def __init__(self, prob, device=0):
            self.o, self.prob = Oracle(prob, threads=2), prob

        def drive_eps_sig(self, step, it):
            self.o.drive_eps_sig(step, it)

        def FFT_nr3(self, nstep=1, first=0):
            o, n = self.o, nstep
            bc = np.ascontiguousarray(self.prob.BC_all()[first:first + n])
            nbc = np.ascontiguousarray(self.prob.isNBC, dtype=np.int32)
            nr = np.zeros(n, dtype=np.int32); cg = np.full((n, 64), -1, dtype=np.int32)
            pb = np.zeros((n, 9)); bk = np.zeros(3); cnt = np.zeros(5, dtype=np.int64)
            dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
            rc = o.L.orc_FFT_nr3_from(o.h, first + 1, n, bc.ctypes.data_as(dp), nbc.ctypes.data_as(ip), nr.ctypes.data_as(ip),
                                      cg.ctypes.data_as(ip), 64, pb.ctypes.data_as(dp), bk.ctypes.data_as(dp),
                                      cnt.ctypes.data_as(C.POINTER(C.c_int64)))
            assert rc == 0
            cgl = [list(r[:list(r).index(-1)]) if -1 in r else list(r) for r in cg]
            return dict(nr_iters=nr, cg_iters=cgl, Pbar=pb, buckets=bk, counters=cnt,
                        log=f"     Now starting step: {first + 1:7d}\\n")

        def download(self, name, layout=0):
            return {"URCS_N1": self.o.urcs_n1, "EPS_N1": self.o.eps_n1, "FN1": self.o.Fn1}[name]

        def material_failures(self):
            return (0, 0)

    monkeypatch.setattr(cli, "Solver", OracleSolver)
    rc = cli.main([os.path.join(DECKS, "test_mm10.in"), "--outdir", str(tmp_path), "--steps", "4"])
    assert rc == 0
    out = capsys.readouterr().out
    assert out.count("Now starting step") == 4
    from helpers import CLI_MM10_FILES
    assert sorted(os.listdir(tmp_path)) == CLI_MM10_FILES
    assert open(tmp_path / "RM_model_flat.text").read().splitlines()[7] == f"{512:9d}{343:9d}"
    # nodal displacements recovered from F: 8^3 nodes, the far corner carries the applied mean stretch
    nod = open(tmp_path / "wnd00004_text").read().splitlines()[7:]
    assert len(nod) == 512
    bc = deck("test_mm10.in").BC_all()[3]
    far = np.array([float(nod[-1][15 * k:15 * k + 15]) for k in range(3)])
    assert np.abs(far - 100.0 * (bc.reshape(3, 3) - np.eye(3)).sum(axis=1)).max() <= 2e-2 * np.abs(far).max()
    rows = open(tmp_path / "wes00004_text").read().splitlines()[7:]
    assert len(rows) == 343 and all(len(r) == 26 * 15 for r in rows)
    assert "solution timings" in out and "pcg solution vector update" in out and "patran output" in out


def test_crystal_debug_and_tangent_keywords(tmp_path):
    """`tang_calc 0` and the debug print selectors (gpall on|off, gpp / delem / dstep / diter <int>, incrystal.f:
    953-986) are accepted and change nothing; another tang_calc or an unknown property is an error"""
    from cpfft_b200.deck import read_deck, DeckError
    src = open(os.path.join(DECKS, "test_mm10.in")).read()
    assert "harden_n" in src
    ref = deck("test_mm10.in")
    for extra, ok in (("tang_calc 0 gpall off gpp 1 delem 3 dstep 2 diter 1", True), ("tang_calc 2", False), ("rho_0 1.0", False)):
        path = tmp_path / "d.in"
        path.write_text(src.replace("harden_n", extra + " harden_n", 1))
        for f in ("angle_bc.in", "angle2.in"):                      # the deck names its angle file relative to itself
            if os.path.exists(os.path.join(DECKS, f)):
                (tmp_path / f).write_text(open(os.path.join(DECKS, f)).read())
        if ok:
            p = read_deck(str(path))
            assert p.crystals[0] == ref.crystals[0]
        else:
            with pytest.raises(DeckError):
                read_deck(str(path))
