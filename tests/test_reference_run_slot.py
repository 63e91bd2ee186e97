"""External pin of the oracle: when a maintainer with ifort + MKL has dropped the flat result files of a real
reference run under tests/golden/reference_run/<deck>/ (recipe in the README there), the oracle's results for
the same deck are compared with them at the files' precision (format e15.6, ouresult.f:208-226).  Without the
files the test skips: parity stays "unpinned" (DESIGN.md 4)."""
import os
import re

import numpy as np
import pytest

from helpers import deck, DECKS

SLOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run")


def read_flat(path):
    """values of a WARP3D flat text result file: 7 header lines, then rows of e15.6 fields"""
    rows = []
    with open(path) as f:
        for line in f.read().splitlines()[7:]:
            if line.strip():
                rows.append([float(line[i:i + 15].replace("D", "E")) for i in range(0, len(line.rstrip()), 15)])
    return np.array(rows)


def compare_flat(ours, ref, name):
    a, b = read_flat(ours), read_flat(ref)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = np.maximum(np.abs(b).max(axis=0), 1e-30)
    err = (np.abs(a - b) / scale).max()
    assert err <= 1e-5, (name, err)          # e15.6 = 0.dddddd E+xx: 6 significant digits, half a unit of the last one on both sides
    return err


def reference_files(deck_name):
    d = os.path.join(SLOT, deck_name)
    if not os.path.isdir(d):
        return []
    return sorted(f for f in os.listdir(d) if re.fullmatch(r"w(es|ee|nd)\d{5}_text", f))


@pytest.mark.parametrize("deck_name", ["test_mm10.in", "test_mm01.in"])
def test_oracle_matches_reference_run(oracle_built, deck_name, tmp_path):
    files = reference_files(deck_name)
    if not files:
        pytest.skip(f"no reference run under {SLOT}/{deck_name} (needs ifort + MKL, see the README there): parity unpinned")
    from oracle import Oracle
    from cpfft_b200.results import write_step
    p = deck(deck_name)
    steps = sorted({int(f[3:8]) for f in files})
    o = Oracle(p, polar="double")             # the literal double arithmetic, as the reference binary runs it
    o.drive_eps_sig(1, 0)
    done = 0
    for step in steps:
        import ctypes as C
        from bench import _oracle_steps       # continue the committed state step by step
        while done < step:
            bc = np.ascontiguousarray(p.BC_all()[done:done + 1])
            r = _oracle_steps(o, bc, done + 1)
            done += 1
        write_step(str(tmp_path), step, o.urcs_n1, o.eps_n1, p.name, p.N, o.Fn1, p.lengths)
    for f in files:
        compare_flat(str(tmp_path / f), os.path.join(SLOT, deck_name, f), f)


def test_slot_reader_round_trip(tmp_path):
    """the reader used above understands what cpfft_b200/results.py writes (so a skip is the only reason the pin is open)"""
    from cpfft_b200.results import write_flat
    v = np.random.default_rng(0).standard_normal((5, 6)) * 1e3
    write_flat(str(tmp_path / "wes00001_text"), "stresses", 1, v, "x", 8)
    got = read_flat(str(tmp_path / "wes00001_text"))
    assert got.shape == (5, 26) and np.abs(got[:, :6] - v).max() <= 5e-6 * np.abs(v).max() and not got[:, 6:].any()
    assert compare_flat(str(tmp_path / "wes00001_text"), str(tmp_path / "wes00001_text"), "self") == 0.0
