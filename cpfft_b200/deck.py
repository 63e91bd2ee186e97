"""Reader for the reference's free-form input decks (examples/*.in), restricted to the
keywords the shipped decks use.  Semantics follow src/FFT_finite_3d.f:14-160 (command loop),
src/inmat.f:69-300, src/incrystal.f:33-998, src/inelem.f:33-106, src/inlodcase.f:24-146,
src/inlod.f:26-109, src/indypm.f:26-68 and the scanner's conventions in src/scan.f:
  * a line whose first column is ``c`` followed by a blank (or nothing) is a comment;
  * a comma is a token; in the property readers a comma makes the reader fetch the NEXT line
    (``readsc``), discarding whatever followed the comma on the current line
    (incrystal.f:989-990, inmat.f:288-289) -- ``harden_n 5,48`` therefore reads 5;
  * bilinear properties are parsed into REAL*4 (inmat.f:100-127).
"""
from __future__ import annotations

import os
import re
from typing import List

import numpy as np

from .problem import Crystal, Material, Problem, SLIP_TYPES, ELASTIC_TYPES, COMPONENTS, HARDENING


class DeckError(ValueError):
    pass


def _tokens(line: str) -> List[str]:
    # quoted strings stay one token; commas are tokens
    out = []
    for m in re.finditer(r"'[^']*'|\"[^\"]*\"|,|[^\s,]+", line):
        out.append(m.group(0))
    return out


def _is_comment(line: str) -> bool:
    s = line.rstrip("\n")
    return len(s) >= 1 and s[0] in "cC" and (len(s) == 1 or s[1] in " \t")


def _int_list(tokens: List[str]) -> List[int]:
    """Integer lists ``1-100``, ``2-10 by 2``, ``1 2 3`` (user_list.f:13, scan trlist)."""
    vals: List[int] = []
    i = 0
    while i < len(tokens):
        t = tokens[i]
        m = re.fullmatch(r"(\d+)-(\d+)", t)
        if m:
            a, b, step = int(m.group(1)), int(m.group(2)), 1
            if i + 2 < len(tokens) and tokens[i + 1].lower() == "by":
                step = int(tokens[i + 2]); i += 2
            vals.extend(range(a, b + 1, step))
        elif re.fullmatch(r"\d+", t):
            vals.append(int(t))
        else:
            break
        i += 1
    return vals


class _Lines:
    def __init__(self, text: str):
        self.lines = [l for l in text.splitlines() if not _is_comment(l) and l.strip()]
        self.pos = 0

    def peek(self):
        return self.lines[self.pos] if self.pos < len(self.lines) else None

    def next(self):
        l = self.peek()
        self.pos += 1
        return l


def _property_tokens(lines: _Lines, first: List[str]) -> List[str]:
    """Token stream of a ``properties`` card: a comma jumps to the next line."""
    toks: List[str] = []
    cur = first
    while True:
        if "," in cur:
            toks.extend(cur[:cur.index(",")])
            nxt = lines.next()
            if nxt is None:
                break
            cur = _tokens(nxt)
        else:
            toks.extend(cur)
            break
    return toks


_CRYSTAL_NUM = {
    "e": "e", "nu": "nu", "mu": "mu", "harden_n": "harden_n", "theta_0": "theta_0",
    "tau_y": "tau_y", "tau_v": "tau_v", "voche_m": "voche_m", "voce_m": "voche_m",
    "iD_v": "iD_v", "eps_dot_0_y": "eps_dot_0_y", "gamma_bar": "eps_dot_0_y", "k_0": "k_0",
    "b": "burgers", "atol": "atol", "atol1": "atol1", "rtol": "rtol", "rtol1": "rtol1",
    # MTS (incrystal.f:165-236)
    "tau_a": "tau_a", "tau_hat_y": "tau_hat_y", "g_0_y": "g_0_y", "tau_hat_v": "tau_hat_v", "g_0_v": "g_0_v",
    "p_y": "p_y", "q_y": "q_y", "p_v": "p_v", "q_v": "q_v", "boltz": "boltzman", "eps_dot_0_v": "eps_dot_0_v",
    "mu_0": "mu_0", "D_0": "D_0", "T_0": "T_0",
}


def _read_crystal(lines: _Lines, toks: List[str], crystals: dict):
    cnum = int(toks[1])
    c = Crystal()
    first = _tokens(lines.next())
    if not first or not first[0].lower().startswith("prop"):
        raise DeckError("crystal: expected 'properties'")
    pt = _property_tokens(lines, first[1:])
    i = 0
    while i < len(pt):
        k = pt[i]
        if k == "slip_type":
            v = pt[i + 1]
            if v not in SLIP_TYPES:
                raise DeckError(f"slip_type {v} not supported (fcc, bcc, single, roters, bcc12, bcc48)")
            c.slip_type = SLIP_TYPES[v]; i += 2
        elif k == "elastic_type":
            c.elastic_type = ELASTIC_TYPES[pt[i + 1]]; i += 2
        elif k == "hardening":
            if pt[i + 1] not in HARDENING:
                raise DeckError(f"hardening {pt[i + 1]} not supported (voce, mts)")
            c.h_type = HARDENING[pt[i + 1]]; i += 2
        elif k == "alter_mode":
            c.alter_mode = 1 if pt[i + 1].lower() in ("on", "true") else 0; i += 2
        elif k == "miter":
            c.miter = int(pt[i + 1]); i += 2
        elif k == "solver":
            if pt[i + 1] != "nr":
                raise DeckError("only solver nr is supported")
            i += 2
        elif k == "tang_calc":
            # 0 = the analytical tangent (default, mod_crystals.f:1750-1759); 1-4 select finite-difference /
            # complex-step checks of the reference's debugging paths
            if int(float(pt[i + 1])) != 0:
                raise DeckError("only tang_calc 0 is supported")
            i += 2
        elif k in ("gpall", "gpp", "delem", "dstep", "diter"):
            # debug print selectors of mm10, one argument each (incrystal.f:953-986): no effect on the solution
            i += 2
        elif k in _CRYSTAL_NUM:
            setattr(c, _CRYSTAL_NUM[k], float(pt[i + 1])); i += 2
        else:
            raise DeckError(f"crystal: unknown property {k}")
    crystals[cnum] = c


def _read_material(lines: _Lines, toks: List[str], materials: List[Material], base_dir: str):
    m = Material(name=toks[1])
    first = _tokens(lines.next())
    if not first or not first[0].lower().startswith("prop"):
        raise DeckError("material: expected 'properties'")
    kind = first[1]
    if kind == "bilinear":
        m.type = 1
        # inmat.f:97-133: a comma only continues when it ends the line
        pt: List[str] = []
        cur = first[2:]
        while True:
            ends_with_comma = bool(cur) and cur[-1] == ","
            pt.extend(t for t in cur if t != ",")
            if ends_with_comma:
                cur = _tokens(lines.next())
            else:
                break
        i = 0
        while i < len(pt):
            k = pt[i]
            if k in ("e", "nu", "beta", "tan_e", "yld_pt"):
                setattr(m, k, float(np.float32(float(pt[i + 1])))); i += 2
            elif k in ("alpha", "rho"):
                i += 2
            else:
                raise DeckError(f"bilinear: unknown property {k}")
    elif kind == "cp":
        m.type = 10
        pt = _property_tokens(lines, first[2:])
        i = 0
        ncry = 1
        while i < len(pt):
            k = pt[i]
            if k in ("rho", "alpha", "tolerance"):
                i += 2
            elif k == "angle_convention":
                if pt[i + 1] != "kocks":
                    raise DeckError("only kocks angles are supported")
                i += 2
            elif k == "angle_type":
                if pt[i + 1] not in ("degrees", "radians"):
                    raise DeckError("angle_type must be degrees or radians (inmat.f:206-216)")
                m.angle_scale = 1.0 if pt[i + 1] == "degrees" else 180.0 / np.pi   # the C ABI takes degrees
                i += 2
            elif k == "n_crystals":
                ncry = int(pt[i + 1]); i += 2
            elif k == "crystal_input":
                if pt[i + 1] not in ("single", "file"):
                    raise DeckError("crystal_input must be single or file (inmat.f:218-233)")
                m.crystal_input = 1 if pt[i + 1] == "single" else 2; i += 2
            elif k == "crystal_type":
                m.crystal = int(pt[i + 1]); i += 2
            elif k == "orientation_input":
                m.orientation_input = 1 if pt[i + 1] == "single" else 2; i += 2
            elif k == "angles":
                m.angles = (float(pt[i + 1]), float(pt[i + 2]), float(pt[i + 3])); i += 4
            elif k == "filename":
                m.orientation_file = os.path.join(base_dir, pt[i + 1].strip("'\"")); i += 2
            elif k == "debug":
                i += 2
            else:
                raise DeckError(f"cp: unknown property {k}")
        if ncry < 1:
            raise DeckError("n_crystals must be >= 1")
        m.n_crystals = ncry
    else:
        raise DeckError(f"material model {kind} not supported (bilinear, cp)")
    materials.append(m)


def read_orientation_file(path: str, n3: int) -> np.ndarray:
    """``elem, psi, theta, phi`` per line (mod_crystals.f:2233-2318, read sequentially)."""
    return read_crystal_file(path, n3, 1, True, False)[0][:, 0, :]


def read_crystal_file(path: str, n3: int, ncry: int, with_angles: bool, with_crystals: bool):
    """The flat file of ``read_defs`` (mod_crystals.f:2233-2318): ``ncry`` consecutive lines per
    element, each ``elem [psi theta phi] [crystal]`` -- angles when ``orientation_input file``,
    crystal numbers when ``crystal_input file``.  Returns (angles (n3,ncry,3), ids (n3,ncry));
    elements absent from the file keep zeros (the reference aborts when it needs one)."""
    ang = np.zeros((n3, ncry, 3))
    ids = np.zeros((n3, ncry), dtype=np.int32)
    seen = {}
    with open(path) as f:
        rows = [re.split(r"[,\s]+", l.strip()) for l in f if l.strip()]
    for r in rows:
        e = int(r[0])
        if not 1 <= e <= n3:
            continue
        c = seen.get(e, 0)
        if c >= ncry:
            continue                      # read_defs stops after ncry lines of the element
        seen[e] = c + 1
        pos = 1
        if with_angles:
            ang[e - 1, c] = [float(r[1]), float(r[2]), float(r[3])]; pos = 4
        if with_crystals:
            ids[e - 1, c] = int(r[pos])
    short = [e for e, c in seen.items() if c < ncry]
    if short:
        raise DeckError(f"{path}: insufficient data for element {short[0]} (mod_crystals.f:2310)")
    return ang, ids


def read_deck(path: str) -> Problem:
    base_dir = os.path.dirname(os.path.abspath(path))
    with open(path) as f:
        lines = _Lines(f.read())
    N = None
    materials: List[Material] = []
    crystals: dict = {}
    elem_mat = None
    FP_max = np.zeros(9)
    isNBC = np.zeros(9, dtype=np.int32)
    mults: dict = {}
    tolNR, tolPCG, maxIter, tstep = 1e-5, 1e-10, 10, 1.0
    name, out_steps, model_file = "", (), ""
    lengths = [1.0, 1.0, 1.0]                # l_x, l_y, l_z (mod_fft.f:8); only the output mesh uses them
    while True:
        line = lines.next()
        if line is None:
            break
        toks = _tokens(line)
        key = toks[0].lower()
        if key == "number":                                   # number of grid N
            N = int(toks[-1])
        elif key == "crystal":
            _read_crystal(lines, toks, crystals)
        elif key == "material":
            _read_material(lines, toks, materials, base_dir)
        elif key == "elements":
            if N is None:
                raise DeckError("'number of grid' must precede 'elements'")
            elem_mat = np.zeros(N ** 3, dtype=np.int32)
            names = [m.name for m in materials]
            while lines.peek() is not None and re.match(r"\s*\d", lines.peek()):
                t = _tokens(lines.next())
                k = t.index("material")
                ids = _int_list(t[:k])
                elem_mat[np.asarray(ids) - 1] = names.index(t[k + 1]) + 1
        elif key == "strains":
            while lines.peek() is not None:
                t = _tokens(lines.peek())
                mm = re.fullmatch(r"([FP])_([xyz]{2})", t[0])
                if not mm:
                    break
                lines.next()
                idx = COMPONENTS.index(mm.group(2))
                FP_max[idx] = float(t[1])
                isNBC[idx] = 1 if mm.group(1) == "P" else 0
        elif key == "loading":
            while lines.peek() is not None and _tokens(lines.peek())[0].lower() == "step":
                t = _tokens(lines.next())
                k = [x.lower()[:6] for x in t].index("constr")
                for s in _int_list(t[1:k]):
                    mults[s] = float(t[k + 1])
        elif key == "nonlinear":
            while lines.peek() is not None:
                t = _tokens(lines.peek())
                k0 = t[0].lower()
                if k0.startswith("maximum"):
                    maxIter = int(t[-1])
                elif k0.startswith("converge"):
                    for j, x in enumerate(t):
                        if x == "NR":
                            tolNR = float(t[j + 1])
                        if x == "CG":
                            tolPCG = float(t[j + 1])
                elif k0 == "time":
                    tstep = float(t[-1])
                else:
                    break
                lines.next()
        elif key == "project":
            name = toks[1] if len(toks) > 1 else ""
        elif key == "output":
            # `output results steps <integer list>` selects the steps written by ouresult
            low = [t.lower() for t in toks]
            if len(low) > 2 and low[1].startswith("result") and low[2].startswith("step"):
                out_steps = tuple(_int_list(toks[3:]))
            elif len(low) > 2 and low[1] == "model":          # `output model "<file>"` (oudriv.f:22, oumodel.f:63-73)
                model_file = toks[2].strip('"\'')
        elif key == "sizes":
            # `sizes of x_direction <l_x> y_direction <l_y> z_direction <l_z>` (FFT_finite_3d.f:97-114): the
            # cell lengths.  formG ignores them (FFT_init.f:272-340); they size the output mesh of oumodel / f2disp
            low = [t.lower() for t in toks]
            for j, t in enumerate(low[:-1]):
                for ax, word in enumerate(("x_direction", "y_direction", "z_direction")):
                    if len(t) >= 4 and word.startswith(t):
                        lengths[ax] = float(toks[j + 1])
        elif key in ("blocking", "compute"):
            pass
        elif key == "stop":
            break
        else:
            raise DeckError(f"unknown command: {line.strip()}")
    if N is None or elem_mat is None:
        raise DeckError("deck lacks grid size or element list")
    n3 = N ** 3
    ncmax = max([m.n_crystals for m in materials if m.type == 10] or [1])
    from_file = any(m.type == 10 and m.crystal_input == 2 for m in materials)
    angles = np.zeros((n3, ncmax, 3))
    crystal_ids = np.zeros((n3, ncmax), dtype=np.int32) if from_file else None
    for im, m in enumerate(materials):
        if m.type != 10:
            continue
        sel = elem_mat == im + 1
        nc = m.n_crystals
        if m.orientation_input == 2 or m.crystal_input == 2:
            fa, fi = read_crystal_file(m.orientation_file, n3, nc, m.orientation_input == 2, m.crystal_input == 2)
        if m.orientation_input == 2:
            angles[sel, :nc] = fa[sel] * m.angle_scale
        else:
            angles[sel, :nc] = np.asarray(m.angles) * m.angle_scale
        if crystal_ids is not None:
            crystal_ids[sel, :nc] = fi[sel] if m.crystal_input == 2 else m.crystal
    if ncmax == 1:
        angles = angles[:, 0, :]
        crystal_ids = None if crystal_ids is None else crystal_ids
    ncry = max(crystals) if crystals else 0
    cry_list = [crystals.get(i + 1, Crystal()) for i in range(ncry)]
    nstep = max(mults) if mults else 0
    mult_arr = np.array([mults.get(s + 1, 0.0) for s in range(nstep)])
    return Problem(N=N, materials=materials, crystals=cry_list, matlist=elem_mat, angles=angles,
                   FP_max=FP_max, isNBC=isNBC, mults=mult_arr, tolNR=tolNR, tolPCG=tolPCG,
                   maxIter=maxIter, tstep=tstep, name=name, out_steps=out_steps, crystal_ids=crystal_ids,
                   lengths=tuple(lengths), model_file=model_file)
