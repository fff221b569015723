"""Synthetic quotient-polynomial AST with the SHAPE of the reference's TinyRamCircuit<W, 8>:
column counts, constraint counts and degrees are taken from the cited reference code, the concrete polynomial
identities are pseudo-random (seeded) stand-ins of the same degree and fan-in.  Used by bench.py's create_proof
workload model and by the large-size quotient tests; it does not reproduce the circuit's semantics (the witness
generator and the real gate list stay on the CPU side of the boundary).

  advice 263 (exe.rs:540-552, prog.rs:143, even_bits.rs:98-99 x 14, ...), instance 94 (prog.rs:141), fixed 23,
  gates ~139 of degree <= 6 (sprod.rs:65-90 is the maximum), 31 lookups (30 static + 1 dynamic 95-wide,
  even_bits.rs:158-165, out_table.rs:33-74, shift.rs:142-165, circuits/mod.rs:52-57), 188 permutation columns
  in 47 chunks of d - 2 = 4 (prog.rs:151-152).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import poly as P

N_ADVICE, N_INSTANCE, N_FIXED = 263, 94, 23
N_GATES, N_LOOKUPS, N_PERM_COLS, PERM_CHUNK = 139, 31, 188, 4
LOOKUP_WIDTHS = [1] * 28 + [15, 2, 95]          # 28 even-bits, CorrectOut, Shift/pow, dynamic program-line lookup


@dataclass
class Shape:
    ast: P.Ast
    n_columns: int
    n_expressions: int
    groups: dict            # name -> (first column, count)
    lookup_exprs: list = None   # per lookup: (compressed input Ast, compressed table Ast) over the Lagrange basis
    perm_cols: list = None      # column index of every equality-enabled column, in argument order


def build(seed: int = 40, modulus: int = P._MODULUS[0], scale: float = 1.0) -> Shape:
    """scale < 1 shrinks every count proportionally (small tests); scale = 1 is the k = 20 workload."""
    rng = random.Random(seed)
    sc = lambda x: max(1, int(round(x * scale)))
    groups, nxt = {}, 0

    def alloc(name, count):
        nonlocal nxt
        groups[name] = (nxt, count)
        cols = [P.Poly(nxt + i) for i in range(count)]
        nxt += count
        return cols

    n_lookups = sc(N_LOOKUPS)
    n_perm_cols = sc(N_PERM_COLS)
    n_chunks = (n_perm_cols + PERM_CHUNK - 1) // PERM_CHUNK
    advice = alloc("advice", sc(N_ADVICE))
    instance = alloc("instance", sc(N_INSTANCE))
    fixed = alloc("fixed", sc(N_FIXED))
    l0, l_last, l_blind = alloc("lagrange_selectors", 3)
    perm_z = alloc("permutation_z", n_chunks)
    sigma = alloc("permutation_sigma", n_perm_cols)
    look_a = alloc("lookup_permuted_input", n_lookups)
    look_s = alloc("lookup_permuted_table", n_lookups)
    look_z = alloc("lookup_z", n_lookups)
    rnd = lambda: rng.randrange(modulus)
    any_col = advice + instance + fixed
    exprs: List[P.Ast] = []

    # ---- custom gates: selector * (polynomial identity of degree 1..5 over current / next row cells) -------------------
    for g in range(sc(N_GATES)):
        sel = rng.choice(fixed)
        deg = rng.choice([1, 2, 2, 3, 3, 4, 5])
        term = None
        for _ in range(deg):
            c = rng.choice(advice)
            f = (c.with_rotation(1) if rng.random() < 0.2 else c) - (rng.choice(advice) if rng.random() < 0.5 else rng.randrange(1 << 16))
            term = f if term is None else term * f
        if rng.random() < 0.5:
            term = term - rng.choice(advice) * rng.randrange(1 << 32)
        exprs.append(sel * term)

    # ---- permutation argument: boundary rules + one product rule per chunk of 4 columns (degree 6) ---------------------
    beta, gamma, delta = rnd(), rnd(), 5
    perm_cols = [any_col[i % len(any_col)] for i in range(n_perm_cols)]
    active = P.ConstantTerm(1) - (l_last + l_blind)
    exprs.append(l0 * (P.ConstantTerm(1) - perm_z[0]))
    exprs.append(l_last * (perm_z[-1] * perm_z[-1] - perm_z[-1]))
    for i in range(1, n_chunks):
        exprs.append(l0 * (perm_z[i] - perm_z[i - 1].with_rotation(-6)))
    dpow = 1
    for ch in range(n_chunks):
        cols = perm_cols[ch * PERM_CHUNK:(ch + 1) * PERM_CHUNK]
        sig = sigma[ch * PERM_CHUNK:(ch + 1) * PERM_CHUNK]
        left = perm_z[ch].with_rotation(1)
        right = perm_z[ch]
        for c, s in zip(cols, sig):
            left = left * (c + s * beta + gamma)
            right = right * (c + P.LinearTerm(beta * dpow % modulus) + gamma)
            dpow = dpow * delta % modulus
        exprs.append((left - right) * active)

    # ---- lookups: 5 rules each; input / table expressions are theta-compressions of `width` columns --------------------
    theta = rnd()
    widths = (LOOKUP_WIDTHS * ((n_lookups + len(LOOKUP_WIDTHS) - 1) // len(LOOKUP_WIDTHS)))[:n_lookups] if scale < 1.0 else LOOKUP_WIDTHS
    lookup_exprs = []
    for lk in range(n_lookups):
        w = widths[lk]
        sel = rng.choice(fixed)
        inp = P.DistributePowers([sel * rng.choice(advice) for _ in range(w)], P.ConstantTerm(theta))
        tab = P.DistributePowers([rng.choice(fixed + instance) for _ in range(w)], P.ConstantTerm(theta))
        lookup_exprs.append((inp, tab))
        z, a, s = look_z[lk], look_a[lk], look_s[lk]
        exprs.append(l0 * (P.ConstantTerm(1) - z))
        exprs.append(l_last * (z * z - z))
        exprs.append((z.with_rotation(1) * (a + beta) * (s + gamma) - z * (inp + beta) * (tab + gamma)) * active)
        exprs.append(l0 * (a - s))
        exprs.append(((a - s) * (a - a.with_rotation(-1))) * active)

    # ---- fold with powers of y:  h = fold(0, |h, e| h * y + e)  (vanishing::Argument::construct) -------------------------
    y = rnd()
    h = P.ConstantTerm(0)
    for e in exprs:
        h = h * y + e
    return Shape(h, nxt, len(exprs), groups, lookup_exprs, [c.index for c in perm_cols])
