"""TEST INFRASTRUCTURE ONLY: a constraint checker in the role of halo2_proofs::dev::MockProver, the
tool the reference's own circuit tests use (`MockProver::run(k, &circuit, instance)` + `assert_satisfied`,
/root/reference/src/circuits/mod.rs:364-375).  It is STRICTER than halo2's: every gate polynomial must vanish on EVERY usable
row (what a real proof needs), not only on rows of regions with an enabled selector; every lookup input row must occur among
the table rows; every copy constraint must hold.  Values are canonical ints; columns shorter than n are zero-padded."""
from __future__ import annotations

import numpy as np


def _dense(col, n):
    a = np.zeros(n, dtype=object)
    if isinstance(col, dict):
        for r, v in col.items():
            a[r] = v
    else:
        a[:len(col)] = list(col)
    return a


def check(PL, cs, n, p, fixed, advice, instances, copies=(), gate_names=None, max_failures=20):
    """Returns a list of failure strings (empty = satisfied)."""
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    cols = {PL.FIXED: [_dense(c, n) for c in fixed], PL.ADVICE: [_dense(c, n) for c in advice],
            PL.INSTANCE: [_dense(c, n) for c in instances]}
    for kind, num in ((PL.FIXED, cs.num_fixed), (PL.ADVICE, cs.num_advice), (PL.INSTANCE, cs.num_instance)):
        if len(cols[kind]) != num:
            return [f"{kind}: {len(cols[kind])} columns given, the constraint system has {num}"]
    memo = {}

    def ev(e):
        if isinstance(e, PL.Constant): return e.value % p
        if isinstance(e, PL.Query):
            key = (e.kind, e.column, e.rotation)
            if key not in memo:
                memo[key] = np.roll(cols[e.kind][e.column], -e.rotation)
            return memo[key]
        if isinstance(e, PL.Negated): return (-ev(e.a)) % p
        if isinstance(e, PL.Sum): return (ev(e.a) + ev(e.b)) % p
        if isinstance(e, PL.Product): return (ev(e.a) * ev(e.b)) % p
        if isinstance(e, PL.Scaled): return (ev(e.a) * (e.scalar % p)) % p
        raise TypeError(type(e))

    def column_of(e):
        v = ev(e)
        return v if isinstance(v, np.ndarray) else np.full(n, v, dtype=object)

    failures = []
    for gi, g in enumerate(cs.gates):
        v = column_of(g)[:usable]
        bad = np.nonzero(v != 0)[0]
        if len(bad):
            name = gate_names[gi] if gate_names else f"gate {gi}"
            failures.append(f"{name} (poly {gi}): not satisfied on rows {bad[:8].tolist()}")
            if len(failures) >= max_failures:
                return failures
    for li, (inputs, tables) in enumerate(cs.lookups):
        tab = set(zip(*[column_of(t)[:usable].tolist() for t in tables]))
        inp = list(zip(*[column_of(a)[:usable].tolist() for a in inputs]))
        bad = [r for r, row in enumerate(inp) if row not in tab]
        if bad:
            failures.append(f"lookup {li}: input rows {bad[:8]} not in the table, e.g. {inp[bad[0]][:4]}")
            if len(failures) >= max_failures:
                return failures
    for (lk, lc, lr), (rk, rc, rr) in PL.expand_copies(copies):
        if cols[lk][lc][lr] != cols[rk][rc][rr]:
            failures.append(f"copy ({lk} {lc} row {lr}) != ({rk} {rc} row {rr})")
            if len(failures) >= max_failures:
                return failures
    return failures
