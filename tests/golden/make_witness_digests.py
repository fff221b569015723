"""Freezes the witness of tiny-ram-halo2_b200/tinyram.py (TinyRamCircuit.synthesize + program_instance: fixed columns, copy
constraints, advice columns, instance columns) as SHA-256 digests, so that the synthesis code can be made faster without its
output moving by a bit (tests/test_tinyram_cpu.py::test_witness_digests).  Generated with the row-by-row implementation of
round 1 (the one checked against the reference's mock-prover tests):

    python tests/golden/make_witness_digests.py > tests/golden/witness_digests.json
"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases(T, programs):
    import tinyram_programs as TP
    out = {"answer_only_w8": (TP.answer_only(T, 8), 6), "load_and_answer_w8": (TP.load_and_answer(T, 8, 1, 2), 6)}
    rnd = random.Random(77)
    for name in TP.THREE_OPERAND + TP.TWO_OPERAND:
        a = rnd.randrange(1, 8) if name in ("Shl", "Shr") else rnd.randrange(1, 256)      # shifts beyond the word panic in the reference
        out[f"mov_{name}_w8"] = (TP.mov_named(T, 8, name, a, rnd.randrange(256)), 6)
    out["loop_w8"] = (programs.counting_loop(8, 2, programs.mixed_body(8)[:1]), 6)
    out["loop_w16"] = (programs.longest_loop(16), 10)
    out["loop_w16_short"] = (programs.counting_loop(16, 3, programs.mixed_body(16)), 11)
    out["loop_w32_short"] = (programs.counting_loop(32, 12, programs.mixed_body(32)), 17)
    out["loop_w32_full"] = (programs.longest_loop(32), 20)          # bench.py's create_proof_real witness (65 521 steps)
    return out


def digest(TR, PL, trace, k, **kw):
    circ, fixed, copies, adv, inst = TR.build(PL, trace, k, dense=False, **kw)
    h = hashlib.sha256()
    for f in fixed:
        h.update(repr((list(f.prefix), f.fill)).encode())
    h.update(repr([(c.left, c.right, c.rows) if hasattr(c, "rows") else tuple(c) for c in copies]).encode())
    for col in adv:
        h.update(repr(list(col)).encode())
    for col in inst:
        h.update(repr(list(col)).encode())
    return h.hexdigest()


def all_digests():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import plonk as PL, programs, tinyram as TR, trace as T
    out = {}
    for name, (tr, k) in cases(T, programs).items():
        out[name] = digest(TR, PL, tr, k)
        if name in ("load_and_answer_w8", "loop_w8", "loop_w16_short"):
            out[name + "/exe"] = digest(TR, PL, tr, k, with_prog=False)
            out[name + "/reg_index"] = digest(TR, PL, tr, k, reg_operand_value=False)
            out[name + "/empty_keys"] = digest(TR, PL, tr, k, keygen_from_empty_circuit=True)
    return out


if __name__ == "__main__":
    print(json.dumps(all_digests(), indent=1, sort_keys=True))
