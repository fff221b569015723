"""GPU tests of SURVEY.md 8(f) row f4 on the REAL circuit: plonk.create_proof over plonk.GpuBackend (libtrp.so) for the
reference's TinyRamCircuit (tiny-ram-halo2_b200/tinyram.py) on traces of the reference's own test programs, accepted by the
oracle's independent verify_proof -- the criterion of gen_proofs_and_verify (/root/reference/src/test_utils.rs:6-71), whose
`two_programs` caller (src/circuits/mod.rs:377-410) is reproduced first."""
import random

import pytest

from util import O, pm

import tinyram_programs as TP
import verify_util as VU

pytestmark = pytest.mark.gpu
C = pm.Vesta


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from tiny_ram_halo2_b200 import tinyram, trace
    return pkg, pkg.plonk, tinyram, trace, pkg.Context(0, pkg.VESTA)


_backends = {}


def _backend(env, k):
    pkg, PL, TR, T, ctx = env
    if k not in _backends:
        _backends[k] = PL.GpuBackend(ctx, k, 6)
    return _backends[k]


def _prove(env, be, pk, adv, inst, seed=1, debug=True):
    PL = env[1]
    rnd = random.Random(seed)
    return PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p), debug=debug)


def test_two_programs_reference_flow(env):
    """circuits/mod.rs:377-410 through test_utils.rs:6-71: k = 2 + W / 2, ONE key pair generated from the empty circuit, one
    proof per (circuit, program instance) pair, every proof verified against that key"""
    pkg, PL, TR, T, ctx = env
    W, k = 8, 6
    be = _backend(env, k)
    traces = [TP.answer_only(T, W), TP.load_and_answer(T, W, 1, 2)]
    circ, fixed, copies, _, _ = TR.build(PL, traces[0], k, keygen_from_empty_circuit=True)
    assert circ.cs.degree() == 6
    pk = PL.keygen(be, circ.cs, fixed, copies)
    proofs = []
    for tr in traces:
        _, _, _, adv, inst = TR.build(PL, tr, k)
        proof = _prove(env, be, pk, adv, inst)
        ok, err = VU.verify(be, pk.vk, inst, proof)
        assert ok, err
        proofs.append((proof, inst))
    # a proof does not verify against the other program's instance
    ok, _ = VU.verify(be, pk.vk, proofs[1][1], proofs[0][0])
    assert not ok


@pytest.mark.parametrize("name,a,b", [("And", 0xF0, 0x3C), ("Xor", 0xFF, 0x0F), ("Or", 1, 2), ("Add", 200, 77), ("Sub", 5, 3), ("Mull", 0x85, 3),
                                       ("UMulh", 0xFE, 0xFD), ("SMulh", 0x85, 0x7E), ("UDiv", 7, 200), ("UMod", 0, 9), ("Shl", 3, 0x91),
                                       ("Shr", 2, 0x93), ("Cmpe", 4, 4), ("Cmpa", 9, 200), ("Cmpae", 9, 9), ("Cmpg", 0x80, 1), ("Cmpge", 3, 0xFF)])
def test_instruction_proofs_with_live_selectors(env, name, a, b):
    """keys from the circuit WITH its execution-table selectors on (what MockProver::run sees): every gate and lookup of the
    instruction's gadget is then part of the statement the proof is about"""
    pkg, PL, TR, T, ctx = env
    k = 6
    be = _backend(env, k)
    circ, fixed, copies, adv, inst = TR.build(PL, TP.mov_named(T, 8, name, a, b), k)
    key = ("pk", k, 8)
    if key not in _backends:
        _backends[key] = PL.keygen(be, circ.cs, fixed, copies)            # the key does not depend on the program
    pk = _backends[key]
    proof = _prove(env, be, pk, adv, inst, seed=a * 256 + b)
    ok, err = VU.verify(be, pk.vk, inst, proof)
    assert ok, err
    assert len(proof) == len(_prove(env, be, pk, adv, inst, seed=5))


def test_unsatisfied_witness_does_not_verify(env):
    pkg, PL, TR, T, ctx = env
    k = 6
    be = _backend(env, k)
    circ, fixed, copies, adv, inst = TR.build(PL, TP.mov_named(T, 8, "Add", 200, 77), k)
    pk = PL.keygen(be, circ.cs, fixed, copies)
    adv[circ.reg[1]][2] = 22                                              # 77 + 200 != 22 (mod 256)
    proof = _prove(env, be, pk, adv, inst, debug=False)
    ok, _ = VU.verify(be, pk.vk, inst, proof)
    assert not ok
    circ, fixed, copies, adv, inst = TR.build(PL, TP.mov_named(T, 8, "Add", 200, 77), k)
    bad_inst = [list(c) for c in inst]
    bad_inst[0][1] = 5                                                    # Sub's opcode in the public program, Add in the witness
    proof = _prove(env, be, pk, adv, inst, debug=False)
    ok, _ = VU.verify(be, pk.vk, bad_inst, proof)
    assert not ok
    adv[circ.tv_b.even][1] = 2                                            # not an even-bits value: the lookup has no such row
    with pytest.raises(ValueError):
        _prove(env, be, pk, adv, inst, debug=False)


@pytest.mark.parametrize("W,k,iters", [(8, 10, 4), (16, 10, 14), (24, 13, 200)])
def test_loop_programs(env, W, k, iters):
    """BASELINE.json configs[0] (word size 8, k = 10) and wider words: a counting loop whose body runs every gadget"""
    pkg, PL, TR, T, ctx = env
    be = _backend(env, k)
    tr = TP.counting_loop(T, W, iters, TP.mixed_body(T, W) if W > 8 else ())
    circ, fixed, copies, adv, inst = TR.build(PL, tr, k, dense=False)
    pk = PL.keygen(be, circ.cs, TR.device_columns(be, fixed), copies)
    proof = _prove(env, be, pk, TR.device_columns(be, adv), TR.device_columns(be, inst))
    ok, err = VU.verify(be, pk.vk, inst, proof)
    assert ok, err
    bad = bytearray(proof); bad[len(proof) // 2] ^= 1
    ok, _ = VU.verify(be, pk.vk, inst, bytes(bad))
    assert not ok


def test_proof_bytes_equal_cpu_backend(env):
    """the real circuit through the SAME host logic over the oracle's PythonBackend: keys and proof bytes identical"""
    import plonk_model as VM
    pkg, PL, TR, T, ctx = env
    k = 6
    circ, fixed, copies, adv, inst = TR.build(PL, TP.mov_named(T, 8, "SMulh", 0x85, 0x7E), k)
    out = []
    for be in (_backend(env, k), VM.PythonBackend(C, k, 6)):
        pk = PL.keygen(be, circ.cs, fixed, copies)
        proof = _prove(env, be, pk, [list(c) for c in adv], inst, seed=9, debug=False)
        out.append((pk.vk, proof))
    assert out[0][0].fixed_commitments == out[1][0].fixed_commitments
    assert out[0][0].permutation_commitments == out[1][0].permutation_commitments
    assert out[0][0].transcript_repr == out[1][0].transcript_repr
    assert out[0][1] == out[1][1]
