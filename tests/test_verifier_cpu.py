"""CPU tests of the product's own verifier (tiny-ram-halo2_b200/verifier.py: plonk::verify_proof, SingleVerifier,
BatchVerifier, Blake2bRead, the lazily collected MSM) over the oracle's PythonBackend.  The checker is the INDEPENDENT
oracle/plonk_model.verify_proof: the two must agree on accepted, tampered and malformed proofs; the flows are the reference's
gen_proofs_and_verify / gen_proofs_and_verify_should_fail (/root/reference/src/test_utils.rs:6-71, 73-119)."""
import random

import pytest

from util import pm

import plonk_model as VM
import plonk_circuits
import tinyram_programs as TP


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import plonk, tinyram, trace, verifier
    return plonk, verifier, tinyram, trace


def _prove(PL, C, k, seed=1, **kw):
    cs, fixed, copies, adv, inst = plonk_circuits.standard(PL, **kw)
    be = VM.PythonBackend(C, k, cs.degree())
    pk = PL.keygen(be, cs, fixed, copies)
    rnd = random.Random(seed)
    proof = PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p))
    return be, pk, inst, proof


def _accepts(V, be, vk, inst, proof):
    try:
        assert V.verify_proof(be, vk, V.SingleVerifier(be), inst, V.Blake2bRead(proof, be.q, be.p)) is None
        return True
    except V.VerifyError:
        return False


@pytest.mark.parametrize("C", [pm.Vesta, pm.Pallas], ids=["vesta", "pallas"])
@pytest.mark.parametrize("kw", [dict(with_lookup=True), dict(with_lookup=False), dict(with_lookup=True, wide_lookup=True)],
                         ids=["lookup", "no-lookup", "wide-lookup"])
def test_single_verifier_agrees_with_the_oracle(mods, C, kw):
    PL, V = mods[:2]
    be, pk, inst, proof = _prove(PL, C, 4, **kw)
    assert _accepts(V, be, pk.vk, inst, proof) and VM.verify_proof(C, be.params, pk.vk, inst, proof)
    # wrong public input; one flipped bit in every 32-byte word class of the proof (commitments, evaluations, the opening)
    assert not _accepts(V, be, pk.vk, [[inst[0][0] + 1]], proof)
    rnd = random.Random(5)
    positions = {0, 40, len(proof) // 2, len(proof) - 1, len(proof) - 33} | {rnd.randrange(len(proof)) for _ in range(12)}
    for pos in sorted(positions):
        bad = bytearray(proof); bad[pos] ^= 1 << rnd.randrange(8)
        ours, theirs = _accepts(V, be, pk.vk, inst, bytes(bad)), VM.verify_proof(C, be.params, pk.vk, inst, bytes(bad))
        assert not ours and not theirs, pos
    assert not _accepts(V, be, pk.vk, inst, proof[:-32])
    assert not _accepts(V, be, pk.vk, inst, proof[:100])


def test_instance_errors(mods):
    PL, V = mods[:2]
    C = pm.Vesta
    be, pk, inst, proof = _prove(PL, C, 4)
    t = lambda: V.Blake2bRead(proof, be.q, be.p)
    with pytest.raises(V.VerifyError, match="InvalidInstances"):
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), [], t())
    usable = be.n - (pk.vk.cs.blinding_factors() + 1)
    with pytest.raises(V.VerifyError, match="InstanceTooLarge"):
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), [[0] * (usable + 1)], t())
    other = VM.PythonBackend(C, 5, pk.vk.cs_degree, params=None)
    with pytest.raises(ValueError):
        V.verify_proof(other, pk.vk, V.SingleVerifier(other), inst, t())


def test_batch_verifier(mods):
    """test_utils.rs:56-70: all proofs of one key through BatchVerifier, falling back to proof-by-proof on failure"""
    PL, V = mods[:2]
    C = pm.Vesta
    cs, fixed, copies, adv, inst = plonk_circuits.standard(PL)
    be = VM.PythonBackend(C, 4, cs.degree())
    pk = PL.keygen(be, cs, fixed, copies)
    proofs = []
    for seed in (1, 2, 3):
        rnd = random.Random(seed)
        proofs.append(PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p)))
    assert len(set(proofs)) == 3
    bv = V.BatchVerifier()
    for pr in proofs:
        bv.add_proof(inst, pr)
    assert bv.finalize(be, pk.vk)
    # deterministic weights are accepted too (zero draws are skipped)
    draws = iter([0, 0, 7, 11, 13])
    bv = V.BatchVerifier(rand=lambda: next(draws))
    for pr in proofs:
        bv.add_proof(inst, pr)
    assert bv.finalize(be, pk.vk)
    # one bad proof spoils the batch, whether it is malformed or merely wrong
    bad = bytearray(proofs[1]); bad[len(bad) - 1] ^= 1
    for spoiled in (bytes(bad), proofs[1][:-64]):
        bv = V.BatchVerifier()
        for pr in (proofs[0], spoiled, proofs[2]):
            bv.add_proof(inst, pr)
        assert not bv.finalize(be, pk.vk)
    assert V.BatchVerifier().finalize(be, pk.vk)              # an empty batch is vacuously fine


def test_msm_accumulator(mods):
    """poly::commitment::MSM: scale / add_msm / the generator-vector part evaluate to what the terms say"""
    PL, V = mods[:2]
    C = pm.Vesta
    be = VM.PythonBackend(C, 3, 3)
    g0, w, u = be.fixed_points()
    rnd = random.Random(9)
    P1, P2 = C.mul(rnd.randrange(be.p), C.G), C.mul(rnd.randrange(be.p), C.G)
    a, b, f = (rnd.randrange(1, be.p) for _ in range(3))
    m = V.MSM(be)
    m.append_term(a, P1)
    m.append_term(b, P2)
    m.add_constant_term(5)
    m.add_to_w_scalar(3)
    m.add_to_u_scalar(4)
    svec = be.ipa_s_vector([2, 3, 5], 7)
    assert svec == [7 * x % be.p for x in (1, 5, 3, 15, 2, 10, 6, 30)]          # s_i = init * prod u_j^(bit (k-1-j) of i)
    m.add_to_g_scalars(svec)
    m.scale(f)
    want = C.best_multiexp([a * f % be.p, b * f % be.p, 5 * f % be.p, 3 * f % be.p, 4 * f % be.p] + [s * f % be.p for s in svec],
                           [P1, P2, g0, w, u] + be.params["g"])
    assert not m.eval() and want is not None
    m.append_term(be.p - 1, want)                                             # subtract the value: now it is the identity
    assert m.eval()
    m2 = V.MSM(be)
    m2.add_msm(m)
    m2.add_msm(m)
    assert m2.eval()
    m2.add_to_u_scalar(1)
    assert not m2.eval()
    assert V.MSM(be).eval()
    assert V.compute_b(3, [2, 5], 1000003) == (1 + 5 * 3) * (1 + 2 * 9) % 1000003


def test_blake2b_read(mods):
    PL, V = mods[:2]
    C = pm.Vesta
    q, p = C.base.p, C.scalar.p
    rnd = random.Random(3)
    pts = [C.mul(rnd.randrange(p), C.G) for _ in range(6)]
    scalars = [rnd.randrange(p) for _ in range(4)]
    w = PL.Blake2bWrite(q, p)
    w.common_scalar(99)
    for pt in pts:
        w.write_point(pt)
    c1 = w.squeeze_challenge_scalar()
    for s in scalars:
        w.write_scalar(s)
    c2 = w.squeeze_challenge_scalar()
    r = V.Blake2bRead(w.finalize(), q, p)
    r.common_scalar(99)
    assert [r.read_point() for _ in pts] == pts and r.squeeze_challenge_scalar() == c1
    assert [r.read_scalar() for _ in scalars] == scalars and r.squeeze_challenge_scalar() == c2
    with pytest.raises(V.VerifyError):
        r.read_scalar()                                                     # the proof is exhausted
    with pytest.raises(V.VerifyError):
        V.Blake2bRead(bytes(32), q, p).read_point()                         # the identity
    with pytest.raises(V.VerifyError):
        V.Blake2bRead((q + 1).to_bytes(32, "little"), q, p).read_point()    # non-canonical x
    with pytest.raises(V.VerifyError):
        V.Blake2bRead(p.to_bytes(32, "little"), q, p).read_scalar()         # non-canonical scalar
    off_curve = next(x for x in range(1, 50) if C.base.sqrt((x ** 3 + 5) % q) is None)
    with pytest.raises(V.VerifyError):
        V.Blake2bRead(off_curve.to_bytes(32, "little"), q, p).read_point()
    for F in (pm.Fp, pm.Fq):                                                # the Tonelli-Shanks helper against the model's sqrt
        for _ in range(20):
            a = rnd.randrange(F.p)
            got, want = V._sqrt(a, F.p), F.sqrt(a)
            assert (got is None) == (want is None) and (got is None or got * got % F.p == a)
        assert V._sqrt(0, F.p) == 0


def test_reference_test_utils_flow(mods):
    """test_utils.py = src/test_utils.rs: `two_programs` of the execution table (exe.rs:1443-1467: ExeCircuit, no public input)
    through gen_proofs_and_verify, and gen_proofs_and_verify_should_fail (the TinyRamCircuit variant of the same flow, with a
    program instance, is tests/test_tinyram_cpu.py::test_real_proof_on_the_cpu_backend_verifies and, on the device, tests/test_gpu_zz_verifier.py)"""
    PL, V, TR, T = mods
    from tiny_ram_halo2_b200 import test_utils as TU
    C = pm.Vesta
    made = []
    def backend_of(k, degree):
        made.append(VM.PythonBackend(C, k, degree))
        return made[-1]
    rnd = random.Random(11)
    rand = lambda: rnd.randrange(C.scalar.p)
    proofs = TU.gen_proofs_and_verify(backend_of, 8, [TP.answer_only(T, 8), TP.load_and_answer(T, 8, 1, 2)], with_prog=False, rand=rand)
    assert len(proofs) == 2 and made[0].k == 6 and len(proofs[0]) == len(proofs[1])
    # test_utils.rs:73-119 with a circuit that CONSTRAINS its public input (one advice cell, assigned independently, is copied from
    # it): proved with the wrong input and checked against the same input -> rejected by the permutation argument; with the
    # right one the helper must complain
    std = plonk_circuits.StandardCircuit()
    right = plonk_circuits.standard(PL)[4]
    TU.gen_proofs_and_verify_should_fail(backend_of, 8, std, [[right[0][0] + 1]], rand=rand)
    assert made[-1].k == 5
    with pytest.raises(AssertionError, match="Erroneously verified proof"):
        TU.gen_proofs_and_verify_should_fail(backend_of, 8, std, right, rand=rand)
    assert len(TU.gen_proofs_and_verify(backend_of, 6, [std, std], rand=rand, public_inputs=[right, right])) == 2
    with pytest.raises(V.VerifyError):
        TU.gen_proofs_and_verify(backend_of, 6, [std, std], rand=rand, public_inputs=[right, [[0]]])
    # the wrong NUMBER of instance columns is create_proof's own error (halo2: Error::InvalidInstances -> "Failed to create proof")
    with pytest.raises(ValueError, match="InvalidInstances"):
        TU.gen_proofs_and_verify_should_fail(backend_of, 8, TP.answer_only(T, 8), [[1]], with_prog=False, rand=rand, k=6)
