"""GPU test of the product's own verifier (tiny-ram-halo2_b200/verifier.py) over plonk.GpuBackend: the reference's
gen_proofs_and_verify flow end to end on the device (/root/reference/src/test_utils.rs:6-71) -- Params::new, keygen, two
proofs of the real TinyRamCircuit, BatchVerifier, then SingleVerifier proof by proof -- held against the oracle's independent
verifier (the verifier's logic itself is covered on the CPU by tests/test_verifier_cpu.py).  First device run: the driver's
round-1 GPU tier, where it passed; the xfail it carried until then is gone."""
import random

import pytest

from util import pm

import tinyram_programs as TP
import verify_util as VU

pytestmark = pytest.mark.gpu
C = pm.Vesta


def test_gen_proofs_and_verify_on_the_device():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from tiny_ram_halo2_b200 import plonk as PL, tinyram as TR, trace as T, verifier as V
    ctx = pkg.Context(0, pkg.VESTA)
    W, k = 8, 6
    be = PL.GpuBackend(ctx, k, 6)
    try:
        traces = [TP.answer_only(T, W), TP.load_and_answer(T, W, 1, 2)]
        circ, fixed, copies, _, _ = TR.build(PL, traces[0], k, keygen_from_empty_circuit=True)
        pk = PL.keygen(be, circ.cs, fixed, copies)
        rnd = random.Random(3)
        proofs = []
        for tr in traces:
            _, _, _, adv, inst = TR.build(PL, tr, k)
            proofs.append((PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p)), inst))
        # the pieces first, so that a failure names the method
        g0, w, u = be.fixed_points()
        params = VU.oracle_params(be)
        assert (g0, w, u) == (params["g"][0], params["w"], params["u"])
        us = [rnd.randrange(1, be.p) for _ in range(k)]
        s = be._ints(be.ipa_s_vector(us, 7).cpu().numpy().view(be.np.uint64))
        want = [7]
        for u_j in reversed(us):
            want = want + [v * u_j % be.p for v in want]
        assert s == want
        pts = [C.mul(rnd.randrange(be.p), C.G) for _ in range(5)]
        sc = [rnd.randrange(be.p) for _ in range(5)]
        assert be.msm_points(sc, pts) == C.best_multiexp(sc, pts)
        assert be.msm_points([1, be.p - 1], [pts[0], pts[0]]) is None
        # test_utils.rs:56-70
        bv = V.BatchVerifier()
        for proof, inst in proofs:
            bv.add_proof(inst, proof)
        assert bv.finalize(be, pk.vk)
        for proof, inst in proofs:
            assert V.verify_proof(be, pk.vk, V.SingleVerifier(be), inst, V.Blake2bRead(proof, be.q, be.p)) is None
            ok, err = VU.verify(be, pk.vk, inst, proof)
            assert ok, err
        # test_utils.rs:104-118: the wrong public input, and a flipped byte
        with pytest.raises(V.VerifyError):
            V.verify_proof(be, pk.vk, V.SingleVerifier(be), proofs[1][1], V.Blake2bRead(proofs[0][0], be.q, be.p))
        bad = bytearray(proofs[0][0]); bad[len(bad) // 2] ^= 4
        with pytest.raises(V.VerifyError):
            V.verify_proof(be, pk.vk, V.SingleVerifier(be), proofs[0][1], V.Blake2bRead(bytes(bad), be.q, be.p))
        bv = V.BatchVerifier()
        bv.add_proof(proofs[0][1], bytes(bad))
        bv.add_proof(proofs[1][1], proofs[1][0])
        assert not bv.finalize(be, pk.vk)
    finally:
        be.close()
    # and the same flow through the package's mirror of src/test_utils.rs (its own backend, keys and OsRng-style blinding)
    from tiny_ram_halo2_b200 import test_utils as TU
    made = []
    def backend_of(k_, degree):
        made.append(PL.GpuBackend(ctx, k_, degree))
        return made[-1]
    import plonk_circuits
    from tiny_ram_halo2_b200.lookup import ConstraintSystemFailure
    wrong = TR.program_instance([T.Answer(T.Imm(0))], W)
    try:
        out = TU.gen_proofs_and_verify(backend_of, W, traces)
        assert len(out) == 2 and made[0].k == 2 + W // 2
        # keys from the first circuit itself: the execution table's selectors are ON, every exe gate and lookup is live
        assert len(TU.gen_proofs_and_verify(backend_of, W, traces, keygen_from_empty_circuit=False)) == 2
        # test_utils.rs:73-119, proved WITH the wrong input and checked against the same: a circuit that constrains its instance
        std = plonk_circuits.StandardCircuit()
        right = plonk_circuits.standard(PL)[4]
        TU.gen_proofs_and_verify_should_fail(backend_of, 8, std, [[right[0][0] + 1]])
        with pytest.raises(AssertionError, match="Erroneously verified proof"):
            TU.gen_proofs_and_verify_should_fail(backend_of, 8, std, right)
        # TinyRamCircuit with another program as its public input: the program table is assigned from the instance, and the
        # execution table's program-line lookup (gated by the ADVICE selector s_trace, prog.rs:170-192, so it is live even with the
        # reference's keys from the empty circuit) has no match: create_proof itself fails (halo2: "Failed to create proof")
        for from_empty in (False, True):
            with pytest.raises(ConstraintSystemFailure):
                TU.gen_proofs_and_verify_should_fail(backend_of, W, traces[0], wrong, k=6, keygen_from_empty_circuit=from_empty)
    finally:
        for b in made:
            b.close()
