"""GPU parity for SURVEY.md 8(f) row f4: plonk.create_proof over the CUDA backend (GpuBackend: every commitment, transform,
quotient evaluation, grand product, lookup permutation, evaluation, Kate division and the IPA on the device through the C ABI)
must produce the SAME PROOF BYTES as the identical host logic over the oracle's PythonBackend, and the oracle's independent
verify_proof must accept them."""
import random

import numpy as np
import pytest

from util import O, pm

import plonk_model as VM
import plonk_circuits

pytestmark = pytest.mark.gpu

CURVE_OF = {O.VESTA: pm.Vesta, O.PALLAS: pm.Pallas}


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="module")
def ctxs(pkg):
    return {O.VESTA: pkg.Context(0, pkg.VESTA), O.PALLAS: pkg.Context(0, pkg.PALLAS)}


def _params_as_oracle(be):
    """the GPU backend's Params (tested against params_model in test_gpu_params.py) in the oracle's representation"""
    pts = lambda arr: [None if not r.any() else tuple(be._ints(r.reshape(2, 4), be.q, be.Rqinv)) for r in np.asarray(arr).reshape(-1, 8)]
    prm = be.params
    return {"k": be.k, "n": be.n, "g": pts(prm.g_points), "g_lagrange": pts(prm.g_lagrange_points), "w": pts(prm.w)[0], "u": pts(prm.u)[0]}


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("kw", [dict(with_lookup=True), dict(with_lookup=False), dict(with_lookup=True, wide_lookup=True)],
                         ids=["lookup", "no-lookup", "wide-lookup"])
def test_proof_bytes_equal_cpu_backend_and_verify(pkg, ctxs, curve, kw):
    PL = pkg.plonk
    C = CURVE_OF[curve]
    k = 4
    cs, fixed, copies, adv, inst = plonk_circuits.standard(PL, **kw)
    proofs, vks = [], []
    for be in (PL.GpuBackend(ctxs[curve], k, cs.degree()), VM.PythonBackend(C, k, cs.degree())):
        pk = PL.keygen(be, cs, fixed, copies)
        rnd = random.Random(5)
        proofs.append(PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p)))
        vks.append(pk.vk)
    assert vks[0].fixed_commitments == vks[1].fixed_commitments
    assert vks[0].permutation_commitments == vks[1].permutation_commitments
    assert vks[0].transcript_repr == vks[1].transcript_repr
    assert proofs[0] == proofs[1]
    assert VM.verify_proof(C, be.params, vks[0], inst, proofs[0])
    assert not VM.verify_proof(C, be.params, vks[0], [[inst[0][0] + 1]], proofs[0])


@pytest.mark.parametrize("k", [7, 10])
def test_larger_domains_verify(pkg, ctxs, k):
    """the same circuit padded to 2^k rows with more witness rows: GPU proof accepted by the oracle verifier"""
    PL = pkg.plonk
    C = pm.Vesta
    cs, fixed, copies, adv, inst = plonk_circuits.standard(PL)
    be = PL.GpuBackend(ctxs[O.VESTA], k, cs.degree())
    usable = be.n - (cs.blinding_factors() + 1)
    rnd = random.Random(k)
    # fill the remaining usable rows with random satisfied add / mul rows whose `a` stays inside the lookup table
    for r in range(4, usable):
        a, b = rnd.randrange(8), rnd.randrange(C.scalar.p)
        add = rnd.random() < 0.5
        adv[0].append(a); adv[1].append(b); adv[2].append((a + b) % C.scalar.p if add else a * b % C.scalar.p)
        fixed[0].append(int(add)); fixed[1].append(int(not add)); fixed[2].append(0); fixed[4].append(1)
    copies = list(copies) + [((PL.ADVICE, 0, r), (PL.ADVICE, 0, r2)) for r, r2 in [(10, 11)] if adv[0][10] == adv[0][11]]
    pk = PL.keygen(be, cs, fixed, copies)
    proof = PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p))
    params = _params_as_oracle(be)
    assert VM.verify_proof(C, params, pk.vk, inst, proof)
    bad = bytearray(proof); bad[100] ^= 4
    assert not VM.verify_proof(C, params, pk.vk, inst, bytes(bad))


def test_lookup_failure_and_argument_errors(pkg, ctxs):
    PL = pkg.plonk
    C = pm.Vesta
    cs, fixed, copies, adv, inst = plonk_circuits.standard(PL)
    be = PL.GpuBackend(ctxs[O.VESTA], 4, cs.degree())
    pk = PL.keygen(be, cs, fixed, copies)
    rnd = random.Random(1)
    adv[0][3] = 9; adv[1][3] = 9
    with pytest.raises(ValueError):          # lookup.ConstraintSystemFailure
        PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p))
    with pytest.raises(ValueError):
        PL.create_proof(be, pk, [], adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p))
    with pytest.raises(ValueError):
        PL.keygen(PL.GpuBackend(ctxs[O.VESTA], 2, cs.degree()), cs, fixed, copies)      # NotEnoughRowsAvailable (n = 4 < 8)


