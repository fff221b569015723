"""Mirror of the reference's src/test_utils.rs -- the only place where the reference runs the real prover:

    gen_proofs_and_verify::<WORD_BITS, C>(inputs)                        test_utils.rs:6-71
    gen_proofs_and_verify_should_fail::<WORD_BITS, C>(circuit, input)    test_utils.rs:73-119

Same steps in the same order: k from the word size, Params::new(k), ONE key pair generated from the EMPTY circuit
(`C::default()`: for TinyRamCircuit that leaves the execution table's selectors off, tinyram.build(keygen_from_empty_circuit=
True)), one proof per (circuit, public input) pair with a fresh Blake2b transcript, all proofs through BatchVerifier, and, if the
batch does not verify, proof by proof through SingleVerifier (which raises).  `backend_of(k, cs_degree)` supplies the backend:
plonk.GpuBackend in the product, the oracle's PythonBackend in the CPU tests.  The circuits are given as traces, the public
input of a trace is program_instance(trace.prog) (circuits/mod.rs:391-406), or nothing for ExeCircuit (exe.rs:1459-1467)."""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

from . import plonk as PL, tinyram as TR, verifier as V


def _os_rng(p: int) -> Callable[[], int]:
    return lambda: int.from_bytes(os.urandom(64), "little") % p          # OsRng + Field::random (from_bytes_wide)


def _setup(backend_of, traces, k: int, with_prog: bool):
    circ, fixed, copies, _, _ = TR.build(PL, traces[0], k, keygen_from_empty_circuit=True, with_prog=with_prog)
    be = backend_of(k, circ.cs.degree())
    return be, PL.keygen(be, circ.cs, fixed, copies)


def _prove(be, pk, trace, k: int, with_prog: bool, rand):
    _, _, _, advice, instances = TR.build(PL, trace, k, with_prog=with_prog)
    proof = PL.create_proof(be, pk, instances, advice, rand, PL.Blake2bWrite(be.q, be.p))
    return proof, instances


def gen_proofs_and_verify(backend_of: Callable, word_bits: int, traces: Sequence, with_prog: bool = True,
                          rand: Optional[Callable[[], int]] = None, public_inputs: Optional[Sequence] = None,
                          k: Optional[int] = None) -> List[bytes]:
    """Returns the proofs.  Raises verifier.VerifyError ("could not verify_proof") if one of them does not verify.
    public_inputs overrides the instance columns a proof is CHECKED against (default: the ones it was made for); k overrides
    the reference's 2 + WORD_BITS / 2 (BASELINE.json's k = 20 for word size 32)."""
    k = k or 2 + word_bits // 2
    be, pk = _setup(backend_of, traces, k, with_prog)
    rand = rand or _os_rng(be.p)
    proofs = [_prove(be, pk, tr, k, with_prog, rand) for tr in traces]
    checked = [(proof, inst if public_inputs is None else public_inputs[i]) for i, (proof, inst) in enumerate(proofs)]
    batch = V.BatchVerifier()
    for proof, inst in checked:
        batch.add_proof(inst, proof)
    if not batch.finalize(be, pk.vk):
        for proof, inst in checked:
            V.verify_proof(be, pk.vk, V.SingleVerifier(be), inst, V.Blake2bRead(proof, be.q, be.p))
        raise V.VerifyError("the batch was rejected although every proof verifies on its own")
    return [proof for proof, _ in proofs]


def gen_proofs_and_verify_should_fail(backend_of: Callable, word_bits: int, trace, public_input, with_prog: bool = True,
                                      rand: Optional[Callable[[], int]] = None, k: Optional[int] = None) -> None:
    """One proof checked against `public_input`; raises AssertionError("Erroneously verified proof") if it verifies.
    k = 1 + WORD_BITS / 2 as in the reference (test_utils.rs:88), which is what its standalone gadget circuits need;
    TinyRamCircuit needs the 2 + WORD_BITS / 2 of gen_proofs_and_verify (pass k)."""
    k = k or 1 + word_bits // 2
    be, pk = _setup(backend_of, [trace], k, with_prog)
    proof, _ = _prove(be, pk, trace, k, with_prog, rand or _os_rng(be.p))
    try:
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), public_input, V.Blake2bRead(proof, be.q, be.p))
    except V.VerifyError:
        return
    raise AssertionError("Erroneously verified proof")
