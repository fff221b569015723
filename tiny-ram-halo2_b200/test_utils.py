"""Mirror of the reference's src/test_utils.rs -- the only place where the reference runs the real prover:

    gen_proofs_and_verify::<WORD_BITS, C>(inputs)                        test_utils.rs:6-71
    gen_proofs_and_verify_should_fail::<WORD_BITS, C>(circuit, input)    test_utils.rs:73-119

Same steps in the same order: k from the word size, Params::new(k), ONE key pair generated from the EMPTY circuit
(`C::default()`: for TinyRamCircuit that leaves the execution table's selectors off, tinyram.build(keygen_from_empty_circuit=
True)), one proof per (circuit, public input) pair with a fresh Blake2b transcript, all proofs through BatchVerifier, and, if the
batch does not verify, proof by proof through SingleVerifier (which raises).  `backend_of(k, cs_degree)` supplies the backend:
plonk.GpuBackend in the product, the oracle's PythonBackend in the CPU tests.  The circuits are given as traces (the public
input of a trace is program_instance(trace.prog), circuits/mod.rs:391-406, or nothing for ExeCircuit, exe.rs:1459-1467) or as
any object with `build(PL, k, public_input=None, keygen_from_empty_circuit=False) -> (cs, fixed, copies, advice, instances)`
(the reference's helpers are generic over `C: Circuit<Fp> + Default + Clone`).

A public input given explicitly is what create_proof RECEIVES as the instance columns (test_utils.rs:41-49, 96-104), with the
cells a circuit assigns from the instance (assign_advice_from_instance: TinyRamCircuit's program table, prog.rs:206-216) filled
from it, and the proof is verified against the same input -- so a rejection shows that the CIRCUIT constrains its instance, not
merely that the verifier absorbs it into the transcript."""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

from . import plonk as PL, tinyram as TR, verifier as V


def _os_rng(p: int) -> Callable[[], int]:
    return lambda: int.from_bytes(os.urandom(64), "little") % p          # OsRng + Field::random (from_bytes_wide)


class TraceCircuit:
    """`TinyRamCircuit { trace }` (with_prog) or `ExeCircuit { trace }` as a subject of the helpers below"""

    def __init__(self, trace, with_prog: bool = True):
        self.trace, self.with_prog = trace, with_prog

    def build(self, PL_, k: int, public_input=None, keygen_from_empty_circuit: bool = False):
        circ, fixed, copies, advice, instances = TR.build(PL_, self.trace, k, keygen_from_empty_circuit=keygen_from_empty_circuit,
                                                          with_prog=self.with_prog)
        if public_input is not None:
            instances = [list(c) for c in public_input]
            if self.with_prog and len(instances) == circ.cs.num_instance:
                circ.assign_instance(advice, instances)      # the program table is assigned FROM the instance
        return circ.cs, fixed, copies, advice, instances


def _subject(x, with_prog: bool):
    return x if hasattr(x, "build") else TraceCircuit(x, with_prog)


def _setup(backend_of, subject, k: int, keygen_from_empty_circuit: bool = True):
    cs, fixed, copies, _, _ = subject.build(PL, k, keygen_from_empty_circuit=keygen_from_empty_circuit)
    be = backend_of(k, cs.degree())
    return be, PL.keygen(be, cs, fixed, copies)


def _prove(be, pk, subject, k: int, rand, public_input=None):
    _, _, _, advice, instances = subject.build(PL, k, public_input=public_input)
    proof = PL.create_proof(be, pk, instances, advice, rand, PL.Blake2bWrite(be.q, be.p))      # .expect("Failed to create proof")
    return proof, instances


def gen_proofs_and_verify(backend_of: Callable, word_bits: int, traces: Sequence, with_prog: bool = True,
                          rand: Optional[Callable[[], int]] = None, public_inputs: Optional[Sequence] = None,
                          k: Optional[int] = None, keygen_from_empty_circuit: bool = True) -> List[bytes]:
    """Returns the proofs.  Raises verifier.VerifyError ("could not verify_proof") if one of them does not verify.
    public_inputs[i], if given, is the instance create_proof receives for circuit i and the one the proof is checked against
    (default: the circuit's own); k overrides the reference's 2 + WORD_BITS / 2 (BASELINE.json's k = 20 for word size 32);
    keygen_from_empty_circuit=False generates the keys from the first circuit itself instead of `C::default()` (for
    TinyRamCircuit the empty circuit leaves the execution table's selectors off, so only then are the exe gates live)."""
    k = k or 2 + word_bits // 2
    subjects = [_subject(t, with_prog) for t in traces]
    be, pk = _setup(backend_of, subjects[0], k, keygen_from_empty_circuit)
    rand = rand or _os_rng(be.p)
    proofs = [_prove(be, pk, sub, k, rand, None if public_inputs is None else public_inputs[i]) for i, sub in enumerate(subjects)]
    batch = V.BatchVerifier()
    for proof, inst in proofs:
        batch.add_proof(inst, proof)
    if not batch.finalize(be, pk.vk):
        for proof, inst in proofs:
            V.verify_proof(be, pk.vk, V.SingleVerifier(be), inst, V.Blake2bRead(proof, be.q, be.p))
        raise V.VerifyError("the batch was rejected although every proof verifies on its own")
    return [proof for proof, _ in proofs]


def gen_proofs_and_verify_should_fail(backend_of: Callable, word_bits: int, trace, public_input, with_prog: bool = True,
                                      rand: Optional[Callable[[], int]] = None, k: Optional[int] = None,
                                      keygen_from_empty_circuit: bool = True) -> None:
    """One proof MADE WITH `public_input` as its instance and checked against the same input (test_utils.rs:96-118); raises
    AssertionError("Erroneously verified proof") if it verifies.  An exception out of create_proof (a lookup input missing
    from its table, the wrong number of instance columns) propagates, as the reference's `.expect("Failed to create proof")`.
    k = 1 + WORD_BITS / 2 as in the reference (test_utils.rs:88), which is what its standalone gadget circuits need;
    TinyRamCircuit needs the 2 + WORD_BITS / 2 of gen_proofs_and_verify (pass k)."""
    k = k or 1 + word_bits // 2
    subject = _subject(trace, with_prog)
    be, pk = _setup(backend_of, subject, k, keygen_from_empty_circuit)
    proof, inst = _prove(be, pk, subject, k, rand or _os_rng(be.p), public_input)
    try:
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), inst, V.Blake2bRead(proof, be.q, be.p))
    except V.VerifyError:
        return
    raise AssertionError("Erroneously verified proof")
