"""Golden vectors for the FIRST PERSON WITH CARGO: proofs of two circuits made
by THIS repo's create_proof under a blinding RNG that Rust can reproduce (tiny-ram-halo2_b200/rng.py: AES-256-CTR keystream,
64 bytes per scalar), frozen as SHA-256 digests in tests/golden/parity_vectors.json.  rust/parity/parity.rs makes the same two
proofs with halo2_proofs 0.2.0 (the reference's pinned fork) and prints the same fields; equal digests pin this repo's prover
to the Rust prover byte for byte.

Two values differ by construction and are handled explicitly:
  * transcript_repr: halo2 hashes the `{:?}` rendering of its PinnedVerificationKey, which cannot be restated without the
    crate.  parity.rs prints halo2's value; `--transcript-repr 0x..` makes this script use it instead of its own.
  * lookup_dynamic: the fork's dynamic lookup is modelled as [s, s * e_i] in [tag, col_i] (tinyram.py docstring).  If the
    digests differ although transcript_repr was supplied, this is the first thing to check (parity.rs prints the number of
    advice / fixed / instance columns, lookups and the degree halo2 sees, to compare with the `shape` recorded here).

  python tests/golden/make_parity_vectors.py [--transcript-repr HEX] [--write]
Runs on the CPU (the oracle's PythonBackend under plonk.create_proof; the GPU backend produces the same bytes:
tests/test_gpu_tinyram.py)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.load_package()
from tiny_ram_halo2_b200 import plonk as PL, tinyram as TR, trace as T, rng as RNG
import pasta_model as pm
import plonk_model as VM
import tinyram_programs as TP

SEED = bytes(range(32))
OUT = os.path.join(HERE, "parity_vectors.json")


def prove(name, trace, k, repr_override):
    C = pm.Vesta
    circ, fixed, copies, adv, inst = TR.build(PL, trace, k, keygen_from_empty_circuit=True)       # test_utils.rs:22-25
    be = VM.PythonBackend(C, k, circ.cs.degree())
    pk = PL.keygen(be, circ.cs, fixed, copies)
    own_repr = pk.vk.transcript_repr
    if repr_override is not None:
        pk.vk.transcript_repr = repr_override
    rng = RNG.ScalarStreamRng(C.scalar.p, SEED)
    proof = PL.create_proof(be, pk, inst, adv, rng, PL.Blake2bWrite(C.base.p, C.scalar.p))
    assert VM.verify_proof(C, be.params, pk.vk, inst, proof), VM.verify_proof.last_error
    cs = circ.cs
    first = lambda pt: None if pt is None else "%064x" % pt[0]
    return {"circuit": name, "word_bits": trace.word_bits, "k": k, "seed_hex": SEED.hex(), "keygen": "from TinyRamCircuit::default() (test_utils.rs:22-25)",
            "transcript_repr_used": hex(pk.vk.transcript_repr), "transcript_repr_of_this_repo": hex(own_repr),
            "proof_bytes": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest(), "scalars_drawn": rng.draws,
            "first_proof_point_le_hex": proof[:32].hex(),
            "fixed_commitment_x": [first(pt) for pt in pk.vk.fixed_commitments[:3]],
            "shape": {"advice": cs.num_advice, "instance": cs.num_instance, "fixed": cs.num_fixed, "lookups": len(cs.lookups),
                      "gates": len(cs.gates), "equality_columns": len(cs.permutation), "degree": cs.degree(), "blinding_factors": cs.blinding_factors()}}


def main():
    rep = None
    if "--transcript-repr" in sys.argv:
        rep = int(sys.argv[sys.argv.index("--transcript-repr") + 1], 16)
    vectors = [prove("answer_only", TP.answer_only(T, 8), 6, rep), prove("load_and_answer(1, 2)", TP.load_and_answer(T, 8, 1, 2), 6, rep)]
    text = json.dumps(vectors, indent=1)
    print(text)
    if "--write" in sys.argv:
        with open(OUT, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
