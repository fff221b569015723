//! parity.rs -- pins this repo's prover to the Rust prover byte for byte.  FOR THE FIRST PERSON WITH CARGO.
//!
//! NOT COMPILED in the build image (no rustc there) and written from the public API of halo2_proofs 0.2.0:
//! expect to fix an import or two.  What it does is small and fixed:
//!
//!   1. build the reference's `TinyRamCircuit<8, 8>` for the two programs of its own `two_programs` test
//!      (/root/reference/src/circuits/mod.rs:377-410): `Answer 1` and `load_and_answer(1, 2)`;
//!   2. Params::new(6), keygen from `TinyRamCircuit::default()` -- exactly src/test_utils.rs:20-25;
//!   3. create_proof with a FIXED-SEED blinding RNG instead of OsRng (src/test_utils.rs:46): `ScalarStream`, the keystream of
//!      AES-256-CTR(key = bytes 0..31, counter block 0) read 8 bytes per next_u64, little endian -- the generator
//!      tiny-ram-halo2_b200/rng.py implements on the other side;
//!   4. print, per circuit, the fields of tests/golden/parity_vectors.json: proof length, SHA-256, transcript_repr, the first
//!      32 proof bytes, the column / lookup / degree counts halo2 sees.
//!
//! How to run: copy this file to `<tiny-ram-halo2>/tests/parity.rs`, add to that crate's [dev-dependencies]
//!     aes = "0.8"   ctr = "0.9"   sha2 = "0.10"   rand_core = "0.6"   hex = "0.4"
//! and `cargo test --test parity -- --nocapture`.
//!
//! Reading the result against tests/golden/parity_vectors.json (made by tests/golden/make_parity_vectors.py):
//!   * `shape` differs            -> the restated circuit (tinyram.py) has a different constraint system than the fork builds;
//!                                   the fork-only `lookup_dynamic` (src/circuits/tables/prog.rs:163-193), modelled there as
//!                                   [s, s * e_i] in [tag, col_i], is the first suspect.
//!   * shape equal, digests differ -> rerun `python tests/golden/make_parity_vectors.py --transcript-repr <the value printed
//!                                   here>`: transcript_repr hashes Rust's `{:?}` of the pinned key and cannot be restated
//!                                   without the crate.  Still different: compare `first_proof_point_le_hex` (the first advice
//!                                   commitment: wrong => witness, blinding-row order or Params differ) and then walk the proof.
//!   * digests equal               -> parity is pinned; record it in DESIGN.md section 4.
use aes::cipher::{KeyIvInit, StreamCipher};
use halo2_proofs::pasta::{vesta, EqAffine, Fp};
use halo2_proofs::plonk::{create_proof, keygen_pk, keygen_vk, verify_proof, SingleVerifier};
use halo2_proofs::poly::commitment::Params;
use halo2_proofs::transcript::{Blake2bRead, Blake2bWrite, Challenge255};
use rand_core::{CryptoRng, Error, RngCore};
use sha2::{Digest, Sha256};
use tiny_ram_halo2::circuits::tables::prog::program_instance;
use tiny_ram_halo2::circuits::TinyRamCircuit;
use tiny_ram_halo2::instructions::*;
use tiny_ram_halo2::trace::*;

type Aes256Ctr = ctr::Ctr128BE<aes::Aes256>;

/// The keystream of AES-256-CTR as an RngCore: next_u64 = the next 8 bytes, little endian.
/// pasta_curves 0.4.1 `Field::random` = from_u512([next_u64(); 8]) = 64 keystream bytes, little endian, mod p.
struct ScalarStream { c: Aes256Ctr, draws: u64 }
impl ScalarStream {
    fn new(seed: [u8; 32]) -> Self { ScalarStream { c: Aes256Ctr::new(&seed.into(), &[0u8; 16].into()), draws: 0 } }
}
impl RngCore for ScalarStream {
    fn next_u32(&mut self) -> u32 { self.next_u64() as u32 }   // never used by Field::random; consumes 8 bytes like next_u64
    fn next_u64(&mut self) -> u64 {
        let mut b = [0u8; 8];
        self.c.apply_keystream(&mut b);
        self.draws += 1;
        u64::from_le_bytes(b)
    }
    fn fill_bytes(&mut self, dest: &mut [u8]) { for d in dest.iter_mut() { *d = 0; } self.c.apply_keystream(dest); }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), Error> { self.fill_bytes(dest); Ok(()) }
}
impl CryptoRng for ScalarStream {}

fn answer_only() -> Trace<8, 8> {
    let t = Program(vec![Instruction::Answer(Answer { a: ImmediateOrRegName::Immediate(Word(1)) })]).eval::<8, 8>(Mem::new(&[], &[]));
    assert_eq!(t.ans.0, 1);
    t
}

fn load_and_answer(a: u32, b: u32) -> Trace<8, 8> {          // src/circuits/mod.rs:88-110
    let prog = Program(vec![
        Instruction::LoadW(LoadW { ri: RegName(0), a: ImmediateOrRegName::Immediate(Word(b)) }),
        Instruction::And(And { ri: RegName(1), rj: RegName(0), a: ImmediateOrRegName::Immediate(Word(a)) }),
        Instruction::Answer(Answer { a: ImmediateOrRegName::Immediate(Word(1)) }),
    ]);
    let t = prog.eval::<8, 8>(Mem::new(&[Word(0b1)], &[]));
    assert_eq!(t.ans.0, 1);
    t
}

#[test]
fn parity_vectors() {
    let k = 2 + 8 / 2;                                        // src/test_utils.rs:20
    let params: Params<EqAffine> = Params::new(k);
    let empty = TinyRamCircuit::<8, 8>::default();
    let vk = keygen_vk(&params, &empty).unwrap();
    let pk = keygen_pk(&params, vk.clone(), &empty).unwrap();
    // what halo2 hashes first into every transcript; Debug prints the pinned key, whose last field is transcript_repr
    println!("pinned verification key: {:?}", vk.pinned());
    let cs = vk.cs();
    println!("shape: advice {} instance {} fixed {} degree {} blinding_factors {}  (lookups / gates: see the pinned key above)",
             cs.num_advice_columns(), cs.num_instance_columns(), cs.num_fixed_columns(), cs.degree(), cs.blinding_factors());
    for (name, trace) in [("answer_only", answer_only()), ("load_and_answer(1, 2)", load_and_answer(1, 2))] {
        let instance: Vec<Vec<Fp>> = program_instance::<8, 8, Fp>(trace.prog.clone());
        let instance_refs: Vec<&[Fp]> = instance.iter().map(|c| c.as_slice()).collect();
        let circuit = TinyRamCircuit::<8, 8> { trace: Some(trace) };
        let mut seed = [0u8; 32];
        for (i, s) in seed.iter_mut().enumerate() { *s = i as u8; }
        let mut rng = ScalarStream::new(seed);
        let mut transcript = Blake2bWrite::<_, vesta::Affine, Challenge255<_>>::init(vec![]);
        create_proof(&params, &pk, &[circuit], &[instance_refs.as_slice()], &mut rng, &mut transcript).expect("Failed to create proof");
        let proof: Vec<u8> = transcript.finalize();
        let mut rd = Blake2bRead::<_, vesta::Affine, Challenge255<_>>::init(&proof[..]);
        verify_proof(&params, pk.get_vk(), SingleVerifier::new(&params), &[instance_refs.as_slice()], &mut rd).expect("could not verify_proof");
        println!("{{\"circuit\": \"{}\", \"k\": {}, \"proof_bytes\": {}, \"proof_sha256\": \"{}\", \"scalars_drawn\": {}, \"first_proof_point_le_hex\": \"{}\"}}",
                 name, k, proof.len(), hex::encode(Sha256::digest(&proof)), rng.draws / 8, hex::encode(&proof[..32]));
    }
}
