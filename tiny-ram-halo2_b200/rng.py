"""A blinding RNG that a Rust run can reproduce: the parity hook of DESIGN.md section 4.

halo2 draws every random scalar as `Scalar::random(&mut rng)`, which pasta_curves 0.4.1 implements as from_u512 of eight
`rng.next_u64()` words -- 64 bytes of the generator's output, little endian, reduced mod p.  ScalarStreamRng is such a generator
with a portable definition: the keystream of AES-256-CTR (key = the 32-byte seed, 16-byte big-endian counter starting at 0)
read sequentially, 64 bytes per scalar.  rust/parity/parity.rs implements the same stream as a `rand_core::RngCore`
(next_u64 = the next 8 keystream bytes, little endian), so that plonk.create_proof here and halo2's create_proof there see the
same blinding scalars in the same order -- IF the draw order restated in plonk.py is halo2's, which is
exactly what a byte comparison of the two proofs pins.

It deliberately has no bulk interface (`vector`, `bulk_key`): the random polynomials are then drawn scalar by scalar, as
halo2 draws them.  sharded_backend.ShardedRng is the fast variant for multi-GPU proving (not reproducible from Rust)."""
from __future__ import annotations


class ScalarStreamRng:
    def __init__(self, p: int, seed: bytes):
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
        if len(seed) != 32:
            raise ValueError("the seed is 32 bytes")
        self.p, self.draws = p, 0
        self._enc = Cipher(algorithms.AES(seed), modes.CTR(bytes(16))).encryptor()
        self._buf, self._off = b"", 0

    def next_bytes(self, k: int) -> bytes:
        if self._off + k > len(self._buf):
            self._buf = self._buf[self._off:] + self._enc.update(bytes(max(1 << 16, k)))
            self._off = 0
        out = self._buf[self._off:self._off + k]
        self._off += k
        return out

    def __call__(self) -> int:
        self.draws += 1
        return int.from_bytes(self.next_bytes(64), "little") % self.p
