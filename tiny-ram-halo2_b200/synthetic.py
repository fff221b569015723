"""Synthetic MSM bases for benchmarks and large-size tests: an arithmetic progression
P_i = P0 + i*D of two fixed pseudo-random multiples of the curve generator, generated ON THE DEVICE by
trp_dev_points_progression.  P0 and D below are k*G for k = blake2b("tinyram-b200 synthetic {P0,D} <curve>") mod r,
stored as affine Montgomery limbs (x[4], y[4]); they were produced once with oracle/pasta_model.py."""
import numpy as np

from ._lib import PALLAS, VESTA

_PALLAS_P0 = [0x2ba3a2b0d2669b4b, 0xd31e3789bcccf5e7, 0xaf5226447a0d9929, 0x7f63a19905f28f4, 0xacb5101b88d08c50, 0x2cdc358d938d6d78, 0xf7e5b817c0c404b1, 0x20f80cbfd482e84f]
_PALLAS_D = [0x48dbe303ca8ccd1d, 0xbde22858dd83293e, 0x5c0d33eaeadc830b, 0x1370a30999151b94, 0x1ff42d1750a6fa47, 0xcd5cc17cb40888a9, 0x6da4af1a8b0f6dec, 0x3ede18f76ae59715]
_VESTA_P0 = [0x58cb2bc95e6cb585, 0xd14b3591158325ca, 0xfadb8c3832d2f3ef, 0x9ce523183bd695d, 0x624a74fedac6bb99, 0x225a53cb3223d027, 0xc3be6fe4ca290519, 0x3c44285f2be60a32]
_VESTA_D = [0x8eb6abedb1b70397, 0x9f0a950e59161a18, 0x3e3c28aafcb75175, 0x5bed9b88efa112, 0x51a8030aed95b0df, 0xa76e040011fa7f60, 0x36d449eecce4a470, 0x3b4ad1d022f5afb9]

SEEDS = {
    PALLAS: (np.array(_PALLAS_P0, dtype=np.uint64), np.array(_PALLAS_D, dtype=np.uint64)),
    VESTA: (np.array(_VESTA_P0, dtype=np.uint64), np.array(_VESTA_D, dtype=np.uint64)),
}


def device_points(ctx, n, d_out_ptr):
    """Fill device memory at d_out_ptr (n x 64 B) with P0 + i*D, i < n, on ctx's curve."""
    from ._lib import ptr
    p0, d = SEEDS[ctx.curve]
    ctx.check(ctx.lib.trp_dev_points_progression(ctx.handle, ptr(p0), ptr(d), n, d_out_ptr))


def random_scalars(n, seed, m=None):
    """Uniform field elements in [0, 2^254) as (n, 4) (or (m, n, 4)) uint64 limbs, PCG64(seed).  Any value below the
    modulus is a valid Montgomery representation, so no conversion is needed for uniform inputs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    shape = (n, 4) if m is None else (m, n, 4)
    a = rng.integers(0, 1 << 64, size=shape, dtype=np.uint64)
    a[..., 3] &= np.uint64((1 << 62) - 1)
    return a
