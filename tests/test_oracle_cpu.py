"""CPU tests (no GPU): the C++ oracle against the committed known-answer vectors and the Python big-int model,
the pinned pasta constants, and host-side logic of the package."""
import json
import os
import random

import numpy as np
import pytest

from util import O, pm, make_points, scalars_uniform, scalars_tinyram

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "pasta_kat.json")))
FID = {"Fp": O.FP, "Fq": O.FQ}
CID = {"pallas": O.PALLAS, "vesta": O.VESTA}


def H(x):
    return int(x, 16)


def mont(fid, ints):
    return O.to_mont(fid, O.ints_to_limbs(ints))


def unmont(fid, arr):
    return O.limbs_to_ints(O.from_mont(fid, arr))


def test_appendix_a_constants():
    """SURVEY.md Appendix A (pasta_curves 0.4.1 constants, re-derived): moduli, R, R2, R3, INV, ROOT_OF_UNITY, ZETA."""
    fp = KAT["fields"]["Fp"]; fq = KAT["fields"]["Fq"]
    assert H(fp["modulus"]) == 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
    assert H(fq["modulus"]) == 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
    assert pm.Fp.limbs(H(fp["R"])) == [0x34786d38fffffffd, 0x992c350be41914ad, 0xffffffffffffffff, 0x3fffffffffffffff]
    assert pm.Fp.limbs(H(fp["R2"])) == [0x8c78ecb30000000f, 0xd7d30dbd8b0de0e7, 0x7797a99bc3c95d18, 0x096d41af7b9cb714]
    assert pm.Fp.limbs(H(fp["R3"])) == [0xf185a5993a9e10f9, 0xf6a68f3b6ac5b1d1, 0xdf8d1014353fd42c, 0x2ae309222d2d9910]
    assert H(fp["INV64"]) == 0x992d30ecffffffff and H(fq["INV64"]) == 0x8c46eb20ffffffff
    assert H(fp["ROOT_OF_UNITY"]) == 0x2bce74deac30ebda362120830561f81aea322bf2b7bb7584bdad6fabd87ea32f
    assert H(fq["ROOT_OF_UNITY"]) == 0x2de6a9b8746d3f589e5c4dfd492ae26e9bb97ea3c106f049a70e2c1102b6d05f
    assert H(fp["ZETA"]) == 0x12ccca834acdba712caad5dc57aab1b01d1f8bd237ad31491dad5ebdfdfe4ab9
    assert H(fq["ZETA"]) == 0x06819a58283e528e511db4d81cf70f5a0fed467d47c033af2aa9d2e050aa0e4f
    assert pm.Fq.limbs(H(fq["R"])) == [0x5b2b3e9cfffffffd, 0x992c350be3420567, 0xffffffffffffffff, 0x3fffffffffffffff]
    for F, f in ((pm.Fp, fp), (pm.Fq, fq)):
        assert H(f["modulus"]) == F.p and H(f["R"]) == F.R and H(f["DELTA"]) == F.DELTA
        assert pow(H(f["ROOT_OF_UNITY"]), 1 << 32, F.p) == 1 and pow(H(f["ROOT_OF_UNITY"]), 1 << 31, F.p) != 1
    # curve orders: [r]G = identity on both curves
    for C in (pm.Pallas, pm.Vesta):
        assert C.mul(C.scalar.p - 1, C.G) == C.neg(C.G)


@pytest.mark.parametrize("fname", ["Fp", "Fq"])
def test_oracle_field_and_ntt_vs_golden(fname):
    f = KAT["fields"][fname]; fid = FID[fname]
    a = [H(x) for x in f["arith"]["a"]]; b = [H(x) for x in f["arith"]["b"]]
    assert unmont(fid, O.field_op(fid, "mul", mont(fid, a), mont(fid, b))) == [H(x) for x in f["arith"]["mul"]]
    assert unmont(fid, O.field_op(fid, "inv", mont(fid, a))) == [H(x) for x in f["arith"]["inv"]]
    for key, case in f["ntt"].items():
        vin = [H(x) for x in case["in"]]
        log_n = len(vin).bit_length() - 1
        got = unmont(fid, O.fft(fid, mont(fid, vin), log_n, mont(fid, [H(case["omega"])])[0], threads=2))
        assert got == [H(x) for x in case["out"]], key


@pytest.mark.parametrize("fname", ["Fp", "Fq"])
def test_oracle_domain_vs_golden(fname):
    d = KAT["fields"][fname]["domain_j6_k3"]; fid = FID[fname]
    ek, om, eom = O.domain_info(fid, 6, 3)
    assert ek == d["extended_k"]
    assert unmont(fid, om) == [H(d["omega"])] and unmont(fid, eom) == [H(d["extended_omega"])]
    lag = mont(fid, [H(x) for x in d["lagrange"]])
    coeff = O.lagrange_to_coeff(fid, 6, 3, lag)
    assert unmont(fid, coeff) == [H(x) for x in d["coeff"]]
    assert unmont(fid, O.coeff_to_extended(fid, 6, 3, coeff)) == [H(x) for x in d["extended"]]
    h = mont(fid, [H(x) for x in d["h_ext"]])
    assert unmont(fid, O.extended_to_coeff(fid, 6, 3, h, divide=True)) == [H(x) for x in d["h_coeff_divided"]]


@pytest.mark.parametrize("cname", ["pallas", "vesta"])
def test_oracle_curve_vs_golden(cname):
    c = KAT["curves"][cname]; cid = CID[cname]; bf = O.BASE_FIELD[cid]; sf = O.SCALAR_FIELD[cid]
    g = mont(bf, [H(c["generator"][0]), H(c["generator"][1])]).reshape(8)
    for k, enc in c["multiples_compressed"].items():
        p = O.point_mul(cid, O.ints_to_limbs([H(k)]), g)
        assert O.point_compress(cid, p)[0].tobytes().hex() == enc, k
    assert O.point_compress(cid, np.zeros(8, dtype=np.uint64))[0].tobytes().hex() == c["identity_compressed"]
    m = c["msm8"]
    bases = mont(bf, [H(v) for P in m["bases"] for v in P]).reshape(8, 8)
    sc = mont(sf, [H(s) for s in m["scalars"]])
    for th in (1, 3):
        assert O.point_compress(cid, O.msm(cid, sc, bases, threads=th))[0].tobytes().hex() == m["result_compressed"]


def test_oracle_msm_matches_python_model_random():
    rnd = random.Random(3)
    for cid, C in ((O.PALLAS, pm.Pallas), (O.VESTA, pm.Vesta)):
        bf, sf = O.BASE_FIELD[cid], O.SCALAR_FIELD[cid]
        n = 90
        pts = make_points(cid, n, seed=rnd.randrange(1 << 30))
        sc_int = [rnd.randrange(C.scalar.p) for _ in range(n)]
        sc_int[0] = 0; sc_int[1] = 1; sc_int[2] = C.scalar.p - 1
        pts_int = []
        for i in range(n):
            x, y = unmont(bf, pts[i].reshape(2, 4))
            pts_int.append((x, y))
            assert C.on_curve((x, y))
        want = C.best_multiexp(sc_int, pts_int)
        got = O.msm(cid, mont(sf, sc_int), pts, threads=4)
        gx, gy = unmont(bf, got.reshape(2, 4))
        assert (gx, gy) == want


def test_oracle_thread_count_invariance_and_tinyram_scalars():
    n = 5000
    pts = make_points(O.VESTA, n)
    for sc in (scalars_uniform(O.VESTA, n), scalars_tinyram(O.VESTA, n)):
        r1 = O.msm(O.VESTA, sc, pts, threads=1)
        assert np.array_equal(r1, O.msm(O.VESTA, sc, pts, threads=7))
    a = O.random_field_mont(O.FP, 1 << 12, 1)
    om = mont(O.FP, [pm.Fp.root_of_unity(12)])[0]
    assert np.array_equal(O.fft(O.FP, a, 12, om, threads=1), O.fft(O.FP, a, 12, om, threads=8))


def test_sampler_primitives_against_the_python_model():
    """the three routines the sampled create_proof baseline adds to the C++ oracle (quotient-program interpreter, Horner
    evaluation, one generator-collapse round) against the big-int model"""
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import poly as P
    import random
    p = O.MODULUS[O.FP]
    rnd = random.Random(4)
    rows = 16
    cols_int = [[rnd.randrange(p) for _ in range(rows)] for _ in range(3)]
    A, B, C = P.Poly(0), P.Poly(1), P.Poly(2)
    ast = (A * B - C.with_rotation(1)) * 7 + P.LinearTerm(3) * A.with_rotation(-1) + B * B + 5
    prog = P.compile_ast(ast, p)
    xs_int = [rnd.randrange(p) for _ in range(rows)]
    mont = lambda v: O.to_mont(O.FP, O.ints_to_limbs(v))
    got = O.limbs_to_ints(O.from_mont(O.FP, O.quotient_vm(O.FP, prog.code, prog.n_regs, mont(prog.consts), [mont(c) for c in cols_int], rows,
                                                             mont(xs_int), rows, threads=3)))
    want = [((cols_int[0][r] * cols_int[1][r] - cols_int[2][(r + 1) % rows]) * 7 + 3 * xs_int[r] * cols_int[0][(r - 1) % rows]
             + cols_int[1][r] ** 2 + 5) % p for r in range(rows)]
    assert got == want
    coeffs = [rnd.randrange(p) for _ in range(1000)]
    x = rnd.randrange(p)
    for threads in (1, 7):
        got = O.limbs_to_ints(O.from_mont(O.FP, O.eval_polynomial(O.FP, mont(coeffs), mont([x])[0], threads=threads).reshape(1, 4)))[0]
        assert got == sum(c * pow(x, i, p) for i, c in enumerate(coeffs)) % p
    V = pm.Vesta
    pts = [V.mul(rnd.randrange(1, V.scalar.p), V.G) for _ in range(6)]
    u = rnd.randrange(V.scalar.p)
    qm = lambda v: O.to_mont(O.FQ, O.ints_to_limbs(v))
    g = np.stack([np.concatenate([qm([x_])[0], qm([y_])[0]]) for x_, y_ in pts])
    got = O.generator_collapse(O.VESTA, g, O.ints_to_limbs([u])[0], threads=2)
    for i in range(3):
        wx, wy = V.add(pts[i], V.mul(u, pts[i + 3]))
        assert O.limbs_to_ints(O.from_mont(O.FQ, got[i].reshape(2, 4))) == [wx, wy]
