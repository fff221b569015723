"""Small PLONKish circuits for the create_proof / verify_proof tests (tests/test_plonk_cpu.py, tests/test_gpu_plonk.py):
custom gates with rotations, a lookup, a two-chunk permutation and a public input -- every argument of the protocol that the
reference's TinyRamCircuit uses (gates: src/circuits/*.rs; lookups: tables/even_bits.rs:158-165; equality: tables/prog.rs:151-152;
instance: tables/prog.rs:141)."""


def standard(PL, with_lookup=True, wide_lookup=False):
    """PL = the product's plonk module.  Returns (cs, fixed_values, copies, advice, instances) for n >= 16."""
    cs = PL.ConstraintSystem()
    a, b, c = cs.advice_column(), cs.advice_column(), cs.advice_column()
    q_add, q_mul, q_next, tbl, q_lk = (cs.fixed_column() for _ in range(5))
    pi = cs.instance_column()
    A, F, I = PL.ADVICE, PL.FIXED, PL.INSTANCE
    qa, qb, qc = cs.query(A, a), cs.query(A, b), cs.query(A, c)
    cs.create_gate([cs.query(F, q_add) * (qa + qb - qc),
                    cs.query(F, q_mul) * (qa * qb - qc)])
    cs.create_gate([cs.query(F, q_next) * (cs.query(A, a, 1) - qc) * 3,
                    cs.query(F, q_next) * (cs.query(A, b, -1) + 1 - cs.query(A, b, -1) - 1)])
    tbl2 = None
    if with_lookup:
        if wide_lookup:
            tbl2 = cs.fixed_column()
            cs.lookup([(cs.query(F, q_lk) * qa, cs.query(F, tbl)), (cs.query(F, q_lk) * qb, cs.query(F, tbl2))])
        else:
            cs.lookup([(cs.query(F, q_lk) * qa, cs.query(F, tbl))])
    for col in (a, b, c):
        cs.enable_equality(A, col)
    cs.enable_equality(I, pi)
    # witness
    adv = [[2, 5, 1, 7], [3, 4, 6, 7], [5, 20, 7, 49]]           # a, b, c on rows 0..3
    fixed = [[1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 0, 0], list(range(8)), [1, 1, 1, 1]]
    copies = [((A, c, 0), (A, a, 1)), ((A, a, 3), (A, b, 3)), ((I, pi, 0), (A, c, 2)), ((A, b, 3), (A, a, 3))]   # last one: already merged
    if tbl2 is not None:
        fixed.append([3 * v % 8 for v in range(8)])              # row r of the table is (r, 3r mod 8)
        adv[1] = [6, 7, 3, 5]                                     # b = 3a mod 8 on the looked-up rows
        adv[2] = [8, 35, 4, 35]
        fixed[2] = [0, 0, 0, 0]                                   # the next-row gate is off in this variant
        copies = [((A, c, 1), (A, c, 3)), ((I, pi, 0), (A, c, 2)), ((A, c, 3), (A, c, 1))]
    instances = [[adv[2][2]]]
    return cs, fixed, copies, adv, instances


class StandardCircuit:
    """`standard` as a subject of tiny_ram_halo2_b200.test_utils (the role of the reference's gadget circuits LogicCircuit /
    SumCircuit / ShiftCircuit in its gen_proofs_and_verify_should_fail tests, e.g. /root/reference/src/circuits/logic.rs:515-527:
    the advice is assigned independently of the public input and one cell is copy-constrained to it, so a wrong input is caught by
    the permutation argument).  The empty circuit (`C::default()`) has the same fixed columns."""

    def __init__(self, **kw):
        self.kw = kw

    def build(self, PL, k, public_input=None, keygen_from_empty_circuit=False):
        cs, fixed, copies, adv, inst = standard(PL, **self.kw)
        return cs, fixed, copies, adv, (inst if public_input is None else [list(c) for c in public_input])
