"""CPU tests of the TinyRAM interpreter and circuit mirrors (tiny-ram-halo2_b200/{trace,tinyram}.py).

The interpreter is pinned by the reference's own unit tests (src/trace.rs:36-62, 565-624: exact answers, access counts and a
recorded Store).  The circuit is checked the way the reference checks it -- every `*_mock_prover` property test of
src/circuits/mod.rs:404-505 -- with oracle/mock_prover.check in the role of MockProver::assert_satisfied (stricter: all rows)."""
import random

import pytest

import mock_prover as MP
import tinyram_programs as TP


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import plonk, tinyram, trace
    return plonk, tinyram, trace


def _check(mods, tr, k=None, mutate=None, **kw):
    PL, TR, T = mods
    k = k or 2 + tr.word_bits // 2                                # mock_prover_test: k = 2 + WORD_BITS / 2 (circuits/mod.rs:367)
    circ, fixed, copies, adv, inst = TR.build(PL, tr, k, **kw)
    if mutate:
        mutate(circ, fixed, adv, inst)
    return MP.check(PL, circ.cs, 1 << k, TR.PL_FIELD_MODULUS, fixed, adv, inst, copies, circ.gate_names)


# ---- interpreter: the reference's unit tests -------------------------------------------------------------------------------------
def test_word_signed_conversions(mods):
    T = mods[2]
    for s in range(-128, 127):                                    # from_signed_test / to_signed_test (trace.rs:38-40, 53-57)
        w = T.try_from_signed(s, 8)
        assert w == s & 0xFF and T.into_signed(w, 8) == s
    for s in (128, 129, 1 << 20, (1 << 31) - 1):                  # from_signed_test_too_high
        assert T.try_from_signed(s, 8) is None
    for s in (-130, -(1 << 20), -(1 << 31)):                      # from_signed_test_too_low
        assert T.try_from_signed(s, 8) is None
    assert T.try_from_signed(127, 8) == 127 and T.try_from_signed(-128, 8) == 128


def test_trace_load_and_store_ans(mods):
    """trace.rs:565-601"""
    T = mods[2]
    prog = [T.LoadW(0, T.Imm(0)), T.And(1, 0, T.Imm(0b1)), T.StoreW(1, T.Imm(8)), T.Answer(T.Reg(1))]
    tr = T.eval_program(prog, T.Mem(8, [0b1]))
    assert tr.ans == 0b1
    assert tr.mem.accesses[8][1] == T.Access("Store", 8, 0b1, time=3, pc=2)
    assert tr.mem.access_count() == 4                             # init 0, load 0, init 8, store 8


def test_trace_load_and_answer(mods):
    """trace.rs:603-624"""
    T = mods[2]
    prog = [T.LoadW(0, T.Imm(16)), T.And(1, 0, T.Imm(128)), T.Answer(T.Imm(1))]
    tr = T.eval_program(prog, T.Mem(8, [0b1]))
    assert tr.mem.access_count() == 3 and tr.ans == 1             # Init tape, Init 16, Load 16


def test_interpreter_semantics(mods):
    T = mods[2]
    run = lambda prog, W=8: T.eval_program(prog, T.Mem(W, [1]))
    I = T.Imm
    s = run([T.Mov(0, I(200)), T.Add(1, 0, I(100)), T.Sub(2, 0, I(201)), T.Mull(3, 0, I(2)), T.UMulh(4, 0, I(2)), T.Answer(I(1))]).exe
    assert s[2].regs[1] == 44 and s[2].flag is True               # carry
    assert s[3].regs[2] == 255 and s[3].flag is True              # borrow
    assert s[4].regs[3] == 144 and s[4].flag is False             # Mull's flag: product < 2^W (trace.rs:447-453)
    assert s[5].regs[4] == 1
    s = run([T.Mov(0, I(0x85)), T.SMulh(1, 0, I(3)), T.UDiv(2, 0, I(0)), T.UMod(3, 0, I(7)), T.Shl(4, 0, I(1)), T.Shr(5, 0, I(1)),
             T.Cmpg(0, I(1)), T.Cmpa(0, I(1)), T.Answer(I(1))]).exe
    assert s[2].regs[1] == 0xFE                                   # -123 * 3 = -369 = 0xFE8F: upper byte
    assert s[3].regs[2] == 0 and s[3].flag is True                # division by zero
    assert s[4].regs[3] == 0x85 % 7
    assert s[5].regs[4] == 0x0A and s[5].flag is True             # Shl: flag = msb of the operand
    assert s[6].regs[5] == 0x42 and s[6].flag is True             # Shr: flag = lsb of the operand
    assert s[7].flag is False and s[8].flag is True               # signed: -123 > 1 is false; unsigned: 0x85 > 1
    with pytest.raises(IndexError):
        run([T.Mov(0, I(1))])                                     # "Program did not Answer 0 or 1."
    with pytest.raises(ValueError):
        T.Instruction("Add", I(1), ri=0)                          # missing rj


# ---- even bits (tables/even_bits.rs:219-297) ----------------------------------------------------------------------------------------
def test_even_bits(mods):
    TR = mods[1]
    assert [TR.even_bits_at(i) for i in range(4)] == [0b0, 0b1, 0b100, 0b101]        # even_bits_at_test
    assert TR.decompose(0xAAAA) == (0, 0xAAAA >> 1)                                   # decompose_test_even_odd
    assert TR.decompose(0x5555) == (0x5555, 0)
    rnd = random.Random(3)
    for _ in range(200):
        w = rnd.randrange(1 << 32)
        e, o = TR.decompose(w)
        assert e + 2 * o == w and e & 0xAAAAAAAA == 0 and o & 0xAAAAAAAA == 0


# ---- the circuit's shape ----------------------------------------------------------------------------------------
def test_circuit_shape(mods):
    PL, TR, T = mods
    c = TR.TinyRamCircuit(PL, 8)
    cs = c.cs
    assert (cs.num_advice, cs.num_instance, cs.num_fixed) == (263, 94, 24)           # 23 + the dynamic table's tag column
    assert len(cs.gates) == 138 and len(cs.lookups) == 31 and len(cs.permutation) == 188
    assert len(c.intermediate) == 45
    assert cs.degree() == 6 and cs.blinding_factors() == 5
    assert [len(i) for i, _ in cs.lookups].count(1) == 28
    assert sorted(len(i) for i, _ in cs.lookups)[-3:] == [2, 15, 96]
    assert {r for kind in cs.queries.values() for _, r in kind} == {0, 1}            # rotations: cur and next only
    assert len(TR.program_instance([T.Answer(T.Imm(1))], 8)) == 94
    for name, out in TR.OUT.items():
        assert set(out) <= set(TR.OUT_NAMES)
    assert set(TR.OUT) == set(T.OPCODES) == set(TR.OUT_TABLE_ORDER)


# ---- the reference's mock-prover tests (circuits/mod.rs:404-505) -------------------------------------------------------------------------
def _operands(name, rnd, W=8):
    if name in ("Shl", "Shr"):
        return rnd.randrange(W), rnd.randrange(1 << W)                               # a in 0..8 (mod.rs:497, 502)
    if name in ("Mull", "UMulh", "UMod", "UDiv", "Cmpg", "Cmpge", "SMulh"):          # signed_word(8): -128 .. 126
        return rnd.randrange(-128, 127) & 0xFF, rnd.randrange(-128, 127) & 0xFF
    return rnd.randrange(1 << W), rnd.randrange(1 << W)


@pytest.mark.parametrize("name", TP.THREE_OPERAND + TP.TWO_OPERAND)
def test_mov_ins_answer_mock_prover(mods, name):
    T = mods[2]
    rnd = random.Random(hash(name) & 0xFFFF)
    cases = [_operands(name, rnd) for _ in range(24)]
    cases += [(0, 0), (1, 0), (0, 1), (7, 255), (7, 128), (3, 127)] if name in ("Shl", "Shr") else [(0, 0), (255, 255), (0, 255), (255, 0), (128, 128), (1, 128)]
    for a, b in cases:
        if name == "SMulh" and 127 in (a, b):
            continue                                              # outside signed_word(8)
        assert _check(mods, TP.mov_named(T, 8, name, a, b)) == [], (name, a, b)


def test_load_and_answer_mock_prover(mods):
    T = mods[2]
    rnd = random.Random(11)
    for _ in range(16):
        assert _check(mods, TP.load_and_answer(T, 8, rnd.randrange(256), rnd.randrange(256))) == []
    assert _check(mods, TP.answer_only(T, 8)) == []


@pytest.mark.parametrize("W,k", [(16, 10), (24, 13)])
def test_wider_words(mods, W, k):
    T = mods[2]
    tr = TP.counting_loop(T, W, 5, TP.mixed_body(T, W))
    assert len(tr.exe) > 80
    assert _check(mods, tr, k=k) == []


def test_loop_fills_the_table(mods):
    T = mods[2]
    tr = TP.counting_loop(T, 8, 4)                                # 2 + 3 * 4 + 1 = 15 steps = TABLE_LEN - 1
    assert len(tr.exe) == 15
    assert _check(mods, tr) == []
    PL, TR, _ = mods
    with pytest.raises(ValueError):
        TR.build(PL, TP.counting_loop(T, 8, 5), 6)                # 18 steps do not fit 2^(W/2) rows


# ---- soundness of the checker / of the constraints ---------------------------------------------------------------------------------------
def test_tampered_witnesses_are_rejected(mods):
    T = mods[2]
    tr = TP.mov_named(T, 8, "Add", 200, 77)                       # 77 + 200 = 277: r1 = 21, carry
    assert _check(mods, tr) == []
    def wrong_sum(c, f, a, i): a[c.reg[1]][2] = 22
    def wrong_flag(c, f, a, i): a[c.flag][2] = 0
    def wrong_opcode(c, f, a, i): a[c.line.opcode][1] = 5
    def wrong_instance(c, f, a, i): i[0][1] = 5
    def trace_goes_on(c, f, a, i): a[c.s_trace][3] = 1
    def pc_skips(c, f, a, i): a[c.pc][1] = 2
    def not_even_bits(c, f, a, i): a[c.tv_b.even][1] += 2; a[c.tv_b.odd][1] -= 1      # still even + 2 odd = word
    for mut, where in ((wrong_sum, "reg_next"), (wrong_flag, "sum"), (wrong_opcode, "lookup 30"), (wrong_instance, "copy"),
                       (trace_goes_on, "unchanged"), (pc_skips, "unchanged"), (not_even_bits, "lookup")):
        fails = _check(mods, tr, mutate=mut)
        assert fails and any(where in f for f in fails), (mut.__name__, fails)


def test_register_operand_quirk(mods):
    """aux.rs:419-427 takes a register operand's INDEX as the temp-var value; the reg[r] gate wants the register's value"""
    T = mods[2]
    tr = T.eval_program([T.Mov(0, T.Imm(7)), T.Mov(1, T.Imm(9)), T.Add(2, 0, T.Reg(1)), T.Answer(T.Imm(1))], T.Mem(8, [1]))
    assert _check(mods, tr) == []
    assert any("tv.a.reg[1]" in f for f in _check(mods, tr, reg_operand_value=False))
    tr = T.eval_program([T.Mov(1, T.Imm(1)), T.Add(2, 0, T.Reg(1)), T.Answer(T.Imm(1))], T.Mem(8, [1]))
    assert _check(mods, tr, reg_operand_value=False) == []        # value == index: the reference's witness happens to be right


def test_keygen_from_the_empty_circuit(mods):
    """test_utils.rs:22-25 generates the keys from C::default() (trace: None): the execution table's selectors stay off"""
    PL, TR, T = mods
    c = TR.TinyRamCircuit(PL, 8)
    fixed, copies, advice = c.synthesize(None, 64)
    fixed = [f.dense(64) for f in fixed]
    assert not any(fixed[c.s_table]) and not any(fixed[c.first_line]) and not any(fixed[c.time])
    assert sum(fixed[c.s_prog]) == 16 and fixed[c.prog_pc][:17] == list(range(16)) + [0]
    assert len(copies) == 94 and len(list(PL.expand_copies(copies))) == 94 * 16
    assert fixed[c.t_even][:16] == [TR.even_bits_at(i) for i in range(16)]
    assert fixed[c.t_pow_powers][:10] == [1, 2, 4, 8, 16, 32, 64, 128, 0, 1]          # row 8 = (W, 0), then the default row (0, 1)
    assert fixed[c.t_out_opcode][25:28] == [32, 0, 1]                                 # Answer + 1, the all-zero default row, fill = row 0


def test_real_proof_on_the_cpu_backend_verifies(mods):
    """gen_proofs_and_verify (test_utils.rs:6-71) with the oracle's PythonBackend under plonk.create_proof and the oracle's
    independent verifier: Params::new(6), keygen, one proof of `Answer 1`, verify; the wrong program is rejected"""
    import pasta_model as pm
    import plonk_model as VM
    PL, TR, T = mods
    C = pm.Vesta
    circ, fixed, copies, adv, inst = TR.build(PL, TP.answer_only(T, 8), 6)
    be = VM.PythonBackend(C, 6, circ.cs.degree())
    pk = PL.keygen(be, circ.cs, fixed, copies)
    rnd = random.Random(1)
    proof = PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p))
    assert VM.verify_proof(C, be.params, pk.vk, inst, proof), VM.verify_proof.last_error
    other = TR.program_instance([T.Answer(T.Imm(0))], 8)
    assert not VM.verify_proof(C, be.params, pk.vk, other, proof)
    import hashlib, json, os                 # the proof bytes are frozen too (fixed-seed blinding): tests/golden/proof_digests.json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proof_digests.json")) as f:
        want = json.load(f)["tinyram_answer_only_w8_k6_seed1_vesta"]
    assert (len(proof), hashlib.sha256(proof).hexdigest(), hex(pk.vk.transcript_repr)) == (want["bytes"], want["sha256"], want["transcript_repr"])
    # the package's own verifier (verifier.py) on the same proof: BatchVerifier, then SingleVerifier (test_utils.rs:56-70)
    from tiny_ram_halo2_b200 import verifier as V
    bv = V.BatchVerifier()
    bv.add_proof(inst, proof)
    assert bv.finalize(be, pk.vk)
    assert V.verify_proof(be, pk.vk, V.SingleVerifier(be), inst, V.Blake2bRead(proof, be.q, be.p)) is None
    with pytest.raises(V.VerifyError):
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), other, V.Blake2bRead(proof, be.q, be.p))
    with pytest.raises(ValueError):
        PL.create_proof(be, pk, inst, adv, lambda: 1, PL.Blake2bWrite(C.base.p, C.scalar.p), debug=True)


def test_word_size_32_short_trace(mods):
    """BASELINE.json configs[3] uses word size 32 (untested in the reference): 2^16-row tables, k >= 17"""
    PL, TR, T = mods
    tr = TP.counting_loop(T, 32, 2, TP.mixed_body(T, 32))
    circ, fixed, copies, adv, inst = TR.build(PL, tr, 17, dense=False)
    assert circ.table_len == 1 << 16 and len(inst[0]) == 1 << 16
    n = 1 << 17
    dense = [f.dense(n) for f in fixed]
    assert MP.check(PL, circ.cs, n, TR.PL_FIELD_MODULUS, dense, adv, inst, copies, circ.gate_names) == []


def test_random_programs_mock_prover(mods):
    """beyond the reference's three-instruction programs: random straight-line programs over the instruction set its witness
    generation covers (immediate and register operands, all eight registers), W = 8, up to 14 steps"""
    PL, TR, T = mods
    rnd = random.Random(2024)
    three = ("And", "Xor", "Or", "Add", "Sub", "Mull", "UMulh", "SMulh", "UDiv", "UMod")
    two = ("Cmpe", "Cmpa", "Cmpae", "Cmpg", "Cmpge", "Mov", "CMov")
    for trial in range(60):
        prog = [T.Mov(rnd.randrange(8), T.Imm(rnd.randrange(256))) for _ in range(2)]
        for _ in range(rnd.randrange(1, 11)):
            operand = T.Imm(rnd.randrange(256)) if rnd.random() < 0.6 else T.Reg(rnd.randrange(8))
            kind = rnd.random()
            if kind < 0.55:
                prog.append(getattr(T, rnd.choice(three))(rnd.randrange(8), rnd.randrange(8), operand))
            elif kind < 0.9:
                prog.append(getattr(T, rnd.choice(two))(rnd.randrange(8), operand))
            else:
                prog.append(getattr(T, rnd.choice(("Shl", "Shr")))(rnd.randrange(8), rnd.randrange(8), T.Imm(rnd.randrange(8))))
        prog.append(T.Answer(T.Imm(1)))
        tr = T.eval_program(prog, T.Mem(8, [1]))
        fails = _check(mods, tr)
        assert fails == [], (trial, [(i.name, i.ri, i.rj, i.a) for i in prog], fails[:3])


def test_exe_circuit_mock_prover(mods):
    """the reference's second circuit, ExeCircuit (tables/exe.rs:1082-1116): the execution table without the program table; its
    tests run the same programs with no instance (exe.rs:1434-1570)"""
    PL, TR, T = mods
    c = TR.TinyRamCircuit(PL, 8, with_prog=False)
    assert (c.cs.num_advice, c.cs.num_instance, c.cs.num_fixed) == (263 - 94, 0, 24 - 3)
    assert len(c.cs.gates) == 138 and len(c.cs.lookups) == 30 and len(c.cs.permutation) == 0
    rnd = random.Random(5)
    traces = [TP.answer_only(T, 8), TP.load_and_answer(T, 8, 1, 2)]
    traces += [TP.mov_named(T, 8, name, *_operands(name, rnd)) for name in TP.THREE_OPERAND + TP.TWO_OPERAND]
    for tr in traces:
        circ, fixed, copies, adv, inst = TR.build(PL, tr, 6, with_prog=False)
        assert inst == [] and copies == []
        assert MP.check(PL, circ.cs, 64, TR.PL_FIELD_MODULUS, fixed, adv, inst, copies, circ.gate_names) == []


def test_witness_digests(mods):
    """The witness (fixed columns, copy constraints, advice, instance) of 33 circuit / trace / option combinations, including
    bench.py's 65 521-step trace at word size 32, against the SHA-256 digests frozen from the row-by-row implementation that
    followed ExeChip::assign_trace line by line (tests/golden/make_witness_digests.py): the column-wise synthesis must not
    move a bit."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_witness_digests", os.path.join(here, "golden", "make_witness_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, "golden", "witness_digests.json")) as f:
        want = json.load(f)
    got = mod.all_digests()
    assert sorted(got) == sorted(want)
    assert {k: v for k, v in got.items() if v != want[k]} == {}


def test_batch_inverse(mods):
    TR = mods[1]
    p = TR.PL_FIELD_MODULUS
    rnd = random.Random(4)
    vals = [rnd.randrange(p) for _ in range(50)] + [0, 1, p - 1, 0, 0, 7]
    rnd.shuffle(vals)
    assert TR._batch_inverse(vals, p) == [pow(v, -1, p) if v else 0 for v in vals]
    assert TR._batch_inverse([], p) == [] and TR._batch_inverse([0, 0], p) == [0, 0]


def test_array_columns_equal_list_columns(mods):
    """build(arrays=True) -- the upload path of bench.py -- returns the same witness: every column that comes back as a uint64
    numpy array holds exactly the ints of the list version (and goes through device_columns' fast path, _column_u64)"""
    import numpy as np
    PL, TR, T = mods
    from tiny_ram_halo2_b200 import programs
    for tr, k, kw in ((TP.load_and_answer(T, 8, 1, 2), 6, {}), (programs.counting_loop(16, 3, programs.mixed_body(16)), 11, {}),
                      (programs.counting_loop(16, 3, programs.mixed_body(16)), 11, {"with_prog": False}),
                      (programs.counting_loop(32, 12, programs.mixed_body(32)), 17, {})):
        _, f0, c0, a0, i0 = TR.build(PL, tr, k, dense=False, **kw)
        _, f1, c1, a1, i1 = TR.build(PL, tr, k, dense=False, arrays=True, **kw)
        assert c0 == c1 and len(a0) == len(a1) and len(i0) == len(i1) and len(f0) == len(f1)
        n_arrays = 0
        for x, y in list(zip(a0, a1)) + list(zip(i0, i1)) + [(f.prefix, g.prefix) for f, g in zip(f0, f1)]:
            if isinstance(y, np.ndarray):
                n_arrays += 1
                assert y.dtype == np.uint64 and TR._column_u64(y) is not None
                assert [int(v) for v in y] == list(x)
            else:
                assert y == x
        assert [f.fill for f in f0] == [g.fill for g in f1]
        assert n_arrays >= (200 if kw.get("with_prog", True) else 100)
    assert TR._column_u64([1, 1 << 64]) is None and TR._column_u64([]).shape == (0,)
    assert TR._column_u64(np.arange(10, dtype=np.uint64).reshape(5, 2)[:, 1]).flags.c_contiguous
    assert TR.even_bits_table(1 << 10) == [TR.even_bits_at(i) for i in range(1 << 10)]


def test_device_columns_host_logic_on_a_fake_backend(mods, monkeypatch):
    """tinyram.device_columns with the CUDA calls stubbed out (torch on the CPU, the Montgomery conversion kernel a no-op): list
    columns and the uint64-array columns of build(arrays=True) must stage the same limb-0 values, fill included"""
    import types
    import numpy as np
    import torch
    PL, TR, T = mods
    from tiny_ram_halo2_b200 import programs
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    fake_torch = types.SimpleNamespace(zeros=lambda shape, dtype=None, device=None: torch.zeros(shape, dtype=dtype), int64=torch.int64,
                                       from_numpy=torch.from_numpy)
    calls = []
    be = types.SimpleNamespace(torch=fake_torch, n=1 << 11, R=5, _limbs=lambda vals: np.zeros((len(vals), 4), dtype=np.uint64),
                               _dev=lambda limbs: torch.zeros((len(limbs), 4), dtype=torch.int64), _sync=lambda: None,
                               ctx=types.SimpleNamespace(check=lambda rc: None, handle=None),
                               lib=types.SimpleNamespace(trp_dev_field_op=lambda *a: calls.append(a[-1]) or 0))
    tr = programs.counting_loop(16, 3, programs.mixed_body(16))
    _, f0, _, a0, i0 = TR.build(PL, tr, 11, dense=False)
    _, f1, _, a1, i1 = TR.build(PL, tr, 11, dense=False, arrays=True)
    small = lambda col: all(v < 1 << 64 for v in (col.prefix if isinstance(col, TR.FixedColumn) else col))
    for lists, arrs in ((f0, f1), (a0, a1), (i0, i1)):
        keep = [j for j, c in enumerate(lists) if small(c)]             # the others go through be.vec, which the fake does not have
        got0 = TR.device_columns(be, [lists[j] for j in keep])
        got1 = TR.device_columns(be, [arrs[j] for j in keep])
        assert len(got0) == len(got1) == len(keep) > 0
        for j, x, y in zip(keep, got0, got1):
            assert torch.equal(x, y), j
            col = lists[j]
            prefix, fill = (col.prefix, col.fill) if isinstance(col, TR.FixedColumn) else (col, 0)
            want = np.full(be.n, fill, dtype=np.uint64)
            want[:len(prefix)] = np.array(list(prefix), dtype=np.uint64)
            assert np.array_equal(x[:, 0].numpy().view(np.uint64), want) and not x[:, 1:].any()
    assert calls and set(calls) == {be.n}


def test_parity_vector_of_the_reproducible_rng(mods):
    """tests/golden/parity_vectors.json (what rust/parity/parity.rs is compared with): the first vector -- `Answer 1`, keys from the
    empty circuit, blinding scalars from rng.ScalarStreamRng (AES-256-CTR, 64 bytes per scalar) -- is reproduced byte for byte, and
    the RNG itself is pinned by a known answer"""
    import hashlib, json, os
    import pasta_model as pm
    import plonk_model as VM
    PL, TR, T = mods
    from tiny_ram_halo2_b200 import rng as RNG
    C = pm.Vesta
    r = RNG.ScalarStreamRng(C.scalar.p, bytes(range(32)))
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    ks = Cipher(algorithms.AES(bytes(range(32))), modes.CTR(bytes(16))).encryptor().update(bytes(192))
    assert [r(), r(), r()] == [int.from_bytes(ks[i:i + 64], "little") % C.scalar.p for i in (0, 64, 128)] and r.draws == 3
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parity_vectors.json")) as f:
        want = json.load(f)[0]
    circ, fixed, copies, adv, inst = TR.build(PL, TP.answer_only(T, 8), 6, keygen_from_empty_circuit=True)
    be = VM.PythonBackend(C, 6, circ.cs.degree())
    pk = PL.keygen(be, circ.cs, fixed, copies)
    rng = RNG.ScalarStreamRng(C.scalar.p, bytes.fromhex(want["seed_hex"]))
    proof = PL.create_proof(be, pk, inst, adv, rng, PL.Blake2bWrite(C.base.p, C.scalar.p))
    assert (len(proof), hashlib.sha256(proof).hexdigest(), rng.draws, hex(pk.vk.transcript_repr), proof[:32].hex()) == \
        (want["proof_bytes"], want["proof_sha256"], want["scalars_drawn"], want["transcript_repr_of_this_repo"], want["first_proof_point_le_hex"])
    cs = circ.cs
    assert want["shape"] == {"advice": cs.num_advice, "instance": cs.num_instance, "fixed": cs.num_fixed, "lookups": len(cs.lookups),
                             "gates": len(cs.gates), "equality_columns": len(cs.permutation), "degree": cs.degree(),
                             "blinding_factors": cs.blinding_factors()}


def test_constraint_degrees(mods):
    """DESIGN.md section 3 / 9 (what a split of h(X) by constraint degree could save): the degree of every gate polynomial, lookup
    and of the permutation argument of TinyRamCircuit<32, 8>, and per column the highest degree of any constraint that reads it"""
    import collections
    PL, TR, T = mods
    cs = TR.TinyRamCircuit(PL, 32).cs

    def cols_of(e, acc):
        if isinstance(e, PL.Query):
            acc.add((e.kind, e.column))
        for name in ("a", "b"):
            v = getattr(e, name, None)
            if isinstance(v, PL.Expression):
                cols_of(v, acc)

    colmax = collections.defaultdict(int)
    gate_deg = collections.Counter()
    for g in cs.gates:
        d = g.degree()
        gate_deg[d] += 1
        acc = set(); cols_of(g, acc)
        for c in acc:
            colmax[c] = max(colmax[c], d)
    assert cs.degree() == 6 and dict(gate_deg) == {2: 11, 3: 24, 4: 102, 6: 1}
    lookup_deg = collections.Counter()
    for inputs, tables in cs.lookups:
        d = max(4, 2 + max([1] + [e.degree() for e in inputs]) + max([1] + [e.degree() for e in tables]))     # lookup::Argument::required_degree
        lookup_deg[d] += 1
        acc = set()
        for e in inputs + tables:
            cols_of(e, acc)
        for c in acc:
            colmax[c] = max(colmax[c], d)
    assert dict(lookup_deg) == {5: 1, 6: 30}
    chunk = cs.degree() - 2
    for c in cs.permutation:                                  # chunk columns + z(omega X) + the active-rows factor
        colmax[c] = max(colmax[c], chunk + 2)
    assert len(cs.permutation) == 188 and chunk + 2 == 6
    by_deg = collections.Counter(colmax.values())
    assert cs.num_advice + cs.num_instance + cs.num_fixed == 381 and sum(by_deg.values()) == 378      # 3 columns are read by no constraint
    assert dict(by_deg) == {6: 260, 5: 96, 4: 16, 3: 5, 2: 1}
    per_kind = {kind: dict(collections.Counter(d for (k_, _), d in colmax.items() if k_ == kind)) for kind in (PL.ADVICE, PL.INSTANCE, PL.FIXED)}
    assert per_kind == {PL.ADVICE: {3: 5, 4: 16, 5: 94, 6: 147}, PL.INSTANCE: {6: 94}, PL.FIXED: {2: 1, 5: 2, 6: 19}}
