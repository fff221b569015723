"""The programs of the reference's own circuit tests (/root/reference/src/circuits/mod.rs:88-360, tables/exe.rs:1121-1570) and of
its interpreter tests (src/trace.rs:565-624), shared by the CPU (mock prover) and GPU (real proof) tests."""


def load_and_answer(T, W, a, b):
    """circuits/mod.rs:88-110"""
    prog = [T.LoadW(0, T.Imm(b)), T.And(1, 0, T.Imm(a)), T.Answer(T.Imm(1))]
    tr = T.eval_program(prog, T.Mem(W, [1]))
    assert tr.ans == 1
    return tr


def mov_ins_answer(T, W, ins, b):
    """circuits/mod.rs:111-129"""
    prog = [T.Mov(0, T.Imm(b)), ins, T.Answer(T.Imm(1))]
    tr = T.eval_program(prog, T.Mem(W, [1]))
    assert tr.ans == 1
    return tr


THREE_OPERAND = ("And", "Xor", "Or", "Add", "Sub", "Mull", "UMulh", "SMulh", "UMod", "UDiv", "Shl", "Shr")   # mov_*_answer, ri = 1, rj = 0
TWO_OPERAND = ("Cmpe", "Cmpa", "Cmpae", "Cmpg", "Cmpge")                                                     # ri = 0


def mov_named(T, W, name, a, b):
    ctor = getattr(T, name)
    ins = ctor(1, 0, T.Imm(a)) if name in THREE_OPERAND else ctor(0, T.Imm(a))
    return mov_ins_answer(T, W, ins, b)


def answer_only(T, W):
    """circuits/mod.rs:379-386 (two_programs)"""
    tr = T.eval_program([T.Answer(T.Imm(1))], T.Mem(W))
    assert tr.ans == 1
    return tr


def counting_loop(T, W, iterations, body=()):
    """not from the reference: the long-trace workload of BASELINE.json configs[3] -- r0 counts to `iterations`, the body runs
    every pass; only instructions the reference's witness generation covers (immediate operands, CnJmp for the back edge)"""
    prog = [T.Mov(0, T.Imm(0)), T.Mov(1, T.Imm(1))] + list(body) + [T.Add(0, 0, T.Imm(1)), T.Cmpe(0, T.Imm(iterations)), T.CnJmp(T.Imm(2)),
                                                                    T.Answer(T.Imm(1))]
    return T.eval_program(prog, T.Mem(W, [1]))


def mixed_body(T, W):
    m = (1 << W) - 1
    return [T.Add(1, 1, T.Imm(3 & m)), T.Xor(2, 1, T.Imm(0x5A & m)), T.And(3, 2, T.Imm(0x3C & m)), T.Or(4, 3, T.Imm(0x81 & m)),
            T.Mull(5, 1, T.Imm(7)), T.UMulh(6, 1, T.Imm(m)), T.Sub(7, 1, T.Imm(9)), T.Shr(2, 1, T.Imm(3)), T.Shl(3, 1, T.Imm(2)),
            T.UDiv(4, 1, T.Imm(5)), T.UMod(5, 1, T.Imm(6)), T.Cmpa(1, T.Imm(100 & m)), T.Cmpge(1, T.Imm(17)), T.SMulh(6, 1, T.Imm(m - 2))]
