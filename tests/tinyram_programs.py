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
    """the long-trace workload of BASELINE.json configs[3] (tiny-ram-halo2_b200/programs.py; not from the reference)"""
    from tiny_ram_halo2_b200 import programs
    return programs.counting_loop(W, iterations, body)


def mixed_body(T, W):
    from tiny_ram_halo2_b200 import programs
    return programs.mixed_body(W)
