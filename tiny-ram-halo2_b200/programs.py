"""Long-trace TinyRAM workloads for the measurements (BASELINE.json configs[3] / [4]: "long execution trace").  Not from the
reference, whose test programs are three instructions long (/root/reference/src/circuits/mod.rs:88-129): a counting loop whose
body runs every gadget of the execution table.  Only instructions the reference's witness generation covers are used (immediate
operands; CnJmp for the back edge -- Jmp / StoreW / Not are assigned no logic decomposition by exe.rs:979-1043 and CJmp's
SelectionD::PcPlusOne sets two contradictory selectors, aux.rs:1024-1027)."""
from . import trace as T


def mixed_body(word_bits: int):
    m = (1 << word_bits) - 1
    I = T.Imm
    return [T.Add(1, 1, I(3 & m)), T.Xor(2, 1, I(0x5A & m)), T.And(3, 2, I(0x3C & m)), T.Or(4, 3, I(0x81 & m)), T.Mull(5, 1, I(7)),
            T.UMulh(6, 1, I(m)), T.Sub(7, 1, I(9)), T.Shr(2, 1, I(3)), T.Shl(3, 1, I(2)), T.UDiv(4, 1, I(5)), T.UMod(5, 1, I(6)),
            T.Cmpa(1, I(100 & m)), T.Cmpge(1, I(17)), T.SMulh(6, 1, I(m - 2))]


def counting_loop(word_bits: int, iterations: int, body=()):
    """r0 counts to `iterations`; the body runs every pass: 2 + (len(body) + 3) * iterations + 1 steps"""
    I = T.Imm
    prog = [T.Mov(0, I(0)), T.Mov(1, I(1))] + list(body) + [T.Add(0, 0, I(1)), T.Cmpe(0, I(iterations)), T.CnJmp(I(2)), T.Answer(I(1))]
    return T.eval_program(prog, T.Mem(word_bits, [1]))


def longest_loop(word_bits: int, steps: int = None):
    """the longest mixed-body loop whose trace fits `steps` (default: TABLE_LEN - 1 = 2^(W/2) - 1 rows of the execution table)"""
    steps = steps if steps is not None else (1 << (word_bits // 2)) - 1
    body = mixed_body(word_bits)
    return counting_loop(word_bits, max(1, (steps - 3) // (len(body) + 3)), body)
