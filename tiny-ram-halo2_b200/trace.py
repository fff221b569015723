"""TinyRAM interpreter: mirror of /root/reference/src/trace.rs and src/instructions.rs (the witness SOURCE of the circuit the
prover proves -- SURVEY.md 8(f) row f4, "witness synthesis ... from a Trace").  Host code, as in the reference.

Program.eval (trace.rs:378-551) runs a program over a memory initialised from the two input tapes (Mem::new, trace.rs:160-179)
and records one Step per executed instruction (the state BEFORE the instruction executes: trace.rs:409-416) plus the
time-ordered memory accesses per address.  Instruction semantics follow the reference line by line, including the places where
it differs from the TinyRAM 2.0 specification (Shl/Shr flags, Mull's flag, no read / byte accesses)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, NamedTuple, Optional, Union


# ---- operands (trace.rs:10-157) -----------------------------------------------------------------------------------------------------
class Imm(NamedTuple):
    """ImmediateOrRegName::Immediate(Word(value))"""
    value: int


class Reg(NamedTuple):
    """ImmediateOrRegName::RegName(RegName(index)) as the `a` operand; ri / rj are plain ints"""
    index: int


Operand = Union[Imm, Reg]


def try_from_signed(s: int, word_bits: int) -> Optional[int]:
    """Word::try_from_signed (trace.rs:15-26)"""
    lo = -(1 << (word_bits - 1))
    if s > -lo - 1 or s < lo:
        return None
    return s if s >= 0 else s + (1 << word_bits)


def into_signed(w: int, word_bits: int) -> int:
    """Word::into_signed (trace.rs:28-34)"""
    return w if not (w >> (word_bits - 1)) & 1 else w - (1 << word_bits)


def decode_signed(w: int, word_bits: int) -> int:
    """signed_arithmetic::decode_signed (trace.rs:557-563)"""
    m = 1 << (word_bits - 1)
    return (w & (m - 1)) - (w & m)


# ---- instructions (instructions.rs:9-118, opcode.rs) --------------------------------------------------------------------------------
OPCODES = {
    "And": 0b00000, "Or": 0b00001, "Xor": 0b00010, "Not": 0b00011, "Add": 0b00100, "Sub": 0b00101, "Mull": 0b00110,
    "UMulh": 0b00111, "SMulh": 0b01000, "UDiv": 0b01001, "UMod": 0b01010, "Shl": 0b01011, "Shr": 0b01100, "Cmpe": 0b01101,
    "Cmpa": 0b01110, "Cmpae": 0b01111, "Cmpg": 0b10000, "Cmpge": 0b10001, "Mov": 0b10010, "CMov": 0b10011, "Jmp": 0b10100,
    "CJmp": 0b10101, "CnJmp": 0b10110, "StoreW": 0b11100, "LoadW": 0b11101, "Answer": 0b11111,
}
_HAS_RI = {"And", "Or", "Xor", "Not", "Add", "Sub", "Mull", "UMulh", "SMulh", "UDiv", "UMod", "Shl", "Shr", "Cmpe", "Cmpa", "Cmpae",
           "Cmpg", "Cmpge", "Mov", "CMov", "StoreW", "LoadW"}                                             # instructions.rs:120-149
_HAS_RJ = {"And", "Or", "Xor", "Add", "Sub", "Mull", "UMulh", "SMulh", "UDiv", "UMod", "Shl", "Shr"}      # instructions.rs:152-181


@dataclass(frozen=True)
class Instruction:
    """Instruction<RegName, ImmediateOrRegName>: name in OPCODES, ri / rj register indices (None where the variant has none)"""
    name: str
    a: Operand
    ri: Optional[int] = None
    rj: Optional[int] = None

    def __post_init__(self):
        if self.name not in OPCODES:
            raise ValueError(f"unknown instruction {self.name}")
        if (self.ri is not None) != (self.name in _HAS_RI) or (self.rj is not None) != (self.name in _HAS_RJ):
            raise ValueError(f"{self.name}: wrong register operands")

    @property
    def opcode(self) -> int:
        return OPCODES[self.name]

    def immediate(self) -> int:
        """inst.a().immediate().unwrap_or_default() (prog.rs:96-101, exe.rs:893-898)"""
        return self.a.value if isinstance(self.a, Imm) else 0


def _mk(name):
    def ctor(*args):
        if name in _HAS_RJ:
            ri, rj, a = args
            return Instruction(name, a, ri, rj)
        if name in _HAS_RI:
            ri, a = args
            return Instruction(name, a, ri)
        (a,) = args
        return Instruction(name, a)
    ctor.__name__ = name
    return ctor


And, Or, Xor, Not, Add, Sub, Mull, UMulh, SMulh, UDiv, UMod, Shl, Shr = (_mk(n) for n in (
    "And", "Or", "Xor", "Not", "Add", "Sub", "Mull", "UMulh", "SMulh", "UDiv", "UMod", "Shl", "Shr"))
Cmpe, Cmpa, Cmpae, Cmpg, Cmpge, Mov, CMov, Jmp, CJmp, CnJmp, StoreW, LoadW, Answer = (_mk(n) for n in (
    "Cmpe", "Cmpa", "Cmpae", "Cmpg", "Cmpge", "Mov", "CMov", "Jmp", "CJmp", "CnJmp", "StoreW", "LoadW", "Answer"))


def smulh_eval(a: int, b: int, word_bits: int):
    """SMulh::eval (instructions.rs:330-347): (upper, lower, flag) of the signed product"""
    f = into_signed(a, word_bits) * into_signed(b, word_bits)
    mask = (1 << word_bits) - 1
    lower, upper = f & mask, (f >> word_bits) & mask
    m = 1 << (word_bits - 1)
    assert (f < 0) == (into_signed(upper, word_bits) < 0)
    return upper, lower, f >= m or f < -m


# ---- memory (trace.rs:150-300) ------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Access:
    kind: str                    # "Init" | "Store" | "Load"
    address: int
    value: int
    time: Optional[int] = None
    pc: Optional[int] = None


class Mem:
    """Mem<WORD_BITS>: the tapes are written to memory, word i at byte address i * W / 8 (trace.rs:160-179)"""

    def __init__(self, word_bits: int, primary_tape=(), auxiliary_tape=()):
        if word_bits % 8:
            raise ValueError("WORD_BITS % 8 != 0")
        self.word_bits = word_bits
        self.accesses: Dict[int, List[Access]] = {}
        for i, w in enumerate(list(primary_tape) + list(auxiliary_tape)):
            addr = i * word_bits // 8
            self.accesses[addr] = [Access("Init", addr, w)]

    def _access(self, address):
        return self.accesses.setdefault(address, [Access("Init", address, 0)])

    def load(self, address, time, pc) -> int:
        acc = self._access(address)
        value = acc[-1].value
        acc.append(Access("Load", address, value, time, pc))
        return value

    def store(self, address, time, pc, value):
        assert value <= 1 << self.word_bits
        self._access(address).append(Access("Store", address, value, time, pc))

    def access_count(self) -> int:
        return sum(len(v) for v in self.accesses.values())


# ---- the interpreter (trace.rs:367-551) ---------------------------------------------------------------------------------------------
@dataclass
class Step:
    time: int
    pc: int
    instruction: Instruction
    regs: tuple
    flag: bool
    v_addr: Optional[int]


@dataclass
class Trace:
    word_bits: int
    reg_count: int
    prog: List[Instruction]
    exe: List[Step]
    mem: Mem
    ans: int


def get_word_size_bit_mask_msb(word_bits: int) -> int:
    m = 1 << word_bits
    return m * (m - 1)


def eval_program(prog: List[Instruction], mem: Mem, reg_count: int = 8, max_steps: Optional[int] = None) -> Trace:
    """Program::eval::<WORD_BITS, REG_COUNT> (trace.rs:378-551).  Raises IndexError where the reference panics with
    "Program did not Answer 0 or 1." (pc runs off the program); max_steps bounds runaway programs (not in the reference)."""
    W = mem.word_bits
    mod = 1 << W
    msb_mask = get_word_size_bit_mask_msb(W)
    regs = [0] * reg_count
    pc, time, flag = 0, 1, False
    exe: List[Step] = []

    def val(a: Operand) -> int:
        return a.value if isinstance(a, Imm) else regs[a.index]

    while True:
        if not 0 <= pc < len(prog):
            raise IndexError("Program did not Answer 0 or 1.")
        ins = prog[pc]
        n, ri, rj, a = ins.name, ins.ri, ins.rj, ins.a
        v_addr = None
        if n == "LoadW":
            v_addr = mem.load(val(a), time, pc)
        elif n == "StoreW":
            mem.store(val(a), time, pc, regs[ri])
            v_addr = regs[ri]
        exe.append(Step(time, pc, ins, tuple(regs), flag, v_addr))
        if n == "And":
            regs[ri] = regs[rj] & val(a); flag = regs[ri] == 0
        elif n == "Or":
            regs[ri] = regs[rj] | val(a); flag = regs[ri] == 0
        elif n == "Xor":
            regs[ri] = regs[rj] ^ val(a); flag = regs[ri] == 0
        elif n == "Not":
            regs[ri] = ~val(a) & 0xFFFFFFFF                       # Word(!a.0) on a u32: NOT truncated to W bits (trace.rs:431-434)
            flag = regs[ri] == 0
        elif n == "Add":
            r = regs[rj] + val(a)
            regs[ri] = r & (mod - 1); flag = (r & msb_mask) != 0
        elif n == "Sub":
            r = regs[rj] + mod - val(a)
            regs[ri] = r & (mod - 1); flag = (r & msb_mask) == 0
        elif n == "Mull":
            r = regs[rj] * val(a)
            regs[ri] = r % mod; flag = r < mod                     # trace.rs:447-453 (inverted w.r.t. the specification)
        elif n == "UMulh":
            r = regs[rj] * val(a)
            regs[ri] = (r >> W) & (mod - 1); flag = regs[ri] == 0
        elif n == "SMulh":
            upper, _lower, _f = smulh_eval(val(a), regs[rj], W)
            regs[ri] = upper; flag = upper == 0
        elif n == "UDiv":
            av = val(a)
            regs[ri] = 0 if av == 0 else regs[rj] // av; flag = av == 0
        elif n == "UMod":
            av = val(a)
            regs[ri] = 0 if av == 0 else regs[rj] % av; flag = av == 0
        elif n == "Shl":
            av, b = val(a), regs[rj]
            if av >= 32:
                raise OverflowError("attempt to shift left with overflow")      # u32 `rj << a` panics in the reference's debug build
            regs[ri] = ((b << av) & 0xFFFFFFFF) & (mod - 1); flag = (b & (1 << (W - 1))) != 0
        elif n == "Shr":
            av, b = val(a), regs[rj]
            if av >= 32:
                raise OverflowError("attempt to shift right with overflow")
            regs[ri] = b >> av; flag = (b & 1) != 0
        elif n == "Cmpe":
            flag = val(a) == regs[ri]
        elif n == "Cmpa":
            flag = regs[ri] > val(a)
        elif n == "Cmpae":
            flag = regs[ri] >= val(a)
        elif n == "Cmpg":
            flag = decode_signed(regs[ri], W) > decode_signed(val(a), W)
        elif n == "Cmpge":
            flag = decode_signed(regs[ri], W) >= decode_signed(val(a), W)
        elif n == "Mov":
            regs[ri] = val(a)
        elif n == "CMov":
            if flag:
                regs[ri] = val(a)
        elif n == "Jmp":
            pc = val(a)
        elif n == "CJmp":
            pc = val(a) if flag else pc + 1
        elif n == "CnJmp":
            pc = val(a) if not flag else pc + 1
        elif n == "LoadW":
            regs[ri] = v_addr
        elif n == "StoreW":
            pass
        elif n == "Answer":
            return Trace(W, reg_count, list(prog), exe, mem, val(a))
        time += 1
        if n not in ("Jmp", "CJmp", "CnJmp"):
            pc += 1
        if max_steps is not None and len(exe) >= max_steps:
            raise RuntimeError("max_steps exceeded")
