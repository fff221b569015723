"""Mirror of halo2_proofs::poly::evaluator (poly/evaluator.rs, halo2_proofs 0.2.0 @ a95945254dcc, the un-vendored
dependency of the reference, Cargo.lock:619-621): ``new_evaluator`` / ``Evaluator::register_poly`` / ``AstLeaf::with_rotation``
/ the ``Ast`` node set / ``Evaluator::evaluate(&ast, domain)``.  create_proof builds one big ``Ast<ExtendedLagrangeCoeff>``
(gates, permutation and lookup constraints folded with powers of y) and evaluates it over the extended domain; that
evaluation is the third kernel family of the hot path.

Here ``evaluate`` lowers the AST once into a straight-line program (include/tr_prover.h, trp_dev_quotient_eval) and
the GPU runs it for every row.  Scalars are Python ints in canonical form; polynomials are numpy uint64 (rows, 4)
Montgomery arrays or raw device pointers.
"""
from __future__ import annotations

import ctypes
import sys
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from ._lib import as_u64, ptr

# opcodes of include/tr_prover.h
LOAD, CONST, ADD, SUB, MUL, NEG, SQR, DBL, COSETX, STORE, MULC, ADDC, SUBC = range(13)


class Ast:
    """Base of the node set of halo2's ``poly::Ast`` with the same operator overloads (Add, Sub, Neg, Mul, Mul<F>)."""

    def __add__(self, other):
        return Add(self, _lift(other))

    __radd__ = __add__

    def __neg__(self):
        return Scale(self, -1)               # impl Neg for Ast: Ast::Scale(Arc::new(self), -F::one())

    def __sub__(self, other):
        return Add(self, -_lift(other))      # impl Sub for Ast: self + (-other)

    def __rsub__(self, other):
        return Add(_lift(other), -self)

    def __mul__(self, other):
        if isinstance(other, int):
            return Scale(self, other)        # impl Mul<F> for Ast
        return Mul(self, other)

    def __rmul__(self, other):
        return Scale(self, other)


def _lift(x):
    return ConstantTerm(x) if isinstance(x, int) else x


@dataclass(eq=False)
class Poly(Ast):
    """Ast::Poly(AstLeaf { index, rotation })"""
    index: int
    rotation: int = 0

    def with_rotation(self, rotation: int) -> "Poly":
        return Poly(self.index, rotation)


@dataclass(eq=False)
class Add(Ast):
    a: Ast
    b: Ast


@dataclass(eq=False)
class Mul(Ast):
    a: Ast
    b: Ast


@dataclass(eq=False)
class Scale(Ast):
    a: Ast
    scalar: int


@dataclass(eq=False)
class DistributePowers(Ast):
    """Ast::DistributePowers(terms, base): fold(0, |acc, term| acc * base + term)"""
    terms: Sequence[Ast]
    base: Ast


@dataclass(eq=False)
class LinearTerm(Ast):
    """Ast::LinearTerm(scalar): scalar * X, X running over the extended coset zeta * extended_omega^row"""
    scalar: int


@dataclass(eq=False)
class ConstantTerm(Ast):
    scalar: int


@dataclass
class Program:
    code: np.ndarray            # (n_instr, 4) uint32
    consts: List[int]           # canonical ints
    n_regs: int
    n_cols: int

    def counts(self):
        ops = self.code[:, 0]
        return {"instr": len(ops), "mul": int(np.isin(ops, (MUL, SQR, MULC, COSETX)).sum()),
                "addsub": int(np.isin(ops, (ADD, SUB, NEG, DBL, ADDC, SUBC)).sum()), "load": int((ops == LOAD).sum())}


class _Compiler:
    def __init__(self, modulus: int):
        self.p = modulus
        self.code: List[tuple] = []
        self.consts: List[int] = []
        self.cidx = {}
        self.free: List[int] = []
        self.n_regs = 0
        self.n_cols = 0
        self.need_memo = {}

    def const(self, v: int) -> int:
        v %= self.p
        if v not in self.cidx:
            self.cidx[v] = len(self.consts)
            self.consts.append(v)
        return self.cidx[v]

    def alloc(self) -> int:
        if self.free:
            return self.free.pop()
        self.n_regs += 1
        return self.n_regs - 1

    def release(self, r: int):
        self.free.append(r)

    def emit(self, op, dst=0, a=0, b=0):
        self.code.append((op, dst, a, b & 0xffffffff))

    def need(self, node) -> int:
        """Sethi-Ullman register need: evaluate the hungrier operand first."""
        key = id(node)
        if key in self.need_memo:
            return self.need_memo[key]
        if isinstance(node, (Poly, ConstantTerm, LinearTerm)):
            r = 1
        elif isinstance(node, Scale):
            r = self.need(node.a)
        elif isinstance(node, (Add, Mul)):
            if isinstance(node.b, ConstantTerm) or node.a is node.b:
                r = self.need(node.a)
            else:
                x, y = self.need(node.a), self.need(node.b)
                r = max(x, y) if x != y else x + 1
        elif isinstance(node, DistributePowers):
            r = max([self.need(node.base)] + [2 + self.need(t) for t in node.terms])
        else:
            raise TypeError(f"not an Ast node: {node!r}")
        self.need_memo[key] = r
        return r

    def gen(self, node) -> int:
        if isinstance(node, Poly):
            r = self.alloc()
            self.n_cols = max(self.n_cols, node.index + 1)
            self.emit(LOAD, r, node.index, node.rotation)
            return r
        if isinstance(node, ConstantTerm):
            r = self.alloc()
            self.emit(CONST, r, self.const(node.scalar))
            return r
        if isinstance(node, LinearTerm):
            r = self.alloc()
            self.emit(COSETX, r)
            if node.scalar % self.p != 1:
                self.emit(MULC, r, r, self.const(node.scalar))
            return r
        if isinstance(node, Scale):
            r = self.gen(node.a)
            s = node.scalar % self.p
            if s == self.p - 1:
                self.emit(NEG, r, r)
            elif s == 2:
                self.emit(DBL, r, r)
            elif s != 1:
                self.emit(MULC, r, r, self.const(s))
            return r
        if isinstance(node, (Add, Mul)):
            is_mul = isinstance(node, Mul)
            if node.a is node.b:
                r = self.gen(node.a)
                self.emit(SQR if is_mul else DBL, r, r)
                return r
            a, b = node.a, node.b
            if isinstance(a, ConstantTerm) and not isinstance(b, ConstantTerm):
                a, b = b, a                                        # both operations commute
            if isinstance(b, ConstantTerm):
                r = self.gen(a)
                self.emit(MULC if is_mul else ADDC, r, r, self.const(b.scalar))
                return r
            if self.need(b) > self.need(a):
                a, b = b, a
            ra = self.gen(a)
            rb = self.gen(b)
            self.emit(MUL if is_mul else ADD, ra, ra, rb)
            self.release(rb)
            return ra
        if isinstance(node, DistributePowers):
            rb = self.gen(node.base)
            acc = self.alloc()
            self.emit(CONST, acc, self.const(0))
            for t in node.terms:
                self.emit(MUL, acc, acc, rb)
                rt = self.gen(t)
                self.emit(ADD, acc, acc, rt)
                self.release(rt)
            self.release(rb)
            return acc
        raise TypeError(f"not an Ast node: {node!r}")


def compile_ast(ast: Ast, modulus: int) -> Program:
    """Lower an Ast to the straight-line program of trp_dev_quotient_eval (ends with STORE of the root)."""
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 100000))
    try:
        c = _Compiler(modulus)
        root = c.gen(ast)
        c.emit(STORE, 0, root)
    finally:
        sys.setrecursionlimit(old)
    return Program(np.array(c.code, dtype=np.uint32).reshape(-1, 4), c.consts, c.n_regs, c.n_cols)


_MODULUS = {0: 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001,    # Fp (Vesta scalars)
            1: 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001}    # Fq (Pallas scalars)


def _to_mont_limbs(vals, p):
    out = np.zeros((max(len(vals), 1), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v << 256) % p
        for l in range(4):
            out[i, l] = (m >> (64 * l)) & 0xffffffffffffffff
    return out


class Evaluator:
    """poly::new_evaluator() / Evaluator { polys }.  Polynomials are registered in extended-Lagrange form
    (host arrays of 2^extended_k values) or, for device-resident provers, as device pointers."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.polys = []
        self.modulus = _MODULUS[0 if ctx.curve == 1 else 1]

    def register_poly(self, poly) -> Poly:
        """Evaluator::register_poly(poly) -> AstLeaf (rotation 0)"""
        self.polys.append(poly if isinstance(poly, int) else as_u64(poly))
        return Poly(len(self.polys) - 1, 0)

    def compile(self, ast: Ast) -> Program:
        prog = compile_ast(ast, self.modulus)
        if prog.n_cols > len(self.polys):
            raise ValueError("the Ast references a polynomial that was never registered")
        return prog

    def evaluate(self, ast: Ast, domain):
        """Evaluator::evaluate(&ast, domain) -> Polynomial<F, ExtendedLagrangeCoeff> (host arrays in, host array out)."""
        prog = self.compile(ast)
        rows = domain.extended_len()
        for p in self.polys:
            if isinstance(p, int) or p.shape != (rows, 4):
                raise ValueError(f"registered polynomials must be host arrays of {rows} extended-domain values")
        consts = _to_mont_limbs(prog.consts, self.modulus)
        colp = (ctypes.c_void_p * max(len(self.polys), 1))(*[p.ctypes.data for p in self.polys])
        out = np.empty((rows, 4), dtype=np.uint64)
        code = np.ascontiguousarray(prog.code)
        self.ctx.check(self.ctx.lib.trp_quotient_eval(domain.handle, ptr(code), len(code), prog.n_regs, ptr(consts),
                                                      len(prog.consts), colp, len(self.polys), ptr(out)))
        return out

    def evaluate_device(self, prog: Program, domain, d_col_ptrs, d_out_ptr, coset: int = -1):
        """Device-resident form: ``d_col_ptrs`` are device addresses (ints) of the columns on the whole extended domain
        (coset = -1) or on the size-n coset ``coset``; the result is written to device memory at ``d_out_ptr``."""
        consts = _to_mont_limbs(prog.consts, self.modulus)
        colp = (ctypes.c_void_p * max(len(d_col_ptrs), 1))(*d_col_ptrs)
        code = np.ascontiguousarray(prog.code)
        self.ctx.check(self.ctx.lib.trp_dev_quotient_eval(domain.handle, ptr(code), len(code), prog.n_regs, ptr(consts),
                                                          len(prog.consts), colp, len(d_col_ptrs), coset, d_out_ptr))


    def evaluate_device_rows(self, prog: Program, domain, d_col_ptrs, d_out_ptr, coset: int, row0: int, nrows: int,
                             halo_before: int, halo_after: int):
        """Row-slice form (trp_dev_quotient_eval_rows): the columns hold rows [row0 - halo_before, row0 + nrows + halo_after)
        of size-n coset ``coset``; the nrows results go to d_out_ptr contiguously."""
        consts = _to_mont_limbs(prog.consts, self.modulus)
        colp = (ctypes.c_void_p * max(len(d_col_ptrs), 1))(*d_col_ptrs)
        code = np.ascontiguousarray(prog.code)
        self.ctx.check(self.ctx.lib.trp_dev_quotient_eval_rows(domain.handle, ptr(code), len(code), prog.n_regs, ptr(consts),
                                                               len(prog.consts), colp, len(d_col_ptrs), coset, row0, nrows,
                                                               halo_before, halo_after, d_out_ptr))


def new_evaluator(ctx) -> Evaluator:
    return Evaluator(ctx)
