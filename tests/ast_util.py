"""Random poly::Ast generator shared by the CPU and GPU tests of the quotient evaluator."""
import random


def random_ast(P, rng: random.Random, n_polys: int, depth: int, p: int):
    """P = the product's poly module (node classes).  Exercises every node type and operator overload."""
    if depth == 0 or rng.random() < 0.15:
        r = rng.random()
        if r < 0.7:
            return P.Poly(rng.randrange(n_polys), rng.choice([0, 0, 1, -1, 2, -3]))
        if r < 0.85:
            return P.ConstantTerm(rng.choice([0, 1, 2, p - 1, rng.randrange(p)]))
        return P.LinearTerm(rng.choice([1, rng.randrange(p)]))
    sub = lambda: random_ast(P, rng, n_polys, depth - 1, p)
    r = rng.random()
    if r < 0.25:
        return sub() + sub()
    if r < 0.35:
        return sub() - sub()
    if r < 0.65:
        return sub() * sub()
    if r < 0.72:
        return sub() * rng.choice([2, p - 1, rng.randrange(p)])
    if r < 0.77:
        return -sub()
    if r < 0.82:
        x = sub()
        return P.Mul(x, x) if rng.random() < 0.5 else P.Add(x, x)
    if r < 0.9:
        return sub() + rng.randrange(p)
    return P.DistributePowers([sub() for _ in range(rng.randrange(1, 4))], sub())


def gate_like_ast(P, leaves, y: int):
    """h = fold(0, |h, e| h * y + e) over a few TinyRAM-style gate expressions (selector * polynomial identity)."""
    a, b, c, s = leaves[:4]
    exprs = [s * (a * b - c),
             s * a * (a - 1),                                   # booleanity
             s * (a.with_rotation(1) - a - b),                   # next-row relation
             (s * (a * b)) * (c * c - a.with_rotation(-1))]      # degree-5 product
    h = P.ConstantTerm(0)
    for e in exprs:
        h = h * y + e
    return h
