"""Rooted-tree machinery for Runge-Kutta order conditions (Butcher theory).

Used by tools/derive_dopri8_dense.py (derivation of the Dopri8 continuous
extension) and by tests/test_tableau.py (order-condition checks of the
tableaus the kernels and the oracle carry).  Pure Python, exact when fed
fractions.Fraction.
"""
from functools import lru_cache


@lru_cache(maxsize=None)
def trees(order):
    """All rooted trees with `order` vertices, as canonical nested tuples."""
    if order == 1:
        return ((),)
    out = set()
    # a tree of order n is a root with a multiset of subtrees of total order n-1
    def parts(n, maxpart):
        if n == 0:
            yield ()
            return
        for p in range(min(n, maxpart), 0, -1):
            for rest in parts(n - p, p):
                yield (p,) + rest
    for part in parts(order - 1, order - 1):
        def build(idx, chosen):
            if idx == len(part):
                out.add(tuple(sorted(chosen)))
                return
            for t in trees(part[idx]):
                build(idx + 1, chosen + [t])
        build(0, [])
    return tuple(sorted(out))


def order(t):
    return 1 + sum(order(s) for s in t)


def gamma(t):
    g = order(t)
    for s in t:
        g *= gamma(s)
    return g


def stage_weights(t, A, one=1):
    """Phi_i(t) for every stage i (list), A = lower-triangular list of rows."""
    s = len(A)
    if t == ():
        return [one] * s
    res = [one] * s
    for sub in t:
        phi = stage_weights(sub, A, one)
        inner = [sum((A[i][j] * phi[j] for j in range(len(A[i])) if A[i][j] != 0), 0 * one) for i in range(s)]
        res = [res[i] * inner[i] for i in range(s)]
    return res
