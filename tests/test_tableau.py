"""Exact (rational) order-condition checks of the Butcher tableaus both the kernels and the oracle carry, and
consistency of the generated C headers with those rationals (CPU only)."""
import os
import re
import sys
from fractions import Fraction as F

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tableaus as T  # noqa: E402
from rk_trees import gamma, order, stage_weights, trees  # noqa: E402


def _residual(A, w, p):
    worst = F(0)
    for r in range(1, p + 1):
        for t in trees(r):
            phi = stage_weights(t, A, F(1))
            worst = max(worst, abs(sum(wi * ph for wi, ph in zip(w, phi)) - F(1, gamma(t))))
    return worst


def test_tree_counts():
    assert [len(trees(k)) for k in range(1, 9)] == [1, 1, 2, 4, 9, 20, 48, 115]


def test_dopri5_order_conditions_exact():
    assert all(sum(row) == c for row, c in zip(T.D5_A, T.D5_C))
    assert _residual(T.D5_A, T.D5_B, 5) == 0                      # 5th-order weights: all 17 trees, exactly
    bhat = [b - e for b, e in zip(T.D5_B, T.D5_E)]
    assert _residual(T.D5_A, bhat, 4) == 0                        # embedded weights: 4th order
    assert _residual(T.D5_A, bhat, 5) != 0
    assert T.nnz(T.D5_A) == 20 and sum(1 for e in T.D5_E if e) == 6
    # mid-point weights of the dense output: quadrature conditions at theta = 1/2 up to degree 3
    for k in range(4):
        assert sum(cm * c ** k for cm, c in zip(T.D5_CMID, T.D5_C)) == F(1, 2) ** (k + 1) / (k + 1)


def test_dopri8_order_conditions():
    # RK8(7)13M is published as rational approximations: conditions hold to ~1e-17, not exactly
    assert max(abs(sum(row) - c) for row, c in zip(T.D8_A, T.D8_C)) < F(1, 10 ** 16)
    assert _residual(T.D8_A, T.D8_B, 8) < F(1, 10 ** 15)          # 8th order: all 200 trees
    assert _residual(T.D8_A, T.D8_BHAT, 7) < F(1, 10 ** 15)       # embedded: 7th order
    assert _residual(T.D8_A, T.D8_BHAT, 8) > F(1, 10 ** 6)
    assert T.nnz(T.D8_A[:13]) == 59 and sum(1 for b in T.D8_B if b) == 9 and sum(1 for e in T.D8_E if e) == 9   # SURVEY 8d counts
    assert T.D8_A[13] == T.D8_B and T.D8_E[13] == 0               # FSAL row, unused in the error estimate


def test_dopri8_dense_output_is_fifth_order_and_c1():
    import json
    B = np.array(json.load(open(os.path.join(ROOT, "tools", "dopri8_dense_coeffs.json")))["B"])
    A = [[float(v) for v in r] for r in T.D8_A]
    b = np.array([float(v) for v in T.D8_B])
    for theta in (0.1, 0.37, 0.5, 0.9, 1.0):
        w = (B * theta ** np.arange(1, 8)).sum(axis=1)
        for r in range(1, 6):
            for t in trees(r):
                phi = np.array(stage_weights(t, A, 1.0))
                assert abs(w @ phi - theta ** r / gamma(t)) < 2e-12
    assert np.abs(B.sum(axis=1) - b).max() < 1e-12                # y(1) = y1
    assert np.abs(B[:, 0] - np.eye(14)[0]).max() < 1e-12           # y'(0) = f0
    assert np.abs((B * np.arange(1, 8)).sum(axis=1) - np.eye(14)[13]).max() < 1e-11   # y'(1) = f(y1)


def _parse_header(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"double (\w+)((?:\[\d+\])+) = \{(.*?)\};", txt, re.S):
        vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(3).replace("{", " ").replace("}", " "))]
        out[m.group(1)] = np.array(vals)
    return out


def test_generated_headers_match_rationals():
    for path in (os.path.join(ROOT, "streamsculptor_b200", "csrc", "ssb_tableau.h"), os.path.join(ROOT, "oracle", "orc_tableau.h")):
        h = _parse_header(path)
        assert np.array_equal(h["d5_a"], np.array([float(v) for r in T.D5_A for v in r]))
        assert np.array_equal(h["d8_a"], np.array([float(v) for r in T.D8_A for v in r]))
        assert np.array_equal(h["d8_e"], np.array([float(v) for v in T.D8_E]))
        assert np.array_equal(h["d5_e"], np.array([float(v) for v in T.D5_E]))
        assert np.array_equal(h["d8_c"], np.array([float(v) for v in T.D8_C]))
    # Nystrom-form arrays of the product header: aa = A.A, ea = e^T A (exact rationals, rounded once)
    h = _parse_header(os.path.join(ROOT, "streamsculptor_b200", "csrc", "ssb_tableau.h"))
    n = 14
    AA = [[sum((T.D8_A[i][j] * T.D8_A[j][l] for j in range(n)), F(0)) for l in range(n)] for i in range(n)]
    assert np.array_equal(h["d8_aa"], np.array([float(v) for r in AA for v in r]))
    ea = [sum((T.D8_E[i] * T.D8_A[i][l] for i in range(n)), F(0)) for l in range(n)]
    assert np.array_equal(h["d8_ea"], np.array([float(v) for v in ea]))
