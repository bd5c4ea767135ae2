"""Derive a continuous extension (dense output) for Dopri8 = RK8(7)13M + FSAL stage.

Why: diffrax.Dopri8's interpolation coefficients are not available in this
environment (diffrax is a third-party dependency of the reference, absent from
/root/reference and not installable offline).  Rooted-tree analysis (this script)
shows that from the 14 stage slopes alone a continuous extension of order 5 exists
(the order-6 conditions are inconsistent), so we construct the C1 5th-order extension

    y(t0 + theta*h) = y0 + h * sum_i b_i(theta) f_i ,   b_i(theta) = sum_{k=1..7} B[i,k] theta^k

with exact constraints
    * all order conditions of order <= 5 for every theta,
    * b_i(1) = b_i  (so y(1) = y1), b_i'(0) = delta_{i,1}, b_i'(1) = delta_{i,14} (C1 across steps)
and the remaining freedom spent on minimising the order-6 residuals (min-norm among the
minimisers).  Output: tools/dopri8_dense_coeffs.json (14 x 7 matrix, column k = theta^k).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import tableaus as T
from rk_trees import gamma, stage_weights, trees

S, K = 14, 7


def build():
    A = [[float(v) for v in r] for r in T.D8_A]
    b = np.array([float(v) for v in T.D8_B])

    def cond_rows(p_lo, p_hi):
        rows, rhs = [], []
        for r in range(p_lo, p_hi + 1):
            for t in trees(r):
                phi = np.array(stage_weights(t, A, 1.0))
                for k in range(1, K + 1):
                    row = np.zeros((S, K))
                    row[:, k - 1] = phi
                    rows.append(row.ravel())
                    rhs.append(1.0 / gamma(t) if r == k else 0.0)
        return np.array(rows), np.array(rhs)

    C, d = cond_rows(1, 5)
    extra_rows, extra_rhs = [], []
    for i in range(S):
        row = np.zeros((S, K)); row[i, :] = 1.0
        extra_rows.append(row.ravel()); extra_rhs.append(b[i])                    # b_i(1) = b_i
        row = np.zeros((S, K)); row[i, 0] = 1.0
        extra_rows.append(row.ravel()); extra_rhs.append(1.0 if i == 0 else 0.0)  # b_i'(0)
        row = np.zeros((S, K)); row[i, :] = np.arange(1, K + 1)
        extra_rows.append(row.ravel()); extra_rhs.append(1.0 if i == S - 1 else 0.0)  # b_i'(1)
    for i in (1, 2, 3, 4):                      # stages 2..5 carry no weight (as in b and b_hat)
        for k in range(K):
            row = np.zeros((S, K)); row[i, k] = 1.0
            extra_rows.append(row.ravel()); extra_rhs.append(0.0)
    C = np.vstack([C, np.array(extra_rows)])
    d = np.concatenate([d, np.array(extra_rhs)])
    P, r = cond_rows(6, 6)

    # particular solution + null space of the constraints
    x0, *_ = np.linalg.lstsq(C, d, rcond=1e-11)
    cres = np.abs(C @ x0 - d).max()
    u, sv, vt = np.linalg.svd(C)
    rank = int((sv > 1e-11 * sv[0]).sum())
    N = vt[rank:].T
    z, *_ = np.linalg.lstsq(P @ N, r - P @ x0, rcond=1e-11)
    x = x0 + N @ z
    # polish the exact constraints in extended precision (keeps the null-space component)
    Cl, dl = C.astype(np.longdouble), d.astype(np.longdouble)
    xl = x.astype(np.longdouble)
    Cp = np.linalg.pinv(C, rcond=1e-11)
    for _ in range(5):
        xl = xl - (Cp @ np.asarray(Cl @ xl - dl, dtype=np.float64)).astype(np.longdouble)
    x = np.asarray(xl, dtype=np.float64)
    x.reshape(S, K)[1:5, :] = 0.0
    info = dict(constraint_rank=rank, n_unknowns=S * K, constraint_residual=float(np.abs(C @ x - d).max()),
                particular_residual=float(cres), order6_residual=float(np.abs(P @ x - r).max()))
    return x.reshape(S, K), info


if __name__ == "__main__":
    B, info = build()
    print(info)
    out = os.path.join(os.path.dirname(__file__), "dopri8_dense_coeffs.json")
    with open(out, "w") as f:
        json.dump(dict(info=info, B=[[float(v) for v in row] for row in B]), f, indent=1)
    print("wrote", out)
