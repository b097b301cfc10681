"""Exact weights of the extrapolation integrator (metrotrpl_b200/csrc/extrapolation.h).

Harmonic sequence n_j = j, j = 1..6.  The result of column j, T_j, has an expansion in h_j = H/j,
so T_kk is the value at 0 of the polynomial through (1/j, T_j): Lagrange weights at x_j = 1/j.
    A_j  : T_66           (order 6)
    B_j  : T_66 - T_65    (error estimate of the order-5 result)
    B5_j : T_55 - T_54    (the estimate one order down)
Checked here against the Aitken-Neville recursion T_{j,l+1} = T_{j,l} + (T_{j,l} - T_{j-1,l}) / (n_j/n_{j-l} - 1).
"""
from fractions import Fraction as F


def lagrange_at_zero(js):
    xs = {j: F(1, j) for j in js}
    out = {}
    for j in js:
        c = F(1)
        for i in js:
            if i != j:
                c *= (-xs[i]) / (xs[j] - xs[i])
        out[j] = c
    return out


def weights():
    a = lagrange_at_zero(range(1, 7))
    t65 = lagrange_at_zero(range(2, 7))
    t55 = lagrange_at_zero(range(1, 6))
    t54 = lagrange_at_zero(range(2, 6))
    b = {j: a[j] - t65.get(j, 0) for j in a}
    b5 = {j: t55.get(j, 0) - t54.get(j, 0) for j in a}
    return a, b, b5


if __name__ == "__main__":
    a, b, b5 = weights()
    print("A ", {j: str(v) for j, v in a.items()}, "sum", sum(a.values()))
    print("B ", {j: str(v) for j, v in b.items()}, "sum", sum(b.values()))
    print("B5", {j: str(v) for j, v in b5.items()}, "sum", sum(b5.values()))
    vals = [F(3, 7) + F(3, 10) / j + F(1, 5) / j ** 2 + F(1, 10) / j ** 3 - F(1, 9) / j ** 5 for j in range(1, 7)]
    T = [[None] * 6 for _ in range(6)]
    for j in range(6):
        T[j][0] = vals[j]
        for k in range(1, j + 1):
            T[j][k] = T[j][k - 1] + (T[j][k - 1] - T[j - 1][k - 1]) / (F(j + 1, j + 1 - k) - 1)
    assert T[5][5] == sum(a[j] * vals[j - 1] for j in a) == F(3, 7)       # exact for degree <= 5
    assert T[5][5] - T[5][4] == sum(b[j] * vals[j - 1] for j in b)
    assert T[4][4] - T[4][3] == sum(b5[j] * vals[j - 1] for j in b5)
    print("Aitken-Neville check passed")
