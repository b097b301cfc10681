"""Verify recalled Rosenbrock coefficients against the order conditions.

Hairer-Wanner transformed form:  (1/(h*gamma) I - J) U_i = f(y + sum a_ij U_j) + sum (c_ij/h) U_j
y1 = y + sum m_i U_i.   With Ginv = diag(1/gamma) - C (lower-tri), Gamma = inv(Ginv),
alpha = A @ Gamma,  b = m @ Gamma.
"""
import numpy as np
import itertools

def to_classic(A, C, gamma, m):
    s = len(m)
    Ginv = np.eye(s) / gamma - C
    G = np.linalg.inv(Ginv)
    alpha = A @ G
    b = m @ G
    return alpha, G, b

def order_residuals(alpha, G, b):
    """Rosenbrock order conditions up to order 5 (autonomous), beta = alpha + Gamma."""
    s = len(b)
    beta = alpha + G
    e = np.ones(s)
    al = alpha @ e          # alpha_i
    be = beta @ e           # beta_i (row sums incl. gamma)
    g = np.diag(G)[0]
    res = {}
    res["1"] = b @ e - 1
    res["2"] = b @ be - 0.5
    res["3a"] = b @ (al**2) - 1/3
    res["3b"] = b @ (beta @ be) - 1/6
    res["4a"] = b @ (al**3) - 1/4
    res["4b"] = b @ (al * (alpha @ be)) - 1/8
    res["4c"] = b @ (beta @ (al**2)) - 1/12
    res["4d"] = b @ (beta @ (beta @ be)) - 1/24
    return res

def rodas4():
    g = 0.25
    A = np.zeros((6, 6)); C = np.zeros((6, 6))
    A[1,0]=0.1544000000000000e+01
    A[2,0]=0.9466785280815826e+00; A[2,1]=0.2557011698983284e+00
    A[3,0]=0.3314825187068521e+01; A[3,1]=0.2896124015972201e+01; A[3,2]=0.9986419139977817e+00
    A[4,0]=0.1221224509226641e+01; A[4,1]=0.6019134481288629e+01; A[4,2]=0.1253708332932087e+02; A[4,3]=-0.6878860361058950e+00
    A[5,:5] = A[4,:5]; A[5,4] = 1.0
    C[1,0]=-0.5668800000000000e+01
    C[2,0]=-0.2430093356833875e+01; C[2,1]=-0.2063599157091915e+00
    C[3,0]=-0.1073529058151375e+00; C[3,1]=-0.9594562251023355e+01; C[3,2]=-0.2047028614809616e+02
    C[4,0]=0.7496443313967647e+01; C[4,1]=-0.1024680431464352e+02; C[4,2]=-0.3399990352819905e+02; C[4,3]=0.1170890893206160e+02
    C[5,0]=0.8083246795921522e+01; C[5,1]=-0.7981132988064893e+01; C[5,2]=-0.3152159432874371e+02; C[5,3]=0.1631930543123136e+02; C[5,4]=-0.6058818238834054e+01
    m = np.array([A[4,0], A[4,1], A[4,2], A[4,3], 1.0, 1.0])
    mhat = np.array([A[4,0], A[4,1], A[4,2], A[4,3], 1.0, 0.0])
    return A, C, g, m, mhat

if __name__ == "__main__":
    A, C, g, m, mhat = rodas4()
    alpha, G, b = to_classic(A, C, g, m)
    print("diag Gamma", np.diag(G))
    print("alpha_i", alpha.sum(1))
    for k, v in order_residuals(alpha, G, b).items():
        print("main", k, f"{v:.3e}")
    alpha, G, bh = to_classic(A, C, g, mhat)
    for k, v in order_residuals(alpha, G, bh).items():
        print("emb ", k, f"{v:.3e}")


def dopri5():
    """Dormand-Prince 5(4) tableau used by the explicit path (csrc/explicit.h)."""
    from fractions import Fraction as F
    c = [F(0), F(1, 5), F(3, 10), F(4, 5), F(8, 9), F(1), F(1)]
    a = [[], [F(1, 5)], [F(3, 40), F(9, 40)], [F(44, 45), F(-56, 15), F(32, 9)],
         [F(19372, 6561), F(-25360, 2187), F(64448, 6561), F(-212, 729)],
         [F(9017, 3168), F(-355, 33), F(46732, 5247), F(49, 176), F(-5103, 18656)],
         [F(35, 384), F(0), F(500, 1113), F(125, 192), F(-2187, 6784), F(11, 84)]]
    b5 = a[6] + [F(0)]
    b4 = [F(5179, 57600), F(0), F(7571, 16695), F(393, 640), F(-92097, 339200), F(187, 2100), F(1, 40)]
    return c, a, b5, b4


def check_dopri5():
    from fractions import Fraction as F
    c, a, b5, b4 = dopri5()
    s = 7
    A = [[(a[i][j] if j < len(a[i]) else F(0)) for j in range(s)] for i in range(s)]
    for i in range(s):
        assert sum(A[i]) == c[i], i
    def dot(x, y): return sum(p * q for p, q in zip(x, y))
    Ac = [dot(A[i], c) for i in range(s)]
    Ac2 = [dot(A[i], [x * x for x in c]) for i in range(s)]
    AAc = [dot(A[i], Ac) for i in range(s)]
    for name, b, order in (("b5", b5, 5), ("b4", b4, 4)):
        conds = {"1": (dot(b, [1] * s), F(1)), "2": (dot(b, c), F(1, 2)), "3a": (dot(b, [x**2 for x in c]), F(1, 3)),
                 "3b": (dot(b, Ac), F(1, 6)), "4a": (dot(b, [x**3 for x in c]), F(1, 4)),
                 "4b": (dot(b, [c[i] * Ac[i] for i in range(s)]), F(1, 8)), "4c": (dot(b, Ac2), F(1, 12)),
                 "4d": (dot(b, AAc), F(1, 24))}
        if order >= 5:
            conds["5a"] = (dot(b, [x**4 for x in c]), F(1, 5))
            conds["5b"] = (dot(b, [c[i]**2 * Ac[i] for i in range(s)]), F(1, 10))
            conds["5c"] = (dot(b, [Ac[i]**2 for i in range(s)]), F(1, 20))
        for k, (lhs, rhs) in conds.items():
            assert lhs == rhs, (name, k, lhs, rhs)
    print("DOPRI5: order conditions hold exactly (rational arithmetic); error weights:",
          [str(x - y) for x, y in zip(b5, b4)])


if __name__ == "__main__":
    check_dopri5()
