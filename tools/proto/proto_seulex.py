"""Prototype behind csrc/extrapolation.h: RODAS4 against extrapolated linearly implicit Euler
(SEULEX scheme, k columns of the harmonic sequence) on fixture curves, dense NumPy linear algebra on
the oracle's excess-variable model with its exact Jacobian.  Counts accepted steps, solves and
factorisations and the length of the longest sequential chain (6 per RODAS4 step, k per extrapolation
step when the columns run in parallel), and compares the signal at the end of the window with a
RODAS4 run at rtol 1e-9.

    python tools/proto/proto_seulex.py 0,1 7,0 15,4        # (state, measurement) pairs of staub6.npz

Measured (rtol 1e-7): steps 240 / 520 / 434 (RODAS4) against 69 / 123 / 107 (k = 6) and 37 / 47 / 57
(k = 8; its steps are too long for the three-point dense output).
"""
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from oracle import excess_model as em, trpl_oracle as orc
from tests import parity_cases as pc

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "staub6.npz"))
names = [str(n) for n in g["names"]]; idx = {n: i for i, n in enumerate(names)}
t_meas = g["t"]; nx = 128

def setup(s_, m):
    args = pc._model_args(g["states"][s_], g["units"], idx, float(g["lengths"][m]), nx)
    dN0 = g["ini"][m] * 1e-21
    y0 = np.concatenate([dN0, dN0, np.zeros(nx + 1)])
    f = lambda y: em.rhs_excess(0.0, y, *args)
    J = lambda y: em.jac_excess(0.0, y, *args)
    n0, p0, ks, dx = args[2], args[3], args[6], args[1]
    def signal(y):
        a, b = y[:nx], y[nx:2*nx]
        return orc.integrate_nodes(dx, ks * (n0*b + p0*a + a*b)) * 1e23
    return y0, f, J, signal

def sig_rate(s_, m, y, f):
    args = pc._model_args(g['states'][s_], g['units'], idx, float(g['lengths'][m]), nx)
    n0, p0, ks, dx = args[2], args[3], args[6], args[1]
    fy = f(y); a, b = y[:nx], y[nx:2*nx]; fa, fb = fy[:nx], fy[nx:2*nx]
    return orc.integrate_nodes(dx, ks * (n0*fb + p0*fa + fa*b + a*fb)) * 1e23


def err_norm(e, y, ynew, rtol, floor):
    sc = rtol * np.maximum(np.maximum(np.abs(y), np.abs(ynew)), floor)
    # field components: scale by max field
    L = nx
    sc[2*L:] = rtol * max(np.abs(y[2*L:]).max(), np.abs(ynew[2*L:]).max(), 1e-6) / 0.03
    return np.sqrt(np.mean((e / sc) ** 2))

A = np.array([[0,0,0,0,0,0],[0.1544e+01,0,0,0,0,0],[0.9466785280815826,0.2557011698983284,0,0,0,0],
 [0.3314825187068521e+01,0.2896124015972201e+01,0.9986419139977817,0,0,0],
 [0.1221224509226641e+01,0.6019134481288629e+01,0.1253708332932087e+02,-0.6878860361058950,0,0],
 [0.1221224509226641e+01,0.6019134481288629e+01,0.1253708332932087e+02,-0.6878860361058950,1.0,0]])
C = np.array([[0,0,0,0,0,0],[-0.56688e+01,0,0,0,0,0],[-0.2430093356833875e+01,-0.2063599157091915,0,0,0,0],
 [-0.1073529058151375,-0.9594562251023355e+01,-0.2047028614809616e+02,0,0,0],
 [0.7496443313967647e+01,-0.1024680431464352e+02,-0.3399990352819905e+02,0.1170890893206160e+02,0,0],
 [0.8083246795921522e+01,-0.7981132988064893e+01,-0.3152159432874371e+02,0.1631930543123136e+02,-0.6058818238834054e+01,0]])

DEBUG = False

HCAP = None

def run(method, s_, m, rtol, k=6, want_ds=False):
    y0, f, J, signal = setup(s_, m)
    floor = 1e-12 * np.abs(y0[:nx]).max()
    tend = t_meas[-1]
    t, y = 0.0, y0.copy()
    n_acc = n_rej = nsolve = nfac = 0
    h = 1e-6
    ts, ys = [0.0], [signal(y)]
    ds = [sig_rate(s_, m, y, f)]
    I = np.eye(len(y))
    order = 4 if method == "rodas4" else k
    while t < tend:
        h = min(h, tend - t)
        if HCAP is not None and ds[-1] != 0: h = min(h, HCAP * abs(ys[-1] / ds[-1]))
        Jy = J(y)
        if method == "rodas4":
            W = I / (0.25 * h) - Jy; nfac += 1
            LU = np.linalg.inv(W)
            K = []
            for s in range(6):
                us = y + sum(A[s][p] * K[p] for p in range(s))
                r = f(us) + sum(C[s][p] / h * K[p] for p in range(s))
                K.append(LU @ r); nsolve += 1
            ynew = y + sum(A[5][p] * K[p] for p in range(5)) + K[5]
            e = K[5]
        else:
            T = []
            for j in range(1, k + 1):
                hj = h / j
                Wi = np.linalg.inv(I - hj * Jy); nfac += 1
                yy = y.copy()
                for mm in range(j):
                    yy = yy + Wi @ (hj * f(yy)); nsolve += 1
                row = [yy]
                for l in range(1, j):
                    row.append(row[l-1] + (row[l-1] - T[j-2][l-1]) / (j / (j - l) - 1))
                T.append(row)
            ynew = T[k-1][k-1]
            e = T[k-1][k-1] - T[k-1][k-2]
        err = err_norm(e, y, ynew, rtol, floor)
        if not np.isfinite(err): err = 1e10
        fac = min(5.0, max(0.2, 0.9 * err ** (-1.0 / (order if method != "rodas4" else 4))))
        if (n_acc + n_rej) % 50 == 0 and DEBUG: print(method, 'step', n_acc, n_rej, 't', t, 'h', h, 'err', err, flush=True)
        if err <= 1.0:
            t += h; y = ynew; n_acc += 1
            ts.append(t); ys.append(signal(y)); ds.append(sig_rate(s_, m, y, f))
        else:
            n_rej += 1
        h *= fac
    if want_ds:
        return np.array(ts), np.array(ys), n_acc, n_rej, nsolve, nfac, np.array(ds)
    return np.array(ts), np.array(ys), n_acc, n_rej, nsolve, nfac


if __name__ == "__main__":
    cases = [(int(a), int(b)) for a, b in (x.split(",") for x in sys.argv[1:])] or [(0, 1)]
    for s_, m in cases:
        ref = run("rodas4", s_, m, 1e-9)
        for method, rtol, k in (("rodas4", 1e-7, 0), ("seulex", 1e-7, 4), ("seulex", 1e-7, 6), ("seulex", 1e-7, 8)):
            t0 = time.time()
            ts, ys, na, nr, ns, nf = run(method, s_, m, rtol, k)
            chain = na * (6 if method == "rodas4" else k)
            print(f"state {s_} meas {m} {method:7s} k={k} rtol={rtol:g}: steps {na:4d}+{nr:2d} solves {ns:6d} "
                  f"factorisations {nf:5d} longest chain {chain:5d}  signal at t_end vs tight {abs(ys[-1] / ref[1][-1] - 1):.1e} "
                  f"({time.time() - t0:.0f}s)", flush=True)
