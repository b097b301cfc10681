"""CPU prototype of the GPU integrator (design tool only; NOT on any product path).

State u_i = (N_i, Q_{i+1}), Q_i = cumulative charge sum_{j<i}(P_j - N_j - (p0-n0)), E_i = Lambda*dx*Q_i.
P_i = N_i + (p0-n0) + Q_{i+1} - Q_i.   Q_0 = 0 (corner_E), Q_L stays 0 (net flux of charge is zero).
"""
import numpy as np
from scipy.linalg import solve_banded
from check_rodas_coeffs import rodas4

kB = 8.61773e-5
eps0 = 8.854e-12 * 1e-9
q_C = 1.602e-19

class Par:
    pass

def make_par(state_conv, idx, thickness, nx):
    p = Par()
    s = state_conv
    p.L = nx; p.dx = thickness / nx
    p.n0 = s[idx["n0"]]; p.p0 = s[idx["p0"]]
    p.mun = s[idx["mu_n"]]; p.mup = s[idx["mu_p"]]
    p.ks = s[idx["ks"]]; p.Cn = s[idx["Cn"]]; p.Cp = s[idx["Cp"]]
    p.Sf = s[idx["Sf"]]; p.Sb = s[idx["Sb"]]
    p.tauN = s[idx["tauN"]]; p.tauP = s[idx["tauP"]]
    p.Lam = q_C / (s[idx["eps"]] * eps0); p.Tm = s[idx["Tm"]]
    return p

def unpack(p, y):
    L = p.L
    N = y[0::2]; Qr = y[1::2]           # Qr[i] = Q_{i+1}
    Ql = np.concatenate(([0.0], Qr[:-1]))  # Q_i
    P = N + (p.p0 - p.n0) + (Qr - Ql)
    return N, P, Ql, Qr

def rhs(p, y, want_jac=False):
    L = p.L; dx = p.dx
    N, P, Ql, Qr = unpack(p, y)
    kT = kB * p.Tm
    LD = p.Lam * dx
    # interior faces i=1..L-1 : between node i-1 and i ; E_i = LD*Q_i = LD*Ql[i]
    E = LD * Ql[1:]
    Nm = N[:-1]; Np_ = N[1:]; Pm = P[:-1]; Pp = P[1:]
    Jn = np.zeros(L + 1); Jp = np.zeros(L + 1)
    Jn[1:L] = p.mun * (0.5 * (Nm + Np_) * E + kT * (Np_ - Nm) / dx)
    Jp[1:L] = p.mup * (0.5 * (Pm + Pp) * E - kT * (Pp - Pm) / dx)
    NP = N * P - p.n0 * p.p0
    Sft = p.Sf * NP[0] / (N[0] + P[0])
    Sbt = p.Sb * NP[-1] / (N[-1] + P[-1])
    Jn[0] = Sft; Jp[0] = -Sft; Jn[L] = -Sbt; Jp[L] = Sbt
    den = p.tauN * P + p.tauP * N
    rate = p.Cn * N + p.Cp * P + p.ks + 1.0 / den
    R = rate * NP
    fN = (Jn[1:] - Jn[:-1]) / dx - R
    fQ = -(Jn[1:] + Jp[1:]) / dx        # dQ_{i+1}/dt, i=0..L-1 (last is exactly 0)
    f = np.empty(2 * L); f[0::2] = fN; f[1::2] = fQ
    if not want_jac:
        return f
    # ---- analytic Jacobian in banded storage (l=u=3) -----
    # partials wrt raw (N,P,E-face) then chain rule P_i = N_i + c + Q_{i+1} - Q_i
    # face i (1..L-1): dJn/dNm, dJn/dNp, dJn/dQ_i ; dJp/dPm, dJp/dPp, dJp/dQ_i
    dJn_dNm = p.mun * (0.5 * E - kT / dx)
    dJn_dNp = p.mun * (0.5 * E + kT / dx)
    dJn_dQ = p.mun * 0.5 * (Nm + Np_) * LD
    dJp_dPm = p.mup * (0.5 * E + kT / dx)
    dJp_dPp = p.mup * (0.5 * E - kT / dx)
    dJp_dQ = p.mup * 0.5 * (Pm + Pp) * LD
    # recombination partials
    dR_dN = (p.Cn - p.tauP / den**2) * NP + rate * P
    dR_dP = (p.Cp - p.tauN / den**2) * NP + rate * N
    J = np.zeros((2 * L, 2 * L))
    def add(r, c, v):
        J[r, c] += v
    for i in range(L):
        rN = 2 * i; rQ = 2 * i + 1
        cN = 2 * i; cQr = 2 * i + 1; cQl = 2 * i - 1   # Q_{i+1}, Q_i (col of block i-1)
        # R_i depends on N_i and P_i
        # dP_i: dN_i (1), dQ_{i+1} (+1), dQ_i (-1)
        add(rN, cN, -(dR_dN[i] + dR_dP[i])); add(rN, cQr, -dR_dP[i])
        if i > 0: add(rN, cQl, +dR_dP[i])
        # flux right face (i+1) into fN_i: +Jn_{i+1}/dx ; fQ_{i+1} = -(Jn_{i+1}+Jp_{i+1})/dx
        if i < L - 1:
            k = i   # index into interior arrays for face i+1 (faces 1..L-1 -> 0..L-2)
            # Jn_{i+1}: Nm=N_i, Np=N_{i+1}, Q_{i+1}
            add(rN, cN, dJn_dNm[k] / dx); add(rN, 2 * (i + 1), dJn_dNp[k] / dx); add(rN, cQr, dJn_dQ[k] / dx)
            add(rQ, cN, -dJn_dNm[k] / dx); add(rQ, 2 * (i + 1), -dJn_dNp[k] / dx); add(rQ, cQr, -dJn_dQ[k] / dx)
            # Jp_{i+1}: Pm=P_i (N_i, Q_{i+1}, -Q_i), Pp=P_{i+1} (N_{i+1}, Q_{i+2}, -Q_{i+1}), Q_{i+1}
            add(rQ, cN, -dJp_dPm[k] / dx); add(rQ, cQr, -dJp_dPm[k] / dx)
            if i > 0: add(rQ, cQl, +dJp_dPm[k] / dx)
            add(rQ, 2 * (i + 1), -dJp_dPp[k] / dx); add(rQ, 2 * (i + 1) + 1, -dJp_dPp[k] / dx); add(rQ, cQr, +dJp_dPp[k] / dx)
            add(rQ, cQr, -dJp_dQ[k] / dx)
        else:
            # right boundary: Jn_L = -Sbt ; fQ_L = 0
            den_b = N[i] + P[i]
            dS_dN = p.Sb * (P[i] / den_b - NP[i] / den_b**2)
            dS_dP = p.Sb * (N[i] / den_b - NP[i] / den_b**2)
            add(rN, cN, -(dS_dN + dS_dP) / dx); add(rN, cQl, +dS_dP / dx)   # Q_L const(=0) -> no cQr term
        # flux left face (i) : -Jn_i/dx
        if i > 0:
            k = i - 1
            add(rN, 2 * (i - 1), -dJn_dNm[k] / dx); add(rN, cN, -dJn_dNp[k] / dx); add(rN, cQl, -dJn_dQ[k] / dx)
        else:
            den_b = N[0] + P[0]
            dS_dN = p.Sf * (P[0] / den_b - NP[0] / den_b**2)
            dS_dP = p.Sf * (N[0] / den_b - NP[0] / den_b**2)
            add(rN, cN, -(dS_dN + dS_dP) / dx); add(rN, cQr, -dS_dP / dx)
    # last Q (Q_L) is a dummy: in R_{L-1} and Sbt the P_{L-1} dependence on Q_L kept consistent (Q_L==0 const)
    J[:, 2 * L - 1] = 0.0; J[2 * L - 1, :] = 0.0
    return f, J

def PL_of(p, y):
    N, P, _, _ = unpack(p, y)
    return p.ks * p.dx * np.sum(N * P - p.n0 * p.p0) * 1e23

def dPL_of(p, y, f):
    N, P, _, _ = unpack(p, y)
    fN = f[0::2]; fQr = f[1::2]; fQl = np.concatenate(([0.0], fQr[:-1]))
    fP = fN + fQr - fQl
    return p.ks * p.dx * np.sum(fN * P + N * fP) * 1e23

def to_banded(M, l=3, u=3):
    n = M.shape[0]
    ab = np.zeros((l + u + 1, n))
    for d in range(-l, u + 1):
        diag = np.diagonal(M, d)
        if d >= 0: ab[u - d, d:] = diag
        else: ab[u - d, :n + d] = diag
    return ab

def scale_vec(p, y, rtol, atol):
    N, P, _, _ = unpack(p, y)
    sc = np.empty_like(y)
    sc[0::2] = atol + rtol * np.abs(N)
    sc[1::2] = atol + rtol * np.maximum(np.abs(N), np.abs(P))
    return sc

def integrate(p, y0, tout, rtol=1e-7, atol=1e-16, h0=None, hmax=np.inf, stats=None, land=False):
    """Adaptive RODAS4; returns PL at tout via 3-point quintic Hermite on ln(PL) (or exact landing if land)."""
    A, C, g, m, mhat = rodas4()
    t = 0.0; y = y0.copy(); tend = tout[-1]
    n = len(y)
    nsteps = nrej = 0
    f0 = rhs(p, y)
    pl = PL_of(p, y); dpl = dPL_of(p, y, f0)
    hist = [(t, pl, dpl)]
    out = np.zeros(len(tout)); out[0] = pl; io = 1
    if h0 is None:
        sc = scale_vec(p, y, rtol, atol)
        d0 = np.sqrt(np.mean((y / sc) ** 2)); d1 = np.sqrt(np.mean((f0 / sc) ** 2))
        h = 0.01 * d0 / d1 if d1 > 0 else 1e-6
        h = min(h, 1e-3)
    else:
        h = h0
    errold = 1e-4; hacc = h; first = True
    while io < len(tout):
        h = min(h, hmax, tend - t)
        if land and t + h > tout[io] - 1e-12 * tend:
            h = tout[io] - t
        elif land and t + 1.8*h > tout[io]:
            h = (tout[io] - t)/2
        f0, J = rhs(p, y, want_jac=True)
        M = np.eye(n) / (g * h) - J
        ab = to_banded(M)
        U = np.zeros((6, n))
        for i in range(6):
            if i == 0:
                fi = f0
            else:
                yi = y + A[i, :i] @ U[:i]
                fi = rhs(p, yi)
            r = fi + (C[i, :i] / h) @ U[:i]
            U[i] = solve_banded((3, 3), ab, r)
        ynew = y + m @ U
        sc = scale_vec(p, np.where(np.abs(ynew) > np.abs(y), ynew, y), rtol, atol)
        err = np.sqrt(np.mean((U[5] / sc) ** 2))
        fac = max(0.2, min(6.0, err ** 0.25 / 0.9))
        hnew = h / fac
        if err <= 1.0 and np.all(np.isfinite(ynew)):
            nsteps += 1
            # Gustafsson
            if not first:
                facgus = (hacc / h) * (err ** 2 / errold) ** 0.25 / 0.9
                facgus = max(1/6.0, min(5.0, facgus))
                fac = max(fac, facgus); hnew = h / fac
            first = False
            hacc = h; errold = max(1e-2, err)
            t += h; y = ynew
            fnew = rhs(p, y)
            pl = PL_of(p, y); dpl = dPL_of(p, y, fnew)
            hist.append((t, pl, dpl))
            while io < len(tout) and tout[io] <= t * (1 + 1e-14):
                out[io] = hermite_eval(hist, tout[io]); io += 1
            h = hnew
        else:
            nrej += 1
            h = hnew if np.isfinite(err) else h * 0.1
            first = True  # no gustafsson after reject growth
    if stats is not None:
        stats["nsteps"] = nsteps; stats["nrej"] = nrej
    return out

def hermite_eval(hist, tq):
    """Quintic Hermite through last 3 points (value+derivative) of ln PL; cubic if only 2."""
    pts = hist[-3:] if len(hist) >= 3 else hist[-2:]
    if abs(tq - pts[-1][0]) <= 1e-14 * max(1.0, abs(tq)):
        return pts[-1][1]
    use_log = all(pp[1] > 0 for pp in pts)
    ts = np.array([pp[0] for pp in pts])
    if use_log:
        v = np.array([np.log(pp[1]) for pp in pts]); d = np.array([pp[2] / pp[1] for pp in pts])
    else:
        v = np.array([pp[1] for pp in pts]); d = np.array([pp[2] for pp in pts])
    # Newton divided differences with doubled nodes
    z = np.repeat(ts, 2); k = len(z)
    Qd = np.zeros((k, k)); Qd[:, 0] = np.repeat(v, 2)
    for i in range(1, k):
        for j in range(1, i + 1):
            if j == 1 and i % 2 == 1:
                Qd[i, j] = d[i // 2]
            else:
                Qd[i, j] = (Qd[i, j - 1] - Qd[i - 1, j - 1]) / (z[i] - z[i - j])
    r = Qd[k - 1, k - 1]
    for i in range(k - 2, -1, -1):
        r = r * (tq - z[i]) + Qd[i, i]
    return np.exp(r) if use_log else r
