"""Excess-variable restatement of the reference's semi-discrete model, with its exact Jacobian.

TEST INFRASTRUCTURE ONLY (see oracle/trpl_oracle.py for the rules).  Same equations as
forward_solver.py:332-418 of the reference, written for y = [dN, dP, E (, Ntrap)] with
dN = N - n0, dP = P - p0, so that N P - n0 p0 = n0 dP + p0 dN + dN dP has no cancellation however
far the signal has decayed.  Pinned: tests/test_oracle_golden.py checks rhs_excess against the
bit-exact rhs_std / rhs_traps and jac_excess against central differences.

Used for the *linear-regime* check of deep decays: at low injection the model is linear, every
curve ends as exp(-lambda_0 t), and lambda_0 is the smallest eigenvalue of -J at equilibrium
(`slowest_decay_rate`), which needs no time integrator at all.
"""
import numpy as np
from numba import njit
KB = 8.61773e-5

@njit(cache=False)
def rhs_excess(t, y, L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm):
    out = np.zeros(3 * L + 1)
    dN = y[:L]; dP = y[L:2*L]; E = y[2*L:]
    kT = KB * Tm
    jn = np.zeros(L + 1); jp = np.zeros(L + 1)
    ex0 = n0*dP[0] + p0*dN[0] + dN[0]*dP[0]
    exL = n0*dP[L-1] + p0*dN[L-1] + dN[L-1]*dP[L-1]
    sf = Sf * ex0 / ((n0 + dN[0]) + (p0 + dP[0]))
    sb = Sb * exL / ((n0 + dN[L-1]) + (p0 + dP[L-1]))
    jn[0] = sf; jp[0] = -sf; jn[L] = -sb; jp[L] = sb
    for k in range(1, L):
        Nm = n0 + 0.5*(dN[k-1] + dN[k]); Pm = p0 + 0.5*(dP[k-1] + dP[k])
        jn[k] = mu_n * (Nm * E[k]) + mu_n * kT * ((dN[k] - dN[k-1]) / dx)
        jp[k] = mu_p * (Pm * E[k]) - mu_p * kT * ((dP[k] - dP[k-1]) / dx)
    for k in range(L + 1):
        out[2*L + k] = -(jn[k] + jp[k]) * Lam
    for k in range(L):
        N = n0 + dN[k]; P = p0 + dP[k]
        ex = n0*dP[k] + p0*dN[k] + dN[k]*dP[k]
        loss = ((Cn*N + Cp*P) + ks + 1/((tauN*P) + (tauP*N))) * ex
        out[k] = (jn[k+1] - jn[k]) / dx - loss
        out[L+k] = -(jp[k+1] - jp[k]) / dx - loss
    return out

@njit(cache=False)
def jac_excess(t, y, L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm):
    n = 3*L + 1
    J = np.zeros((n, n))
    dN = y[:L]; dP = y[L:2*L]; E = y[2*L:]
    kT = KB * Tm
    # d jn[k] / d(dN[k-1]), d(dN[k]), dE[k];  same for jp wrt dP
    jn_a = np.zeros(L+1); jn_b = np.zeros(L+1); jn_e = np.zeros(L+1)
    jp_a = np.zeros(L+1); jp_b = np.zeros(L+1); jp_e = np.zeros(L+1)
    for k in range(1, L):
        Nm = n0 + 0.5*(dN[k-1] + dN[k]); Pm = p0 + 0.5*(dP[k-1] + dP[k])
        jn_a[k] = mu_n*0.5*E[k] - mu_n*kT/dx; jn_b[k] = mu_n*0.5*E[k] + mu_n*kT/dx; jn_e[k] = mu_n*Nm
        jp_a[k] = mu_p*0.5*E[k] + mu_p*kT/dx; jp_b[k] = mu_p*0.5*E[k] - mu_p*kT/dx; jp_e[k] = mu_p*Pm
    # contacts: s = S*ex/(N+P)
    N0 = n0 + dN[0]; P0 = p0 + dP[0]; ex0 = n0*dP[0] + p0*dN[0] + dN[0]*dP[0]
    sfn = Sf*(P0/(N0+P0) - ex0/(N0+P0)**2); sfp = Sf*(N0/(N0+P0) - ex0/(N0+P0)**2)
    NL = n0 + dN[L-1]; PL = p0 + dP[L-1]; exL = n0*dP[L-1] + p0*dN[L-1] + dN[L-1]*dP[L-1]
    sbn = Sb*(PL/(NL+PL) - exL/(NL+PL)**2); sbp = Sb*(NL/(NL+PL) - exL/(NL+PL)**2)
    # E rows
    for k in range(1, L):
        r = 2*L + k
        J[r, k-1] = -jn_a[k]*Lam; J[r, k] = -jn_b[k]*Lam
        J[r, L+k-1] = -jp_a[k]*Lam; J[r, L+k] = -jp_b[k]*Lam
        J[r, 2*L+k] = -(jn_e[k] + jp_e[k])*Lam
    # E[0], E[L]: jn+jp == 0 identically -> zero rows
    for k in range(L):
        N = n0 + dN[k]; P = p0 + dP[k]
        ex = n0*dP[k] + p0*dN[k] + dN[k]*dP[k]
        den = tauN*P + tauP*N
        rate = (Cn*N + Cp*P) + ks + 1/den
        ln = (Cn - tauP/den**2)*ex + rate*P
        lp = (Cp - tauN/den**2)*ex + rate*N
        # N row: (jn[k+1]-jn[k])/dx - loss
        J[k, k] -= ln; J[k, L+k] -= lp
        J[L+k, k] -= ln; J[L+k, L+k] -= lp
        # right face k+1
        if k+1 < L:
            J[k, k] += jn_a[k+1]/dx; J[k, k+1] += jn_b[k+1]/dx; J[k, 2*L+k+1] += jn_e[k+1]/dx
            J[L+k, L+k] -= jp_a[k+1]/dx; J[L+k, L+k+1] -= jp_b[k+1]/dx; J[L+k, 2*L+k+1] -= jp_e[k+1]/dx
        else:
            # jn[L] = -sb, jp[L] = +sb
            J[k, k] += -sbn/dx; J[k, L+k] += -sbp/dx
            J[L+k, k] -= sbn/dx; J[L+k, L+k] -= sbp/dx
        if k >= 1:
            J[k, k-1] -= jn_a[k]/dx; J[k, k] -= jn_b[k]/dx; J[k, 2*L+k] -= jn_e[k]/dx
            J[L+k, L+k-1] += jp_a[k]/dx; J[L+k, L+k] += jp_b[k]/dx; J[L+k, 2*L+k] += jp_e[k]/dx
        else:
            # jn[0] = sf, jp[0] = -sf
            J[k, k] -= sfn/dx; J[k, L+k] -= sfp/dx
            J[L+k, k] += -sfn/dx; J[L+k, L+k] += -sfp/dx
    return J


@njit(cache=False)
def rhs_excess_traps(t, y, L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm, kC, Nt, tauE):
    """y = [dN, dP, E, Ntrap]; forward_solver.py:374-418."""
    out = np.zeros(4 * L + 1)
    out[:3 * L + 1] = rhs_excess(t, y[:3 * L + 1], L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm)
    for k in range(L):
        N = n0 + y[k]
        tr = y[3 * L + 1 + k]
        capture = kC * N * (Nt - tr)
        release = tr / tauE
        out[k] += release - capture
        out[3 * L + 1 + k] = capture - release
    return out


@njit(cache=False)
def jac_excess_traps(t, y, L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm, kC, Nt, tauE):
    n = 4 * L + 1
    J = np.zeros((n, n))
    J[:3 * L + 1, :3 * L + 1] = jac_excess(t, y[:3 * L + 1], L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm)
    for k in range(L):
        N = n0 + y[k]
        tr = y[3 * L + 1 + k]
        cap_n = kC * (Nt - tr)
        cap_t = -kC * N
        r = 3 * L + 1 + k
        J[k, k] -= cap_n
        J[k, r] += 1 / tauE - cap_t
        J[r, k] = cap_n
        J[r, r] = cap_t - 1 / tauE
    return J


def decay_rates(L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm):
    """Decay rates [1/ns], ascending, of the model linearised about equilibrium (dN = dP = E = 0).

    The Jacobian there has exact zero eigenvalues for the two contact fields (they never change)
    and for the conserved total charge; the returned rates are the strictly positive ones.  At low
    injection every simulated curve ends as a sum of exp(-rate t) terms with these rates.
    """
    n = 3 * L + 1
    J = jac_excess(0.0, np.zeros(n), L, dx, n0, p0, mu_n, mu_p, ks, Cn, Cp, Sf, Sb, tauN, tauP, Lam, Tm)
    keep = np.ones(n, dtype=np.bool_)
    keep[2 * L] = False
    keep[3 * L] = False
    ev = np.linalg.eigvals(J[np.ix_(keep, keep)])
    rates = -ev.real
    scale = np.abs(rates).max()
    return np.sort(rates[rates > 1e-13 * scale])
