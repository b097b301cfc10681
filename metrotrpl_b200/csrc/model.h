// model.h - drift-diffusion-recombination right-hand side and its exact block-tridiagonal Jacobian.
//
// Physics: the reference's semi-discrete carrier model, forward_solver.py:332-372 (dydt_numba, 'std')
// and :374-418 (dydt_numba_traps).  The reference integrates [N, P, E] with dE/dt = -Lambda (Jn+Jp).
// Because Jn+Jp vanishes on both contacts, E stays equal to Gauss's law (E_field(),
// forward_solver.py:26-38) for all time, so the same trajectory is described by
//
//     u_i = ( N_i , Q_{i+1} ),   Q_k = sum_{j<k} (P_j - N_j [- Ntrap_j] - (p0 - n0))   (running net charge)
//     E_k = Lambda * dx * Q_k ,  P_i = N_i [+ Ntrap_i] + (p0 - n0) + Q_{i+1} - Q_i
//
// In these variables every equation touches only nodes i-1, i, i+1: the Jacobian is block
// tridiagonal with 2x2 blocks, which is what the in-warp solver (blocktri.h) factorises.
//
// Node ownership: lane l owns the NPL consecutive nodes i = l*NPL + j, j = 0..NPL-1.  Nodes with
// i >= L are padding (frozen, decoupled).
#pragma once
#include "simt.h"

namespace trpl {
using namespace simt;

// Parameter vector layout (model units: nm, ns, V), one per parameter set.  The host mirror fills
// it from `state * units` exactly as forward_solver.py:119-138 does.
enum ParamSlot {
  P_N0 = 0, P_P0, P_MUN, P_MUP, P_KS, P_CN, P_CP, P_SF, P_SB, P_TAUN, P_TAUP, P_EPS, P_TM,
  P_KC, P_NT, P_TAUE, NPARAM = 16
};

constexpr double KB_EV = 8.61773e-5;                 // forward_solver.py:24
constexpr double EPS0_NM = 8.854 * 1e-12 * 1e-9;     // forward_solver.py:21
constexpr double Q_COULOMB = 1.602e-19;              // forward_solver.py:23

enum Model { MODEL_STD = 0, MODEL_TRAPS = 1 };
enum MeasType { MEAS_TRPL = 0, MEAS_TRTS = 1 };

// warp-uniform coefficients derived once per trajectory
struct Coef {
  double n0, p0, d0;          // d0 = p0 - n0
  double an, ap;              // mu/2
  double dn, dp;              // mu * kB*T / dx
  double ld;                  // Lambda * dx  (E_k = ld * Q_k)
  double ix;                  // 1/dx
  double ks, cn, cp, taun, taup, sf, sb, n0p0;
  double kc, nt, itaue;       // traps
  double mun, mup;            // for the TRTS readout
  double dx;
  int L;                      // real node count
  // flux coefficients pre-multiplied by 1/dx (and Lambda dx for the drift terms): the right-hand
  // side then needs no per-node scaling of the current differences
  double anl, apl;            // (mu/2) * Lambda dx / dx
  double dnx, dpx;            // mu kT / dx^2
  double sfx, sbx;            // S / dx
};

TRPL_FN Coef make_coef(const double* p, double thickness, int L) {
  Coef c;
  c.L = L;
  c.dx = thickness / L;                                   // sim_utils.py:268
  c.ix = 1.0 / c.dx;
  c.n0 = p[P_N0]; c.p0 = p[P_P0]; c.d0 = c.p0 - c.n0; c.n0p0 = c.n0 * c.p0;
  const double kT = KB_EV * p[P_TM];
  c.mun = p[P_MUN]; c.mup = p[P_MUP];
  c.an = 0.5 * p[P_MUN]; c.ap = 0.5 * p[P_MUP];
  c.dn = p[P_MUN] * kT * c.ix; c.dp = p[P_MUP] * kT * c.ix;
  c.ld = (Q_COULOMB / (p[P_EPS] * EPS0_NM)) * c.dx;       // forward_solver.py:131
  c.ks = p[P_KS]; c.cn = p[P_CN]; c.cp = p[P_CP];
  c.taun = p[P_TAUN]; c.taup = p[P_TAUP]; c.sf = p[P_SF]; c.sb = p[P_SB];
  c.kc = p[P_KC]; c.nt = p[P_NT]; c.itaue = 1.0 / p[P_TAUE];
  c.anl = c.an * c.ld * c.ix; c.apl = c.ap * c.ld * c.ix;
  c.dnx = c.dn * c.ix; c.dpx = c.dp * c.ix;
  c.sfx = c.sf * c.ix; c.sbx = c.sb * c.ix;
  return c;
}

// The coefficients are warp-uniform.  Kept in registers they would pin ~46 registers for the whole
// trajectory; parked in one shared-memory slot they cost a handful of broadcast loads per use.
TRPL_FN void park_coef(LaneMem& sm, int slot, const Coef& c) {
  const double v[30] = {c.n0, c.p0, c.d0, c.an, c.ap, c.dn, c.dp, c.ld, c.ix, c.ks, c.cn, c.cp, c.taun,
                        c.taup, c.sf, c.sb, c.n0p0, c.kc, c.nt, c.itaue, c.mun, c.mup, c.dx, (double)c.L,
                        c.anl, c.apl, c.dnx, c.dpx, c.sfx, c.sbx};
  TRPL_UNROLL for (int i = 0; i < 30; ++i) sm.ust(slot, i, v[i]);
  warp_sync();
}
TRPL_FN Coef fetch_coef(const LaneMem& sm, int slot) {
  Coef c;
  c.n0 = sm.uld(slot, 0); c.p0 = sm.uld(slot, 1); c.d0 = sm.uld(slot, 2); c.an = sm.uld(slot, 3);
  c.ap = sm.uld(slot, 4); c.dn = sm.uld(slot, 5); c.dp = sm.uld(slot, 6); c.ld = sm.uld(slot, 7);
  c.ix = sm.uld(slot, 8); c.ks = sm.uld(slot, 9); c.cn = sm.uld(slot, 10); c.cp = sm.uld(slot, 11);
  c.taun = sm.uld(slot, 12); c.taup = sm.uld(slot, 13); c.sf = sm.uld(slot, 14); c.sb = sm.uld(slot, 15);
  c.n0p0 = sm.uld(slot, 16); c.kc = sm.uld(slot, 17); c.nt = sm.uld(slot, 18); c.itaue = sm.uld(slot, 19);
  c.mun = sm.uld(slot, 20); c.mup = sm.uld(slot, 21); c.dx = sm.uld(slot, 22); c.L = (int)sm.uld(slot, 23);
  c.anl = sm.uld(slot, 24); c.apl = sm.uld(slot, 25); c.dnx = sm.uld(slot, 26); c.dpx = sm.uld(slot, 27);
  c.sfx = sm.uld(slot, 28); c.sbx = sm.uld(slot, 29);
  return c;
}

// Per-lane node classification (constant for a trajectory).
template <int NPL>
struct NodeMask {
  mask real_node[NPL];   // i <  L
  mask last_node[NPL];   // i == L-1
  mask inner_face[NPL];  // right face of node i is an interior face (i < L-1)
  mask first_lane;       // lane 0 (owns node 0)
};

// FULL: L == LANES*NPL is known at compile time (no padding), so every mask but the two contact
// lanes folds to a constant and the selects disappear from the unrolled node loops.
template <int NPL, bool FULL>
TRPL_FN NodeMask<NPL> make_mask(int L) {
  NodeMask<NPL> m;
  const ivec base = imul(lane_id(), NPL);
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    if (FULL) {
      m.real_node[j] = mconst(true);
      if (j == NPL - 1) {
        m.last_node[j] = lane_id() == LANES - 1;
        m.inner_face[j] = lane_id() < LANES - 1;
      } else {
        m.last_node[j] = mconst(false);
        m.inner_face[j] = mconst(true);
      }
    } else {
      const ivec i = iadd(base, j);
      m.real_node[j] = i < L;
      m.last_node[j] = i == (L - 1);
      m.inner_face[j] = i < (L - 1);
    }
  }
  m.first_lane = lane_id() == 0;
  return m;
}

// One per-lane state slice.  NT is only used by the traps model.
template <int NPL, int MODEL>
struct Vec {
  real n[NPL];
  real q[NPL];
  real t[MODEL == MODEL_TRAPS ? NPL : 1];
};

// Everything the RHS computes that later stages want to reuse.
template <int NPL>
struct RhsAux {
  real p[NPL];     // hole density
  // node-local recombination terms and neighbour values, reused by the Jacobian when it is
  // evaluated at the same state (stage 1 of every step)
  real npx[NPL];   // N P - n0 p0
  real inv[NPL];   // 1 / (tauN P + tauP N)
  real rate[NPL];  // Cn N + Cp P + ks + inv
  real ql0;        // running charge left of this lane's first node
  real n_next, p_next;   // first node of the next lane
  real bn, bp, bx, isum; // the contact node this lane owns: N, P, N P - n0 p0, 1 / (N + P)
  real snl[NPL], spl[NPL];   // (N_i + N_{i+1}) anl and (P_i + P_{i+1}) apl of the right face: d(flux)/dQ
};

// hole density from the state: P_i = N_i [+ Ntrap_i] + (p0 - n0) + Q_{i+1} - Q_i
template <int NPL, int MODEL>
TRPL_FN void holes(const Coef& c, const NodeMask<NPL>& m, const Vec<NPL, MODEL>& u, real (&P)[NPL]) {
  real ql0 = shfl_up(u.q[NPL - 1], 1);
  ql0 = sel(m.first_lane, 0.0, ql0);
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real ql = (j == 0) ? ql0 : u.q[j - 1];
    P[j] = u.n[j] + c.d0 + (u.q[j] - ql);
    if (MODEL == MODEL_TRAPS) P[j] = P[j] + u.t[j];
  }
}

// Right-hand side f(u).  Also returns P for the readout / error scale.
template <int NPL, int MODEL>
TRPL_FN void rhs(const Coef& c, const NodeMask<NPL>& m, const Vec<NPL, MODEL>& u,
                 Vec<NPL, MODEL>& f, RhsAux<NPL>& aux) {
  // left running charge of my first node comes from my left neighbour
  real ql0 = shfl_up(u.q[NPL - 1], 1);
  ql0 = sel(m.first_lane, 0.0, ql0);
  real P[NPL];
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real ql = (j == 0) ? ql0 : u.q[j - 1];
    P[j] = u.n[j] + c.d0 + (u.q[j] - ql);
    if (MODEL == MODEL_TRAPS) P[j] = P[j] + u.t[j];
    aux.p[j] = P[j];
  }
  const real n_next = shfl_down(u.n[0], 1);
  const real p_next = shfl_down(P[0], 1);
  aux.ql0 = ql0; aux.n_next = n_next; aux.p_next = p_next;

  // node-local recombination, and the surface term of whichever contact this lane owns
  real np_ex[NPL], loss[NPL];
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    np_ex[j] = fmadd(u.n[j], P[j], -c.n0p0);
    const real inv = rcp(fmadd(c.taun, P[j], c.taup * u.n[j]));
    const real rate = fmadd(c.cn, u.n[j], fmadd(c.cp, P[j], c.ks)) + inv;
    loss[j] = rate * np_ex[j];
    aux.npx[j] = np_ex[j]; aux.inv[j] = inv; aux.rate[j] = rate;
  }
  // one division serves both contacts: lane 0 evaluates the front (node 0), the lane holding node
  // L-1 the back.  NPL is chosen minimal by the host so these are different lanes.
  real bn = u.n[0], bp = P[0], bx = np_ex[0];
  mask has_last = mconst(false);
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    bn = sel(m.last_node[j], u.n[j], bn);
    bp = sel(m.last_node[j], P[j], bp);
    bx = sel(m.last_node[j], np_ex[j], bx);
    has_last = mor(has_last, m.last_node[j]);
  }
  // currents carry the 1/dx of the divergence (coefficients pre-multiplied in make_coef)
  const real svel = sel(has_last, c.sbx, c.sfx);
  const real isum = rcp(bn + bp);
  const real surf = svel * bx * isum;               // forward_solver.py:346-347, times 1/dx
  aux.bn = bn; aux.bp = bp; aux.bx = bx; aux.isum = isum;
  real jn[NPL], jp[NPL];
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real nn = (j == NPL - 1) ? n_next : u.n[j + 1];
    const real pn = (j == NPL - 1) ? p_next : P[j + 1];
    const real snl = (u.n[j] + nn) * c.anl, spl = (P[j] + pn) * c.apl;
    aux.snl[j] = snl; aux.spl[j] = spl;
    real a = fmadd(snl, u.q[j], c.dnx * (nn - u.n[j]));                        // forward_solver.py:356-357
    real b = fmadd(spl, u.q[j], -(c.dpx * (pn - P[j])));                       // forward_solver.py:358-359
    // back contact: Jn = -S, Jp = +S (sum exactly zero); padding: no current
    a = sel(m.inner_face[j], a, sel(m.last_node[j], -surf, 0.0));
    jn[j] = a;
    jp[j] = sel(m.inner_face[j], b, -a);
  }
  if (MODEL != MODEL_TRAPS) f.t[0] = splat(0.0);
  real jl0 = shfl_up(jn[NPL - 1], 1);
  jl0 = sel(m.first_lane, surf, jl0);                                   // forward_solver.py:349
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real jl = (j == 0) ? jl0 : jn[j - 1];
    real fn = (jn[j] - jl) - loss[j];                                   // forward_solver.py:369
    real fq = -jn[j] - jp[j];                                           // forward_solver.py:363 (/ (Lambda dx))
    if (MODEL == MODEL_TRAPS) {
      const real capture = c.kc * u.n[j] * (c.nt - u.t[j]);             // forward_solver.py:410
      const real release = u.t[j] * c.itaue;                            // forward_solver.py:411
      fn = fn + (release - capture);
      f.t[j] = sel(m.real_node[j], capture - release, 0.0);
    }
    f.n[j] = sel(m.real_node[j], fn, 0.0);
    f.q[j] = fq;
  }
}

// 2x2 block, row-major {a00, a01, a10, a11}
struct Blk { real a00, a01, a10, a11; };

// Jacobian blocks of f with respect to u (traps: the Ntrap unknown is condensed out by the caller,
// see trajectory.h; here we return the extra couplings it needs).
template <int NPL>
struct JacTraps {
  real fn_t[NPL];   // d fN / d Ntrap   (total, including through P)
  real fq_t[NPL];   // d fQ_{i+1} / d Ntrap_i
  real fq_tn[NPL];  // d fQ_{i+1} / d Ntrap_{i+1}
  real ft_n[NPL];   // d fT / d N
  real ft_t[NPL];   // d fT / d Ntrap
};

// `aux` is what rhs() left behind for the SAME state u (the Jacobian is only ever evaluated right
// after f(u), at stage 1 of a step): hole densities, recombination terms, contact terms, neighbour
// values and the d(flux)/dQ factors are not recomputed.  Every flux partial carries the 1/dx of the
// divergence (coefficients pre-multiplied in make_coef), like the fluxes in rhs().
template <int NPL, int MODEL>
TRPL_FN void jacobian(const Coef& c, const NodeMask<NPL>& m, const Vec<NPL, MODEL>& u, const RhsAux<NPL>& aux,
                      Blk (&A)[NPL], Blk (&B)[NPL], Blk (&C)[NPL], JacTraps<NPL>& jt) {
  const real n_prev = shfl_up(u.n[NPL - 1], 1);

  // contact term partials (one lane each, same trick as in rhs)
  mask has_last = mconst(false);
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) has_last = mor(has_last, m.last_node[j]);
  const real svel = sel(has_last, c.sbx, c.sfx);
  const real common = aux.bx * aux.isum * aux.isum;
  const real s_n = svel * (aux.bp * aux.isum - common);   // dS/dN at fixed P
  const real s_p = svel * (aux.bn * aux.isum - common);   // dS/dP at fixed N

  // interior form of the right-face partials of every node (the left face of node j is the right
  // face of node j-1; the face left of this lane's first node is formed from the shuffled values)
  real r_ni[NPL], r_nn[NPL], p_pi[NPL], p_pn[NPL];
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    r_ni[j] = fmadd(c.anl, u.q[j], -c.dnx);         // d jnR / d N_i
    r_nn[j] = fmadd(c.anl, u.q[j], c.dnx);          // d jnR / d N_{i+1}
    p_pi[j] = fmadd(c.apl, u.q[j], c.dpx);          // d jpR / d P_i
    p_pn[j] = fmadd(c.apl, u.q[j], -c.dpx);         // d jpR / d P_{i+1}
  }
  const real l0_nm = fmadd(c.anl, aux.ql0, -c.dnx), l0_ni = fmadd(c.anl, aux.ql0, c.dnx);
  const real l0_q = (n_prev + u.n[0]) * c.anl;

  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real nj = u.n[j], pj = aux.p[j];
    // recombination partials
    const real inv2 = aux.inv[j] * aux.inv[j];
    const real r_n = fmadd(c.cn - c.taup * inv2, aux.npx[j], aux.rate[j] * pj);
    const real r_p = fmadd(c.cp - c.taun * inv2, aux.npx[j], aux.rate[j] * nj);
    // right face (between i and i+1); the back contact replaces it at node L-1: jnR = -S(N_i, P_i)
    const real jr_ni = sel(m.inner_face[j], r_ni[j], sel(m.last_node[j], -s_n, 0.0));
    const real jr_pi = sel(m.last_node[j], -s_p, 0.0);      // d jnR / d P_i (only through the back contact)
    const real jr_nn = sel(m.inner_face[j], r_nn[j], 0.0);
    const real jr_q = sel(m.inner_face[j], aux.snl[j], 0.0); // d jnR / d Q_{i+1}
    const real jp_pi = p_pi[j], jp_pn = p_pn[j];
    const real jp_q = aux.spl[j];                           // d jpR / d Q_{i+1} (explicit field)
    // left face (between i-1 and i); front contact: jnL = +S(N_0, P_0)
    real jl_nm = (j == 0) ? l0_nm : r_ni[j > 0 ? j - 1 : 0];    // d jnL / d N_{i-1}
    real jl_ni = (j == 0) ? l0_ni : r_nn[j > 0 ? j - 1 : 0];    // d jnL / d N_i
    real jl_q = (j == 0) ? l0_q : aux.snl[j > 0 ? j - 1 : 0];   // d jnL / d Q_i
    real jl_pi = splat(0.0);
    if (j == 0) {
      jl_nm = sel(m.first_lane, 0.0, jl_nm);
      jl_q = sel(m.first_lane, 0.0, jl_q);
      jl_ni = sel(m.first_lane, s_n, jl_ni);
      jl_pi = sel(m.first_lane, s_p, jl_pi);
    }
    // chain coefficient of P_i in fN_i
    const real c_p = (jr_pi - jl_pi) - r_p;
    real b00 = ((jr_ni - jl_ni) - r_n) + c_p;
    real b01 = jr_q + c_p;
    real a00 = -jl_nm;
    real a01 = -(jl_q + c_p);
    real c00 = jr_nn;
    // fQ_{i+1} = -(jnR + jpR) on interior faces, identically zero otherwise
    real b10 = -(jr_ni + jp_pi);
    real b11 = -((jr_q + jp_q) + (jp_pi - jp_pn));
    real a11 = jp_pi;
    real c10 = -(jr_nn + jp_pn);
    real c11 = -jp_pn;
    b10 = sel(m.inner_face[j], b10, 0.0);
    b11 = sel(m.inner_face[j], b11, 0.0);
    a11 = sel(m.inner_face[j], a11, 0.0);
    c10 = sel(m.inner_face[j], c10, 0.0);
    c11 = sel(m.inner_face[j], c11, 0.0);
    if (MODEL == MODEL_TRAPS) {
      const real cap_n = c.kc * (c.nt - u.t[j]);    // d capture / d N
      const real cap_t = -(c.kc * nj);              // d capture / d Ntrap
      b00 = b00 - cap_n;
      // Ntrap enters fN directly (release - capture) and through P_i (+1)
      jt.fn_t[j] = sel(m.real_node[j], (c.itaue - cap_t) + c_p, 0.0);
      jt.fq_t[j] = sel(m.inner_face[j], -jp_pi, 0.0);
      jt.fq_tn[j] = sel(m.inner_face[j], -jp_pn, 0.0);
      jt.ft_n[j] = sel(m.real_node[j], cap_n, 0.0);
      jt.ft_t[j] = sel(m.real_node[j], cap_t - c.itaue, 0.0);
    }
    // padding nodes are frozen
    A[j].a00 = sel(m.real_node[j], a00, 0.0); A[j].a01 = sel(m.real_node[j], a01, 0.0);
    A[j].a10 = splat(0.0);                     A[j].a11 = a11;
    B[j].a00 = sel(m.real_node[j], b00, 0.0); B[j].a01 = sel(m.real_node[j], b01, 0.0);
    B[j].a10 = b10;                            B[j].a11 = b11;
    C[j].a00 = sel(m.real_node[j], c00, 0.0); C[j].a01 = splat(0.0);
    C[j].a10 = c10;                            C[j].a11 = c11;
  }
}

}  // namespace trpl
