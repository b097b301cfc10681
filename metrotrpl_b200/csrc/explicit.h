// explicit.h - embedded explicit Runge-Kutta path for non-stiff trajectories.
//
// When transport is absent or slow (e.g. the mobility-free cases of the reference's own unit
// tests, Tests/test_eval_trial_move.py) the system is a set of mildly coupled recombination ODEs
// and an explicit pair is both cheaper per step (no Jacobian, no factorisation, no solves) and of
// higher order.  Dormand-Prince 5(4), FSAL; the tableau is verified against the order conditions
// in exact rational arithmetic by tools/proto/check_rodas_coeffs.py.
//
// Which trajectories come here is decided per trajectory at t = 0 from a Gershgorin-type bound on
// the spectral radius of the Jacobian (stiffness_bound): explicit iff the step count stability
// alone would force, rho * t_end / 3.3, is below EXPLICIT_MAX_STABILITY_STEPS.
#pragma once
#include "trajectory.h"

namespace trpl {
using namespace simt;

constexpr double EXPLICIT_MAX_STABILITY_STEPS = 200.0;
constexpr double DOPRI5_STABILITY = 3.3;

// upper bound on the spectral radius of df/du at the state u (rows of |J| summed, per block row)
template <int NPL, int MODEL>
TRPL_FN double stiffness_bound(const Coef& c, const NodeMask<NPL>& m, const Vec<NPL, MODEL>& u) {
  real P[NPL];
  holes<NPL, MODEL>(c, m, u, P);
  real worst = splat(0.0);
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real n = vabs(u.n[j]), p = vabs(P[j]);
    const real npx = vabs(fmadd(u.n[j], P[j], -c.n0p0));
    const real inv = rcp(fmadd(c.taun, p, c.taup * n));
    const real rate = fmadd(c.cn, n, fmadd(c.cp, p, c.ks)) + inv;
    real rho = rate * (n + p) + (c.cn + c.cp + (c.taun + c.taup) * inv * inv) * npx;   // recombination
    rho = rho + 4.0 * c.ix * (c.dn + c.dp);                                               // diffusion
    rho = rho + 4.0 * c.ix * c.ld * (c.an * n + c.ap * p) * 2.0;                          // drift / dielectric relaxation
    rho = rho + 2.0 * c.ix * (c.an + c.ap) * c.ld * vabs(u.q[j]);                         // field-driven drift
    rho = rho + c.ix * (c.sf + c.sb);                                                     // contacts
    if (MODEL == MODEL_TRAPS) rho = rho + c.kc * (c.nt + n + vabs(u.t[j])) + c.itaue;
    worst = vmax(worst, sel(m.real_node[j], rho, 0.0));
  }
  return uni(warp_max(worst));
}

template <int NPL, int MODEL>
TRPL_FN bool is_nonstiff(const Coef& c, const NodeMask<NPL>& m, const Vec<NPL, MODEL>& u, double tend) {
  const double rho = stiffness_bound<NPL, MODEL>(c, m, u);
  return rho * tend <= EXPLICIT_MAX_STABILITY_STEPS * DOPRI5_STABILITY;
}

// initial condition shared with the implicit path (forward_solver.py:100-122)
template <int NPL, int MODEL>
TRPL_FN void initial_state(const TrajIn& in, const Coef& c, const NodeMask<NPL>& m, Vec<NPL, MODEL>& u) {
  const MeasDesc& md = *in.md;
  const int L = md.nx;
  const ivec node0 = imul(lane_id(), NPL);
  real rho_run = splat(0.0);
  real qloc[NPL];
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const ivec i = iadd(node0, j);
    real dn;
    if (md.ini_mode == 0) {
      dn = gather(in.profile, i, m.real_node[j], 0.0) * 1e-21;
    } else {
      const double fluence = md.ini_a * in.fl_mult * 1e-14;
      const double alpha = md.ini_b * in.al_mult * 1e-7;
      const double x0 = 0.5 * c.dx;
      const double step = (L > 1) ? (md.thickness - c.dx) / (L - 1) : 0.0;
      const real idx = to_real((md.ini_dir < 0) ? irsub(L - 1, i) : i);
      dn = (fluence * alpha) * vexp(-(alpha * fmadd(idx, step, x0)));
    }
    const real n = dn + c.n0, p = dn + c.p0;
    const real rho = (p - c.p0) - (n - c.n0);
    rho_run = rho_run + sel(m.real_node[j], rho, 0.0);
    qloc[j] = rho_run;
    u.n[j] = sel(m.real_node[j], n, 1.0);
    if (MODEL == MODEL_TRAPS) u.t[j] = splat(0.0);
  }
  if (MODEL != MODEL_TRAPS) u.t[0] = splat(0.0);
  const real excl = warp_scan_incl(rho_run) - rho_run;
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) u.q[j] = sel(m.real_node[j], qloc[j] + excl, 0.0);
}

// Dormand-Prince 5(4)
TRPL_CONST double DP_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5.0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
TRPL_CONST double DP_E[7] = {71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0,
                             22.0 / 525.0, -1.0 / 40.0};

template <int NPL, int MODEL, bool FULL>
TRPL_FN void run_trajectory_explicit(const TrajIn& in, const SolverOpts& opt, TrajMem& mem, TrajOut& out,
                                     TrajMid& mid) {
  typedef Slots<NPL, MODEL> SL;
  typedef Vec<NPL, MODEL> V;
  const MeasDesc& md = *in.md;
  const int L = md.nx;
  const Coef c = make_coef(in.par, md.thickness, L);
  const NodeMask<NPL> m = make_mask<NPL, FULL>(L);
  const bool want_ll = !(opt.flags & OPT_NO_LIKELIHOOD);
  const double tend = in.times[md.n_t - 1];
  V u;
  initial_state<NPL, MODEL>(in, c, m, u);
  double ex_floor;
  {
    real dn_max = splat(0.0);
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) dn_max = vmax(dn_max, sel(m.real_node[j], vabs(u.n[j] - c.n0), 0.0));
    ex_floor = EXCESS_RANGE * uni(warp_max(dn_max));
  }
  Emitter em;
  emitter_init(em);
  int n_acc = 0, n_rej = 0;
  double t = 0.0, h = 0.0;
  const double inv_n = 1.0 / (2.0 * L + ((MODEL == MODEL_TRAPS) ? L : 0));
  const double h_min = 1e-14 * fmax(tend, 1e-300);
  static_assert(7 * SL::KSTRIDE <= SL::KCAP, "stage storage does not fit the memory that holds the increments");
  auto& km = kmem<SL>(mem);

  // k1 = f(u)
  V k1; RhsAux<NPL> aux;
  rhs<NPL, MODEL>(c, m, u, k1, aux);
  double val, dval;
  readout<NPL, MODEL>(c, m, md.meas_type, u, k1, aux, val, dval);
  int nh = 0;                       // entries in the step log (trajectory.h, "Deferred emission")
  bool done = log_point(in, want_ll, em, nh, 0.0, val, dval) || 0.0 >= tend || val < md.min_y;
  if (!done) {
    real s0 = splat(0.0), s1 = splat(0.0);
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
      const real iscn = rcp(fmadd(opt.rtol, vabs(u.n[j]), opt.atol));
      const real iscq = rcp(fmadd(opt.rtol, vmax(vabs(u.n[j]), vabs(aux.p[j])), opt.atol));
      const real a = u.n[j] * iscn, b = k1.n[j] * iscn, q = u.q[j] * iscq, g = k1.q[j] * iscq;
      s0 = s0 + sel(m.real_node[j], fmadd(a, a, q * q), 0.0);
      s1 = s1 + sel(m.real_node[j], fmadd(b, b, g * g), 0.0);
    }
    const double d0 = sqrt(uni(warp_sum(s0))), d1 = sqrt(uni(warp_sum(s1)));
    h = (d1 > 0.0 && d0 > 0.0) ? 0.01 * d0 / d1 : 1e-6;
    h = fmin(h, 1e-2 * fmax(tend, 1e-300));
    if (!(h > 0.0)) h = 1e-6;
  }
  while (!done) {
    if (n_acc + n_rej >= opt.max_steps) { em.status |= ST_MAX_STEPS; break; }
    if (opt.hmax > 0.0) h = fmin(h, opt.hmax);
    bool final_step = false;
    if (t + 1.01 * h >= tend) { h = tend - t; final_step = true; }
    if (h < h_min) { em.status |= ST_H_UNDERFLOW; break; }
    store_k<NPL, MODEL>(km, SL::KBASE, k1);
    V us, kk;
    for (int s = 1; s < 7; ++s) {
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        us.n[j] = u.n[j]; us.q[j] = u.q[j];
        if (MODEL == MODEL_TRAPS) us.t[j] = u.t[j];
      }
      if (MODEL != MODEL_TRAPS) us.t[0] = splat(0.0);
      for (int p = 0; p < s; ++p) {
        const double a = DP_A[s][p] * h;
        V kp;
        load_k<NPL, MODEL>(km, SL::KBASE + p * SL::KSTRIDE, kp);
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
          us.n[j] = fmadd(a, kp.n[j], us.n[j]); us.q[j] = fmadd(a, kp.q[j], us.q[j]);
          if (MODEL == MODEL_TRAPS) us.t[j] = fmadd(a, kp.t[j], us.t[j]);
        }
      }
      rhs<NPL, MODEL>(c, m, us, kk, aux);
      store_k<NPL, MODEL>(km, SL::KBASE + s * SL::KSTRIDE, kk);
    }
    // us = u_new (stage 7 argument), kk = f(u_new); error = h sum e_j k_j
    V er;
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) { er.n[j] = splat(0.0); er.q[j] = splat(0.0); if (MODEL == MODEL_TRAPS) er.t[j] = splat(0.0); }
    for (int p = 0; p < 7; ++p) {
      const double e = DP_E[p] * h;
      V kp;
      load_k<NPL, MODEL>(km, SL::KBASE + p * SL::KSTRIDE, kp);
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        er.n[j] = fmadd(e, kp.n[j], er.n[j]); er.q[j] = fmadd(e, kp.q[j], er.q[j]);
        if (MODEL == MODEL_TRAPS) er.t[j] = fmadd(e, kp.t[j], er.t[j]);
      }
    }
    real esum = splat(0.0);
    mask bad = mconst(false);
    real pold[NPL];
    holes<NPL, MODEL>(c, m, u, pold);
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
      // same scales as the Rosenbrock path (trajectory.h): excess density for N, carrier density for Q
      const real mx = vmax(vmax(vabs(u.n[j] - c.n0), vabs(us.n[j] - c.n0)), ex_floor);
      const real iscn = rcp(fmadd(opt.rtol, mx, opt.atol));
      const real iscq = rcp(fmadd(opt.rtol, vmax(vabs(u.n[j]), vabs(pold[j])), opt.atol));
      const real en = er.n[j] * iscn, eq = er.q[j] * iscq;
      real e2 = fmadd(en, en, eq * eq);
      if (MODEL == MODEL_TRAPS) {
        const real isct = rcp(fmadd(opt.rtol, vmax(vabs(u.t[j]), vmax(vabs(us.t[j]), mx)), opt.atol));
        const real et = er.t[j] * isct;
        e2 = fmadd(et, et, e2);
      }
      esum = esum + sel(m.real_node[j], e2, 0.0);
      bad = mor(bad, mand(m.real_node[j], mor(is_nan(us.n[j]), is_nan(us.q[j]))));
    }
    const double err2 = uni(warp_sum(esum)) * inv_n;
    const bool nonfinite = warp_any(bad) || !(err2 == err2) || err2 > 1e300;
    const double err = nonfinite ? 1e10 : sqrt(err2);
    const float errf = (float)fmin(fmax(err, 1e-30), 1e30);
    const float fac = fmaxf(0.2f, fminf(5.0f, 0.9f * powf(errf, -0.2f)));
    if (err <= 1.0) {
      ++n_acc;
      t = final_step ? tend : t + h;
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        u.n[j] = us.n[j]; u.q[j] = us.q[j]; k1.n[j] = kk.n[j]; k1.q[j] = kk.q[j];
        if (MODEL == MODEL_TRAPS) { u.t[j] = us.t[j]; k1.t[j] = kk.t[j]; }
      }
      readout<NPL, MODEL>(c, m, md.meas_type, u, k1, aux, val, dval);     // aux.p belongs to u_new
      done = log_point(in, want_ll, em, nh, t, val, dval) || t >= tend || val < md.min_y;
      h = h * (double)fac;
    } else {
      ++n_rej;
      h = nonfinite ? 0.1 * h : h * (double)fminf(fac, 1.0f);
    }
  }
  warp_sync();
  emit_history(in, want_ll, in.hist, nh, em, false);
  emitter_finish(em, in, want_ll, mid);
  out.status = em.status; out.n_acc = n_acc; out.n_rej = n_rej;
}

}  // namespace trpl
