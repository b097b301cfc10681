// trajectory.h - one (parameter set, measurement) trajectory, start to finish, inside one warp:
// initial condition -> adaptive RODAS4 time integration -> PL/TRTS readout at the measurement
// times -> log-likelihood against the measurement.  Replaces, for one trajectory,
//   forward_solver.py:41-203   solve()            (init, LSODA integration, PL readout, min_y floor)
//   trial_move_evaluation.py:96-166 one_sim_likelihood() (abs/negative test, log10 residual, sum)
//
// Integrator: RODAS4 (Hairer & Wanner, Solving ODEs II, sec. IV.7/IV.10; 6 stages, order 4(3),
// stiffly accurate, L-stable, gamma = 1/4), exact Jacobian every step, error = 6th stage.
// The coefficient set is verified against the order conditions in tools/proto/check_rodas_coeffs.py.
// Step control: Gustafsson predictive controller, all norms reduced with warp shuffles.
// Output: quintic Hermite interpolation of ln(signal) through the last three step points
// (value + time derivative, both by warp reductions), evaluated by the lanes in parallel.
#pragma once
#include <float.h>
#include "simt.h"
#include "model.h"
#include "blocktri.h"
#include "irf.h"

namespace trpl {
using namespace simt;

#if defined(TRPL_FN) && defined(__CUDACC__) && !defined(TRPL_HOST_EMU)
#define TRPL_CONST __constant__ const
#define TRPL_NOINLINE __device__ __noinline__
#else
#define TRPL_CONST static const
#define TRPL_NOINLINE static
#endif

// RODAS4 in the Hairer-Wanner "transformed" form:
//   (1/(gamma h) I - J) K_i = f(u + sum_j a_ij K_j) + sum_j (c_ij / h) K_j ,  u_new = u + sum m_i K_i
TRPL_CONST double RODAS4_A[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.1544000000000000e+01, 0, 0, 0, 0, 0},
    {0.9466785280815826e+00, 0.2557011698983284e+00, 0, 0, 0, 0},
    {0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00, 0, 0, 0},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 0, 0},
    {0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 1.0, 0}};
TRPL_CONST double RODAS4_C[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {-0.5668800000000000e+01, 0, 0, 0, 0, 0},
    {-0.2430093356833875e+01, -0.2063599157091915e+00, 0, 0, 0, 0},
    {-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02, 0, 0, 0},
    {0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02, 0, 0},
    {0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
     -0.6058818238834054e+01, 0}};
constexpr double RODAS4_GAMMA = 0.25;

// Weight of the running-charge components in the error norm (see DESIGN.md section 2,
// "Tolerances"): Q is slaved to the densities by dielectric relaxation (sub-picosecond), the
// stiffly accurate integrator resolves it like an algebraic variable, and the embedded estimate
// for such components is known to be pessimistic.
#ifndef TRPL_Q_ERR_WEIGHT
#define TRPL_Q_ERR_WEIGHT 0.03
#endif
constexpr double Q_ERR_WEIGHT = TRPL_Q_ERR_WEIGHT;

enum StatusBits {
  ST_OK = 0,
  ST_MAX_STEPS = 1,     // step budget exhausted before the last measurement time
  ST_H_UNDERFLOW = 2,   // step size collapsed
  ST_NONFINITE = 4,     // NaN/Inf state
  ST_FLOORED = 8,       // signal fell below DBL_MIN and was floored (forward_solver.py:190-192)
  ST_NEG_FRAC = 16,     // too many negative values (trial_move_evaluation.py:117-123) -> -inf
  ST_NAN_LL = 32,       // likelihood was NaN -> -inf (trial_move_evaluation.py:159-165)
  ST_CONV_FAIL = 64,    // IRF convolution failed -> -inf (trial_move_evaluation.py:83-87, :103-106)
  ST_EXPLICIT = 128     // informational: integrated by the explicit Runge-Kutta path (explicit.h)
};

enum OptFlags { OPT_FORCE_MIN_Y = 1, OPT_NO_LIKELIHOOD = 2, OPT_LADDER = 4, OPT_NO_EXPLICIT = 8 };

struct SolverOpts {
  double rtol, atol;
  double hmax;          // <= 0: steps limited by the error controller only
  int max_steps;
  int flags;
};

// one measurement (shared by every parameter set)
struct MeasDesc {
  double thickness;
  double ini_a, ini_b;   // fluence mode: fluence [cm^-2], absorption [cm^-1]
  int nx;
  int meas_type;         // MeasType
  int ini_mode;          // 0 density profile, 1 fluence/absorption/direction
  int ini_dir;           // fluence mode: <0 reverses the profile
  int n_t;               // number of measurement times (times[0] == 0)
  int t_off;             // offset of this measurement in times/vals/uncs
  int prof_off;          // offset of this measurement's profile (density mode)
  int irf_nk;            // rows of this measurement's IRF moment table, 0 = no convolution
  double irf_dt;         // mean IRF time step [ns]
  int irf_off;           // first row of the table in the moments array
  int pad_;
  double min_y;          // floor of the simulated signal (Grid.min_y, sim_utils.py:281)
};

struct TrajIn {
  const double* par;     // TRPL_NPARAM model-unit parameters
  const MeasDesc* md;
  const double* times;   // [n_t]
  const double* vals;    // [n_t] log10 measurement (may be null with OPT_NO_LIKELIHOOD)
  const double* uncs;    // [n_t]
  const double* profile; // [nx] cm^-3 (density mode)
  double scale_shift;    // log10 of the curve's scale factor
  double s2T[3];         // model_uncertainty^2 * T for up to three temperatures
  double fl_mult, al_mult;
  double* curve;         // optional [n_t] simulated signal in measurement units
  bool post_pass;        // likelihood is taken in finalize_trajectory (min_y floor, IRF, ladder)
  double* hist;          // per-warp scratch: (t, S, dS/dt) of accepted steps, HIST_CAP entries
};

constexpr int HIST_CAP = 1024;

// Everything only the final pass needs.  Built AFTER the integration loop so that none of it is
// live (and spilled) across the hot loop.
struct TailIn {
  IrfDesc irf;           // irf.nk == 0: no convolution
  // tempering ladder (OPT_LADDER): likelihood at every ladder temperature; s2T[1] holds sigma^2
  const double* ladder_T;
  int ladder_n;
  double* ladder_out;    // [ladder_n]
  double* r2_scratch;    // [n_t] per-warp
  double* u2_scratch;    // [n_t] per-warp
};

// what the integration loop hands to the final pass
struct TrajMid {
  double l[3];           // streaming likelihood sums (valid when !post_pass)
  double n_neg;
};

struct TrajOut {
  double logll[3];
  int status, n_acc, n_rej;
};

// shared-memory budget of one warp, in PAIRS (16 bytes per lane)
template <int NPL, int MODEL>
struct Slots {
  static constexpr int NKS = 5;                                // stages kept (the 6th is consumed in registers)
  static constexpr int TPAIRS = (MODEL == MODEL_TRAPS) ? (NPL + 1) / 2 : 0;   // trap component, two nodes per pair
  static constexpr int KSTRIDE = NPL + TPAIRS;                 // pairs per stage: (K_N, K_Q) per node [+ K_T]
  static constexpr int KBASE = 0;
  static constexpr int FAC = KBASE + NKS * KSTRIDE;
  static constexpr int TRAP = FAC + FacSlots<NPL>::COUNT;      // traps: 5 condensation coefficients per node (3 pairs)
  // lane-exchange scratch: 2 pairs for the solve; the factorisation needs 12 and borrows the K
  // region (dead at that point) when that is large enough, else it gets its own
  static constexpr int XCH = TRAP + ((MODEL == MODEL_TRAPS) ? 3 * NPL : 0);
  static constexpr int XCH_FACTOR = (NKS * KSTRIDE >= 12) ? KBASE : XCH;
  static constexpr int UNI = XCH + ((NKS * KSTRIDE >= 12) ? 2 : 12);   // one slot of warp-uniform scalars (Coef)
  static constexpr int COUNT = UNI + 1;
  static constexpr int BYTES = COUNT * 32 * 16;
};

// stage increment K_s <-> shared memory
template <int NPL, int MODEL>
TRPL_FN void store_k(LaneMem& sm, int kb, const Vec<NPL, MODEL>& k) {
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) sm.st2(kb + j, k.n[j], k.q[j]);
  if (MODEL == MODEL_TRAPS) {
    TRPL_UNROLL for (int j = 0; j < NPL; j += 2) sm.st2(kb + NPL + j / 2, k.t[j], (j + 1 < NPL) ? k.t[j + 1] : k.t[j]);
  }
}
template <int NPL, int MODEL>
TRPL_FN void load_k(const LaneMem& sm, int kb, Vec<NPL, MODEL>& k) {
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) sm.ld2(kb + j, k.n[j], k.q[j]);
  if (MODEL == MODEL_TRAPS) {
    TRPL_UNROLL for (int j = 0; j < NPL; j += 2) {
      real a, b;
      sm.ld2(kb + NPL + j / 2, a, b);
      k.t[j] = a;
      if (j + 1 < NPL) k.t[j + 1] = b;
    }
  } else {
    k.t[0] = splat(0.0);
  }
}

// ---- readout: signal and its time derivative, reduced over the warp --------------------------
template <int NPL, int MODEL>
TRPL_FN void readout(const Coef& c, const NodeMask<NPL>& m, int meas_type, const Vec<NPL, MODEL>& u,
                     const Vec<NPL, MODEL>& f, const RhsAux<NPL>& aux, double& val, double& dval) {
  real fql0 = shfl_up(f.q[NPL - 1], 1);
  fql0 = sel(m.first_lane, 0.0, fql0);
  real acc = splat(0.0), dacc = splat(0.0);
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
    const real fql = (j == 0) ? fql0 : f.q[j - 1];
    real fp = f.n[j] + (f.q[j] - fql);
    if (MODEL == MODEL_TRAPS) fp = fp + f.t[j];
    real a, d;
    if (meas_type == MEAS_TRPL) {                       // forward_solver.py:228-236,267-269
      a = fmadd(u.n[j], aux.p[j], -c.n0p0);
      d = fmadd(f.n[j], aux.p[j], u.n[j] * fp);
    } else {                                            // forward_solver.py:239-247,272-274
      a = fmadd(c.mun, u.n[j] - c.n0, c.mup * (aux.p[j] - c.p0));
      d = fmadd(c.mun, f.n[j], c.mup * fp);
    }
    acc = acc + sel(m.real_node[j], a, 0.0);
    dacc = dacc + sel(m.real_node[j], d, 0.0);
  }
  const double scale = (meas_type == MEAS_TRPL) ? c.ks * c.dx * 1e23 : Q_COULOMB * c.dx * 1e9;
  val = uni(warp_sum(acc)) * scale;
  dval = uni(warp_sum(dacc)) * scale;
}

// ---- Hermite history (warp-uniform scalars) -------------------------------------------------
struct History {
  double t[3], v[3], d[3];   // index 2 = newest
  int n;
};

struct HermiteCoef {
  double t1, t2, t0;
  double c0, c1, c2, c3, c4, c5;
  bool in_log;
};

// Newton form on the doubled nodes [t1,t1,t2,t2,t0,t0] (t1,t2 = last step; t0 = the step before)
TRPL_FN HermiteCoef hermite_setup(const History& H) {
  HermiteCoef k;
  const bool three = H.n >= 3;
  k.in_log = H.v[1] > 0.0 && H.v[2] > 0.0 && (!three || H.v[0] > 0.0);
  double v0, v1, v2, d0, d1, d2;
  if (k.in_log) {
    v1 = log(H.v[1]); v2 = log(H.v[2]); d1 = H.d[1] / H.v[1]; d2 = H.d[2] / H.v[2];
    v0 = three ? log(H.v[0]) : 0.0; d0 = three ? H.d[0] / H.v[0] : 0.0;
  } else {
    v1 = H.v[1]; v2 = H.v[2]; d1 = H.d[1]; d2 = H.d[2]; v0 = H.v[0]; d0 = H.d[0];
  }
  k.t1 = H.t[1]; k.t2 = H.t[2]; k.t0 = H.t[0];
  const double i21 = 1.0 / (k.t2 - k.t1);
  const double f12 = (v2 - v1) * i21;
  const double f112 = (f12 - d1) * i21;
  const double f122 = (d2 - f12) * i21;
  const double f1122 = (f122 - f112) * i21;
  k.c0 = v1; k.c1 = d1; k.c2 = f112; k.c3 = f1122; k.c4 = 0.0; k.c5 = 0.0;
  if (three) {
    const double i01 = 1.0 / (k.t0 - k.t1), i02 = 1.0 / (k.t0 - k.t2);
    const double f20 = (v0 - v2) * i02;
    const double f220 = (f20 - d2) * i02;
    const double f200 = (d0 - f20) * i02;
    const double f1220 = (f220 - f122) * i01;
    const double f2200 = (f200 - f220) * i02;
    const double f11220 = (f1220 - f1122) * i01;
    const double f12200 = (f2200 - f1220) * i01;
    k.c4 = f11220;
    k.c5 = (f12200 - f11220) * i01;
  }
  return k;
}

TRPL_FN real hermite_eval(const HermiteCoef& k, real tq) {
  const real a = tq - k.t1, b = tq - k.t2, e = tq - k.t0;
  const real a2 = a * a, b2 = b * b;
  real p = fmadd(e, k.c5, k.c4);          // c4 + c5 (t-t0)
  p = fmadd(p, b2, fmadd(b, k.c3, k.c2)); // c2 + c3 (t-t2) + (t-t2)^2 (...)
  p = fmadd(p, a2, fmadd(a, k.c1, k.c0)); // c0 + c1 (t-t1) + (t-t1)^2 (...)
  return k.in_log ? vexp(p) : p;
}

// Guard for hermite_eval: a smooth signal stays inside the band spanned by the step's two end
// values (plus one span of margin).  When the signal has decayed into rounding noise the
// controller takes huge steps and the high-order interpolant of noisy data can overshoot by tens
// of decades; those points fall back to (log-)linear interpolation between the two ends.  The
// fallback arithmetic sits behind a warp-uniform branch that is almost never taken.
TRPL_FN real hermite_guard(const History& H, const real& tq, const real& y) {
  const double a = H.v[1], b = H.v[2];
  const double mn = fmin(a, b), mx = fmax(a, b);
  const double margin = (mx - mn) + 1e-6 * fabs(mx);
  const mask outside = mnot(mand(y >= mn - margin, y <= mx + margin));    // true for NaN
  if (!warp_any(outside)) return y;
  const bool ends_log = a > 0.0 && b > 0.0;
  const double e1 = ends_log ? log(a) : a, e2 = ends_log ? log(b) : b;
  const real theta = (tq - H.t[1]) * (1.0 / (H.t[2] - H.t[1]));
  const real lin = fmadd(theta, e2 - e1, e1);
  const real fb = ends_log ? vexp(lin) : lin;
  return sel(outside, fb, y);
}

// ---- emission: measurement times inside an accepted step, likelihood sums -------------------------
struct Emitter {
  int io;
  bool floored;
  int status;
  History H;
  real ll0, ll1, ll2, nneg;
};

TRPL_FN void emitter_init(Emitter& e) {
  e.io = 0; e.floored = false; e.status = ST_OK; e.H.n = 0;
  for (int k = 0; k < 3; ++k) { e.H.t[k] = 0; e.H.v[k] = 0; e.H.d[k] = 0; }
  e.ll0 = splat(0.0); e.ll1 = splat(0.0); e.ll2 = splat(0.0); e.nneg = splat(0.0);
}

TRPL_FN void emitter_accumulate(Emitter& e, const TrajIn& in, bool want_ll, const ivec& k, const mask& take, const real& y) {
  if (in.curve) scatter(in.curve, k, take, y);
  if (want_ll && !in.post_pass) {
    e.nneg = e.nneg + sel(mand(take, y < 0.0), 1.0, 0.0);
    const real vk = gather(in.vals, k, take, 0.0);
    const real uk = gather(in.uncs, k, take, 1.0);
    const real r = (vlog10(vabs(y)) + in.scale_shift) - vk;
    const real r2 = r * r;
    const real u2 = 2.0 * (uk * uk);
    e.ll0 = e.ll0 + sel(take, r2 * rcp(in.s2T[0] + u2), 0.0);
    e.ll1 = e.ll1 + sel(take, r2 * rcp(in.s2T[1] + u2), 0.0);
    e.ll2 = e.ll2 + sel(take, r2 * rcp(in.s2T[2] + u2), 0.0);
  }
}

// returns true when the trajectory is finished (all times emitted, or the signal hit its floor)
TRPL_FN bool emitter_step(Emitter& e, const TrajIn& in, bool want_ll, double t, double val, double dval) {
  const MeasDesc& md = *in.md;
  const int n_t = md.n_t;
  const ivec lane = lane_id();
  History& H = e.H;
  H.t[0] = H.t[1]; H.v[0] = H.v[1]; H.d[0] = H.d[1];
  H.t[1] = H.t[2]; H.v[1] = H.v[2]; H.d[1] = H.d[2];
  H.t[2] = t; H.v[2] = val; H.d[2] = dval;
  if (H.n < 3) ++H.n;
  HermiteCoef hc;
  bool have_hc = false;
  while (e.io < n_t) {
    const ivec k = iadd(lane, e.io);
    const mask in_range = k < n_t;
    const real tq = gather(in.times, k, in_range, DBL_MAX);
    const unsigned bits = warp_ballot(mand(in_range, tq <= t));
    if (bits == 0u) break;
    int cnt = 0;
    { unsigned b = bits; while (b & 1u) { ++cnt; b >>= 1; } }
    real y;
    if (H.n < 2) {
      y = splat(val);
    } else {
      if (!have_hc) { hc = hermite_setup(H); have_hc = true; }
      y = hermite_guard(H, tq, hermite_eval(hc, tq));
      y = sel(tq >= t, val, y);
    }
    const mask take = lane < cnt;
    const unsigned low = warp_ballot(mand(take, y < md.min_y));
    if (low != 0u) {
      int firstlow = 0; { unsigned b = low; while (!(b & 1u)) { ++firstlow; b >>= 1; } }
      y = sel(lane >= firstlow, md.min_y, y);
      e.floored = true; e.status |= ST_FLOORED;
    }
    emitter_accumulate(e, in, want_ll, k, take, y);
    e.io += cnt;
    if (cnt < 32 || e.floored) break;
  }
  return e.io >= n_t || e.floored;
}

TRPL_FN void emitter_finish(Emitter& e, const TrajIn& in, bool want_ll, TrajMid& mid) {
  const int n_t = in.md->n_t;
  const ivec lane = lane_id();
  while (e.io < n_t) {                              // floor reached or integrator failure
    const ivec k = iadd(lane, e.io);
    emitter_accumulate(e, in, want_ll, k, k < n_t, splat(in.md->min_y));
    e.io += 32;
  }
  if (want_ll && !in.post_pass) {
    mid.l[0] = -uni(warp_sum(e.ll0)); mid.l[1] = -uni(warp_sum(e.ll1)); mid.l[2] = -uni(warp_sum(e.ll2));
    mid.n_neg = uni(warp_sum(e.nneg));
  } else {
    mid.l[0] = mid.l[1] = mid.l[2] = 0.0; mid.n_neg = 0.0;
  }
}

// Deferred emission.  The integration loop only appends (t, S, dS/dt) of every accepted step to a
// per-warp history buffer; this routine replays the buffer.  It is deliberately NOT inlined: the
// interpolation and likelihood arithmetic (log, exp, log10, gathers, ballots) then cannot disturb
// the register allocation of the hot loop, and nothing it needs is live there.
TRPL_NOINLINE void emit_history(const TrajIn& in, bool want_ll, const double* hist, int n, Emitter& e) {
  for (int i = 0; i < n; ++i) {
    if (emitter_step(e, in, want_ll, hist[3 * i], hist[3 * i + 1], hist[3 * i + 2])) break;
  }
}

// ---- the trajectory -------------------------------------------------------------------------
// Control flow is a small state machine so that the right-hand side, the readout/emit block and the
// linear solve each exist at exactly ONE code site (the kernel is instruction-cache sensitive):
//   PH_ACCEPTED  evaluate f(u) at the newly accepted state, read the signal out, emit measurement
//                times, then start a step (Jacobian + factorisation), stage 1 uses f(u)
//   PH_STAGE     evaluate f(stage argument), add the c-combination, solve
//   PH_RETRY     step rejected: same u, same f(u) (kept in shared memory), new h
enum Phase { PH_ACCEPTED = 0, PH_STAGE = 1, PH_RETRY = 2 };

template <int NPL, int MODEL>
TRPL_FN bool is_nonstiff(const Coef& c, const NodeMask<NPL>& m, const Vec<NPL, MODEL>& u, double tend);   // explicit.h

// Returns true (and does nothing else) when `allow_defer` is set and the trajectory is classified
// non-stiff at t = 0: the caller hands it to the explicit Runge-Kutta path instead.
template <int NPL, int MODEL, bool FULL>
TRPL_FN bool run_trajectory(const TrajIn& in, const SolverOpts& opt, LaneMem& sm, TrajOut& out,
                            TrajMid& mid, bool allow_defer) {
  typedef Slots<NPL, MODEL> SL;
  typedef Vec<NPL, MODEL> V;
  const MeasDesc& md = *in.md;
  const int L = md.nx;
  // The coefficients are formed once, parked in the uniform shared-memory slot and re-fetched by
  // each block of the loop that needs them (broadcast loads), instead of pinning registers.
  const Coef c = make_coef(in.par, md.thickness, L);
  park_coef(sm, SL::UNI, c);
  const NodeMask<NPL> m = make_mask<NPL, FULL>(L);
  const ivec lane = lane_id();
  const ivec node0 = imul(lane, NPL);
  const int n_t = md.n_t;
  const bool want_ll = !(opt.flags & OPT_NO_LIKELIHOOD);
  // min_y floor and IRF convolution need the whole curve: likelihood in a final pass over it
  const double min_y = md.min_y;

  // ---- initial condition (forward_solver.py:100-122) ----
  V u;
  {
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
        const double step = (L > 1) ? (md.thickness - c.dx) / (L - 1) : 0.0;   // np.linspace, sim_utils.py:269
        const real idx = to_real((md.ini_dir < 0) ? irsub(L - 1, i) : i);
        const real x = fmadd(idx, step, x0);
        dn = (fluence * alpha) * vexp(-(alpha * x));
      }
      const real n = dn + c.n0, p = dn + c.p0;
      const real rho = (p - c.p0) - (n - c.n0);                  // forward_solver.py:28-29
      rho_run = rho_run + sel(m.real_node[j], rho, 0.0);
      qloc[j] = rho_run;
      u.n[j] = sel(m.real_node[j], n, 1.0);
      if (MODEL == MODEL_TRAPS) u.t[j] = splat(0.0);
    }
    if (MODEL != MODEL_TRAPS) u.t[0] = splat(0.0);
    // Gauss's law: running net charge = in-lane running sum + exclusive warp scan of lane totals
    const real incl = warp_scan_incl(rho_run);
    const real excl = incl - rho_run;
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) u.q[j] = sel(m.real_node[j], qloc[j] + excl, 0.0);
  }

  const double tend = in.times[n_t - 1];
  if (allow_defer && is_nonstiff<NPL, MODEL>(c, m, u, tend)) return true;

  // ---- bookkeeping ----
  double t = 0.0;
  int status = ST_OK, n_acc = 0, n_rej = 0;
  int nh = 0;                       // entries in the history buffer
  Emitter em;                       // lives in local memory: only the (cold) emission touches it
  emitter_init(em);

  double h = 0.0, h_new = 0.0, gi = 0.0, ih = 0.0;
  float err_old = 1e-4f;
  double h_acc = 0.0;
  bool first = true, last_rejected = false, final_step = false;
  const double inv_n = 1.0 / (2.0 * L + ((MODEL == MODEL_TRAPS) ? L : 0));
  const double h_min = 1e-14 * fmax(tend, 1e-300);
  int phase = PH_ACCEPTED;
  int s = 0;
  V us = u;       // argument of the next right-hand-side evaluation
  V cs;           // sum_j c_sj / h K_j of the current stage
  TRPL_UNROLL for (int j = 0; j < NPL; ++j) { cs.n[j] = splat(0.0); cs.q[j] = splat(0.0); }
  TRPL_UNROLL for (int j = 0; j < (MODEL == MODEL_TRAPS ? NPL : 1); ++j) cs.t[j] = splat(0.0);
  PcrFac pf;

  for (;;) {
    V r;
    {
      // PH_RETRY re-evaluates f(u) (us == u): rejections are rare (~0.5% of steps) and this keeps
      // f(u) out of shared memory
      RhsAux<NPL> aux;
      const Coef cr = fetch_coef(sm, SL::UNI);
      rhs<NPL, MODEL>(cr, m, us, r, aux);
      if (phase == PH_ACCEPTED) {
        // ---- newly accepted state (us == u): read the signal out and log it ----
        double val, dval;
        readout<NPL, MODEL>(cr, m, md.meas_type, u, r, aux, val, dval);
        if (nh == HIST_CAP) {
          warp_sync();
          emit_history(in, want_ll, in.hist, nh, em);
          nh = 0;
          if (em.floored) break;
        }
        scatter(in.hist, iadd(lane, 3 * nh), lane < 3, sel(lane == 0, t, sel(lane == 1, val, dval)));
        ++nh;
        // done when the last measurement time is reached, or the signal fell through its floor
        // (forward_solver.py:190-192: the rest of the curve is min_y by definition)
        if (t >= tend || val < min_y) break;
        if (n_acc == 0) {
          // ---- initial step (Hairer's d0/d1 rule on the scaled norms) ----
          real s0 = splat(0.0), s1 = splat(0.0);
          TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
            const real iscn = rcp(fmadd(opt.rtol, vabs(u.n[j]), opt.atol));
            const real iscq = rcp(fmadd(opt.rtol, vmax(vabs(u.n[j]), vabs(aux.p[j])), opt.atol));
            const real a = u.n[j] * iscn, b = r.n[j] * iscn, q = u.q[j] * iscq, g = r.q[j] * iscq;
            s0 = s0 + sel(m.real_node[j], fmadd(a, a, q * q), 0.0);
            s1 = s1 + sel(m.real_node[j], fmadd(b, b, g * g), 0.0);
          }
          const double d0 = sqrt(uni(warp_sum(s0))), d1 = sqrt(uni(warp_sum(s1)));
          h = (d1 > 0.0 && d0 > 0.0) ? 0.01 * d0 / d1 : 1e-6;
          h = fmin(h, 1e-3 * fmax(tend, 1e-300));
          if (!(h > 0.0)) h = 1e-6;
          h_acc = h;
        } else {
          h = h_new;
        }
      }
    }
    if (phase != PH_STAGE) {
      // ---- start (or restart) a step from u with step size h; stage 1 right-hand side is f(u) ----
      if (n_acc + n_rej >= opt.max_steps) { status |= ST_MAX_STEPS; break; }
      if (opt.hmax > 0.0) h = fmin(h, opt.hmax);
      final_step = false;
      if (t + 1.01 * h >= tend) { h = tend - t; final_step = true; }
      if (h < h_min) { status |= ST_H_UNDERFLOW; break; }
      gi = 1.0 / (RODAS4_GAMMA * h);
      ih = 1.0 / h;
      {
        // W = 1/(gamma h) I - J, factorised
        Blk A[NPL], B[NPL], C[NPL];
        JacTraps<NPL> jt;
        jacobian<NPL, MODEL>(fetch_coef(sm, SL::UNI), m, u, A, B, C, jt);
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
          A[j] = blk_neg(A[j]); C[j] = blk_neg(C[j]);
          B[j].a00 = gi - B[j].a00; B[j].a01 = -B[j].a01; B[j].a10 = -B[j].a10; B[j].a11 = gi - B[j].a11;
        }
        // the front contact has no left neighbour (Q_0 is the fixed corner field)
        A[0] = blk_sel(m.first_lane, blk_zero(), A[0]);
        if (MODEL == MODEL_TRAPS) {
          // condense the node-local trap occupancy out of the block rows
          real g_n[NPL];
          TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
            const real idt = rcp(gi - jt.ft_t[j]);
            g_n[j] = jt.ft_n[j] * idt;              // K_T = idt * r_T + g_n * K_N
            sm.st2(SL::TRAP + 3 * j + 0, idt, g_n[j]);
            sm.st2(SL::TRAP + 3 * j + 1, jt.fn_t[j], jt.fq_t[j]);
            sm.st2(SL::TRAP + 3 * j + 2, jt.fq_tn[j], jt.fq_tn[j]);
          }
          const real gn_next = shfl_down(g_n[0], 1);
          TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
            const real gnn = (j == NPL - 1) ? gn_next : g_n[j + 1];
            B[j].a00 = B[j].a00 - jt.fn_t[j] * g_n[j];
            B[j].a10 = B[j].a10 - jt.fq_t[j] * g_n[j];
            C[j].a10 = C[j].a10 - jt.fq_tn[j] * gnn;
          }
        }
        bt_factor<NPL>(A, B, C, sm, SL::FAC, SL::XCH_FACTOR, pf);
      }
      s = 0;
    } else {
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        r.n[j] = r.n[j] + cs.n[j]; r.q[j] = r.q[j] + cs.q[j];
        if (MODEL == MODEL_TRAPS) r.t[j] = r.t[j] + cs.t[j];
      }
    }

    // ---- K_s = W^{-1} r ----
    V kk;
    if (MODEL != MODEL_TRAPS) kk.t[0] = splat(0.0);
    {
      V2 b[NPL];
      if (MODEL == MODEL_TRAPS) {
        real w[NPL], gn[NPL], fnt[NPL], fqt[NPL], fqtn[NPL];
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
          real idt, dummy;
          sm.ld2(SL::TRAP + 3 * j + 0, idt, gn[j]);
          sm.ld2(SL::TRAP + 3 * j + 1, fnt[j], fqt[j]);
          sm.ld2(SL::TRAP + 3 * j + 2, fqtn[j], dummy);
          w[j] = idt * r.t[j];                                  // idt * r_T
        }
        const real w_next = shfl_down(w[0], 1);
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
          const real wn = (j == NPL - 1) ? w_next : w[j + 1];
          b[j].x = fmadd(fnt[j], w[j], r.n[j]);
          b[j].y = fmadd(fqt[j], w[j], fmadd(fqtn[j], wn, r.q[j]));
        }
        bt_solve<NPL>(b, sm, SL::FAC, SL::XCH, pf);
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) kk.t[j] = fmadd(gn[j], b[j].x, w[j]);
      } else {
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) { b[j].x = r.n[j]; b[j].y = r.q[j]; }
        bt_solve<NPL>(b, sm, SL::FAC, SL::XCH, pf);
      }
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) { kk.n[j] = b[j].x; kk.q[j] = b[j].y; }
    }

    if (s < 5) {
      // ---- keep K_s, build the next stage argument and c-combination ----
      store_k<NPL, MODEL>(sm, SL::KBASE + s * SL::KSTRIDE, kk);
      ++s;
      // the newest increment is still in registers; older ones come back from shared memory
      {
        const double a = RODAS4_A[s][s - 1], cc = RODAS4_C[s][s - 1] * ih;
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
          us.n[j] = fmadd(a, kk.n[j], u.n[j]); us.q[j] = fmadd(a, kk.q[j], u.q[j]);
          cs.n[j] = cc * kk.n[j]; cs.q[j] = cc * kk.q[j];
          if (MODEL == MODEL_TRAPS) { us.t[j] = fmadd(a, kk.t[j], u.t[j]); cs.t[j] = cc * kk.t[j]; }
        }
      }
      for (int p = 0; p < s - 1; ++p) {
        const double a = RODAS4_A[s][p], cc = RODAS4_C[s][p] * ih;
        V kp;
        load_k<NPL, MODEL>(sm, SL::KBASE + p * SL::KSTRIDE, kp);
        TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
          us.n[j] = fmadd(a, kp.n[j], us.n[j]); us.q[j] = fmadd(a, kp.q[j], us.q[j]);
          cs.n[j] = fmadd(cc, kp.n[j], cs.n[j]); cs.q[j] = fmadd(cc, kp.q[j], cs.q[j]);
          if (MODEL == MODEL_TRAPS) { us.t[j] = fmadd(a, kp.t[j], us.t[j]); cs.t[j] = fmadd(cc, kp.t[j], cs.t[j]); }
        }
      }
      phase = PH_STAGE;
      continue;
    }

    // ---- stage 6 done: u_new = u + sum_j m_j K_j, m = (a_6j, 1); the error estimate is K_6 ----
    // (the stage argument `us` is rebuilt from the stored increments here so that it is not live
    //  across the solves)
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
      us.n[j] = u.n[j] + kk.n[j]; us.q[j] = u.q[j] + kk.q[j];
      if (MODEL == MODEL_TRAPS) us.t[j] = u.t[j] + kk.t[j];
    }
    TRPL_UNROLL for (int p = 0; p < 5; ++p) {
      const double a = RODAS4_A[5][p];
      V kp;
      load_k<NPL, MODEL>(sm, SL::KBASE + p * SL::KSTRIDE, kp);
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        us.n[j] = fmadd(a, kp.n[j], us.n[j]); us.q[j] = fmadd(a, kp.q[j], us.q[j]);
        if (MODEL == MODEL_TRAPS) us.t[j] = fmadd(a, kp.t[j], us.t[j]);
      }
    }
    real esum = splat(0.0);
    mask bad = mconst(false);
    real pold[NPL];
    holes<NPL, MODEL>(fetch_coef(sm, SL::UNI), m, u, pold);
    TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
      const real iscn = rcp(fmadd(opt.rtol, vmax(vabs(u.n[j]), vabs(us.n[j])), opt.atol));
      const real iscq = rcp(fmadd(opt.rtol, vmax(vabs(u.n[j]), vabs(pold[j])), opt.atol));
      const real en = kk.n[j] * iscn, eq = kk.q[j] * (iscq * Q_ERR_WEIGHT);
      real e2 = fmadd(en, en, eq * eq);
      if (MODEL == MODEL_TRAPS) {
        const real isct = rcp(fmadd(opt.rtol, vmax(vabs(u.t[j]), vmax(vabs(us.t[j]), vabs(u.n[j]))), opt.atol));
        const real et = kk.t[j] * isct;
        e2 = fmadd(et, et, e2);
      }
      esum = esum + sel(m.real_node[j], e2, 0.0);
      bad = mor(bad, mand(m.real_node[j], mor(is_nan(us.n[j]), is_nan(us.q[j]))));
    }
    const double err2 = uni(warp_sum(esum)) * inv_n;
    const bool nonfinite = warp_any(bad) || !(err2 == err2) || err2 > 1e300;
    const double err = nonfinite ? 1e10 : sqrt(err2);

    // ---- controller (Hairer's RODAS: standard + Gustafsson predictive) ----
    // step-size factor in single precision (it only steers h)
    const float errf = (float)fmin(err, 1e30);
    float fac = fmaxf(0.2f, fminf(6.0f, sqrtf(sqrtf(errf)) * (1.0f / 0.9f)));
    h_new = h / (double)fac;
    if (err <= 1.0) {
      ++n_acc;
      if (!first) {
        float fg = (float)(h_acc / h) * sqrtf(sqrtf(errf * errf / err_old)) * (1.0f / 0.9f);
        fg = fmaxf(0.2f, fminf(6.0f, fg));
        fac = fmaxf(fac, fg);
        h_new = h / (double)fac;
      }
      first = false; h_acc = h; err_old = fmaxf(1e-2f, errf);
      if (last_rejected) h_new = fmin(h_new, h);
      last_rejected = false;
      t = final_step ? tend : t + h;
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        u.n[j] = us.n[j]; u.q[j] = us.q[j];
        if (MODEL == MODEL_TRAPS) u.t[j] = us.t[j];
      }
      phase = PH_ACCEPTED;
    } else {
      ++n_rej;
      last_rejected = true;
      h = nonfinite ? 0.1 * h : h_new;
      TRPL_UNROLL for (int j = 0; j < NPL; ++j) {
        us.n[j] = u.n[j]; us.q[j] = u.q[j];
        if (MODEL == MODEL_TRAPS) us.t[j] = u.t[j];
      }
      phase = PH_RETRY;
    }
  }
  // replay the logged steps: measurement times, floor, likelihood sums; anything not reached
  // (floor, or integrator failure) is min_y: forward_solver.py:168 + :190-192
  warp_sync();
  emit_history(in, want_ll, in.hist, nh, em);
  emitter_finish(em, in, want_ll, mid);
  out.status = status | em.status; out.n_acc = n_acc; out.n_rej = n_rej;
  return false;
}

// ---- likelihood (trial_move_evaluation.py:117-166): streaming sums, or a pass over the curve ----
TRPL_FN void finalize_trajectory(const TrajIn& in, const TailIn& tl, const SolverOpts& opt,
                                 const TrajMid& mid, TrajOut& out) {
  const bool want_ll = !(opt.flags & OPT_NO_LIKELIHOOD);
  const int n_t = in.md->n_t;
  if (want_ll) {
    double l[3];
    double n_neg;
    int n_c = n_t;
    bool ok = true;
    const bool ladder = in.post_pass && (opt.flags & OPT_LADDER) && tl.ladder_n > 0 && tl.r2_scratch;
    if (in.post_pass) {
      warp_sync();
      const double* sol = in.curve;
      if (tl.irf.nk > 0) {
        ok = irf_convolve_trim(in.times, in.curve, n_t, tl.irf, n_c);
        sol = tl.irf.trim;
        if (!ok) out.status |= ST_CONV_FAIL;
      }
      if (ok) array_loglik(sol, n_c, in.vals, in.uncs, in.scale_shift, in.s2T,
                           (opt.flags & OPT_FORCE_MIN_Y) != 0, l, n_neg,
                           ladder ? tl.r2_scratch : nullptr, ladder ? tl.u2_scratch : nullptr);
    } else {
      l[0] = mid.l[0]; l[1] = mid.l[1]; l[2] = mid.l[2];
      n_neg = mid.n_neg;
    }
    const double ninf = -HUGE_VAL;
    if (!ok) {
      l[0] = l[1] = l[2] = ninf;
    } else {
      if (!(n_neg < 0.2 * n_c)) { out.status |= ST_NEG_FRAC; l[0] = l[1] = l[2] = ninf; }
      if (l[0] != l[0]) { out.status |= ST_NAN_LL; l[0] = ninf; }
      if (l[1] != l[1]) l[1] = ninf;
      if (l[2] != l[2]) l[2] = ninf;
    }
    out.logll[0] = l[0]; out.logll[1] = l[1]; out.logll[2] = l[2];
    if (ladder) {
      const bool failed = !ok || !(n_neg < 0.2 * n_c);
      ladder_loglik(tl.r2_scratch, tl.u2_scratch, failed ? 0 : n_c, in.s2T[1], tl.ladder_T, tl.ladder_n,
                    tl.ladder_out, failed);
    }
  } else {
    out.logll[0] = out.logll[1] = out.logll[2] = 0.0;
  }
}

}  // namespace trpl
