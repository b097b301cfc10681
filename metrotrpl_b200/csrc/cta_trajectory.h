// cta_trajectory.h - RODAS4 with one (parameter set, measurement) trajectory per CTA: the north
// star's mapping taken literally, built in round 2 to find out whether more threads per trajectory
// shorten a tempering iteration (which is as slow as its slowest trajectory).  MEASURED ANSWER: no.
// 6.5 us per integrator step against the one-warp kernel's 6.8, at 0.3x its throughput (DESIGN.md
// section 5, tools/latency_probe.py): a RODAS4 step is a chain of ~50 dependent exchange-and-reduce
// levels however it is decomposed.  The latency path that does work is extrapolation.h (parallelism
// inside the method).  This kernel stays behind TRPL_OPT_CTA_PER_TRAJ, parity-tested against the
// same converged truth, and nothing selects it by default.
//
// Same method as trajectory.h - RODAS4, exact Jacobian, the same controller, error norm, step log,
// dense output and likelihood code - but one space node per THREAD (nx = 128 threads, four warps):
//   * right-hand side / Jacobian: neighbour values come from one shared-memory exchange per
//     evaluation (state written, one bar.sync, neighbours read); the flux through the face between
//     i-1 and i is recomputed by thread i from the same operands thread i-1 uses, bit for bit, so
//     carriers and charge are conserved exactly as in the one-warp kernel;
//   * W = I/(gamma h) - J: no interior elimination at all - the 128 block rows are reduced by
//     parallel cyclic reduction in 7 levels (strides 1..64), rows exchanged through shared memory
//     (double buffered, one bar.sync per level); the 7 x 2 multiplier blocks and the final inverse
//     stay in registers (one node per thread leaves room) and each of the six solves of a step is
//     7 levels on a 2-vector;
//   * step-size control: the error norm, the signal and its derivative are block reductions (warp
//     shuffles, then four partials through shared memory in a fixed order, so every thread holds the
//     same bits and control flow stays CTA-uniform without broadcasts);
//   * emission / likelihood: warp 0 alone runs the step log, emit_history and finalize_trajectory of
//     trajectory.h (they are one-warp routines); the other three warps wait at the next barrier.
// It trades the one-warp kernel's in-lane sweeps (no synchronisation, four FP64 chains in flight) for
// two more exchange levels per solve and ~60 bar.sync per step.
// Restrictions of this instantiation: 'std' model, nx = 128 for every measurement.
#pragma once
#include "trajectory.h"

#if defined(__CUDACC__) && !defined(TRPL_HOST_EMU)
namespace trpl {
namespace cta {

constexpr int NX = 128;
constexpr int NWARP = NX / 32;
constexpr int LEVELS = 7;

struct Smem {
  double sN[NX];
  double sQ[NX];
  double2 xf[2][6][NX];      // factorisation exchange: {B^-1, A, C} of every row, double buffered
  double2 xs[2][NX];         // solve exchange, double buffered
  double sF[NX];             // dQ/dt of the accepted state (readout: dP/dt needs the left neighbour's)
  double red[2][NWARP][2];   // block reductions, double buffered
  Coef coef;                 // warp-uniform model coefficients of the current trajectory
  int flag;
  int traj;
};

struct Red { int buf; };

// block-wide sum of two values, identical bits on every thread; one bar.sync
__device__ __forceinline__ void block_sum2(Smem& s, Red& r, double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s.red[r.buf][w][0] = a; s.red[r.buf][w][1] = b; }
  __syncthreads();
  a = (s.red[r.buf][0][0] + s.red[r.buf][1][0]) + (s.red[r.buf][2][0] + s.red[r.buf][3][0]);
  b = (s.red[r.buf][0][1] + s.red[r.buf][1][1]) + (s.red[r.buf][2][1] + s.red[r.buf][3][1]);
  r.buf ^= 1;
}
// block-wide max of one value and inclusive scan of another (the initial condition only)
__device__ __forceinline__ void block_max_scan(Smem& s, Red& r, double& mx, double& scan) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, scan, o);
    if (lane >= o) scan += y;
  }
  if (lane == 31) { s.red[r.buf][w][0] = mx; s.red[r.buf][w][1] = scan; }
  __syncthreads();
  mx = fmax(fmax(s.red[r.buf][0][0], s.red[r.buf][1][0]), fmax(s.red[r.buf][2][0], s.red[r.buf][3][0]));
  double off = 0.0;
  for (int k = 0; k < w; ++k) off += s.red[r.buf][k][1];
  scan += off;
  r.buf ^= 1;
}

// what the right-hand side leaves behind for the Jacobian at the same state (stage 1 of a step)
struct Aux {
  double P, npx, inv, rate, ql, n_prev, snl, spl, isum;
};

// f(u) at node i (model.h rhs<1, MODEL_STD>, forward_solver.py:332-372), neighbours through shared memory
__device__ __forceinline__ void rhs_node(const Coef& c, Smem& s, int i, double n, double q, double& fn,
                                         double& fq, Aux& x) {
  s.sN[i] = n;
  s.sQ[i] = q;
  __syncthreads();
  const bool first = (i == 0), last = (i == NX - 1);
  const double ql = first ? 0.0 : s.sQ[i - 1];
  const double n_prev = first ? n : s.sN[i - 1];
  const double n_next = last ? n : s.sN[i + 1];
  const double q_next = last ? q : s.sQ[i + 1];
  const double P = n + c.d0 + (q - ql);
  const double p_next = n_next + c.d0 + (q_next - q);
  const double npx = fma(n, P, -c.n0p0);
  const double inv = simt::rcp(fma(c.taun, P, c.taup * n));
  const double rate = fma(c.cn, n, fma(c.cp, P, c.ks)) + inv;
  const double loss = rate * npx;
  // contacts: thread 0 evaluates the front, thread NX-1 the back (forward_solver.py:346-352)
  const double isum = simt::rcp(n + P);
  const double surf = (last ? c.sbx : c.sfx) * npx * isum;
  // right face
  const double snl = (n + n_next) * c.anl, spl = (P + p_next) * c.apl;
  double jn = fma(snl, q, c.dnx * (n_next - n));                        // forward_solver.py:356-357
  double jp = fma(spl, q, -(c.dpx * (p_next - P)));                     // forward_solver.py:358-359
  if (last) { jn = -surf; jp = surf; }
  // left face: the same expression thread i-1 evaluates for its right face, same operands
  double jl = fma((n_prev + n) * c.anl, ql, c.dnx * (n - n_prev));
  if (first) jl = surf;
  fn = (jn - jl) - loss;                                                // forward_solver.py:369
  fq = -jn - jp;                                                        // forward_solver.py:363
  x.P = P; x.npx = npx; x.inv = inv; x.rate = rate; x.ql = ql; x.n_prev = n_prev; x.snl = snl; x.spl = spl;
  x.isum = isum;
}

// block row (A, B, C) of the Jacobian at node i (model.h jacobian<1, MODEL_STD>)
__device__ __forceinline__ void jac_node(const Coef& c, int i, double n, double q, const Aux& x, Blk& A, Blk& B,
                                         Blk& C) {
  const bool first = (i == 0), last = (i == NX - 1), inner = !last;
  const double svel = last ? c.sbx : c.sfx;
  const double common = x.npx * x.isum * x.isum;
  const double s_n = svel * (x.P * x.isum - common);
  const double s_p = svel * (n * x.isum - common);
  const double r_ni = fma(c.anl, q, -c.dnx), r_nn = fma(c.anl, q, c.dnx);
  const double p_pi = fma(c.apl, q, c.dpx), p_pn = fma(c.apl, q, -c.dpx);
  const double inv2 = x.inv * x.inv;
  const double r_n = fma(c.cn - c.taup * inv2, x.npx, x.rate * x.P);
  const double r_p = fma(c.cp - c.taun * inv2, x.npx, x.rate * n);
  const double jr_ni = inner ? r_ni : -s_n;
  const double jr_pi = last ? -s_p : 0.0;
  const double jr_nn = inner ? r_nn : 0.0;
  const double jr_q = inner ? x.snl : 0.0;
  double jl_nm = fma(c.anl, x.ql, -c.dnx), jl_ni = fma(c.anl, x.ql, c.dnx), jl_q = (x.n_prev + n) * c.anl;
  double jl_pi = 0.0;
  if (first) { jl_nm = 0.0; jl_q = 0.0; jl_ni = s_n; jl_pi = s_p; }
  const double c_p = (jr_pi - jl_pi) - r_p;
  A.a00 = -jl_nm;                       A.a01 = -(jl_q + c_p);
  A.a10 = 0.0;                          A.a11 = inner ? p_pi : 0.0;
  B.a00 = ((jr_ni - jl_ni) - r_n) + c_p; B.a01 = jr_q + c_p;
  B.a10 = inner ? -(jr_ni + p_pi) : 0.0; B.a11 = inner ? -((jr_q + x.spl) + (p_pi - p_pn)) : 0.0;
  C.a00 = jr_nn;                        C.a01 = 0.0;
  C.a10 = inner ? -(jr_nn + p_pn) : 0.0; C.a11 = inner ? -p_pn : 0.0;
}

struct Factor {
  Blk al[LEVELS], ga[LEVELS];
  Blk binv;
};

__device__ __forceinline__ void st_blk3(Smem& s, int b, int i, const Blk& x, const Blk& y, const Blk& z) {
  s.xf[b][0][i] = make_double2(x.a00, x.a01); s.xf[b][1][i] = make_double2(x.a10, x.a11);
  s.xf[b][2][i] = make_double2(y.a00, y.a01); s.xf[b][3][i] = make_double2(y.a10, y.a11);
  s.xf[b][4][i] = make_double2(z.a00, z.a01); s.xf[b][5][i] = make_double2(z.a10, z.a11);
}
__device__ __forceinline__ Blk ld_blk(const Smem& s, int b, int p, int i) {
  const double2 u = s.xf[b][p][i], v = s.xf[b][p + 1][i];
  Blk r; r.a00 = u.x; r.a01 = u.y; r.a10 = v.x; r.a11 = v.y; return r;
}

// PCR factorisation of the block-tridiagonal W (rows ra | rb | rc).  No masking at the ends: ra is an
// exact zero block on rows < stride and rc on rows >= NX - stride (products with the zero
// sub/super-diagonal of the first/last row), and out-of-range reads are clamped to the own row.
__device__ __forceinline__ void factor(Smem& s, int i, Blk ra, Blk rb, Blk rc, Factor& F) {
#pragma unroll
  for (int k = 0; k < LEVELS; ++k) {
    const int st = 1 << k, b = k & 1;
    const Blk bi = blk_inv(rb);
    st_blk3(s, b, i, bi, ra, rc);
    __syncthreads();
    const int up = i >= st ? i - st : i, dn = i + st < NX ? i + st : i;
    const Blk bi_up = ld_blk(s, b, 0, up), ra_up = ld_blk(s, b, 2, up), rc_up = ld_blk(s, b, 4, up);
    const Blk bi_dn = ld_blk(s, b, 0, dn), ra_dn = ld_blk(s, b, 2, dn), rc_dn = ld_blk(s, b, 4, dn);
    const Blk alpha = blk_mul_neg(ra, bi_up);
    const Blk gamma = blk_mul_neg(rc, bi_dn);
    rb = blk_fma(gamma, ra_dn, blk_fma(alpha, rc_up, rb));
    ra = blk_mul(alpha, ra_up);
    rc = blk_mul(gamma, rc_dn);
    F.al[k] = alpha; F.ga[k] = gamma;
  }
  F.binv = blk_inv(rb);
}

// W x = r, in place
__device__ __forceinline__ void solve(Smem& s, int i, const Factor& F, double& x, double& y) {
#pragma unroll
  for (int k = 0; k < LEVELS; ++k) {
    const int st = 1 << k, b = k & 1;
    s.xs[b][i] = make_double2(x, y);
    __syncthreads();
    const double2 up = s.xs[b][i >= st ? i - st : i];
    const double2 dn = s.xs[b][i + st < NX ? i + st : i];
    const double lx = fma(F.al[k].a00, up.x, F.al[k].a01 * up.y), ly = fma(F.al[k].a10, up.x, F.al[k].a11 * up.y);
    const double hx = fma(F.ga[k].a01, dn.y, fma(F.ga[k].a00, dn.x, x));
    const double hy = fma(F.ga[k].a11, dn.y, fma(F.ga[k].a10, dn.x, y));
    x = lx + hx; y = ly + hy;
  }
  const double zx = fma(F.binv.a00, x, F.binv.a01 * y), zy = fma(F.binv.a10, x, F.binv.a11 * y);
  x = zx; y = zy;
}

// stage combinations with immediate coefficients: us += a_SP K_P, cs += (c_SP / h) K_P for P = PMAX..0
template <int S, int P>
__device__ __forceinline__ void combine(const double (&Kn)[5], const double (&Kq)[5], double ih, double& usn,
                                        double& usq, double& csn, double& csq) {
  constexpr double a = Rodas4::A[S][P];
  const double cc = Rodas4::C[S][P] * ih;
  usn = fma(a, Kn[P], usn); usq = fma(a, Kq[P], usq);
  csn = fma(cc, Kn[P], csn); csq = fma(cc, Kq[P], csq);
  if constexpr (P > 0) combine<S, P - 1>(Kn, Kq, ih, usn, usq, csn, csq);
}

// stage ST (0-based) of the Rosenbrock step: K_ST = W^-1 (f(us) + cs), then the argument and
// c-combination of stage ST+1.  On return of stage 5, (rn, rq) hold K_6 and (usn, usq) the
// stage-6 argument, so that u_new = us + K_6 and the error estimate is K_6.
template <int ST>
__device__ __forceinline__ void stage(Smem& s, int i, const Factor& F, double n, double q, double ih,
                                      double (&Kn)[5], double (&Kq)[5], double& usn, double& usq, double& csn,
                                      double& csq, double& rn, double& rq) {
  if constexpr (ST > 0) {
    Aux unused;
    rhs_node(s.coef, s, i, usn, usq, rn, rq, unused);
    rn += csn; rq += csq;
  }
  solve(s, i, F, rn, rq);
  if constexpr (ST < 5) {
    Kn[ST] = rn; Kq[ST] = rq;
    usn = n; usq = q; csn = 0.0; csq = 0.0;
    combine<ST + 1, ST>(Kn, Kq, ih, usn, usq, csn, csq);
  }
}

// One trajectory, all 128 threads.  `out` / `mid` are meaningful on warp 0, which ran the emission.
__device__ __forceinline__ void run_trajectory_cta(const TrajIn& in, const SolverOpts& opt, Smem& s, TrajOut& out,
                                                   TrajMid& mid) {
  const MeasDesc& md = *in.md;
  const int i = threadIdx.x;
  const int warp = i >> 5, lane = i & 31;
  // warp-uniform model coefficients live in shared memory (30 doubles; broadcast loads on use)
  __syncthreads();                         // the previous trajectory is done with s.coef
  if (i == 0) s.coef = make_coef(in.par, md.thickness, NX);
  __syncthreads();
  const Coef& c = s.coef;
  const int n_t = md.n_t;
  const bool want_ll = !(opt.flags & OPT_NO_LIKELIHOOD);
  const double min_y = md.min_y;
  Red red{0};

  // ---- initial condition (forward_solver.py:100-122), Gauss's law as a block prefix sum ----
  double n, q, ex_floor;
  {
    double dn;
    if (md.ini_mode == 0) {
      dn = in.profile[i] * 1e-21;
    } else {
      const double fluence = md.ini_a * in.fl_mult * 1e-14;
      const double alpha = md.ini_b * in.al_mult * 1e-7;
      const double step = (md.thickness - c.dx) / (NX - 1);
      const double idx = (double)((md.ini_dir < 0) ? (NX - 1 - i) : i);
      dn = (fluence * alpha) * exp(-(alpha * fma(idx, step, 0.5 * c.dx)));
    }
    n = dn + c.n0;
    const double p = dn + c.p0;
    double scan = (p - c.p0) - (n - c.n0);
    double mx = fabs(dn);
    block_max_scan(s, red, mx, scan);
    q = scan;
    ex_floor = EXCESS_RANGE * mx;
  }

  const double tend = in.times[n_t - 1];
  double t = 0.0;
  int status = ST_OK, n_acc = 0, n_rej = 0, nh = 0;
  Emitter em;
  emitter_init(em);
  double h = 0.0, h_new = 0.0;
  float err2_old = 1e-8f;
  double ih_acc = 0.0;
  bool first = true, last_rejected = false;
  const double inv_n = 1.0 / (2.0 * NX);
  const double h_min = 1e-14 * fmax(tend, 1e-300);
  const double sig_scale = (md.meas_type == MEAS_TRPL) ? c.ks * c.dx * 1e23 : Q_COULOMB * c.dx * 1e9;
  bool done = false;

  while (!done) {
    // ---- newly accepted (or initial) state: f(u), signal and its time derivative, step log ----
    double f0n, f0q;
    Aux ax;
    rhs_node(c, s, i, n, q, f0n, f0q, ax);
    double val, dval;
    {
      s.sF[i] = f0q;
      __syncthreads();
      const double fql = (i == 0) ? 0.0 : s.sF[i - 1];
      const double fp = f0n + (f0q - fql);                             // dP_i/dt
      double a, d;
      if (md.meas_type == MEAS_TRPL) {                                 // forward_solver.py:228-236,267-269
        a = ax.npx;
        d = fma(f0n, ax.P, n * fp);
      } else {                                                         // forward_solver.py:239-247,272-274
        a = fma(c.mun, n - c.n0, c.mup * (ax.P - c.p0));
        d = fma(c.mun, f0n, c.mup * fp);
      }
      block_sum2(s, red, a, d);
      val = a * sig_scale; dval = d * sig_scale;
    }
    if (nh == HIST_CAP) {                  // log full: warp 0 emits what it covers and keeps the last two entries
      if (warp == 0) {
        __syncwarp();
        emit_history(in, want_ll, in.hist, nh, em, true);
        if (lane == 0) s.flag = em.floored ? 1 : 0;
      }
      __syncthreads();
      nh = 2;
      if (s.flag) break;
    }
    if (warp == 0 && lane < 3) in.hist[3 * nh + lane] = (lane == 0) ? t : (lane == 1 ? val : dval);
    ++nh;
    // done when the last measurement time is reached, or the signal fell through its floor
    // (forward_solver.py:190-192: the rest of the curve is min_y by definition)
    if (t >= tend || val < min_y) break;
    if (n_acc == 0) {
      // ---- initial step (Hairer's d0/d1 rule on the scaled norms, as trajectory.h) ----
      const double iscn = simt::rcp(fma(opt.rtol, fabs(n), opt.atol));
      const double iscq = simt::rcp(fma(opt.rtol, fmax(fabs(n), fabs(ax.P)), opt.atol));
      const double a0 = n * iscn, b0 = f0n * iscn, q0 = q * iscq, g0 = f0q * iscq;
      double s0 = fma(a0, a0, q0 * q0), s1 = fma(b0, b0, g0 * g0);
      block_sum2(s, red, s0, s1);
      const double d0 = sqrt(s0), d1 = sqrt(s1);
      h = (d1 > 0.0 && d0 > 0.0) ? 0.01 * d0 / d1 : 1e-6;
      h = fmin(h, 1e-3 * fmax(tend, 1e-300));
      if (!(h > 0.0)) h = 1e-6;
    } else {
      h = h_new;
    }
    // ---- attempt steps from u until one is accepted ----
    for (;;) {
      if (n_acc + n_rej >= opt.max_steps) { status |= ST_MAX_STEPS; done = true; break; }
      if (opt.hmax > 0.0) h = fmin(h, opt.hmax);
      bool final_step = false;
      if (t + 1.01 * h >= tend) { h = tend - t; final_step = true; }
      if (h < h_min) { status |= ST_H_UNDERFLOW; done = true; break; }
      const double ih = simt::rcp(h);
      const double gi = (1.0 / RODAS4_GAMMA) * ih;
      Factor F;
      {
        Blk A, B, C;
        jac_node(c, i, n, q, ax, A, B, C);
        A = blk_neg(A); C = blk_neg(C);
        B.a00 = gi - B.a00; B.a01 = -B.a01; B.a10 = -B.a10; B.a11 = gi - B.a11;
        if (i == 0) A = blk_zero();          // the front contact has no left neighbour (Q_0 is fixed)
        factor(s, i, A, B, C, F);
      }
      double Kn[5], Kq[5];
      double usn = n, usq = q, csn = 0.0, csq = 0.0, rn = f0n, rq = f0q;
      stage<0>(s, i, F, n, q, ih, Kn, Kq, usn, usq, csn, csq, rn, rq);
      stage<1>(s, i, F, n, q, ih, Kn, Kq, usn, usq, csn, csq, rn, rq);
      stage<2>(s, i, F, n, q, ih, Kn, Kq, usn, usq, csn, csq, rn, rq);
      stage<3>(s, i, F, n, q, ih, Kn, Kq, usn, usq, csn, csq, rn, rq);
      stage<4>(s, i, F, n, q, ih, Kn, Kq, usn, usq, csn, csq, rn, rq);
      stage<5>(s, i, F, n, q, ih, Kn, Kq, usn, usq, csn, csq, rn, rq);
      const double nn = usn + rn, nq = usq + rq;       // u_new; the error estimate is K_6 = (rn, rq)
      // ---- error norm: same scales as trajectory.h (excess density for N, carrier density for Q) ----
      const double mx = fmax(fmax(fabs(n - c.n0), fabs(nn - c.n0)), ex_floor);
      const double mq = fmax(fabs(n), fabs(ax.P));
      const double en = rn * simt::rcp_approx(fma(opt.rtol, mx, opt.atol));
      const double eq = rq * (simt::rcp_approx(fma(opt.rtol, mq, opt.atol)) * Q_ERR_WEIGHT);
      double e2 = fma(en, en, eq * eq);
      double bad = (nn != nn || nq != nq) ? 1.0 : 0.0;
      block_sum2(s, red, e2, bad);
      const double err2 = e2 * inv_n;
      const bool nonfinite = bad > 0.0 || !(err2 == err2) || err2 > 1e300;
      // ---- controller (trajectory.h: Hairer's RODAS standard + Gustafsson predictive) ----
      const float e2f = nonfinite ? 1e20f : (float)fmax(fmin(err2, 1e30), 1e-30);
      float ifac = fmaxf(1.0f / 6.0f, fminf(5.0f, CTL_SAFETY * simt::ctl_powf(e2f, -0.125f)));
      h_new = h * (double)ifac;
      if (!nonfinite && err2 <= 1.0) {
        ++n_acc;
        if (!first) {
          float ifg = CTL_SAFETY * (float)(h * ih_acc) * simt::ctl_powf(e2f, -0.25f) * simt::ctl_powf(err2_old, 0.125f);
          ifg = fmaxf(1.0f / 6.0f, fminf(5.0f, ifg));
          ifac = fminf(ifac, ifg);
          h_new = h * (double)ifac;
        }
        first = false; ih_acc = ih; err2_old = fmaxf(1e-4f, e2f);
        if (last_rejected) h_new = fmin(h_new, h);
        last_rejected = false;
        t = final_step ? tend : t + h;
        n = nn; q = nq;
        break;
      }
      ++n_rej;
      last_rejected = true;
      h = nonfinite ? 0.1 * h : h_new;
    }
  }
  // measurement times against the step log, floor, likelihood sums: warp 0 (one-warp routines)
  __syncthreads();
  if (warp == 0) {
    emit_history(in, want_ll, in.hist, nh, em, false);
    emitter_finish(em, in, want_ll, mid);
    out.status = status | em.status; out.n_acc = n_acc; out.n_rej = n_rej;
  }
}

}  // namespace cta
}  // namespace trpl
#endif
