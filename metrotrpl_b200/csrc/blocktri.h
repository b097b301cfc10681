// blocktri.h - in-warp direct solver for the block-tridiagonal Rosenbrock matrix
//
//     W = 1/(gamma h) I - J ,   W_i = [ A_i  B_i  C_i ]  with 2x2 blocks, 32*NPL block rows.
//
// Two-level elimination, all of it inside one warp:
//   1. partition: lane l owns block rows l*NPL .. l*NPL+NPL-1.  Its first NPL-1 rows ("interior")
//      are eliminated sequentially in registers (block Thomas) against the two neighbouring
//      interface unknowns, which produces the spike blocks V, W (fill-in columns);
//   2. the 32 interface rows (one per lane, z_l = u at the lane's last node) form a reduced
//      block-tridiagonal system that is solved by parallel cyclic reduction: 5 levels, strides
//      1,2,4,8,16, neighbour rows exchanged through shared memory;
//   3. interiors are recovered as x = g - V z_{l-1} - W z_l.
// The factorisation is done once per step; each Rosenbrock stage then costs one `solve`.
// No pivoting: W is a shifted M-matrix-like operator and the prototype
// (tools/proto/proto_test3.py) shows <=3e-11 relative error against a pivoted dense solve over the
// whole prior box, step sizes 1e-6..1e3 ns.
//
// Storage: the interior factors, spikes and interface blocks (FacSlots: two runs of consecutive
// pairs) and the PCR multipliers (PmRun: one run per level) are lane-private and live in the
// warp's tensor-memory slice (or shared memory, whatever `FM`/`PM` are); only the PCR lane exchange
// goes through shared memory.
#pragma once
#include "simt.h"
#include "model.h"

// Lane exchange of the solves' PCR levels: shuffles (4 x 64 bit per level) or two shared-memory
// pairs.  Measured on B200 in DESIGN.md section 5.
#ifndef TRPL_SOLVE_PCR_SHFL
#define TRPL_SOLVE_PCR_SHFL 0
#endif

namespace trpl {
using namespace simt;

TRPL_FN Blk blk_inv(const Blk& m) {
  const real idet = rcp(fmadd(m.a00, m.a11, -(m.a01 * m.a10)));
  Blk r;
  r.a00 = m.a11 * idet; r.a01 = -(m.a01 * idet);
  r.a10 = -(m.a10 * idet); r.a11 = m.a00 * idet;
  return r;
}
TRPL_FN Blk blk_mul(const Blk& x, const Blk& y) {
  Blk r;
  r.a00 = fmadd(x.a00, y.a00, x.a01 * y.a10); r.a01 = fmadd(x.a00, y.a01, x.a01 * y.a11);
  r.a10 = fmadd(x.a10, y.a00, x.a11 * y.a10); r.a11 = fmadd(x.a10, y.a01, x.a11 * y.a11);
  return r;
}
// -(x*y), signs folded into the operands
TRPL_FN Blk blk_mul_neg(const Blk& x, const Blk& y) {
  Blk r;
  r.a00 = fmadd(-x.a00, y.a00, -(x.a01 * y.a10)); r.a01 = fmadd(-x.a00, y.a01, -(x.a01 * y.a11));
  r.a10 = fmadd(-x.a10, y.a00, -(x.a11 * y.a10)); r.a11 = fmadd(-x.a10, y.a01, -(x.a11 * y.a11));
  return r;
}
// z + x*y
TRPL_FN Blk blk_fma(const Blk& x, const Blk& y, const Blk& z) {
  Blk r;
  r.a00 = fmadd(x.a01, y.a10, fmadd(x.a00, y.a00, z.a00)); r.a01 = fmadd(x.a01, y.a11, fmadd(x.a00, y.a01, z.a01));
  r.a10 = fmadd(x.a11, y.a10, fmadd(x.a10, y.a00, z.a10)); r.a11 = fmadd(x.a11, y.a11, fmadd(x.a10, y.a01, z.a11));
  return r;
}
TRPL_FN Blk blk_sub(const Blk& x, const Blk& y) {
  Blk r; r.a00 = x.a00 - y.a00; r.a01 = x.a01 - y.a01; r.a10 = x.a10 - y.a10; r.a11 = x.a11 - y.a11; return r;
}
TRPL_FN Blk blk_add(const Blk& x, const Blk& y) {
  Blk r; r.a00 = x.a00 + y.a00; r.a01 = x.a01 + y.a01; r.a10 = x.a10 + y.a10; r.a11 = x.a11 + y.a11; return r;
}
TRPL_FN Blk blk_neg(const Blk& x) { Blk r; r.a00 = -x.a00; r.a01 = -x.a01; r.a10 = -x.a10; r.a11 = -x.a11; return r; }
TRPL_FN Blk blk_zero() { Blk r; r.a00 = splat(0.0); r.a01 = splat(0.0); r.a10 = splat(0.0); r.a11 = splat(0.0); return r; }
TRPL_FN Blk blk_sel(mask m, const Blk& x, const Blk& y) {
  Blk r; r.a00 = sel(m, x.a00, y.a00); r.a01 = sel(m, x.a01, y.a01); r.a10 = sel(m, x.a10, y.a10); r.a11 = sel(m, x.a11, y.a11); return r;
}
struct V2 { real x, y; };
TRPL_FN V2 blk_mv(const Blk& m, const V2& v) { V2 r; r.x = fmadd(m.a00, v.x, m.a01 * v.y); r.y = fmadd(m.a10, v.x, m.a11 * v.y); return r; }
// r - M v
TRPL_FN V2 sub_mv(const V2& r, const Blk& m, const V2& v) {
  V2 o; o.x = fmadd(-m.a01, v.y, fmadd(-m.a00, v.x, r.x)); o.y = fmadd(-m.a11, v.y, fmadd(-m.a10, v.x, r.y)); return o;
}
// r + M v
TRPL_FN V2 add_mv(const V2& r, const Blk& m, const V2& v) {
  V2 o; o.x = fmadd(m.a01, v.y, fmadd(m.a00, v.x, r.x)); o.y = fmadd(m.a11, v.y, fmadd(m.a10, v.x, r.y)); return o;
}

// Shared-memory map of one factorisation, in PAIRS (16 bytes per lane each).  A 2x2 block is two
// pairs {a00,a01},{a10,a11}; triangular blocks (super-diagonal blocks have a01 == 0: dN_i/dt does
// not see Q_{i+2}; sub-diagonal blocks have a10 == 0: dQ_{i+1}/dt does not see N_{i-1}) are stored
// as full blocks so that every access stays a 128-bit one.
template <int NPL>
struct FacSlots {
  static constexpr int NI = NPL - 1;                 // interior rows per lane
  // first run: what the forward/backward interior sweep and the reduced right-hand side read
  static constexpr int DINV = 0;                     // NI blocks
  static constexpr int LMUL = DINV + 2 * NI;         // NI-1 blocks (rows 1..NI-1)
  static constexpr int CSUP = LMUL + 2 * (NI > 0 ? NI - 1 : 0);   // NI blocks
  static constexpr int AZ = CSUP + 2 * NI;           // 1 block
  static constexpr int CZ = AZ + 2;                  // 1 block
  static constexpr int RUN1 = CZ + 2;                // pairs in the first run
  // second run: the spikes, read after the reduced solve
  static constexpr int VSPK = RUN1;                  // NI blocks
  static constexpr int WSPK = VSPK + 2 * NI;         // NI blocks
  static constexpr int RUN2 = 4 * NI;
  static constexpr int COUNT = RUN1 + RUN2;          // pairs
};

// a block of the lane exchange (shared memory)
TRPL_FN void st_blk(LaneMem& sm, int p, const Blk& b) { sm.st2(p, b.a00, b.a01); sm.st2(p + 1, b.a10, b.a11); }
// a block inside a run of factor values staged in registers (pair index p of the run)
TRPL_FN void put_blk(real* v, int p, const Blk& b) { v[2 * p] = b.a00; v[2 * p + 1] = b.a01; v[2 * p + 2] = b.a10; v[2 * p + 3] = b.a11; }
TRPL_FN Blk get_blk(const real* v, int p) { Blk b; b.a00 = v[2 * p]; b.a01 = v[2 * p + 1]; b.a10 = v[2 * p + 2]; b.a11 = v[2 * p + 3]; return b; }
TRPL_FN Blk ld_blk_from(const LaneMem& sm, int p, const ivec& src) {
  Blk b; sm.ld2_from(p, src, b.a00, b.a01); sm.ld2_from(p + 1, src, b.a10, b.a11); return b;
}
// r - M v for M lower triangular (a01 == 0) / upper triangular (a10 == 0)
TRPL_FN V2 sub_mv_lower(const V2& r, const Blk& m, const V2& v) {
  V2 o; o.x = fmadd(-m.a00, v.x, r.x); o.y = fmadd(-m.a11, v.y, fmadd(-m.a10, v.x, r.y)); return o;
}
TRPL_FN V2 sub_mv_upper(const V2& r, const Blk& m, const V2& v) {
  V2 o; o.x = fmadd(-m.a01, v.y, fmadd(-m.a00, v.x, r.x)); o.y = fmadd(-m.a11, v.y, r.y); return o;
}

// PCR multipliers of the reduced system (alpha_k, gamma_k of the five levels, final inverse): 44
// values per lane.  Register-resident (PmRegs), or kept in the trajectory's memories and fetched
// level by level (PmRun: 4 pairs per level + 2 for the inverse, first TM_PAIRS pairs in `tm`, the
// rest in `sm`), which takes 88 registers per thread out of the integration loop.
struct PmRegs {
  Blk al[LOG2_LANES], ga[LOG2_LANES];
  Blk binv_;
  TRPL_FN void put(int k, const Blk& a, const Blk& g) { al[k] = a; ga[k] = g; }
  TRPL_FN void put_binv(const Blk& b) { binv_ = b; }
  TRPL_FN void done_storing() const {}
  struct Level { Blk al, ga; };
  TRPL_FN Level fetch(int k) const { Level l; l.al = al[k]; l.ga = ga[k]; return l; }
  TRPL_FN void ready(Level&) const {}
  struct Inverse {};
  TRPL_FN Inverse fetch_binv() const { return Inverse(); }
  TRPL_FN Blk binv(const Inverse&) const { return binv_; }
};
template <class TM, class SM, int TM_BASE, int SM_BASE, int TM_PAIRS>
struct PmRun {
  TM& tm;
  SM& sm;
  static constexpr int BINV = 4 * LOG2_LANES;        // pair index of the inverse
  TRPL_FN void put(int k, const Blk& a, const Blk& g) {
    real v[8];
    put_blk(v, 0, a); put_blk(v, 2, g);
    if (4 * k < TM_PAIRS) mem_st_pairs<4>(tm, TM_BASE + 4 * k, v);
    else mem_st_pairs<4>(sm, SM_BASE + 4 * k - TM_PAIRS, v);
  }
  TRPL_FN void put_binv(const Blk& b) {
    real v[4];
    put_blk(v, 0, b);
    if (BINV < TM_PAIRS) mem_st_pairs<2>(tm, TM_BASE + BINV, v);
    else mem_st_pairs<2>(sm, SM_BASE + BINV - TM_PAIRS, v);
  }
  TRPL_FN void done_storing() const { if (TM_PAIRS > 0) mem_wait_st(tm); }
  struct Level { real v[8]; Blk al, ga; bool in_tm; };
  // issue the loads of level k (asynchronous when it lives in tensor memory)
  TRPL_FN Level fetch(int k) const {
    Level l;
    l.in_tm = 4 * k < TM_PAIRS;
    if (4 * k < TM_PAIRS) mem_ld_pairs<4>(tm, TM_BASE + 4 * k, l.v);
    else mem_ld_pairs<4>(sm, SM_BASE + 4 * k - TM_PAIRS, l.v);
    return l;
  }
  TRPL_FN void ready(Level& l) const {
    if (l.in_tm) mem_wait_ld(tm);
    l.al = get_blk(l.v, 0); l.ga = get_blk(l.v, 2);
  }
  struct Inverse { real v[4]; };
  TRPL_FN Inverse fetch_binv() const {
    Inverse i;
    if (BINV < TM_PAIRS) mem_ld_pairs<2>(tm, TM_BASE + BINV, i.v);
    else mem_ld_pairs<2>(sm, SM_BASE + BINV - TM_PAIRS, i.v);
    return i;
  }
  TRPL_FN Blk binv(const Inverse& i) const {
    if (BINV < TM_PAIRS) mem_wait_ld(tm);
    return get_blk(i.v, 0);
  }
};

// Factorise W given by (A, B, C) blocks of this lane's rows.  `fm`/`base`: lane-private factor
// storage and its first pair; `sm`/`xch`: 12 shared-memory scratch pairs for the lane exchange.
template <int NPL, class FM, class PM>
TRPL_FN void bt_factor(const Blk (&A)[NPL], const Blk (&B)[NPL], const Blk (&C)[NPL], FM& fm,
                       int base, LaneMem& sm, int xch, PM& pf) {
  typedef FacSlots<NPL> S;
  constexpr int NI = NPL - 1;
  Blk ra, rb, rc;
  if constexpr (NI > 0) {
    Blk dinv[NI > 0 ? NI : 1], lm[NI > 0 ? NI : 1];
    dinv[0] = blk_inv(B[0]);
    TRPL_UNROLL for (int j = 1; j < NI; ++j) {
      lm[j] = blk_mul(A[j], dinv[j - 1]);
      dinv[j] = blk_inv(blk_sub(B[j], blk_mul(lm[j], C[j - 1])));
    }
    // spikes: T V = [A_0; 0; ...], T W = [...; 0; C_{NI-1}]
    Blk v[NI > 0 ? NI : 1], w[NI > 0 ? NI : 1];
    v[0] = A[0];
    TRPL_UNROLL for (int j = 1; j < NI; ++j) v[j] = blk_mul_neg(lm[j], v[j - 1]);
    v[NI - 1] = blk_mul(dinv[NI - 1], v[NI - 1]);
    w[NI - 1] = blk_mul(dinv[NI - 1], C[NI - 1]);
    TRPL_UNROLL for (int j = NI - 2; j >= 0; --j) {
      v[j] = blk_mul(dinv[j], blk_sub(v[j], blk_mul(C[j], v[j + 1])));
      w[j] = blk_mul_neg(dinv[j], blk_mul(C[j], w[j + 1]));
    }
    // the factor blocks leave in two runs of consecutive pairs (widest stores the memory has)
    {
      real f1[2 * S::RUN1];
      TRPL_UNROLL for (int j = 0; j < NI; ++j) {
        put_blk(f1, S::DINV + 2 * j, dinv[j]);
        if (j > 0) put_blk(f1, S::LMUL + 2 * (j - 1), lm[j]);
        put_blk(f1, S::CSUP + 2 * j, C[j]);
      }
      put_blk(f1, S::AZ, A[NPL - 1]);
      put_blk(f1, S::CZ, C[NPL - 1]);
      mem_st_pairs<S::RUN1>(fm, base, f1);
      real f2[2 * S::RUN2];
      TRPL_UNROLL for (int j = 0; j < NI; ++j) {
        put_blk(f2, 2 * j, v[j]);
        put_blk(f2, 2 * NI + 2 * j, w[j]);
      }
      mem_st_pairs<S::RUN2>(fm, base + S::RUN1, f2);
    }
    // reduced (interface) row of this lane
    // first spikes of the next lane (one exchange of eight values)
    Blk v0n, w0n;
    {
      const real mine[8] = {v[0].a00, v[0].a01, v[0].a10, v[0].a11, w[0].a00, w[0].a01, w[0].a10, w[0].a11};
      real next[8];
      nbr_down(mine, next);
      v0n.a00 = next[0]; v0n.a01 = next[1]; v0n.a10 = next[2]; v0n.a11 = next[3];
      w0n.a00 = next[4]; w0n.a01 = next[5]; w0n.a10 = next[6]; w0n.a11 = next[7];
    }
    ra = blk_mul_neg(A[NPL - 1], v[NI - 1]);
    rb = blk_sub(blk_sub(B[NPL - 1], blk_mul(A[NPL - 1], w[NI - 1])), blk_mul(C[NPL - 1], v0n));
    rc = blk_mul_neg(C[NPL - 1], w0n);
  } else {
    ra = A[0]; rb = B[0]; rc = C[0];
  }
  // Parallel cyclic reduction on (ra, rb, rc) across the 32 lanes.  Neighbour rows travel through
  // 2 x 6 scratch pairs (double buffered, one warp_sync per level) instead of 24 64-bit shuffles.
  // No masking at the ends: ra is an exact zero block on lanes < stride and rc on lanes >= LANES -
  // stride (they are products with the zero sub/super-diagonal of the first/last row), and
  // out-of-range reads are clamped to the lane's own (finite) row.
  TRPL_UNROLL for (int k = 0; k < LOG2_LANES; ++k) {
    const int s = 1 << k;
    const int xb = xch + 6 * (k & 1);
    const Blk bi = blk_inv(rb);
    st_blk(sm, xb, bi); st_blk(sm, xb + 2, ra); st_blk(sm, xb + 4, rc);
    warp_sync();
    const ivec up = lane_minus(s), dn = lane_plus(s);
    const Blk bi_up = ld_blk_from(sm, xb, up), ra_up = ld_blk_from(sm, xb + 2, up), rc_up = ld_blk_from(sm, xb + 4, up);
    const Blk bi_dn = ld_blk_from(sm, xb, dn), ra_dn = ld_blk_from(sm, xb + 2, dn), rc_dn = ld_blk_from(sm, xb + 4, dn);
    const Blk alpha = blk_mul_neg(ra, bi_up);     // -(ra * bi_up)
    const Blk gamma = blk_mul_neg(rc, bi_dn);
    rb = blk_fma(gamma, ra_dn, blk_fma(alpha, rc_up, rb));
    ra = blk_mul(alpha, ra_up);
    rc = blk_mul(gamma, rc_dn);
    pf.put(k, alpha, gamma);
  }
  pf.put_binv(blk_inv(rb));
  pf.done_storing();
  mem_wait_st(fm);          // one fence per step: everything the six solves read has landed
  warp_sync();
}

// Solve W x = r in place.  r[j] / x[j] are this lane's NPL block rows; `xch` = 2 scratch pairs.
template <int NPL, class FM, class PM>
TRPL_FN void bt_solve(V2 (&r)[NPL], const FM& fm, int base, LaneMem& sm, int xch, const PM& pf) {
  typedef FacSlots<NPL> S;
  constexpr int NI = NPL - 1;
  V2 g[NI > 0 ? NI : 1];
  V2 rr;
  if constexpr (NI > 0) {
    // every factor block of the interior sweep is fetched up front (independent of r)
    Blk dinv[NI], csup[NI], lm[NI > 1 ? NI - 1 : 1];
    Blk az, cz;
    {
      real f1[2 * S::RUN1];
      mem_ld_pairs<S::RUN1>(fm, base, f1);          // stores were fenced once, at the end of bt_factor
      mem_wait_ld(fm);
      TRPL_UNROLL for (int j = 0; j < NI; ++j) {
        dinv[j] = get_blk(f1, S::DINV + 2 * j);
        csup[j] = get_blk(f1, S::CSUP + 2 * j);
        if (j > 0) lm[j - 1] = get_blk(f1, S::LMUL + 2 * (j - 1));
      }
      az = get_blk(f1, S::AZ); cz = get_blk(f1, S::CZ);
    }
    g[0] = r[0];
    TRPL_UNROLL for (int j = 1; j < NI; ++j) g[j] = sub_mv(r[j], lm[j - 1], g[j - 1]);
    g[NI - 1] = blk_mv(dinv[NI - 1], g[NI - 1]);
    TRPL_UNROLL for (int j = NI - 2; j >= 0; --j) g[j] = blk_mv(dinv[j], sub_mv_lower(g[j], csup[j], g[j + 1]));
    V2 g0n;
    { const real mine[2] = {g[0].x, g[0].y}; real next[2]; nbr_down(mine, next); g0n.x = next[0]; g0n.y = next[1]; }
    rr = sub_mv_upper(r[NPL - 1], az, g[NI - 1]);
    rr = sub_mv_lower(rr, cz, g0n);                          // cz is zero on the last lane
  } else {
    rr = r[0];
  }
  // Everything the tail of the solve reads (spikes, final inverse) is requested two levels before
  // it is needed, so that its latency hides behind the last lane exchanges.
  real f2[2 * (S::RUN2 > 0 ? S::RUN2 : 1)];
  typename PM::Inverse inv;
  TRPL_UNROLL for (int k = 0; k < LOG2_LANES; ++k) {
    const int s = 1 << k;
    V2 up, dn;
    typename PM::Level lv = pf.fetch(k);                    // in flight during the lane exchange
    if (k == LOG2_LANES - 2) {
      if constexpr (NI > 0) mem_ld_pairs<S::RUN2>(fm, base + S::RUN1, f2);
      inv = pf.fetch_binv();
    }
#if TRPL_SOLVE_PCR_SHFL
    up.x = shfl_up(rr.x, s); up.y = shfl_up(rr.y, s);       // own row where there is no neighbour
    dn.x = shfl_down(rr.x, s); dn.y = shfl_down(rr.y, s);
#else
    const int xb = xch + (k & 1);
    sm.st2(xb, rr.x, rr.y);
    warp_sync();
    sm.ld2_from(xb, lane_minus(s), up.x, up.y);
    sm.ld2_from(xb, lane_plus(s), dn.x, dn.y);
#endif
    // two independent chains (multipliers are exact zeros where there is no neighbour)
    pf.ready(lv);
    const V2 lo = blk_mv(lv.al, up), hi = add_mv(rr, lv.ga, dn);
    rr.x = lo.x + hi.x; rr.y = lo.y + hi.y;
  }
  const V2 z = blk_mv(pf.binv(inv), rr);
  r[NPL - 1] = z;
  if constexpr (NI > 0) {
    Blk vs[NI], ws[NI];
    mem_wait_ld(fm);
    TRPL_UNROLL for (int j = 0; j < NI; ++j) { vs[j] = get_blk(f2, 2 * j); ws[j] = get_blk(f2, 2 * NI + 2 * j); }
    V2 zl;                                                   // lane 0: V is zero there
    { const real mine[2] = {z.x, z.y}; real prev[2]; nbr_up(mine, prev); zl.x = prev[0]; zl.y = prev[1]; }
    TRPL_UNROLL for (int j = 0; j < NI; ++j) r[j] = sub_mv(sub_mv(g[j], vs[j], zl), ws[j], z);
  }
}

}  // namespace trpl
