// proposals.h - host side of a Metropolis iteration: every chain's proposal and acceptance draw
// from one PCG64 stream, in the reference's serial order (include/metrotrpl_b200.h
// trpl_make_trial_moves).  Replaces, for whole ensembles, trial_move_generation.py:54-96
// (make_trial_move), :4-52 (approve_move) and the draw order of metropolis.py:118-127.
// Plain host C++.  Input and output are in the sampler's own scale (log10 of the parameters that
// are sampled in log space): the transcendental conversions stay in NumPy on both sides of the
// call, what happens here is one multiply-add per draw, so the proposals equal the Python mirror's
// bit for bit (tests/test_metropolis_batched.py, tests/test_golden_chains.py).  pow(10, x) is only
// used for the bounds test and for the mobility constraint.
#pragma once
#include <math.h>
#include <stdint.h>

namespace trpl_host {

// NumPy's PCG64: 128-bit LCG with the XSL-RR 128/64 output function
struct Pcg64 {
  unsigned __int128 state, inc;
  uint64_t next64() {
    const unsigned __int128 mult = ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
    state = state * mult + inc;
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned r = (unsigned)(hi >> 58);
    return (x >> r) | (x << ((64u - r) & 63u));
  }
  double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }   // Generator.random()
};

inline int make_trial_moves(int n_chains, int n_par, const double* cur, const double* moves,
                            const uint8_t* do_log, const uint8_t* active, const double* lo, const double* hi,
                            int idx_p0, int idx_n0, int idx_taun, int idx_taup, int hard_bounds, int max_tries,
                            const uint64_t pcg_state[2], const uint64_t pcg_inc[2], double* proposals, double* u,
                            int64_t* n_draws, int32_t* n_failed, uint32_t* fail_masks, int max_logged,
                            int idx_mun, int idx_mup, double ambi_lo, double ambi_hi, const double* ambi_u,
                            int n_ambi_u, int32_t* n_ambi_used, double* mu_arg) {
  Pcg64 g;
  g.state = ((unsigned __int128)pcg_state[0] << 64) | pcg_state[1];
  g.inc = ((unsigned __int128)pcg_inc[0] << 64) | pcg_inc[1];
  int64_t draws = 0;
  int ambi_used = 0;
  double cand[32];
  // The reference tests lo < 10^x < hi in linear units.  pow() is the most expensive thing in this
  // loop (hot chains need tens of attempts), so the test is made in log space first and pow() only
  // decides candidates within 1e-9 decades of a bound (its own error is 4e-17 decades).
  double loglo[32], loghi[32];
  for (int i = 0; i < n_par; ++i) {
    loglo[i] = lo[i] > 0 ? log10(lo[i]) : -HUGE_VAL;
    loghi[i] = hi[i] > 0 ? log10(hi[i]) : -HUGE_VAL;
  }
  const double LOG_MARGIN = 1e-9;
  const int tries = hard_bounds ? max_tries : 1;
  for (int m = 0; m < n_chains; ++m) {
    const double* logcur = cur + (size_t)m * n_par;       // already log-scaled where do_log
    const double* mv = moves + (size_t)m * n_par;
    int failed = 0;
    double mu_last = 0.0;
    for (int t = 0; t < tries; ++t) {
      for (int i = 0; i < n_par; ++i) cand[i] = logcur[i] + mv[i] * (2 * g.next_double() - 1);
      draws += n_par;
      if (idx_mun >= 0) {
        // do_mu_constraint (trial_move_generation.py:77-83): the ambipolar mobility is drawn from the
        // caller's pre-drawn np.random uniforms and mu_p follows from it and the proposed mu_n; the
        // same libm calls NumPy's scalar math makes, in the same order
        if (ambi_used >= n_ambi_u) return 2;
        const double new_ambi = ambi_lo + (ambi_hi - ambi_lo) * ambi_u[ambi_used++];
        // NumPy's log10 is its own implementation (7% of arguments differ from libm's in the last
        // bit): the argument goes back to the caller, which applies np.log10 to the final attempt's;
        // here the libm value only feeds the bounds test
        mu_last = pow(2 / new_ambi - 1 / pow(10.0, cand[idx_mun]), -1.0);
        cand[idx_mup] = log10(mu_last);
      }
      if (!hard_bounds) break;                       // no checks without hard bounds (one attempt)
      uint32_t mask = 0;
      for (int i = 0; i < n_par; ++i) {
        if (!active[i]) continue;
        if (do_log[i]) {
          const double x = cand[i];
          if (x > loglo[i] + LOG_MARGIN && x < loghi[i] - LOG_MARGIN) continue;                    // inside
          if (x < loglo[i] - LOG_MARGIN || x > loghi[i] + LOG_MARGIN) { mask |= (1u << i); continue; }  // outside
        }
        const double lin = do_log[i] ? pow(10.0, cand[i]) : cand[i];
        if (!(lo[i] < lin && lin < hi[i])) mask |= (1u << i);
      }
      if (idx_p0 >= 0 && idx_n0 >= 0 && !(cand[idx_p0] > cand[idx_n0])) mask |= (1u << 30);
      if (idx_taun >= 0 && idx_taup >= 0) {
        const double ltn = do_log[idx_taun] ? cand[idx_taun] : log10(cand[idx_taun]);
        const double ltp = do_log[idx_taup] ? cand[idx_taup] : log10(cand[idx_taup]);
        if (!(fabs(ltn - ltp) <= 2)) mask |= (1u << 31);
      }
      if (mask == 0) break;
      if (failed < max_logged) fail_masks[(size_t)m * max_logged + failed] = mask;
      ++failed;
    }
    n_failed[m] = failed;
    double* p = proposals + (size_t)m * n_par;
    for (int i = 0; i < n_par; ++i) p[i] = cand[i];          // the last attempt, admissible or not
    if (idx_mun >= 0 && mu_arg) mu_arg[m] = mu_last;
    u[m] = g.next_double();
    draws += 1;
  }
  *n_draws = draws;
  if (n_ambi_used) *n_ambi_used = ambi_used;
  return 0;
}

}  // namespace trpl_host
