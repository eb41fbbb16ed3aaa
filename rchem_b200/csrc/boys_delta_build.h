// boys_delta_build.h -- host-side construction of the tables of boys_delta.h (setup only).
// Everything is evaluated in long double; the iteration counts come from the faithful loops
// (boys_reference_order), so the tables describe the reference's Fgamma exactly.
#pragma once
#include <cmath>
#include <vector>

#include "boys_delta.h"
#include "pair_build.h"

namespace rchem {

inline long double ld_boys_exact(int m, long double x) {
  if (x < 40.0L) {  // all-positive series
    long double term = 1.0L / (2 * m + 1), sum = term;
    for (int k = 1; k < 4000; ++k) {
      term *= 2.0L * x / (2 * m + 2 * k + 1);
      sum += term;
      if (term < 1e-23L * sum) break;
    }
    return expl(-x) * sum;
  }
  long double f = 0.5L * sqrtl(3.14159265358979323846264338327950288L / x);
  const long double ex = expl(-x);
  for (int k = 0; k < m; ++k) f = ((2 * k + 1) * f - ex) / (2.0L * x);
  return f;
}

// the reference's truncated functions themselves, n = its iteration count on this side of the step
inline long double ld_series_ref(int m, int n, long double x) {
  const long double a = m + 0.5L;
  long double term = 1.0L / a, sum = term;
  for (int k = 1; k <= n; ++k) {
    term *= x / (a + k);
    sum += term;
  }
  return 0.5L * expl(-x) * sum;
}
inline long double ld_fraction_ref(int m, int n, long double x) {
  const long double a = m + 0.5L;
  long double gam = sqrtl(3.14159265358979323846264338327950288L);
  for (int k = 0; k < m; ++k) gam *= (k + 0.5L);
  const long double b0 = x + 1.0L - a;
  long double A0 = 0.0L, A1 = 1.0L, B0 = 1.0L, B1 = b0;
  for (int j = 1; j <= n; ++j) {
    const long double aj = -(long double)j * (j - a), bj = b0 + 2.0L * j;
    const long double A2 = bj * A1 + aj * A0, B2 = bj * B1 + aj * B0;
    A0 = A1; A1 = A2; B0 = B1; B1 = B2;
  }
  return 0.5L * (gam * powl(x, -a) - expl(-x) * (A1 / B1));
}

// degree-(NC-1) interpolant at NC Chebyshev nodes of [xc-h, xc+h] as coefficients of (x - xc)^k
template <int NC, class Fn, class Out>
inline void cheb_fit(Fn f, long double xc, long double h, Out* out) {
  long double V[NC][NC + 1];
  for (int j = 0; j < NC; ++j) {
    const long double t = cosl((2 * j + 1) * 3.14159265358979323846264338327950288L / (2.0L * NC));
    long double p = 1.0L;
    for (int k = 0; k < NC; ++k) { V[j][k] = p; p *= t; }
    V[j][NC] = f(xc + h * t);
  }
  for (int c = 0; c < NC; ++c) {  // Gaussian elimination with partial pivoting
    int piv = c;
    for (int r = c + 1; r < NC; ++r) if (fabsl(V[r][c]) > fabsl(V[piv][c])) piv = r;
    for (int k = 0; k <= NC; ++k) std::swap(V[c][k], V[piv][k]);
    for (int r = 0; r < NC; ++r) {
      if (r == c) continue;
      const long double fct = V[r][c] / V[c][c];
      for (int k = c; k <= NC; ++k) V[r][k] -= fct * V[c][k];
    }
  }
  long double hk = 1.0L;
  for (int k = 0; k < NC; ++k) {
    out[k] = (Out)(V[k][NC] / V[k][k] / hk);
    hk *= h;
  }
}

// thr: [kRefMaxM+1][kDeltaCells] doubles (first x above the cell's step; 1e300 = no step);
// direct: [kRefMaxM+1][kDeltaCells][2][kDirectRowLen] doubles, the reference's Fgamma_m per
// (cell, side of the step).  Returns false if a cell holds more than one step.
inline bool build_boys_delta_tables(std::vector<double>* thr, std::vector<double>* direct) {
  thr->assign((size_t)(kRefMaxM + 1) * kDeltaCells, 1e300);
  direct->assign((size_t)(kRefMaxM + 1) * kDeltaCells * 2 * kDirectRowLen, 0.0);
  bool ok = true;
  for (int m = 0; m <= kRefMaxM; ++m) {
    const double a = m + 0.5;
    const double cut = ref_exact_from_order(m) + 1.0;  // the reference's own steps matter below cut
    for (int c = 0; c < kDeltaCells; ++c) {
      const double lo = c < kDeltaFineCells ? c / 64.0 : (c - 192) / 16.0;
      const double hi = c < kDeltaFineCells ? (c + 1) / 64.0 : (c - 191) / 16.0;
      const long double xc = 0.5L * ((long double)lo + hi), h = 0.5L * ((long double)hi - lo);
      double* rows = direct->data() + ((size_t)(m * kDeltaCells + c) * 2) * kDirectRowLen;
      if (c == 0 || lo >= cut) {  // converged F_m (cell 0 always takes the faithful loops)
        cheb_fit<kDirectRowLen>([&](long double x) { return ld_boys_exact(m, x); }, xc, h, rows);
        for (int k = 0; k < kDirectRowLen; ++k) rows[kDirectRowLen + k] = rows[k];
        continue;
      }
      const bool series = lo < a + 1.0;
      const int n_lo = ref_iterations(m, lo), n_hi = ref_iterations(m, std::nextafter(hi, 0.0));
      const bool step = n_lo != n_hi;
      if (step && n_hi != n_lo + (series ? 1 : -1)) ok = false;
      if (step) (*thr)[(size_t)m * kDeltaCells + c] = ref_step(m, lo, std::nextafter(hi, 0.0), n_lo);
      for (int side = 0; side < 2; ++side) {
        const int n = side ? n_hi : n_lo;
        double* out = rows + side * kDirectRowLen;
        if (series) cheb_fit<kDirectRowLen>([&](long double x) { return ld_series_ref(m, n, x); }, xc, h, out);
        else cheb_fit<kDirectRowLen>([&](long double x) { return ld_fraction_ref(m, n, x); }, xc, h, out);
      }
    }
  }
  return ok;
}

}  // namespace rchem
