// boys_delta.h -- reference-flavour Boys function from per-cell polynomial rows of Fgamma itself.
//
// libpyquante2's Fgamma (cints.c:302-373) truncates its series / continued fraction at a
// relative change of 3e-7, so it differs from the converged Boys function by up to 3e-7 F, and
// the number of terms n it takes is a step function of x.  BETWEEN two steps Fgamma_m is smooth:
//   series branch   (x <  m+3/2):  0.5 e^-x sum_{k<=n} x^k / (a (a+1) .. (a+k)),      a = m + 1/2
//   fraction branch (x >= m+3/2):  0.5 (Gamma(a) x^-a - e^-x h_n(x)),  h_n the n-th Lentz convergent
// so every (order, cell, side of the step) carries a degree-7 polynomial in double precision
// fitted to that function in long double at Chebyshev nodes (boys_delta_build.h; truncation
// < 1e-17 relative on cells this narrow).  Cells: 1/64 wide below x = 4, 1/16 wide up to 37; a
// cell holds at most one step (checked when the tables are built).  Past
// ref_exact_from_order(m) + 1 the reference differs from the converged F_m by < 2e-15 F and the
// rows hold F_m.
//
// Device cost per order: one 8-byte threshold and one 64-byte row, 7 FMAs -- no exp, no
// converged Boys value, no downward recursion.  Every cell has TWO row slots (below / above its
// step; identical without one), so the row address does not depend on the threshold: the
// thresholds of all orders and the below-step rows are requested together (one memory round trip
// for the whole evaluation; the above-step row is fetched only by the lanes that need it).
// The faithful loops (boys_reference_order) run only for x < 1/64, within 2^17 ulp of a step, or
// within 1e-10 of the branch switch x = m+3/2, where the result depends on the exact rounding
// of x.  (Round 1 evaluated Fgamma = F_exact - delta with a tabulated truncation error delta:
// an exp, float rows, x^(n+1) from shared squarings, ~4x the instructions per order; A/B in
// profiles/r02_ab_dynamic_chunks_direct_boys.txt.)
#pragma once
// (included from the middle of eri_core.h, after the Boys helpers it uses)

namespace rchem {

constexpr int kDeltaFineCells = 256;                    // x in [0, 4) at 1/64
constexpr int kDeltaCells = kDeltaFineCells + 33 * 16;  // + x in [4, 37) at 1/16  -> 784
constexpr int kDeltaNearUlps = 1 << 17;
constexpr int kDirectRowLen = 8;  // degree-7 polynomial in (x - cell centre), 64 bytes

// Device-side view: thr[m][cell] (double: first x above the cell's step, 1e300 without one),
// direct[m][cell][side][kDirectRowLen] (double).
struct BoysDeltaTables {
  const double* thr;
  const double* direct;
};

RCHEM_HD int delta_cell(double x) { return x < 4.0 ? (int)(x * 64.0) : 192 + (int)(x * 16.0); }
RCHEM_HD double delta_center(int cell) {
  return cell < kDeltaFineCells ? ((double)cell + 0.5) * (1.0 / 64.0)
                                : ((double)(cell - 192) + 0.5) * (1.0 / 16.0);
}

RCHEM_HD void direct_row_load(const double* __restrict__ row, double* __restrict__ c) {
#if defined(__CUDA_ARCH__)
  ldg256(row, c);  // (eri_core.h: one 32-byte load per lane)
  ldg256(row + 4, c + 4);
#else
  for (int j = 0; j < kDirectRowLen; ++j) c[j] = row[j];
#endif
}

// Reference-flavour F[0..L] at xa < 37.  xa is the (possibly few-ulp inexact) argument;
// exact_x() recomputes the bit-exact reference argument, only called on the slow path.
template <int L, class ExactX>
RCHEM_HD void boys_reference_direct(double xa, const BoysDeltaTables& tab, ExactX exact_x,
                                    double* __restrict__ F) {
  const int cell = delta_cell(xa);
  const double dx = xa - delta_center(cell);
  const long long xb = ref_bits(xa);
  long long tb[L + 1];
  bool near = cell == 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m <= L; ++m) {
    tb[m] = ref_bits(ref_tab(tab.thr + m * kDeltaCells + cell));
    near = near || (unsigned long long)(xb - tb[m] + kDeltaNearUlps) < 2ULL * kDeltaNearUlps ||
           fabs(xa - (m + 1.5)) < 1e-10;
  }
  if (near) {  // rare: the faithful loops for every order (they all see the same exact x)
    double x = exact_x();
    if (fabs(x) < 0.00000001) x = 0.00000001;  // cints.c:304
    const double rx = 1.0 / x, ex = exp(-x);
    double xpow = sqrt(rx);
    for (int m = 0; m <= L; ++m) {
      F[m] = boys_reference_order_slow(m, x, ex, xpow);
      xpow *= rx;
    }
    return;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m <= L; ++m) {
    const double* row = tab.direct + ((size_t)(m * kDeltaCells + cell) * 2) * kDirectRowLen;
    double c[kDirectRowLen];
    direct_row_load(row, c);                                      // below the step: no dependence on tb
    if (xb >= tb[m]) direct_row_load(row + kDirectRowLen, c);     // above it (few lanes)
    double p = c[7];
    p = fma(p, dx, c[6]);
    p = fma(p, dx, c[5]);
    p = fma(p, dx, c[4]);
    p = fma(p, dx, c[3]);
    p = fma(p, dx, c[2]);
    p = fma(p, dx, c[1]);
    F[m] = fma(p, dx, c[0]);
  }
}

}  // namespace rchem
