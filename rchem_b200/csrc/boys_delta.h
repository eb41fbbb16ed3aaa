// boys_delta.h -- reference-flavour Boys function as "converged Boys minus a tabulated
// correction", branch-free in the common path.
//
// libpyquante2's Fgamma (cints.c:302-373) truncates its series / continued fraction at a
// relative change of 3e-7, so  Fgamma_m(x) = F_m(x) - delta_m(x)  with |delta| <= 3e-7 F.
// Between two x values at which the reference's iteration count n changes, delta is smooth:
//   series branch   (x <  m+3/2):  delta = 0.5 e^-x x^(n+1) G_n(x),
//                                  G_n(x) = sum_{j>=0} x^j / ((a)(a+1)..(a+n+1+j)),  a = m+1/2
//   fraction branch (x >= m+3/2):  delta = F_m(x) - 0.5 (Gamma(a) x^-a - e^-x h_n(x))
// and past ref_exact_from(L) (eri_core.h) it is below 2e-15 F.  Because delta only has to be
// known to ~3e-8 RELATIVE for a 1e-14 relative result, G_n (series) and delta (fraction) are
// stored per cell as degree-5 polynomials with float coefficients (fitted in long double at
// Chebyshev nodes by the host).  Cells: 1/64 wide below x = 4, 1/16 wide up to 36; a cell with
// an iteration-count step has two rows (below / above the step).
//
// Device cost per order: one 8-byte and one 24-byte table read, a 5-FMA polynomial and (series
// only) x^(n+1) from four shared squarings.  The faithful loops (boys_reference_order) run only
// for x < 1/64, within 2^17 ulp of a step, or within 1e-10 of the branch switch x = m+3/2,
// where the result depends on the exact rounding of x.
#pragma once
// (included from the middle of eri_core.h, after the Boys helpers it uses)

namespace rchem {

constexpr int kDeltaFineCells = 256;                    // x in [0, 4) at 1/64
constexpr int kDeltaCells = kDeltaFineCells + 33 * 16;  // + x in [4, 37) at 1/16  -> 784
constexpr int kDeltaRowLen = 8;                         // 6 coefficients + 2 pad (32 bytes)
constexpr int kDeltaMaxRows = kDeltaCells + 128;        // cells + steps, per order
constexpr int kDeltaNearUlps = 1 << 17;

// Device-side view: thr[m][cell] (double, low 12 bits = n_lo | steps_before << 5),
// rows[m][row][8] (float).
struct BoysDeltaTables {
  const double* thr;
  const float* rows;
  const double* direct;  // boys_reference_direct: rows[m][row][kDirectRowLen] (double)
};
constexpr int kDirectRowLen = 8;  // degree-7 polynomial in (x - cell centre), 64 bytes

RCHEM_HD int delta_cell(double x) { return x < 4.0 ? (int)(x * 64.0) : 192 + (int)(x * 16.0); }
RCHEM_HD double delta_center(int cell) {
  return cell < kDeltaFineCells ? ((double)cell + 0.5) * (1.0 / 64.0)
                                : ((double)(cell - 192) + 0.5) * (1.0 / 16.0);
}

// F[0..L] holds the converged Boys values at xa on entry and the reference-flavour values on
// exit.  xa is the (possibly few-ulp inexact) argument; exact_x() recomputes the bit-exact
// reference argument, only called on the slow path.
template <int L, class ExactX>
RCHEM_HD void boys_reference_from_exact(double xa, double ex, const BoysDeltaTables& tab,
                                        ExactX exact_x, double* __restrict__ F) {
  const int cell = delta_cell(xa);
  const double dx = xa - delta_center(cell);
  const long long xb = ref_bits(xa);
  const double x2 = xa * xa, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
  const double hex = 0.5 * ex;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m <= L; ++m) {
    if (xa >= ref_exact_from_order(m) + 0.5) continue;  // delta < 2e-15 F for this order
    const long long tb = ref_bits(ref_tab(tab.thr + m * kDeltaCells + cell));
    const bool near = (unsigned long long)(xb - tb + kDeltaNearUlps) < 2ULL * kDeltaNearUlps ||
                      cell == 0 || fabs(xa - (m + 1.5)) < 1e-10;
    if (near) {
      double x = exact_x();
      if (fabs(x) < 0.00000001) x = 0.00000001;  // cints.c:304
      const double rx = 1.0 / x;
      double xpow = sqrt(rx);
      for (int k = 0; k < m; ++k) xpow *= rx;
      F[m] = boys_reference_order_slow(m, x, exp(-x), xpow);
      continue;
    }
    int n = (int)(tb & 31);
    const int before = (int)((tb >> 5) & 127);
    const bool series = xa < m + 1.5;
    const int side = xb >= tb ? 1 : 0;
    n += side ? (series ? 1 : -1) : 0;
    const float* row = tab.rows + ((size_t)m * kDeltaMaxRows + cell + before + side) * kDeltaRowLen;
#if defined(__CUDA_ARCH__)
    const float4 lo = __ldg(reinterpret_cast<const float4*>(row));
    const float2 hi = __ldg(reinterpret_cast<const float2*>(row + 4));
    const double c0 = lo.x, c1 = lo.y, c2 = lo.z, c3 = lo.w, c4 = hi.x, c5 = hi.y;
#else
    const double c0 = row[0], c1 = row[1], c2 = row[2], c3 = row[3], c4 = row[4], c5 = row[5];
#endif
    double p = fma(c5, dx, c4);
    p = fma(p, dx, c3);
    p = fma(p, dx, c2);
    p = fma(p, dx, c1);
    p = fma(p, dx, c0);
    if (series) {
      const int e = n + 1;  // x^(n+1), n+1 <= 31
      double w = (e & 1) ? xa : 1.0;
      if (e & 2) w *= x2;
      if (e & 4) w *= x4;
      if (e & 8) w *= x8;
      if (e & 16) w *= x16;
      p *= hex * w;
    }
    F[m] -= p;
  }
}

// ---------------------------------------------------------------------------------------
// DIRECT form (the one the kernels use, RCHEM_BOYS_DIRECT): between two iteration-count steps the
// reference's Fgamma_m itself is smooth --
//   series branch:   0.5 e^-x sum_{k<=n} x^k / (a (a+1) .. (a+k)),      a = m + 1/2
//   fraction branch: 0.5 (Gamma(a) x^-a - e^-x h_n(x)),  h_n the n-th Lentz convergent
// -- so every (order, cell, side of the step) carries a degree-7 polynomial in double precision
// fitted to that function in long double (truncation < 1e-17 relative on cells this narrow).
// One 8-byte threshold read (shared with the delta form: same cells, same row numbering), one
// 64-byte row and 7 FMAs per order; no exp, no converged Boys value, no downward recursion.
// Past ref_exact_from_order(m) + 1 the rows hold the converged F_m (the reference differs from it
// by < 2e-15 F there).  Valid for xa < 37; the faithful loops run in the same rare cases.
// ---------------------------------------------------------------------------------------
template <int L, class ExactX>
RCHEM_HD void boys_reference_direct(double xa, const BoysDeltaTables& tab, ExactX exact_x,
                                    double* __restrict__ F) {
  const int cell = delta_cell(xa);
  const double dx = xa - delta_center(cell);
  const long long xb = ref_bits(xa);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m <= L; ++m) {
    const long long tb = ref_bits(ref_tab(tab.thr + m * kDeltaCells + cell));
    const bool near = (unsigned long long)(xb - tb + kDeltaNearUlps) < 2ULL * kDeltaNearUlps ||
                      cell == 0 || fabs(xa - (m + 1.5)) < 1e-10;
    if (near) {
      double x = exact_x();
      if (fabs(x) < 0.00000001) x = 0.00000001;  // cints.c:304
      const double rx = 1.0 / x;
      double xpow = sqrt(rx);
      for (int k = 0; k < m; ++k) xpow *= rx;
      F[m] = boys_reference_order_slow(m, x, exp(-x), xpow);
      continue;
    }
    const int before = (int)((tb >> 5) & 127);
    const int side = xb >= tb ? 1 : 0;
    const double* row = tab.direct + ((size_t)m * kDeltaMaxRows + cell + before + side) * kDirectRowLen;
    double c[kDirectRowLen];
#if defined(__CUDA_ARCH__)
    const double2* r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int j = 0; j < kDirectRowLen / 2; ++j) {
      const double2 v = __ldg(r2 + j);
      c[2 * j] = v.x;
      c[2 * j + 1] = v.y;
    }
#else
    for (int j = 0; j < kDirectRowLen; ++j) c[j] = row[j];
#endif
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
