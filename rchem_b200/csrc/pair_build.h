// pair_build.h -- primitive-pair quantities of one shell pair (host, setup only).
//
// Every expression that feeds the REFERENCE Boys argument is formed exactly as
// coulomb_repulsion does (cints.c:85-96): product centres by product_center_1D
// (cints.c:391-394), 1./gamma by an IEEE division.  This file must be compiled without
// FMA contraction (-ffp-contract=off), like the reference's x86-64 build.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "basis_model.h"
#include "eri_core.h"

namespace rchem {

// Primitive pairs whose prefactor (contraction coefficients x norms x Gaussian-product
// exponential / zeta) is below this are dropped: each contributes < ~1e-15 to any integral,
// three orders below the 1e-12 parity tolerance.  Tight core primitives on different atoms
// are the typical case (exp(-alpha beta R^2 / zeta) underflows).
constexpr double kPrimPairEps = 1e-20;

// Primitive pairs of (A,B), A-primitive major: k = i*nprim(B) + j.  Every weight variant of the
// pair (fused sp shells: s|p part of A times s|p part of B) shares zeta, P and the Gaussian-
// product exponential; only the contraction weights differ.
inline void build_prim_pairs(const Shell& A, const Shell& B, std::vector<PrimPair>* out) {
  out->clear();
  const double dx = A.ctr[0] - B.ctr[0], dy = A.ctr[1] - B.ctr[1], dz = A.ctr[2] - B.ctr[2];
  const double rab2 = dx * dx + dy * dy + dz * dz;  // dist2, cints.c:282-285
  const int nva = nvariants(A.l), nvb = nvariants(B.l);
  for (size_t i = 0; i < A.exps.size(); ++i)
    for (size_t j = 0; j < B.exps.size(); ++j) {
      const double aa = A.exps[i], ab = B.exps[j];
      PrimPair pp;
      pp.zeta = aa + ab;
      pp.rzeta = 1. / pp.zeta;
      pp.Px = (aa * A.ctr[0] + ab * B.ctr[0]) / (aa + ab);
      pp.Py = (aa * A.ctr[1] + ab * B.ctr[1]) / (aa + ab);
      pp.Pz = (aa * A.ctr[2] + ab * B.ctr[2]) / (aa + ab);
      // exp(-aa*ab*rab2/gamma) as in cints.c:113, the 1/gamma of cints.c:112, and the
      // contraction coefficient x norm of both primitives
      const double e = std::exp(-aa * ab * rab2 / pp.zeta);
      pp.nv = nva * nvb;
      for (int v = 0; v < kMaxPairVariants; ++v) pp.w[v] = 0.0;
      for (int va = 0; va < nva; ++va)
        for (int vb = 0; vb < nvb; ++vb)
          pp.w[va * nvb + vb] = (va ? A.cn2[i] : A.cn[i]) * (vb ? B.cn2[j] : B.cn[j]) * e / pp.zeta;
      pp.fsc = 5.5683279968317078453 / std::sqrt(pp.zeta);               // pi^(3/2)
      pp.pfar = 5.5683279968317078453 * pp.w[0] / std::sqrt(pp.zeta);
      out->push_back(pp);
    }
}

// The significant primitive pairs of (A,B): those whose largest weight is >= eps (at least one
// is kept), ordered by increasing zeta.  Returns the count.
inline int build_significant_prim_pairs(const Shell& A, const Shell& B, double eps,
                                        std::vector<PrimPair>* out) {
  build_prim_pairs(A, B, out);
  std::stable_sort(out->begin(), out->end(), [](const PrimPair& x, const PrimPair& y) {
    return x.wmax() > y.wmax();
  });
  size_t n = out->size();
  while (n > 1 && !((*out)[n - 1].wmax() >= eps)) --n;
  out->resize(n);
  // most diffuse pair first: its Boys argument is (nearly) the smallest of the shell pair,
  // which is what the kernels' near/far scheduling looks at
  std::stable_sort(out->begin(), out->end(),
                   [](const PrimPair& x, const PrimPair& y) { return x.zeta < y.zeta; });
  return (int)n;
}

// Bounding sphere of the primitive-pair centres of a shell pair and its most diffuse exponent
// sum: every primitive quartet between two shell pairs has
//   x = rho |P - Q|^2 >= rho(zmin_b, zmin_k) * (|M_b - M_k| - rad_b - rad_k)^2,
// rho(z, e) = z e/(z + e) being increasing in both arguments.  The block kernel uses it to
// PROVE a shell quartet far-field (eri_kernel.cuh).
struct PairBound {
  double M[3];
  double rad;
  double zmin;
};
inline PairBound bound_prim_pairs(const std::vector<PrimPair>& pps) {
  PairBound pb{};
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  pb.zmin = 1e300;
  for (const PrimPair& pp : pps) {
    const double P[3] = {pp.Px, pp.Py, pp.Pz};
    for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], P[d]); hi[d] = std::max(hi[d], P[d]); }
    pb.zmin = std::min(pb.zmin, pp.zeta);
  }
  for (int d = 0; d < 3; ++d) pb.M[d] = 0.5 * (lo[d] + hi[d]);
  double r2 = 0.0;
  for (const PrimPair& pp : pps) {
    const double dx = pp.Px - pb.M[0], dy = pp.Py - pb.M[1], dz = pp.Pz - pb.M[2];
    r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
  }
  pb.rad = r2 > 0.0 ? std::sqrt(r2) * (1.0 + 1e-12) + 1e-300 : 0.0;
  return pb;
}

// exact Boys tables for boys_exact (eri_core.h): one table per total angular momentum
// L = 0..kBoysMaxL, rows x_i = i/kBoysPerUnit, row = {F_{L+k}(x_i)/k! (k=0..7), exp(-x_i), 0}.
inline void build_boys_tables(std::vector<double>* tables) {
  const int mtop = kBoysMaxL + 7;
  tables->assign((size_t)(kBoysMaxL + 1) * kBoysTableLen, 0.0);
  std::vector<long double> F(mtop + 1);
  for (int i = 0; i < kBoysRows; ++i) {
    const long double x = (long double)i / kBoysPerUnit;
    // top order by the all-positive series e^-x sum_k (2x)^k / ((2m+1)(2m+3)..(2m+2k+1)),
    // lower orders by the stable downward recursion
    long double term = 1.0L / (2 * mtop + 1), sum = term;
    for (int k = 1; k < 2000; ++k) {
      term *= 2.0L * x / (2 * mtop + 2 * k + 1);
      sum += term;
      if (term < 1e-24L * sum) break;
    }
    const long double ex = expl(-x);
    F[mtop] = ex * sum;
    for (int m = mtop; m > 0; --m) F[m - 1] = (2.0L * x * F[m] + ex) / (2 * m - 1);
    for (int L = 0; L <= kBoysMaxL; ++L) {
      double* row = tables->data() + (size_t)L * kBoysTableLen + (size_t)i * kBoysRowLen;
      long double fact = 1.0L;
      for (int k = 0; k < 8; ++k) {
        if (k > 0) fact *= k;
        row[k] = (double)(F[L + k] / fact);
      }
      row[8] = (double)ex;
      row[9] = 0.0;
    }
  }
}

// Iteration count of the reference's series / continued-fraction loop as a step function of
// x (for the tables of boys_delta.h), from the FAITHFUL loops (boys_reference_order).
inline int ref_iterations(int m, double x) {
  int n = 0;
  boys_reference_order(m, x, 0.0, 0.0, &n);
  return n;
}

// smallest double x in (lo, hi] with ref_iterations(m, x) == n_hi, given it is n_lo at lo
inline double ref_step(int m, double lo, double hi, int n_lo) {
  while (true) {
    const double mid = lo + 0.5 * (hi - lo);
    if (mid <= lo || mid >= hi) break;
    if (ref_iterations(m, mid) == n_lo) lo = mid; else hi = mid;
  }
  return hi;
}

}  // namespace rchem
