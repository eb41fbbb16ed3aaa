// tho_device.cuh -- small device helpers shared by the Taketa-Huzinaga-O-ohata style kernels
// (prim_generic.cu: tier-1 coulomb_repulsion; one_electron.cu: S, T, V).
#pragma once
#include <cuda_runtime.h>

namespace rchem {

__device__ __forceinline__ double dfact(int n) {
  double r = 1.0;
  for (int k = 2; k <= n; ++k) r *= k;
  return r;
}
__device__ __forceinline__ double ipow(double x, int n) {  // n >= 0
  double r = 1.0;
  for (int k = 0; k < n; ++k) r *= x;
  return r;
}
__device__ __forceinline__ double binom(int a, int b) { return dfact(a) / (dfact(b) * dfact(a - b)); }
__device__ __forceinline__ double fact_ratio2(int a, int b) {
  return dfact(a) / dfact(b) / dfact(a - 2 * b);
}

// sum_t C(ia,s-t) C(ib,t) xpa^(ia-s+t) xpb^(ib-t)            (cints.c:291-298)
static __device__ double binomial_prefactor(int s, int ia, int ib, double xpa, double xpb) {
  double sum = 0.0;
  for (int t = 0; t <= s; ++t)
    if (s - ia <= t && t <= ib)
      sum += binom(ia, s - t) * binom(ib, t) * ipow(xpa, ia - s + t) * ipow(xpb, ib - t);
  return sum;
}

// F_m(x), converged (used for boys = exact): positive series below 36, asymptotic above
static __device__ double boys_converged(int m, double x) {
  if (x < 36.0) {
    double term = 1.0 / (2 * m + 1), sum = term;
    for (int k = 1; k < 400; ++k) {
      term *= 2.0 * x / (2 * m + 2 * k + 1);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    return exp(-x) * sum;
  }
  // F_0 = sqrt(pi/x)/2, then F_{k+1} = ((2k+1) F_k - e^-x) / 2x (the e^-x term matters at the
  // 1e-8 level for high orders near x = 36)
  double f = 0.88622692545275801365 * rsqrt(x);
  const double ex = x < 200.0 ? exp(-x) : 0.0;
  for (int k = 0; k < m; ++k) f = ((2 * k + 1) * f - ex) / (2.0 * x);
  return f;
}

}  // namespace rchem
