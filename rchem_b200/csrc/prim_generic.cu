// prim_generic.cu -- tier-1 boundary: the libpyquante2 `coulomb_repulsion` symbol
// (cints.h:23-30) evaluated on the GPU for arbitrary Cartesian powers (l+m+n <= 3 per centre).
//
// This is the Taketa-Huzinaga-O-ohata explicit-sum formula the reference executes
// (cints.c:72-149), one primitive integral per thread.  It exists so that the reference's
// own unit vectors (tho66.rs:288-321, os86.rs:882-937) can be pushed through the C ABI, and
// as an independent on-device cross-check of the class kernels; the production path is the
// per-class HGP kernels (eri_kernel.cuh).
#include <cuda_runtime.h>

#include <cmath>
#include <limits>
#include <string>

#include "../../include/rchem_eri.h"
#include "eri_core.h"
#include "tho_device.cuh"

namespace rchem {

constexpr int kMaxAxis = 12;  // 12! is the last factorial the reference's int fact() can hold

// fB (cints.c:33-40); (4g)^(r-i) has a non-positive exponent
__device__ double f_b(int i, int l1, int l2, double p, double a, double b, int r, double g) {
  return binomial_prefactor(i, l1, l2, p - a, p - b) * (fact_ratio2(i, r) / ipow(4.0 * g, i - r));
}

// one axis of THO eq. 2.22 (cints.c:117-149)
__device__ void b_axis(double* B, int l1, int l2, int l3, int l4, double p, double a, double b,
                       double q, double c, double d, double g1, double g2, double delta) {
  const int n = l1 + l2 + l3 + l4 + 1;
  for (int i = 0; i < n; ++i) B[i] = 0.0;
  for (int i1 = 0; i1 <= l1 + l2; ++i1)
    for (int i2 = 0; i2 <= l3 + l4; ++i2)
      for (int r1 = 0; r1 <= i1 / 2; ++r1)
        for (int r2 = 0; r2 <= i2 / 2; ++r2) {
          const int top = i1 + i2 - 2 * (r1 + r2);
          const double fb = f_b(i1, l1, l2, p, a, b, r1, g1) * ((i2 & 1) ? -1.0 : 1.0) *
                            f_b(i2, l3, l4, q, c, d, r2, g2);
          for (int u = 0; u <= top / 2; ++u)
            B[top - u] += fb * ((u & 1) ? -1.0 : 1.0) * fact_ratio2(top, u) *
                          ipow(q - p, top - 2 * u) / ipow(delta, top - u);
        }
}

__global__ void prim_batch_kernel(long long n, const double* __restrict__ centres,
                                  const double* __restrict__ norms,
                                  const int* __restrict__ powers,
                                  const double* __restrict__ alphas, int boys,
                                  double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* R = centres + 12 * t;
  const int* pw = powers + 12 * t;
  const double aa = alphas[4 * t], ab = alphas[4 * t + 1], ac = alphas[4 * t + 2], ad = alphas[4 * t + 3];
  const double g1 = RN_ADD(aa, ab), g2 = RN_ADD(ac, ad);
  double P[3], Q[3];
  for (int k = 0; k < 3; ++k) {  // product_center_1D, cints.c:391-394 (no FMA)
    P[k] = RN_DIV(RN_ADD(RN_MUL(aa, R[k]), RN_MUL(ab, R[3 + k])), g1);
    Q[k] = RN_DIV(RN_ADD(RN_MUL(ac, R[6 + k]), RN_MUL(ad, R[9 + k])), g2);
  }
  double rab2 = 0, rcd2 = 0;
  for (int k = 0; k < 3; ++k) {
    rab2 += (R[k] - R[3 + k]) * (R[k] - R[3 + k]);
    rcd2 += (R[6 + k] - R[9 + k]) * (R[6 + k] - R[9 + k]);
  }
  const double dx = P[0] - Q[0], dy = P[1] - Q[1], dz = P[2] - Q[2];
  const double rpq2 = RN_ADD(RN_ADD(RN_MUL(dx, dx), RN_MUL(dy, dy)), RN_MUL(dz, dz));
  const double rsum = RN_ADD(RN_DIV(1.0, g1), RN_DIV(1.0, g2));
  const double delta = 0.25 * rsum;
  const double x = RN_DIV(rpq2, rsum);  // == 0.25*rpq2/delta exactly (cints.c:96,106)

  double B[3][kMaxAxis + 1];
  int nB[3];
  for (int k = 0; k < 3; ++k) {
    nB[k] = pw[k] + pw[3 + k] + pw[6 + k] + pw[9 + k] + 1;
    b_axis(B[k], pw[k], pw[3 + k], pw[6 + k], pw[9 + k], P[k], R[k], R[3 + k], Q[k], R[6 + k],
           R[9 + k], g1, g2, delta);
  }
  const int L = nB[0] + nB[1] + nB[2] - 3;
  double F[3 * kMaxAxis + 1];
  if (boys == kBoysReference) {
    // Fgamma is a pure function of (m, x); the reference re-evaluates it per (I,J,K) triple
    // (cints.c:106), here once per order.
    double xx = x;
    if (fabs(xx) < 0.00000001) xx = 0.00000001;  // cints.c:304
    const double ex = exp(-xx), rx = 1.0 / xx;
    double xpow = sqrt(rx);
    for (int m = 0; m <= L; ++m) {
      F[m] = boys_reference_order(m, xx, ex, xpow);
      xpow *= rx;
    }
  } else {
    for (int m = 0; m <= L; ++m) F[m] = boys_converged(m, x);
  }
  double sum = 0.0;
  for (int I = 0; I < nB[0]; ++I)
    for (int J = 0; J < nB[1]; ++J)
      for (int K = 0; K < nB[2]; ++K) sum += B[0][I] * B[1][J] * B[2][K] * F[I + J + K];
  const double pref = kTwoPi52 / (g1 * g2 * sqrt(g1 + g2)) * exp(-aa * ab * rab2 / g1) *
                      exp(-ac * ad * rcd2 / g2);
  out[t] = pref * sum * norms[4 * t] * norms[4 * t + 1] * norms[4 * t + 2] * norms[4 * t + 3];
}

}  // namespace rchem

namespace rchem { int fail_public(int code, const std::string& msg); }

extern "C" int rchem_coulomb_repulsion_batch(int64_t n, const double* centres, const double* norms,
                                             const int32_t* powers, const double* alphas, int boys,
                                             double* out) {
  using namespace rchem;
  if (n < 0 || (n > 0 && (!centres || !norms || !powers || !alphas || !out)))
    return fail_public(RCHEM_ERR_INVALID_ARG, "rchem_coulomb_repulsion_batch: null argument");
  if (n == 0) return RCHEM_OK;
  for (int64_t t = 0; t < n; ++t) {
    int axis[3] = {0, 0, 0};
    for (int c = 0; c < 4; ++c) {
      int L = 0;
      for (int k = 0; k < 3; ++k) {
        const int v = powers[12 * t + 3 * c + k];
        if (v < 0) return fail_public(RCHEM_ERR_INVALID_ARG, "negative Cartesian power");
        L += v;
        axis[k] += v;
      }
      if (L > 3) return fail_public(RCHEM_ERR_UNSUPPORTED_AM, "tier-1 kernel: l+m+n <= 3 per centre");
      if (!(alphas[4 * t + c] > 0.0)) return fail_public(RCHEM_ERR_INVALID_ARG, "non-positive exponent");
    }
    for (int k = 0; k < 3; ++k)
      if (axis[k] > kMaxAxis) return fail_public(RCHEM_ERR_UNSUPPORTED_AM, "axis power sum > 12");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail_public(RCHEM_ERR_NO_DEVICE, "no CUDA device: librchem_b200 has no CPU path");
  double *dC = nullptr, *dN = nullptr, *dA = nullptr, *dO = nullptr;
  int* dP = nullptr;
  cudaError_t e = cudaMalloc(&dC, 12 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dN, 4 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dA, 4 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dO, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dP, 12 * n * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(dC, centres, 12 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dN, norms, 4 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dA, alphas, 4 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dP, powers, 12 * n * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    prim_batch_kernel<<<(unsigned)((n + 63) / 64), 64>>>(n, dC, dN, dP, dA, boys, dO);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, dO, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dC); cudaFree(dN); cudaFree(dA); cudaFree(dO); cudaFree(dP);
  if (e != cudaSuccess) return fail_public(RCHEM_ERR_CUDA, cudaGetErrorString(e));
  return RCHEM_OK;
}

extern "C" double coulomb_repulsion(double xa, double ya, double za, double norma, int la, int ma,
                                    int na, double alphaa, double xb, double yb, double zb,
                                    double normb, int lb, int mb, int nb, double alphab, double xc,
                                    double yc, double zc, double normc, int lc, int mc, int nc,
                                    double alphac, double xd, double yd, double zd, double normd,
                                    int ld, int md, int nd, double alphad) {
  const double centres[12] = {xa, ya, za, xb, yb, zb, xc, yc, zc, xd, yd, zd};
  const double norms[4] = {norma, normb, normc, normd};
  const int32_t powers[12] = {la, ma, na, lb, mb, nb, lc, mc, nc, ld, md, nd};
  const double alphas[4] = {alphaa, alphab, alphac, alphad};
  double out = std::numeric_limits<double>::quiet_NaN();
  if (rchem_coulomb_repulsion_batch(1, centres, norms, powers, alphas, 0, &out) != RCHEM_OK)
    return std::numeric_limits<double>::quiet_NaN();
  return out;
}
