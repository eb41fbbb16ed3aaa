// one_electron.cu -- overlap, kinetic and nuclear-attraction matrices on the GPU: the step
// before the ERI hot path, mirroring basis::S / T / V (basis.rs:234-338).
//
// The reference evaluates these through its Obara-Saika engine os86.rs with the exact Boys
// function; the kernels below use the equivalent Taketa-Huzinaga-O-ohata closed forms
// (cf. cints.c:152-269) with a converged Boys function.  O(N^2) work, one thread per function
// pair (mu >= nu), primitives looped inside; not a hot path.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rchem_eri.h"
#include "tho_device.cuh"

namespace rchem {
int fail_public(int code, const std::string& msg);

struct FlatBasisDev {
  int n;
  const double* origins;      // [3n]
  const int* powers;          // [3n]
  const int* prim_offset;     // [n+1]
  const double *exps, *coefs, *norms;
};

// THO eq. 2.12, one axis
__device__ double overlap_axis(int l1, int l2, double pa, double pb, double g) {
  double s = 0.0, odd = 1.0, pw = 1.0;  // (2i-1)!! and (2g)^i
  for (int i = 0; i <= (l1 + l2) / 2; ++i) {
    if (i > 0) { odd *= (2 * i - 1); pw *= 2.0 * g; }
    s += binomial_prefactor(2 * i, l1, l2, pa, pb) * odd / pw;
  }
  return s;
}

__device__ double prim_overlap(double a1, double a2, const double* ra, const double* rb,
                               const int* p1, int l2, int m2, int n2) {
  if (l2 < 0 || m2 < 0 || n2 < 0) return 0.0;
  const double g = a1 + a2;
  double r2 = 0.0;
  for (int k = 0; k < 3; ++k) r2 += (ra[k] - rb[k]) * (ra[k] - rb[k]);
  const int p2[3] = {l2, m2, n2};
  double v = 3.14159265358979323846 / g;
  v = v * sqrt(v) * exp(-a1 * a2 * r2 / g);
  for (int k = 0; k < 3; ++k) {
    const double p = (a1 * ra[k] + a2 * rb[k]) / g;
    v *= overlap_axis(p1[k], p2[k], p - ra[k], p - rb[k], g);
  }
  return v;
}

__device__ double prim_kinetic(double a1, double a2, const double* ra, const double* rb,
                               const int* p1, const int* p2) {
  const int l = p2[0], m = p2[1], n = p2[2];
  const double s0 = prim_overlap(a1, a2, ra, rb, p1, l, m, n);
  const double up = prim_overlap(a1, a2, ra, rb, p1, l + 2, m, n) +
                    prim_overlap(a1, a2, ra, rb, p1, l, m + 2, n) +
                    prim_overlap(a1, a2, ra, rb, p1, l, m, n + 2);
  const double dn = l * (l - 1) * prim_overlap(a1, a2, ra, rb, p1, l - 2, m, n) +
                    m * (m - 1) * prim_overlap(a1, a2, ra, rb, p1, l, m - 2, n) +
                    n * (n - 1) * prim_overlap(a1, a2, ra, rb, p1, l, m, n - 2);
  return a2 * (2 * (l + m + n) + 3) * s0 - 2.0 * a2 * a2 * up - 0.5 * dn;
}

// THO eq. 2.18 / 3.1, one axis
__device__ void nuclear_axis(double* A, int l1, int l2, double pa, double pb, double cp, double g) {
  for (int i = 0; i <= l1 + l2; ++i) A[i] = 0.0;
  for (int i = 0; i <= l1 + l2; ++i) {
    const double bp = binomial_prefactor(i, l1, l2, pa, pb) * dfact(i);
    for (int r = 0; r <= i / 2; ++r)
      for (int u = 0; u <= (i - 2 * r) / 2; ++u) {
        const double sign = ((i + u) & 1) ? -1.0 : 1.0;
        A[i - 2 * r - u] += sign * bp * ipow(cp, i - 2 * r - 2 * u) * ipow(0.25 / g, r + u) /
                            (dfact(r) * dfact(u) * dfact(i - 2 * r - 2 * u));
      }
  }
}

constexpr int kOneElAxis = 9;  // l1 + l2 <= 8 (up to g functions on each centre)

__device__ double prim_nuclear(double a1, double a2, const double* ra, const double* rb,
                               const double* rc, const int* p1, const int* p2) {
  const double g = a1 + a2;
  double r2 = 0.0, P[3], A[3][kOneElAxis], pc2 = 0.0;
  for (int k = 0; k < 3; ++k) {
    r2 += (ra[k] - rb[k]) * (ra[k] - rb[k]);
    P[k] = (a1 * ra[k] + a2 * rb[k]) / g;
    pc2 += (P[k] - rc[k]) * (P[k] - rc[k]);
    nuclear_axis(A[k], p1[k], p2[k], P[k] - ra[k], P[k] - rb[k], P[k] - rc[k], g);
  }
  const int L = p1[0] + p1[1] + p1[2] + p2[0] + p2[1] + p2[2];
  double F[3 * kOneElAxis];
  for (int m = 0; m <= L; ++m) F[m] = boys_converged(m, g * pc2);
  double sum = 0.0;
  for (int I = 0; I <= p1[0] + p2[0]; ++I)
    for (int J = 0; J <= p1[1] + p2[1]; ++J)
      for (int K = 0; K <= p1[2] + p2[2]; ++K) sum += A[0][I] * A[1][J] * A[2][K] * F[I + J + K];
  return -2.0 * 3.14159265358979323846 / g * exp(-a1 * a2 * r2 / g) * sum;
}

// which: 0 = S, 1 = T, 2 = V (sum over atoms of Z_c <mu| -1/|r-C| |nu>)
__global__ void one_electron_kernel(FlatBasisDev b, int which, int natoms,
                                    const double* __restrict__ atomcoords,
                                    const double* __restrict__ charges, double* __restrict__ M) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = b.n;
  if (t >= (long long)n * (n + 1) / 2) return;
  long long mu = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((mu + 1) * (mu + 2) / 2 <= t) ++mu;
  while (mu * (mu + 1) / 2 > t) --mu;
  const long long nu = t - mu * (mu + 1) / 2;
  const double* ra = b.origins + 3 * mu;
  const double* rb = b.origins + 3 * nu;
  const int* p1 = b.powers + 3 * mu;
  const int* p2 = b.powers + 3 * nu;
  double acc = 0.0;
  // contraction order of basis.rs:241-246: outer over the ket primitives, inner over the bra
  for (int j = b.prim_offset[nu]; j < b.prim_offset[nu + 1]; ++j) {
    double left = 0.0;
    for (int i = b.prim_offset[mu]; i < b.prim_offset[mu + 1]; ++i) {
      double v = 0.0;
      if (which == 0) v = prim_overlap(b.exps[i], b.exps[j], ra, rb, p1, p2[0], p2[1], p2[2]);
      else if (which == 1) v = prim_kinetic(b.exps[i], b.exps[j], ra, rb, p1, p2);
      else
        for (int c = 0; c < natoms; ++c)
          v += charges[c] * prim_nuclear(b.exps[i], b.exps[j], ra, rb, atomcoords + 3 * c, p1, p2);
      left += b.coefs[i] * (b.norms[i] * b.norms[j] * v);
    }
    acc += b.coefs[j] * left;
  }
  M[(size_t)mu * n + nu] = acc;
  M[(size_t)nu * n + mu] = acc;
}

// Host driver: flat CGTO arrays in, one N x N matrix out (host buffers).
int one_electron_host(int n, const double* origins, const int* powers, const int* prim_offset,
                      const double* exps, const double* coefs, const double* norms, int which,
                      int natoms, const double* atomcoords, const double* charges, double* M,
                      int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail_public(RCHEM_ERR_NO_DEVICE, "no CUDA device: librchem_b200 has no CPU path");
  if (device < 0 || device >= ndev) return fail_public(RCHEM_ERR_INVALID_ARG, "bad device ordinal");
  if (cudaSetDevice(device) != cudaSuccess) return fail_public(RCHEM_ERR_CUDA, "cudaSetDevice");
  for (int i = 0; i < n; ++i) {
    const int L = powers[3 * i] + powers[3 * i + 1] + powers[3 * i + 2];
    if (L > 4) return fail_public(RCHEM_ERR_UNSUPPORTED_AM, "one-electron kernels: l <= 4");
  }
  const int nprim = prim_offset[n];
  double *dO = nullptr, *dE = nullptr, *dC = nullptr, *dN = nullptr, *dM = nullptr, *dX = nullptr, *dZ = nullptr;
  int *dP = nullptr, *dF = nullptr;
  cudaError_t e = cudaSuccess;
  auto up = [&](auto** dst, const auto* src, size_t count) {
    if (e != cudaSuccess || count == 0) return;
    e = cudaMalloc((void**)dst, count * sizeof(**dst));
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, count * sizeof(**dst), cudaMemcpyHostToDevice);
  };
  up(&dO, origins, 3 * (size_t)n); up(&dP, powers, 3 * (size_t)n); up(&dF, prim_offset, (size_t)n + 1);
  up(&dE, exps, nprim); up(&dC, coefs, nprim); up(&dN, norms, nprim);
  if (which == 2) { up(&dX, atomcoords, 3 * (size_t)natoms); up(&dZ, charges, natoms); }
  if (e == cudaSuccess) e = cudaMalloc(&dM, (size_t)n * n * sizeof(double));
  if (e == cudaSuccess) {
    FlatBasisDev b{n, dO, dP, dF, dE, dC, dN};
    const long long npairs = (long long)n * (n + 1) / 2;
    one_electron_kernel<<<(unsigned)((npairs + 63) / 64), 64>>>(b, which, natoms, dX, dZ, dM);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(M, dM, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dO); cudaFree(dP); cudaFree(dF); cudaFree(dE); cudaFree(dC); cudaFree(dN);
  cudaFree(dM); cudaFree(dX); cudaFree(dZ);
  if (e != cudaSuccess) return fail_public(RCHEM_ERR_CUDA, cudaGetErrorString(e));
  return RCHEM_OK;
}

}  // namespace rchem
