// eri_kernel.cuh -- the sm_100a shell-quartet kernel, one instantiation per angular-momentum
// class, Boys flavour and output mode.
//
// Work decomposition (DESIGN.md "kernel"):
//   * a TASK is a pair of shell-pair batches (bra batch, ket batch); every pair of a batch has
//     the same (la,lb) and the same number of primitive pairs K2, so all quartets of a task
//     run the same straight-line code with the same trip counts;
//   * ket pairs are sorted by Schwarz bound (descending), so the kets that survive screening
//     against bra pair p are the prefix q < nq[p];
//   * one WARP takes (bra pair p, 32 consecutive kets): bra data is warp-uniform (one
//     broadcast load), ket data is read coalesced from [primitive][pair] SoA arrays;
//   * one THREAD owns one contracted shell quartet: loops over K2_ket x K2_bra primitive
//     quartets, accumulating the [e0|f0] VRR targets in registers, then HRR, then either
//     digests the block into J/K (8-fold symmetry, 6 updates per integral), scatters it into
//     the dense tensor, or reduces it to a Schwarz bound.
// All arithmetic is IEEE fp64 on the FP64 pipe.
//
// Reference: the loops this replaces are basis.rs:383-428 (JK_direct) and basis.rs:430-460
// (build_I); the per-primitive arithmetic is chgp.c:113-135,412-586 / cints.c:72-115.
#pragma once
#include <cuda_runtime.h>

#include "eri_core.h"

namespace rchem {

constexpr int kThreads = 128;
constexpr int kWarpsPerBlock = kThreads / 32;

enum EriMode : int { kModeJK = 0, kModeTensor = 1, kModeSchwarz = 2 };

// Device view of one shell-pair batch.  prim holds six [K2][stride] arrays in the order
// zeta, rzeta, Px, Py, Pz, pref; geom holds six [stride] arrays Ax,Ay,Az,ABx,ABy,ABz;
// idx holds three [stride] int arrays bfA, bfB, diag.
struct BatchView {
  const double* prim;
  const double* geom;
  const int* idx;
  int npairs;
  int stride;
  int K2;
};

struct EriTask {
  BatchView bra, ket;
  const long long* warp_prefix;  // [bra.npairs+1] running count of 32-ket chunks
  const int* nq;                 // [bra.npairs]   surviving kets of bra pair p
  long long nwarps;              // warp_prefix[bra.npairs]
  int same;                      // bra batch == ket batch (then ket q <= p, and p==q is diagonal)
  int rank, nranks;              // multi-GPU: this process takes blocks b with b % nranks == rank
  int N;                         // number of basis functions
  const double* D;               // [N][N] density (symmetric)            kModeJK
  double* Jh;                    // [N][N] half-accumulated J             kModeJK
  double* Kh;                    // [N][N] half-accumulated K             kModeJK
  double* I;                     // [N]^4 dense tensor                    kModeTensor
  double* Qout;                  // [bra.npairs] Schwarz bounds           kModeSchwarz
  const double* boys_table;      // exact-Boys grid (eri_core.h)
  double compscale[3][6];        // per-l component norm ratios (basis_model.h)
};

__device__ __forceinline__ PrimPair load_prim(const BatchView& b, int k, int p) {
  const size_t fs = (size_t)b.K2 * b.stride;
  const double* base = b.prim + (size_t)k * b.stride + p;
  PrimPair pp;
  pp.zeta = __ldg(base);
  pp.rzeta = __ldg(base + fs);
  pp.Px = __ldg(base + 2 * fs);
  pp.Py = __ldg(base + 3 * fs);
  pp.Pz = __ldg(base + 4 * fs);
  pp.pref = __ldg(base + 5 * fs);
  return pp;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LA, int LB, int LC, int LD, int BOYS, int MODE>
__global__ void __launch_bounds__(kThreads) eri_kernel(const EriTask t) {
  using C = EriClass<LA, LB, LC, LD>;
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  constexpr bool kUnroll = C::kOut <= 81;

  const int lane = threadIdx.x & 31;
  const long long w =
      ((long long)blockIdx.x * t.nranks + t.rank) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (w >= t.nwarps) return;  // whole warp leaves together

  int p, q;
  bool active;
  if (MODE == kModeSchwarz) {
    p = q = (int)(w * 32 + lane);
    active = p < t.bra.npairs;
  } else {
    int lo = 0, hi = t.bra.npairs;  // warp_prefix[lo] <= w < warp_prefix[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(t.warp_prefix + mid) <= w) lo = mid; else hi = mid;
    }
    p = lo;
    q = (int)(w - __ldg(t.warp_prefix + p)) * 32 + lane;
    active = q < __ldg(t.nq + p);
  }

  double out[C::kOut];
  int bfA = 0, bfB = 0, bfC = 0, bfD = 0;
  double scale = 0.0;

  if (active) {
    const double* gb = t.bra.geom + p;
    const double* gk = t.ket.geom + q;
    const int sb = t.bra.stride, sk = t.ket.stride;
    const double Ax = __ldg(gb), Ay = __ldg(gb + sb), Az = __ldg(gb + 2 * sb);
    const double Cx = __ldg(gk), Cy = __ldg(gk + sk), Cz = __ldg(gk + 2 * sk);

    double acc[C::kTargets];
#pragma unroll
    for (int i = 0; i < C::kTargets; ++i) acc[i] = 0.0;

    const int K2b = t.bra.K2, K2k = t.ket.K2;
    for (int kk = 0; kk < K2k; ++kk) {
      const PrimPair pk = load_prim(t.ket, kk, q);
      for (int kb = 0; kb < K2b; ++kb) {
        const PrimPair pb = load_prim(t.bra, kb, p);
        primitive_quartet<C, BOYS>(pb, pk, Ax, Ay, Az, Cx, Cy, Cz, t.boys_table, acc);
      }
    }
    C::hrr(acc, __ldg(gb + 3 * sb), __ldg(gb + 4 * sb), __ldg(gb + 5 * sb), __ldg(gk + 3 * sk),
           __ldg(gk + 4 * sk), __ldg(gk + 5 * sk), out);

    if (LA >= 2 || LB >= 2 || LC >= 2 || LD >= 2) {  // per-component norm ratios (d and up)
#pragma unroll(kUnroll ? C::kOut : 1)
      for (int i = 0; i < C::kOut; ++i) {
        const int d = i % ND, c = (i / ND) % NC, b = (i / (ND * NC)) % NB, a = i / (ND * NC * NB);
        out[i] *= t.compscale[LA][a] * t.compscale[LB][b] * t.compscale[LC][c] * t.compscale[LD][d];
      }
    }

    bfA = __ldg(t.bra.idx + p);
    bfB = __ldg(t.bra.idx + sb + p);
    bfC = __ldg(t.ket.idx + q);
    bfD = __ldg(t.ket.idx + sk + q);
    // degeneracy of the shell quartet under the 8 index permutations
    scale = 1.0;
    if (__ldg(t.bra.idx + 2 * sb + p)) scale *= 0.5;
    if (__ldg(t.ket.idx + 2 * sk + q)) scale *= 0.5;
    if (t.same && p == q) scale *= 0.5;
  }

  if (MODE == kModeSchwarz) {
    if (active) {
      double m = 0.0;
      for (int a = 0; a < NA; ++a)
        for (int b = 0; b < NB; ++b) m = fmax(m, fabs(out[((a * NB + b) * NA + a) * NB + b]));
      t.Qout[p] = sqrt(m);
    }
    return;
  }

  if (MODE == kModeTensor) {
    if (active) {
      const size_t N = (size_t)t.N;
      double* __restrict__ I = t.I;
#pragma unroll(kUnroll ? C::kOut : 1)
      for (int o = 0; o < C::kOut; ++o) {
        const int d = o % ND, c = (o / ND) % NC, b = (o / (ND * NC)) % NB, a = o / (ND * NC * NB);
        const size_t i = bfA + a, j = bfB + b, k = bfC + c, l = bfD + d;
        const double v = out[o];
        I[((i * N + j) * N + k) * N + l] = v;
        I[((j * N + i) * N + k) * N + l] = v;
        I[((i * N + j) * N + l) * N + k] = v;
        I[((j * N + i) * N + l) * N + k] = v;
        I[((k * N + l) * N + i) * N + j] = v;
        I[((l * N + k) * N + i) * N + j] = v;
        I[((k * N + l) * N + j) * N + i] = v;
        I[((l * N + k) * N + j) * N + i] = v;
      }
    }
    return;
  }

  // ---- kModeJK: digest the block against D -------------------------------------------
  //   Jh[ab] += 2 s v D[cd]     Jh[cd] += 2 s v D[ab]
  //   Kh[ac] += s v D[bd]  Kh[ad] += s v D[bc]  Kh[bc] += s v D[ad]  Kh[bd] += s v D[ac]
  // and J = Jh + Jh^T, K = Kh + Kh^T afterwards (finalize kernel).  s is the degeneracy
  // factor; D is symmetric.  See DESIGN.md "J/K digestion" for the derivation from
  // basis.rs:406-417.
  {
    const int N = t.N;
    const double* __restrict__ D = t.D;
    double jab[NA * NB], jcd[NC * ND], kac[NA * NC], kad[NA * ND], kbc[NB * NC], kbd[NB * ND];
#pragma unroll
    for (int i = 0; i < NA * NB; ++i) jab[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NC * ND; ++i) jcd[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NA * NC; ++i) kac[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NA * ND; ++i) kad[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB * NC; ++i) kbc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB * ND; ++i) kbd[i] = 0.0;

    if (active) {
      double Dab[NA * NB], Dcd[NC * ND], Dac[NA * NC], Dad[NA * ND], Dbc[NB * NC], Dbd[NB * ND];
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) Dab[a * NB + b] = __ldg(D + (size_t)(bfA + a) * N + bfB + b);
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int d = 0; d < ND; ++d) Dcd[c * ND + d] = __ldg(D + (size_t)(bfC + c) * N + bfD + d);
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < NC; ++c) Dac[a * NC + c] = __ldg(D + (size_t)(bfA + a) * N + bfC + c);
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int d = 0; d < ND; ++d) Dad[a * ND + d] = __ldg(D + (size_t)(bfA + a) * N + bfD + d);
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int c = 0; c < NC; ++c) Dbc[b * NC + c] = __ldg(D + (size_t)(bfB + b) * N + bfC + c);
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int d = 0; d < ND; ++d) Dbd[b * ND + d] = __ldg(D + (size_t)(bfB + b) * N + bfD + d);

#pragma unroll(kUnroll ? C::kOut : 1)
      for (int o = 0; o < C::kOut; ++o) {
        const int d = o % ND, c = (o / ND) % NC, b = (o / (ND * NC)) % NB, a = o / (ND * NC * NB);
        const double v = scale * out[o];
        const double v2 = v + v;
        jab[a * NB + b] = fma(v2, Dcd[c * ND + d], jab[a * NB + b]);
        jcd[c * ND + d] = fma(v2, Dab[a * NB + b], jcd[c * ND + d]);
        kac[a * NC + c] = fma(v, Dbd[b * ND + d], kac[a * NC + c]);
        kad[a * ND + d] = fma(v, Dbc[b * NC + c], kad[a * ND + d]);
        kbc[b * NC + c] = fma(v, Dad[a * ND + d], kbc[b * NC + c]);
        kbd[b * ND + d] = fma(v, Dac[a * NC + c], kbd[b * ND + d]);
      }
    }

    // bra-pair row of J: every lane of the warp shares (a,b) -> reduce, one atomic per warp
    const int bfA0 = __shfl_sync(0xffffffffu, bfA, 0), bfB0 = __shfl_sync(0xffffffffu, bfB, 0);
#pragma unroll
    for (int i = 0; i < NA * NB; ++i) {
      const double s = warp_sum(jab[i]);
      if (lane == 0) atomicAdd(t.Jh + (size_t)(bfA0 + i / NB) * N + bfB0 + i % NB, s);
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < NC * ND; ++i)
        atomicAdd(t.Jh + (size_t)(bfC + i / ND) * N + bfD + i % ND, jcd[i]);
#pragma unroll
      for (int i = 0; i < NA * NC; ++i)
        atomicAdd(t.Kh + (size_t)(bfA + i / NC) * N + bfC + i % NC, kac[i]);
#pragma unroll
      for (int i = 0; i < NA * ND; ++i)
        atomicAdd(t.Kh + (size_t)(bfA + i / ND) * N + bfD + i % ND, kad[i]);
#pragma unroll
      for (int i = 0; i < NB * NC; ++i)
        atomicAdd(t.Kh + (size_t)(bfB + i / NC) * N + bfC + i % NC, kbc[i]);
#pragma unroll
      for (int i = 0; i < NB * ND; ++i)
        atomicAdd(t.Kh + (size_t)(bfB + i / ND) * N + bfD + i % ND, kbd[i]);
    }
  }
}

// Host-side launcher signature, one per class (eri_class_launch.cu.in)
typedef cudaError_t (*EriLaunchFn)(int boys, int mode, const EriTask& task, unsigned grid,
                                   cudaStream_t stream);

}  // namespace rchem
