// eri_kernel.cuh -- the sm_100a shell-quartet kernels, instantiated per angular-momentum class
// and Boys flavour.
//
// Work decomposition:
//   * a TASK is a pair of shell-pair batches (bra batch, ket batch); every pair of a batch has
//     the same (la,lb) and the same number of (significant) primitive pairs K2, so all
//     quartets of a task run the same straight-line code with the same trip counts;
//   * ket pairs are sorted by Schwarz bound (descending), so the kets that survive screening
//     against bra pair p are the prefix q < nq[p];
//   * one THREAD owns one contracted shell quartet: loops over K2_ket x K2_bra primitive
//     quartets, accumulating the [e0|f0] VRR targets in registers, then HRR, then digests the
//     block into J/K (8-fold symmetry, 6 updates per integral), writes the canonical tensor
//     element, or reduces it to a Schwarz bound;
//   * eri_kernel (chunk kernel): one WARP per (bra pair p, 32 consecutive kets) -- bra data
//     warp-uniform, ket data coalesced from [primitive][pair] SoA arrays.  Used for the dense
//     tensor, the Schwarz bounds and (fallback) the J/K of bra pairs no other kernel takes;
//   * eri_jk_block_kernel: one BLOCK per (bra pair, <= 8 T kets) for "heavy" bra pairs, with the
//     bra pair's D rows and K accumulators in shared memory (K as 64-bit fixed point on native
//     32-bit shared atomics);
//   * eri_jk_light_kernel / eri_jk_light_multi_kernel: one WARP per "light" bra pair and all
//     its kets, every light task of a class merged into one launch;
//   * both J/K kernels sort their kets by regime -- PROVED far-field / Boys grid / grid +
//     Fgamma correction -- and run the far-field ones through the rho-free point-multipole
//     form primitive_quartet_far (eri_core.h): no Boys table, no 1/sqrt(zeta+eta).
// All integral arithmetic is IEEE fp64 on the FP64 pipe; only the shared-memory K accumulators
// of the block kernel are 64-bit fixed point (krow_add), converted back to fp64 at the flush.
//
// Reference: the loops this replaces are basis.rs:383-428 (JK_direct) and basis.rs:430-460
// (build_I); the per-primitive arithmetic is chgp.c:113-135,412-586 / cints.c:72-115.
#pragma once
#include <cuda_runtime.h>

#include "eri_core.h"

namespace rchem {

// integrals per shell quartet up to which the component scaling and the J/K digestion are fully
// unrolled (static indices: the J/K partial sums stay in registers and out[] is addressed with
// immediates); beyond it they are rolled loops with run-time div/mod over local-memory arrays.
// Measured on (H2O)32/6-31G* (profiles/r02_ab_digest_unroll.txt): 81 -> 33.9 ms, 108 -> 29.8,
// 216 -> 28.5, 324 -> 27.2, 1296 (= every class) -> 26.6 ms, for 50 s more build time.
#ifndef RCHEM_UNROLL_MAX
#define RCHEM_UNROLL_MAX 1296
#endif
constexpr int kThreads = 128;
constexpr int kWarpsPerBlock = kThreads / 32;

enum EriMode : int { kModeJK = 0, kModeTensor = 1, kModeSchwarz = 2 };

// Device view of one shell-pair batch.  prim holds kPrimFieldsBase + NV [K2][stride] arrays in the
// order zeta, rzeta, Px, Py, Pz, fsc, pfar, w[0..NV) (PrimPairV, eri_core.h; NV = weight variants of
// the pair type: 1, or 2 / 4 with fused sp shells); geom holds kGeomFields [stride] arrays Ax,Ay,Az,
// ABx,ABy,ABz, then the pair's bounding data Mx,My,Mz,rad,zmin (pair_build.h PairBound) and its
// Schwarz bound Q; idx holds four [stride] int arrays bfA, bfB, diag and the packed word
// bfA | bfB << 15 | diag << 30 (valid for N < 32768: the J/K block and light kernels).
constexpr int kGeomFields = 12;
struct BatchView {
  const double* prim;
  const double* geom;
  const float4* bnd;    // [stride] {Mx, My, Mz, rad} of the pair's bounding sphere, single precision
  const float* zminf;   // [stride] most diffuse exponent sum (rounded down)
  const int* idx;
  const double* Dp;  // [ncart(la)*ncart(lb)][stride]  D block of every pair, pair-major packed
  double* Jp;        // [ncart(la)*ncart(lb)][stride]  J block of every pair (accumulated)
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
  int far_sched;                 // 1: prove-and-route far-field quartets (0 = tuning/debug)
  const double* dmax;            // device scalar max|D|                  (block kernel, K rows)
  double kbound;                 // 16 max_X sum_Y n_Y Q_XY               (block kernel, K rows)
  int N;                         // number of basis functions
  const double* D;               // [N][N] density (symmetric)            kModeJK
  double* Kh;                    // [N][N] half-accumulated K             kModeJK
  // block-per-bra-pair J/K kernel (eri_jk_block_kernel): "heavy" bra pairs only
  const int* hp;                 // [nheavy]   bra pair index of heavy entry h
  const long long* hblk_prefix;  // [nheavy+1] running block count
  long long nblocks_heavy;
  int nheavy;
  // warp-per-bra-pair J/K kernel (eri_jk_light_kernel): "light" bra pairs with >= 1 ket
  const int* lp;                 // [nlight]   bra pair index of light entry
  int nlight;
  int light_cap;                 // list capacity per warp (upper bound of a light pair's nq)
  double* I;                     // [N]^4 dense tensor                    kModeTensor
  double* Qout;                  // [bra.npairs] Schwarz bounds           kModeSchwarz
  BoysTabs boys;                 // Boys tables (exact grid of this class's L, reference tables)
  double compscale[kNumTypes][6];  // per-type component norm ratios (basis_model.h)
};

template <int NV>
__device__ __forceinline__ PrimPairV<NV> load_prim(const BatchView& b, int k, int p) {
  const size_t fs = (size_t)b.K2 * b.stride;
  const double* base = b.prim + (size_t)k * b.stride + p;
  PrimPairV<NV> pp;
  pp.zeta = __ldg(base);
  pp.rzeta = __ldg(base + fs);
  pp.Px = __ldg(base + 2 * fs);
  pp.Py = __ldg(base + 3 * fs);
  pp.Pz = __ldg(base + 4 * fs);
  pp.fsc = __ldg(base + 5 * fs);
  pp.pfar = __ldg(base + 6 * fs);  // (fields a caller does not use are dead loads, removed)
#pragma unroll
  for (int v = 0; v < NV; ++v) pp.w[v] = __ldg(base + (kPrimFieldsBase + v) * fs);
  return pp;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// One contracted shell quartet (bra pair p | ket pair q) of a task: primitive loop, HRR,
// component norms.  Returns the degeneracy factor of the quartet under the 8 permutations.
// ---------------------------------------------------------------------------------------
struct BraGeom {
  double Ax, Ay, Az, ABx, ABy, ABz;
  int diag;
};

__device__ __forceinline__ BraGeom load_bra_geom(const BatchView& bra, int p) {
  const double* gb = bra.geom + p;
  const int sb = bra.stride;
  BraGeom g;
  g.Ax = __ldg(gb); g.Ay = __ldg(gb + sb); g.Az = __ldg(gb + 2 * sb);
  g.ABx = __ldg(gb + 3 * sb); g.ABy = __ldg(gb + 4 * sb); g.ABz = __ldg(gb + 5 * sb);
  g.diag = __ldg(bra.idx + 2 * sb + p);
  return g;
}

// BRA_SMEM: the bra pair's primitive pairs were staged in shared memory (block kernel);
// otherwise they are read (warp-uniformly) from the SoA arrays.
// FAR: the caller has proved every primitive quartet far-field (primitive_quartet_far).
template <class C, int LA, int LB, int LC, int LD, int BOYS, bool BRA_SMEM, bool FAR = false>
__device__ __forceinline__ double shell_quartet(const EriTask& t, int p, const BraGeom& g,
                                                const PrimPairV<C::kNVb>* __restrict__ s_bra, int q,
                                                double* __restrict__ out, int& bfC, int& bfD) {
  constexpr int NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  constexpr bool kUnroll = C::kOut <= RCHEM_UNROLL_MAX;
  const double* gk = t.ket.geom + q;
  const int sk = t.ket.stride;
  const double Cx = __ldg(gk), Cy = __ldg(gk + sk), Cz = __ldg(gk + 2 * sk);
  // J/K kernels: first functions and diagonal flag of the ket pair in ONE word, requested before
  // the primitive loop so the digestion does not wait for it (bfC | bfD << 15 | diag << 30)
  const unsigned cdx = BRA_SMEM ? __ldg(reinterpret_cast<const unsigned*>(t.ket.idx) + 3 * sk + q) : 0u;

  // Small classes keep the contraction accumulators in registers (fully unrolled init ->
  // scalar replacement).  Past ~100 targets they cannot fit; a rolled init loop indexes the
  // array dynamically, which pins it in (L1-cached) local memory from the start instead of
  // sending ptxas into its spill-everything fallback.
  double acc[C::kTargets];
#pragma unroll(C::kTargets <= 100 ? C::kTargets : 1)
  for (int i = 0; i < C::kTargets; ++i) acc[i] = 0.0;

  const int K2b = t.bra.K2, K2k = t.ket.K2;
  for (int kk = 0; kk < K2k; ++kk) {
    const PrimPairV<C::kNVk> pk = load_prim<C::kNVk>(t.ket, kk, q);
    for (int kb = 0; kb < K2b; ++kb) {
      if (FAR) {
        primitive_quartet_far<C>(s_bra[kb], pk, g.Ax, g.Ay, g.Az, Cx, Cy, Cz, acc);
      } else if (BRA_SMEM) {
        primitive_quartet<C, BOYS>(s_bra[kb], pk, g.Ax, g.Ay, g.Az, Cx, Cy, Cz, t.boys, acc);
      } else {
        const PrimPairV<C::kNVb> pb = load_prim<C::kNVb>(t.bra, kb, p);
        primitive_quartet<C, BOYS>(pb, pk, g.Ax, g.Ay, g.Az, Cx, Cy, Cz, t.boys, acc);
      }
    }
  }
  C::hrr(acc, g.ABx, g.ABy, g.ABz, __ldg(gk + 3 * sk), __ldg(gk + 4 * sk), __ldg(gk + 5 * sk), out);

  if (LA == 2 || LB == 2 || LC == 2 || LD == 2) {  // per-component norm ratios (d shells)
#pragma unroll(kUnroll ? C::kOut : 1)
    for (int i = 0; i < C::kOut; ++i) {
      const int d = i % ND, c = (i / ND) % NC, b = (i / (ND * NC)) % NB, a = i / (ND * NC * NB);
      out[i] *= t.compscale[LA][a] * t.compscale[LB][b] * t.compscale[LC][c] * t.compscale[LD][d];
    }
  }
  int kdiag;
  if (BRA_SMEM) {
    bfC = (int)(cdx & 0x7fffu);
    bfD = (int)((cdx >> 15) & 0x7fffu);
    kdiag = (int)(cdx >> 30);
  } else {
    bfC = __ldg(t.ket.idx + q);
    bfD = __ldg(t.ket.idx + sk + q);
    kdiag = __ldg(t.ket.idx + 2 * sk + q);
  }
  // degeneracy of the shell quartet under the 8 index permutations
  double scale = 1.0;
  if (g.diag) scale *= 0.5;
  if (kdiag) scale *= 0.5;
  if (t.same && p == q) scale *= 0.5;
  return scale;
}

// ---------------------------------------------------------------------------------------
// Warp-task kernel: one warp per (bra pair, 32 consecutive kets).  Used for the dense tensor,
// the Schwarz bounds, and the J/K digestion of "light" bra pairs (few surviving kets).
//
// J/K digestion (derived from basis.rs:406-417 for symmetric D): with
// v = s*(ab|cd), s the degeneracy factor,
//   J[ab] += 2 v D[cd]     J[cd] += 2 v D[ab]
//   Kh[ac] += v D[bd]  Kh[ad] += v D[bc]  Kh[bc] += v D[ad]  Kh[bd] += v D[ac],  K = Kh + Kh^T.
// J and the D blocks it needs live PAIR-PACKED (BatchView::Jp / Dp), so lanes touch
// consecutive addresses; finalize_j_kernel scatters Jp into the N x N matrix.
// ---------------------------------------------------------------------------------------
// Classes emitted in parts (> 100 accumulators, J/K through this kernel): few, long-running
// threads whose time is the latency of one thread's serial work, so resident warps matter more
// than registers (RCHEM_CHUNK_MINB_BIG blocks of 128 threads per SM; measured on (H2O)32/6-31G*:
// 1 = uncapped 36.2 ms, 3: 34.8, 4 (128 registers): 33.9, 6: 33.5, 8: 33.6).
#ifndef RCHEM_CHUNK_MINB_BIG
#define RCHEM_CHUNK_MINB_BIG 4
#endif
template <int LA, int LB, int LC, int LD> struct ChunkCfg {
  static constexpr int kMinBlocks = EriClass<LA, LB, LC, LD>::kTargets > 100 ? RCHEM_CHUNK_MINB_BIG : 1;
};
template <int LA, int LB, int LC, int LD, int BOYS, int MODE>
__global__ void __launch_bounds__(kThreads, ChunkCfg<LA, LB, LC, LD>::kMinBlocks) eri_kernel(const EriTask t) {
  using C = EriClass<LA, LB, LC, LD>;
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  constexpr bool kUnroll = C::kOut <= RCHEM_UNROLL_MAX;

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
  int bfC = 0, bfD = 0;
  double scale = 0.0;
  if (active) {
    const BraGeom g = load_bra_geom(t.bra, min(p, t.bra.npairs - 1));
    scale = shell_quartet<C, LA, LB, LC, LD, BOYS, false>(t, p, g, nullptr, q, out, bfC, bfD);
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

  const int sb = t.bra.stride, sk = t.ket.stride;
  if (MODE == kModeTensor) {
    if (active) {
      const size_t N = (size_t)t.N;
      const int bfA = __ldg(t.bra.idx + p), bfB = __ldg(t.bra.idx + sb + p);
      double* __restrict__ I = t.I;
#pragma unroll(kUnroll ? C::kOut : 1)
      for (int o = 0; o < C::kOut; ++o) {
        const int d = o % ND, c = (o / ND) % NC, b = (o / (ND * NC)) % NB, a = o / (ND * NC * NB);
        const size_t i = bfA + a, j = bfB + b, k = bfC + c, l = bfD + d;
        // only the canonical element (stored pair orientations, bra pair >= ket pair); the other
        // seven permutations are filled in by tensor_fill_kernel (engine.cu)
        I[((i * N + j) * N + k) * N + l] = out[o];
      }
    }
    return;
  }

  // ---- kModeJK ---------------------------------------------------------------------------
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

    const int bfA = __ldg(t.bra.idx + p), bfB = __ldg(t.bra.idx + sb + p);  // warp-uniform
    if (active) {
      double Dab[NA * NB], Dcd[NC * ND], Dac[NA * NC], Dad[NA * ND], Dbc[NB * NC], Dbd[NB * ND];
#pragma unroll
      for (int i = 0; i < NA * NB; ++i) Dab[i] = __ldg(t.bra.Dp + (size_t)i * sb + p);
#pragma unroll
      for (int i = 0; i < NC * ND; ++i) Dcd[i] = __ldg(t.ket.Dp + (size_t)i * sk + q);
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

    // bra-pair block of J: every lane of the warp shares (a,b) -> reduce, one atomic per warp
#pragma unroll
    for (int i = 0; i < NA * NB; ++i) {
      const double s = warp_sum(jab[i]);
      if (lane == 0) atomicAdd(t.bra.Jp + (size_t)i * sb + p, s);
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < NC * ND; ++i) atomicAdd(t.ket.Jp + (size_t)i * sk + q, jcd[i]);
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

// ---------------------------------------------------------------------------------------
// Block-per-bra-pair J/K kernel for "heavy" bra pairs (many surviving kets).  A block owns
// one bra pair (A,B) and a contiguous range of its ket prefix.  The rows D[a,:], D[b,:] and
// the accumulators K[a,:], K[b,:] (a in A, b in B) live in SHARED memory, so the four
// mixed-index K updates and their density factors never touch L1/L2 with scattered 8-byte
// accesses; the bra block of J accumulates in registers over the whole ket range.
// Dynamic shared memory: 2*(NA+NB)*N doubles + K2_bra primitive pairs + the ket index list.
// ---------------------------------------------------------------------------------------
#ifndef RCHEM_BLK_T_SMALL
#define RCHEM_BLK_T_SMALL 512
#endif
#ifndef RCHEM_BLK_PASSES
#define RCHEM_BLK_PASSES 8
#endif
// large classes (> 18 targets) on a wide-row bra pair ((sp sp|: 160 kB of rows, one block per SM)
#ifndef RCHEM_BLK_T_LARGE_WIDE
#define RCHEM_BLK_T_LARGE_WIDE 256
#endif
#ifndef RCHEM_BLK_MINB_SMALL
#define RCHEM_BLK_MINB_SMALL 2
#endif
template <int LA, int LB, int LC, int LD> struct BlockCfg {
  // small classes: RCHEM_BLK_T_SMALL threads, 2 blocks per SM (<= 64 registers); medium ones
  // (<= 18 targets): 256 threads, 2 blocks per SM (<= 128 registers); large ones: 256 threads
  static constexpr int kTargets = EriClass<LA, LB, LC, LD>::kTargets;
  static constexpr bool kSmall = kTargets <= 9;
  static constexpr bool kMedium = !kSmall && kTargets <= 18;
  // a fused (sp sp| bra pair has 8 D rows + 8 K rows in shared memory (160 kB at N = 1248): only
  // one block fits an SM, so medium classes of that kind run 512 threads in it
  static constexpr bool kWideRows = ncart(LA) + ncart(LB) >= 8;
  static constexpr int kThreadsBlk =
      kSmall ? RCHEM_BLK_T_SMALL : ((kMedium && kWideRows) ? 512 : ((!kMedium && kWideRows) ? RCHEM_BLK_T_LARGE_WIDE : 256));
  static constexpr int kMinBlocks = kSmall ? RCHEM_BLK_MINB_SMALL : ((kMedium && !kWideRows) ? 2 : 1);
  static constexpr int kKetsPerBlock = kThreadsBlk * RCHEM_BLK_PASSES;
  static_assert(kKetsPerBlock <= 65535, "the block kernel's ket list holds 16-bit offsets");
};

// K rows of the block kernel in shared memory.  Shared-memory fp64 (and 64-bit integer)
// atomicAdd compile to a compare-and-swap loop (LDS + ATOMS.CAST.SPIN, ~9 LSU wavefronts per
// update; ncu: the LSU data pipe is what bounds the big shallow-contraction launches).  Only
// 32-bit shared atomics are native, so with RCHEM_K_FIXED the rows are 64-bit FIXED-POINT
// accumulators split into two 32-bit words: add the low word (the returned old value gives
// the carry), add high word + carry.  Integer addition commutes, so the 64-bit sum is exact
// whatever the interleaving -- and the shared-memory part of K becomes order-independent.
// Scale: a power of two chosen per block from the rigorous bound
//   sum |terms of one K element| <= 16 Q_ab max|D| max_X sum_Y n_Y Q_XY   (Schwarz),
// so that the accumulator stays below 2^61: resolution ~2^-61 of that bound (<= 1e-16
// absolute for the bounds met in practice), far below the 1e-12 parity tolerance.
#ifndef RCHEM_K_FIXED
#define RCHEM_K_FIXED 1
#endif
__device__ __forceinline__ void krow_add(double* row, int n_row_doubles, int idx, double v,
                                         double kscale) {
#if RCHEM_K_FIXED
  unsigned* lo = reinterpret_cast<unsigned*>(row);
  unsigned* hi = lo + n_row_doubles;
  const long long x = __double2ll_rn(v * kscale);
  const unsigned xl = (unsigned)x;
  unsigned xh = (unsigned)((unsigned long long)x >> 32);
  const unsigned old = atomicAdd(lo + idx, xl);
  xh += (unsigned)((unsigned)(old + xl) < old);  // carry out of the low word
  if (xh) atomicAdd(hi + idx, xh);
#else
  atomicAdd(row + idx, v);
#endif
}
__device__ __forceinline__ double krow_get(const double* row, int n_row_doubles, int idx,
                                           double kinv) {
#if RCHEM_K_FIXED
  const unsigned* lo = reinterpret_cast<const unsigned*>(row);
  const unsigned* hi = lo + n_row_doubles;
  const long long acc = (long long)(((unsigned long long)hi[idx] << 32) | lo[idx]);
  return (double)acc * kinv;
#else
  return row[idx];
#endif
}

// A shell quartet is scheduled as far-field when the bounding spheres of its two shell pairs
// prove x >= far_proven_x(L) for every primitive quartet (eri_core.h: 36 / 40 / 48 by class; the
// far-only code never looks at x again, so the proof must hold).


// Scheduling regime of shell quartet (bra pair | ket pair q): quartet_regime_f (eri_core.h) on the
// pairs' single-precision bounding data (BatchView::bnd: {Mx, My, Mz, rad}, BatchView::zminf).
__device__ __forceinline__ PairBoundF load_bound(const BatchView& b, int p) {
  const float4 v = __ldg(b.bnd + p);
  return PairBoundF{v.x, v.y, v.z, v.w, __ldg(b.zminf + p)};
}

template <int LA, int LB, int LC, int LD, int BOYS>
__global__ void __launch_bounds__(BlockCfg<LA, LB, LC, LD>::kThreadsBlk,
                                  BlockCfg<LA, LB, LC, LD>::kMinBlocks)
eri_jk_block_kernel(const EriTask t) {
  using C = EriClass<LA, LB, LC, LD>;
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  constexpr bool kUnroll = C::kOut <= RCHEM_UNROLL_MAX;
  constexpr int T = BlockCfg<LA, LB, LC, LD>::kThreadsBlk;
  // regimes a block sorts its kets into: far-field (proved), grid, grid + Fgamma correction
  constexpr int kRegimes = BOYS == kBoysReference ? 3 : 2;
  extern __shared__ double smem[];
  __shared__ int s_info[3];
  __shared__ int s_cnt[3];

  const long long blk = (long long)blockIdx.x * t.nranks + t.rank;
  if (blk >= t.nblocks_heavy) return;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    int lo = 0, hi = t.nheavy;  // hblk_prefix[lo] <= blk < hblk_prefix[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(t.hblk_prefix + mid) <= blk) lo = mid; else hi = mid;
    }
    const int p = __ldg(t.hp + lo);
    const int nq = __ldg(t.nq + p);
    const long long first = __ldg(t.hblk_prefix + lo);
    const int nb = (int)(__ldg(t.hblk_prefix + lo + 1) - first);
    const int per = (((nq + nb - 1) / nb) + 31) & ~31;
    const int q0 = (int)(blk - first) * per;
    s_info[0] = p;
    s_info[1] = q0;
    s_info[2] = min(nq, q0 + per);
  }
  if (tid < 3) s_cnt[tid] = 0;
  __syncthreads();
  const int p = s_info[0], q0 = s_info[1], q1 = s_info[2];
  const int N = t.N, sb = t.bra.stride, sk = t.ket.stride;
  const int bfA = __ldg(t.bra.idx + p), bfB = __ldg(t.bra.idx + sb + p);

  double* Drow_a = smem;                 // [NA][N]
  double* Drow_b = Drow_a + NA * N;      // [NB][N]
  double* Krow_a = Drow_b + NB * N;      // [NA][N]
  double* Krow_b = Krow_a + NA * N;      // [NB][N]
  using BraPrim = PrimPairV<C::kNVb>;
  BraPrim* s_bra = reinterpret_cast<BraPrim*>(Krow_b + NB * N);  // [K2_bra]
  for (int k = tid; k < t.bra.K2; k += T) s_bra[k] = load_prim<C::kNVb>(t.bra, k, p);
  BraPrim* s_bra_far = s_bra;
  BraPrim* s_bra_end = s_bra + t.bra.K2;
  const BraGeom g = load_bra_geom(t.bra, p);
  for (int a = 0; a < NA; ++a)
    for (int j = tid; j < N; j += T) {
      Drow_a[a * N + j] = __ldg(t.D + (size_t)(bfA + a) * N + j);
      Krow_a[a * N + j] = 0.0;
    }
  for (int b = 0; b < NB; ++b)
    for (int j = tid; j < N; j += T) {
      Drow_b[b * N + j] = __ldg(t.D + (size_t)(bfB + b) * N + j);
      Krow_b[b * N + j] = 0.0;
    }
  double Dab[NA * NB], jab[NA * NB];
#pragma unroll
  for (int i = 0; i < NA * NB; ++i) {
    Dab[i] = __ldg(t.bra.Dp + (size_t)i * sb + p);
    jab[i] = 0.0;
  }

  // fixed-point scale of the K rows (krow_add): 2^(61 - e), 2^e > bound
  double kscale = 1.0, kinv = 1.0;
  if (RCHEM_K_FIXED) {
    const double bnd = fmax(__ldg(t.bra.geom + 11 * sb + p) * t.kbound * __ldg(t.dmax), 1e-280);
    const int e = min(61 - (ilogb(bnd) + 1), 900);
    kscale = scalbn(1.0, e);
    kinv = scalbn(1.0, -e);
  }

  // Regime scheduling of this block's kets.  89 % of the primitive quartets of a large
  // cluster are far-field (x >= 48: point-multipole form, no Boys table), 8 % need the Boys
  // grid and -- reference flavour -- 3 % the Fgamma truncation correction; but in list
  // (Schwarz) order nearly every warp holds a lane of each regime and pays for all three.
  // The block therefore sorts its kets in shared memory into [far | grid | corrected]
  // (bra pair fixed, so one pass over the kets' bounding data decides):
  //   * far: PROVED from the pairs' bounding data (centre M, radius rad, most diffuse
  //     exponent zmin): rho(zmin_b, zmin_k) (|M_b - M_k| - rad_b - rad_k)^2 >= 48 bounds every
  //     primitive quartet's x from below.  These run the far-only code (primitive_quartet_far);
  //   * grid / corrected: the general code, which handles any x; the split (same lower bound
  //     of x against ref_exact_from(L)) only decides which lanes run together.
  // (16-bit offsets from q0: a block owns at most kKetsPerBlock = 4096 kets, and the list shares
  // the SM's shared memory with up to 200 kB of D/K rows of two resident blocks)
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_bra_end);  // [nk + 32 kRegimes]
  const int nk = q1 - q0;
  // The list is laid out [corrected | grid | far], each regime padded to a whole warp chunk with
  // kEmpty entries, and the warps DRAW 32-ket chunks from a shared counter, most expensive regime
  // first: with a static split the warps that happened to hold the general-code kets reached the
  // final barrier long after the others (ncu: 11-19 % of all warp samples waiting there).
  constexpr unsigned short kEmpty = 0xffffu;
  int far_chunk0, nchunks;
  {
    const PairBoundF bb = load_bound(t.bra, p);
    const float xcorr = (float)(ref_exact_from(C::kL) + 2.0);
    unsigned cls_bits = 0;  // 2 bits per pass: 0 far, 1 grid, 2 corrected, 3 none
    int pass = 0;
    for (int base = 0; base < nk; base += T, ++pass) {
      const int i = base + tid;
      const int cls = i < nk ? quartet_regime_f<kRegimes>(bb, load_bound(t.ket, q0 + i), (float)far_proven_x(C::kL), xcorr, t.far_sched) : 3;
      cls_bits |= (unsigned)cls << (2 * pass);
#pragma unroll
      for (int c = 0; c < kRegimes; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (lane == 0 && m) atomicAdd(&s_cnt[c], __popc(m));
      }
    }
    __syncthreads();
    const int n0 = s_cnt[0], n1 = s_cnt[1], n2 = s_cnt[2];  // (n2 = 0 with two regimes)
    const int p0 = (n0 + 31) & ~31, p1 = (n1 + 31) & ~31, p2 = (n2 + 31) & ~31;
    far_chunk0 = (p2 + p1) >> 5;
    nchunks = (p2 + p1 + p0) >> 5;
    __syncthreads();
    if (tid == 0) { s_cnt[2] = 0; s_cnt[1] = p2; s_cnt[0] = p2 + p1; s_info[0] = 0; }  // cursors, chunk counter
    if (tid < 32) {
      if (n2 + tid < p2) s_list[n2 + tid] = kEmpty;
      if (n1 + tid < p1) s_list[p2 + n1 + tid] = kEmpty;
      if (n0 + tid < p0) s_list[p2 + p1 + n0 + tid] = kEmpty;
    }
    __syncthreads();
    pass = 0;
    for (int base = 0; base < nk; base += T, ++pass) {
      const int i = base + tid;
      const int cls = (cls_bits >> (2 * pass)) & 3;
#pragma unroll
      for (int c = 0; c < kRegimes; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        int at = 0;
        if (lane == 0 && m) at = atomicAdd(&s_cnt[c], __popc(m));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (cls == c) s_list[at + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
      }
    }
  }
  // (The north-star variant -- this class's 61.5 kB exact-Boys grid slice staged in the block's
  // shared memory -- was measured and lost: one block per SM instead of two, 111.6 vs 80.4 ms;
  // profiles/r02_ab_dynamic_chunks_direct_boys.txt.  The grids stay in global memory / L2.)
  const EriTask& tl = t;
  __syncthreads();

  // digestion of one evaluated shell quartet (bra pair p | ket pair q) into J and K
  auto digest = [&](int q, const double* __restrict__ out, double scale, int bfC, int bfD) {
    double jcd[NC * ND], kac[NA * NC], kad[NA * ND], kbc[NB * NC], kbd[NB * ND], Dcd[NC * ND];
#pragma unroll
    for (int i = 0; i < NC * ND; ++i) { jcd[i] = 0.0; Dcd[i] = __ldg(t.ket.Dp + (size_t)i * sk + q); }
#pragma unroll
    for (int i = 0; i < NA * NC; ++i) kac[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NA * ND; ++i) kad[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB * NC; ++i) kbc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB * ND; ++i) kbd[i] = 0.0;
    const double* Da_c = Drow_a + bfC;
    const double* Da_d = Drow_a + bfD;
    const double* Db_c = Drow_b + bfC;
    const double* Db_d = Drow_b + bfD;
#pragma unroll(kUnroll ? C::kOut : 1)
    for (int o = 0; o < C::kOut; ++o) {
      const int d = o % ND, c = (o / ND) % NC, b = (o / (ND * NC)) % NB, a = o / (ND * NC * NB);
      const double v = scale * out[o];
      const double v2 = v + v;
      jab[a * NB + b] = fma(v2, Dcd[c * ND + d], jab[a * NB + b]);
      jcd[c * ND + d] = fma(v2, Dab[a * NB + b], jcd[c * ND + d]);
      kac[a * NC + c] = fma(v, Db_d[b * N + d], kac[a * NC + c]);
      kad[a * ND + d] = fma(v, Db_c[b * N + c], kad[a * ND + d]);
      kbc[b * NC + c] = fma(v, Da_d[a * N + d], kbc[b * NC + c]);
      kbd[b * ND + d] = fma(v, Da_c[a * N + c], kbd[b * ND + d]);
    }
#pragma unroll
    for (int i = 0; i < NC * ND; ++i) atomicAdd(t.ket.Jp + (size_t)i * sk + q, jcd[i]);
#pragma unroll
    for (int i = 0; i < NA * NC; ++i) krow_add(Krow_a, NA * N, (i / NC) * N + bfC + i % NC, kac[i], kscale);
#pragma unroll
    for (int i = 0; i < NA * ND; ++i) krow_add(Krow_a, NA * N, (i / ND) * N + bfD + i % ND, kad[i], kscale);
#pragma unroll
    for (int i = 0; i < NB * NC; ++i) krow_add(Krow_b, NB * N, (i / NC) * N + bfC + i % NC, kbc[i], kscale);
#pragma unroll
    for (int i = 0; i < NB * ND; ++i) krow_add(Krow_b, NB * N, (i / ND) * N + bfD + i % ND, kbd[i], kscale);
  };

  // Every warp draws chunks until the list is exhausted: general-code chunks (corrected, then
  // grid) first, the cheap uniform far-field chunks last, so the warps finish together.
  for (;;) {
    int c = 0;
    if (lane == 0) c = atomicAdd(&s_info[0], 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) break;
    const unsigned short e = s_list[c * 32 + lane];
    const int q = q0 + e;
    double out[C::kOut];
    int bfC, bfD;
    if (c >= far_chunk0) {
      if (e != kEmpty) {
        const double scale =
            shell_quartet<C, LA, LB, LC, LD, BOYS, true, true>(t, p, g, s_bra_far, q, out, bfC, bfD);
        digest(q, out, scale, bfC, bfD);
      }
    } else {
      if (e != kEmpty) {
        const double scale =
            shell_quartet<C, LA, LB, LC, LD, BOYS, true, false>(tl, p, g, s_bra, q, out, bfC, bfD);
        digest(q, out, scale, bfC, bfD);
      }
    }
  }

  // bra block of J: registers -> warp reduce -> one atomic per warp and component
#pragma unroll
  for (int i = 0; i < NA * NB; ++i) {
    const double s = warp_sum(jab[i]);
    if (lane == 0 && s != 0.0) atomicAdd(t.bra.Jp + (size_t)i * sb + p, s);
  }
  __syncthreads();
  // flush the K rows (only touched entries)
  for (int a = 0; a < NA; ++a)
    for (int j = tid; j < N; j += T) {
      const double v = krow_get(Krow_a, NA * N, a * N + j, kinv);
      if (v != 0.0) atomicAdd(t.Kh + (size_t)(bfA + a) * N + j, v);
    }
  for (int b = 0; b < NB; ++b)
    for (int j = tid; j < N; j += T) {
      const double v = krow_get(Krow_b, NB * N, b * N + j, kinv);
      if (v != 0.0) atomicAdd(t.Kh + (size_t)(bfB + b) * N + j, v);
    }
}


// ---------------------------------------------------------------------------------------
// Warp-per-bra-pair J/K kernel for "light" bra pairs (fewer surviving kets than one pass of the
// block kernel).  One warp owns one bra pair and ALL its kets: it sorts them by regime in its
// slice of shared memory exactly like the block kernel ([far | grid | corrected]) and walks the
// list 32 kets at a time, far-field chunks through primitive_quartet_far.  D and K are
// touched in global memory (scattered 8-byte loads / fp64 atomics), the bra block of J
// accumulates in registers over the whole ket list.
// Dynamic shared memory per warp: K2_bra primitive pairs + light_cap ints.
// ---------------------------------------------------------------------------------------
template <int LA, int LB, int LC, int LD, int BOYS>
__device__ __forceinline__ void jk_light_body(const EriTask& t, long long blk, double* smem) {
  using C = EriClass<LA, LB, LC, LD>;
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  constexpr bool kUnroll = C::kOut <= RCHEM_UNROLL_MAX;
  constexpr int kRegimes = BOYS == kBoysReference ? 3 : 2;

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long w = blk * kWarpsPerBlock + wib;
  if (w >= t.nlight) return;  // whole warp leaves together
  const int p = __ldg(t.lp + (int)w);
  const int nq = __ldg(t.nq + p);
  const int N = t.N, sb = t.bra.stride, sk = t.ket.stride;

  const int k2_staged = t.bra.K2;
  using BraPrim = PrimPairV<C::kNVb>;
  const size_t per_warp = (size_t)k2_staged * sizeof(BraPrim) + (size_t)t.light_cap * sizeof(int);
  BraPrim* s_bra = reinterpret_cast<BraPrim*>(reinterpret_cast<char*>(smem) +
                                              (size_t)wib * ((per_warp + 7) & ~(size_t)7));
  int* s_list = reinterpret_cast<int*>(s_bra + k2_staged);
  for (int k = lane; k < t.bra.K2; k += 32) s_bra[k] = load_prim<C::kNVb>(t.bra, k, p);
  BraPrim* s_bra_far = s_bra;
  const BraGeom g = load_bra_geom(t.bra, p);
  const int bfA = __ldg(t.bra.idx + p), bfB = __ldg(t.bra.idx + sb + p);

  // regime sort (warp-local: ballots and running cursors, no atomics)
  int n_far = 0;
  {
    const PairBoundF bb = load_bound(t.bra, p);
    const float xcorr = (float)(ref_exact_from(C::kL) + 2.0);
    int n0 = 0, n1 = 0;
    unsigned long long cls_bits = 0;  // 2 bits per pass of 32 kets (a light pair has < 1024 kets)
    int pass = 0;
    for (int base = 0; base < nq; base += 32, ++pass) {
      const int i = base + lane;
      const int cls = i < nq ? quartet_regime_f<kRegimes>(bb, load_bound(t.ket, i), (float)far_proven_x(C::kL), xcorr, t.far_sched) : 3;
      if (pass < 32) cls_bits |= (unsigned long long)cls << (2 * pass);
      n0 += __popc(__ballot_sync(0xffffffffu, cls == 0));
      n1 += __popc(__ballot_sync(0xffffffffu, cls == 1));
    }
    int cur[3] = {0, n0, n0 + n1};
    pass = 0;
    for (int base = 0; base < nq; base += 32, ++pass) {
      const int i = base + lane;
      const int cls = pass < 32 ? (int)((cls_bits >> (2 * pass)) & 3)
                                : (i < nq ? quartet_regime_f<kRegimes>(bb, load_bound(t.ket, i), (float)far_proven_x(C::kL), xcorr, t.far_sched) : 3);
#pragma unroll
      for (int c = 0; c < kRegimes; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (cls == c) s_list[cur[c] + __popc(m & ((1u << lane) - 1u))] = i;
        cur[c] += __popc(m);
      }
    }
    n_far = n0;
  }
  __syncwarp();

  double Dab[NA * NB], jab[NA * NB];
#pragma unroll
  for (int i = 0; i < NA * NB; ++i) {
    Dab[i] = __ldg(t.bra.Dp + (size_t)i * sb + p);
    jab[i] = 0.0;
  }
  const double* __restrict__ D = t.D;

  auto digest = [&](int q, const double* __restrict__ out, double scale, int bfC, int bfD) {
    double jcd[NC * ND], kac[NA * NC], kad[NA * ND], kbc[NB * NC], kbd[NB * ND];
    double Dcd[NC * ND], Dac[NA * NC], Dad[NA * ND], Dbc[NB * NC], Dbd[NB * ND];
#pragma unroll
    for (int i = 0; i < NC * ND; ++i) { jcd[i] = 0.0; Dcd[i] = __ldg(t.ket.Dp + (size_t)i * sk + q); }
#pragma unroll
    for (int i = 0; i < NA * NC; ++i) {
      kac[i] = 0.0;
      Dac[i] = __ldg(D + (size_t)(bfA + i / NC) * N + bfC + i % NC);
    }
#pragma unroll
    for (int i = 0; i < NA * ND; ++i) {
      kad[i] = 0.0;
      Dad[i] = __ldg(D + (size_t)(bfA + i / ND) * N + bfD + i % ND);
    }
#pragma unroll
    for (int i = 0; i < NB * NC; ++i) {
      kbc[i] = 0.0;
      Dbc[i] = __ldg(D + (size_t)(bfB + i / NC) * N + bfC + i % NC);
    }
#pragma unroll
    for (int i = 0; i < NB * ND; ++i) {
      kbd[i] = 0.0;
      Dbd[i] = __ldg(D + (size_t)(bfB + i / ND) * N + bfD + i % ND);
    }
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
#pragma unroll
    for (int i = 0; i < NC * ND; ++i) atomicAdd(t.ket.Jp + (size_t)i * sk + q, jcd[i]);
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
  };

  int it = lane;
  for (; it < n_far; it += 32) {
    const int q = s_list[it];
    double out[C::kOut];
    int bfC, bfD;
    const double scale =
        shell_quartet<C, LA, LB, LC, LD, BOYS, true, true>(t, p, g, s_bra_far, q, out, bfC, bfD);
    digest(q, out, scale, bfC, bfD);
  }
  for (; it < nq; it += 32) {
    const int q = s_list[it];
    double out[C::kOut];
    int bfC, bfD;
    const double scale =
        shell_quartet<C, LA, LB, LC, LD, BOYS, true, false>(t, p, g, s_bra, q, out, bfC, bfD);
    digest(q, out, scale, bfC, bfD);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NA * NB; ++i) {
    const double s = warp_sum(jab[i]);
    if (lane == 0 && s != 0.0) atomicAdd(t.bra.Jp + (size_t)i * sb + p, s);
  }
}

// Resident blocks per SM the light kernels are compiled for (register cap 65536/(128 n)):
// the kernel waits on scattered global D loads, so occupancy is worth some spills
// (measured on the headline workload: 1 -> 134.7 ms, 5 -> 133.7, 6 -> 133.5, 8 -> 136.0).
#ifndef RCHEM_LIGHT_MINB
#define RCHEM_LIGHT_MINB 6
#endif
// medium classes (10-18 targets, ~220 registers uncapped): measured on (H2O)96/6-31G
// 1 block of 128 threads per SM (uncapped) 80.6 ms, 3 (170 registers) 78.7 ms
#ifndef RCHEM_LIGHT_MINB_MED
#define RCHEM_LIGHT_MINB_MED 3
#endif
template <int LA, int LB, int LC, int LD> struct LightCfg {
  static constexpr int kTargets = EriClass<LA, LB, LC, LD>::kTargets;
  static constexpr int kMinBlocks =
      kTargets <= 9 ? RCHEM_LIGHT_MINB : (kTargets <= 18 ? RCHEM_LIGHT_MINB_MED : 1);
};
template <int LA, int LB, int LC, int LD, int BOYS>
__global__ void __launch_bounds__(kThreads, LightCfg<LA, LB, LC, LD>::kMinBlocks)
eri_jk_light_kernel(const EriTask t) {
  extern __shared__ double smem[];
  jk_light_body<LA, LB, LC, LD, BOYS>(t, (long long)blockIdx.x * t.nranks + t.rank, smem);
}

// All light tasks of one class in ONE launch: the per-task grids are mostly a few dozen
// blocks (ncu: 157 of 293 light launches under 150 blocks, 6 % of the SMs' warp slots in use).
// tasks[] lives in device memory; block b belongs to task ti with
// blk_prefix[ti] <= b < blk_prefix[ti + 1] and copies its descriptor into shared memory.
template <int LA, int LB, int LC, int LD, int BOYS>
__global__ void __launch_bounds__(kThreads, LightCfg<LA, LB, LC, LD>::kMinBlocks)
eri_jk_light_multi_kernel(const EriTask* __restrict__ tasks, const int* __restrict__ blk_prefix,
                          int ntasks) {
  extern __shared__ double smem[];
  __shared__ EriTask s_t;
  int lo = 0, hi = ntasks;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(blk_prefix + mid) <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int* src = reinterpret_cast<const int*>(tasks + lo);
  int* dst = reinterpret_cast<int*>(&s_t);
  for (int i = threadIdx.x; i < (int)(sizeof(EriTask) / sizeof(int)); i += kThreads) dst[i] = __ldg(src + i);
  __syncthreads();
  const long long local = (long long)((int)blockIdx.x - __ldg(blk_prefix + lo));
  jk_light_body<LA, LB, LC, LD, BOYS>(s_t, local * s_t.nranks + s_t.rank, smem);
}

// Host-side launcher signatures, one pair per class (eri_class_tu.cu)
typedef cudaError_t (*EriLaunchFn)(int boys, int mode, const EriTask& task, unsigned grid,
                                   cudaStream_t stream);
typedef cudaError_t (*EriBlockLaunchFn)(int boys, const EriTask& task, unsigned grid,
                                        size_t smem_bytes, cudaStream_t stream);
typedef cudaError_t (*EriLightLaunchFn)(int boys, const EriTask& task, unsigned grid,
                                        size_t smem_bytes, cudaStream_t stream);
typedef cudaError_t (*EriLightMultiLaunchFn)(int boys, const EriTask* tasks, const int* blk_prefix,
                                             int ntasks, unsigned grid, size_t smem_bytes,
                                             cudaStream_t stream);
struct EriBlockInfo { int threads; int kets_per_block; };

}  // namespace rchem
