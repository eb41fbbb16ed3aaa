// engine.cu -- host runtime behind the C ABI (include/rchem_eri.h): shell-pair batches in HBM,
// Schwarz bounds, the implicit screened quartet list, kernel launches, J/K finalisation.
//
// Reference correspondence: this file plays the role of the loop nests in src/basis.rs
// (JK_direct 383-428, build_I 430-460, JK_inmem 462-484) -- re-designed as batched GPU work.
// There is no CPU compute path: without a CUDA device every compute call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rchem_eri.h"
#include "basis_model.h"
#include "eri_kernel.cuh"
#include "gen/eri_class_list.h"
#include "pair_build.h"
#include "boys_delta_build.h"

namespace rchem {

// launchers, one per class translation unit
// (one translation unit per class and Boys flavour: eri_class_tu.cu)
#define X(la, lb, lc, ld, tag)                                                              \
  cudaError_t launch_eri_##tag##_b0(int, const EriTask&, unsigned, cudaStream_t);           \
  cudaError_t launch_eri_##tag##_b1(int, const EriTask&, unsigned, cudaStream_t);           \
  cudaError_t launch_eri_block_##tag##_b0(const EriTask&, unsigned, size_t, cudaStream_t);  \
  cudaError_t launch_eri_block_##tag##_b1(const EriTask&, unsigned, size_t, cudaStream_t);  \
  cudaError_t launch_eri_light_##tag##_b0(const EriTask&, unsigned, size_t, cudaStream_t);  \
  cudaError_t launch_eri_light_##tag##_b1(const EriTask&, unsigned, size_t, cudaStream_t);  \
  cudaError_t launch_eri_light_multi_##tag##_b0(const EriTask*, const int*, int, unsigned,  \
                                                size_t, cudaStream_t);                      \
  cudaError_t launch_eri_light_multi_##tag##_b1(const EriTask*, const int*, int, unsigned,  \
                                                size_t, cudaStream_t);                      \
  EriBlockInfo block_info_##tag();                                                          \
  static cudaError_t launch_eri_##tag(int boys, int mode, const EriTask& t, unsigned g,     \
                                      cudaStream_t s) {                                     \
    return boys == kBoysReference ? launch_eri_##tag##_b0(mode, t, g, s)                    \
                                  : launch_eri_##tag##_b1(mode, t, g, s);                   \
  }                                                                                         \
  static cudaError_t launch_eri_block_##tag(int boys, const EriTask& t, unsigned g,         \
                                            size_t smem, cudaStream_t s) {                  \
    return boys == kBoysReference ? launch_eri_block_##tag##_b0(t, g, smem, s)              \
                                  : launch_eri_block_##tag##_b1(t, g, smem, s);             \
  }                                                                                         \
  static cudaError_t launch_eri_light_##tag(int boys, const EriTask& t, unsigned g,         \
                                            size_t smem, cudaStream_t s) {                  \
    return boys == kBoysReference ? launch_eri_light_##tag##_b0(t, g, smem, s)              \
                                  : launch_eri_light_##tag##_b1(t, g, smem, s);             \
  }                                                                                         \
  static cudaError_t launch_eri_light_multi_##tag(int boys, const EriTask* t, const int* pf, \
                                                  int n, unsigned g, size_t smem,           \
                                                  cudaStream_t s) {                         \
    return boys == kBoysReference ? launch_eri_light_multi_##tag##_b0(t, pf, n, g, smem, s) \
                                  : launch_eri_light_multi_##tag##_b1(t, pf, n, g, smem, s); \
  }
RCHEM_ERI_CLASSES(X)
#undef X

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int fail_public(int code, const std::string& msg) { return fail(code, msg); }
// one_electron.cu
int one_electron_host(int n, const double* origins, const int* powers, const int* prim_offset,
                      const double* exps, const double* coefs, const double* norms, int which,
                      int natoms, const double* atomcoords, const double* charges, double* M,
                      int device);
#define CUDA_OK(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(RCHEM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));     \
  } while (0)

// SURVEY section 8(d) flop model: per primitive quartet P, per contracted quartet H.
struct FlopModel { int la, lb, lc, ld, P, H; };
static const FlopModel kFlopModel[] = {
    {0, 0, 0, 0, 51, 0},       {1, 0, 0, 0, 66, 0},        {1, 0, 1, 0, 118, 0},
    {1, 1, 0, 0, 115, 18},     {1, 1, 1, 0, 281, 54},      {1, 1, 1, 1, 792, 324},
    {2, 0, 0, 0, 112, 0},      {2, 0, 1, 0, 239, 0},       {2, 0, 1, 1, 627, 108},
    {2, 0, 2, 0, 609, 0},      {2, 1, 0, 0, 201, 36},      {2, 1, 1, 0, 516, 108},
    {2, 1, 1, 1, 1486, 648},   {2, 1, 2, 0, 1438, 216},    {2, 1, 2, 1, 3286, 1224},
    {2, 2, 0, 0, 355, 158},    {2, 2, 1, 0, 964, 474},     {2, 2, 1, 1, 2834, 2070},
    {2, 2, 2, 0, 2741, 948},   {2, 2, 2, 1, 6263, 3824},   {2, 2, 2, 2, 12280, 10586},
};
static const FlopModel* flop_model(int la, int lb, int lc, int ld) {
  for (const FlopModel& f : kFlopModel)
    if (f.la == la && f.lb == lb && f.lc == lc && f.ld == ld) return &f;
  return nullptr;
}
// The model for a quartet of shell TYPES: a fused sp shell stands for its s and its p shell, so a
// fused quartet is charged the SURVEY figures of every segmented (s/p/d) quartet it evaluates --
// the algorithmic work of the reference's shells; what the fused kernel saves by sharing the
// primitive quartet between them is the optimisation.  *nseg = segmented quartets covered.
static bool flop_model_types(int ta, int tb, int tc, int td, double* P, double* H, int* nseg) {
  *P = *H = 0.0;
  *nseg = 0;
  for (int va = 0; va < nvariants(ta); ++va)
    for (int vb = 0; vb < nvariants(tb); ++vb)
      for (int vc = 0; vc < nvariants(tc); ++vc)
        for (int vd = 0; vd < nvariants(td); ++vd) {
          int la = variant_l(ta, va), lb = variant_l(tb, vb), lc = variant_l(tc, vc), ld = variant_l(td, vd);
          if (la < lb) std::swap(la, lb);
          if (lc < ld) std::swap(lc, ld);
          if (la * (la + 1) / 2 + lb < lc * (lc + 1) / 2 + ld) { std::swap(la, lc); std::swap(lb, ld); }
          const FlopModel* f = flop_model(la, lb, lc, ld);
          if (!f) return false;
          *P += f->P;
          *H += f->H;
          *nseg += 1;
        }
  return true;
}

static EriLaunchFn find_launcher(int la, int lb, int lc, int ld) {
#define X(a, b, c, d, tag) \
  if (la == a && lb == b && lc == c && ld == d) return launch_eri_##tag;
  RCHEM_ERI_CLASSES(X)
#undef X
  return nullptr;
}
static EriBlockLaunchFn find_block_launcher(int la, int lb, int lc, int ld, EriBlockInfo* info) {
#define X(a, b, c, d, tag)                                  \
  if (la == a && lb == b && lc == c && ld == d) {           \
    *info = block_info_##tag();                             \
    return launch_eri_block_##tag;                          \
  }
  RCHEM_ERI_CLASSES(X)
#undef X
  return nullptr;
}
static EriLightLaunchFn find_light_launcher(int la, int lb, int lc, int ld) {
#define X(a, b, c, d, tag) \
  if (la == a && lb == b && lc == c && ld == d) return launch_eri_light_##tag;
  RCHEM_ERI_CLASSES(X)
#undef X
  return nullptr;
}
static EriLightMultiLaunchFn find_light_multi_launcher(int la, int lb, int lc, int ld) {
#define X(a, b, c, d, tag) \
  if (la == a && lb == b && lc == c && ld == d) return launch_eri_light_multi_##tag;
  RCHEM_ERI_CLASSES(X)
#undef X
  return nullptr;
}
// dynamic shared memory the block kernel may use (227 KB per CTA on sm_100, minus slack)
static constexpr size_t kMaxBlockSmem = 220 * 1024;

struct Batch {
  int la = 0, lb = 0, K2 = 0, npairs = 0, stride = 0;
  // A "sub" batch holds the SEGMENTED twins -- (s s|, (p s|, (p p| -- of the fused (sp sp| shell
  // pairs: a quartet of two such pairs has 289 contraction accumulators, far beyond the register
  // file, so those quartets are evaluated through the segmented classes instead (J/K mode).
  // shA/shB of a sub batch index rchem_basis::seg_shells, Q is the PARENT pair's bound (so the
  // screened set is exactly the fused list's).
  bool sub = false;
  std::vector<int> shA, shB;
  std::vector<double> Q;
  double* d_prim = nullptr;
  double* d_geom = nullptr;
  float4* d_bnd = nullptr;   // single-precision bounding spheres (regime classification)
  float* d_zminf = nullptr;
  int* d_idx = nullptr;
  double* d_Dp = nullptr;  // [ncart(la)*ncart(lb)][stride] packed D blocks
  double* d_Jp = nullptr;  // [ncart(la)*ncart(lb)][stride] packed J blocks
  int ncomp() const { return ncart(la) * ncart(lb); }
  int nv() const { return nvariants(la) * nvariants(lb); }                      // weight variants
  int nfields() const { return kPrimFieldsBase + nv(); }                        // SoA arrays per primitive pair
  size_t prim_bytes() const { return (size_t)nfields() * sizeof(double); }      // sizeof(PrimPairV<nv>)
  BatchView view() const {
    return BatchView{d_prim, d_geom, d_bnd, d_zminf, d_idx, d_Dp, d_Jp, npairs, stride, K2};
  }
};

struct TaskTable {
  int bra = 0, ket = 0;
  bool sub = false;      // both batches are sub batches (J/K mode only)
  bool jk_skip = false;  // (sp sp|sp sp): evaluated through the sub batches in J/K mode
  long long nwarps = 0, nquartets = 0, nquartets_all = 0;
  long long* d_prefix = nullptr;  // warp chunks, every bra pair (tensor mode)
  int* d_nq = nullptr;
  std::vector<long long> h_prefix;  // kept for rchem_quartet_list
  std::vector<int> h_nq;
  // J/K split: heavy bra pairs -> block kernel, light ones -> warp kernel
  long long nwarps_light = 0, nquartets_light = 0, nblocks_heavy = 0;
  int nheavy = 0;
  long long* d_prefix_light = nullptr;
  int* d_nq_light = nullptr;
  int* d_hp = nullptr;
  long long* d_hblk_prefix = nullptr;
  size_t smem_bytes = 0;
  // light bra pairs with >= 1 ket, for the warp-per-bra-pair kernel (0 = use the chunk kernel)
  int* d_lp = nullptr;
  int nlight = 0, light_cap = 0;
  size_t light_smem = 0;
  // host copies of the split, for the exact per-rank statistics (quartets_of_rank)
  std::vector<int> h_nq_light, h_lp, h_hp;
  std::vector<long long> h_hblk;
  int kets_per_block = 0;
  // cache: shell quartets this rank evaluates in each kernel family
  int stat_rank = -1, stat_nranks = -1;
  long long q_tensor_mine = 0, q_light_mine = 0, q_heavy_mine = 0;
};

// Shell quartets of the 32-ket chunks [w0, w0 + nw) of one bra pair with nq kets that belong to
// `rank`: the chunk kernel gives thread block b (4 consecutive chunks of the task) to rank
// b % nranks (eri_kernel.cuh eri_kernel).  All chunks are full except the last one.
static long long owned_chunk_quartets(long long w0, long long nw, int nq, int rank, int nranks) {
  if (nw <= 0) return 0;
  auto owned_below = [&](long long w) {  // chunks in [0, w) owned by rank
    const long long nb = w / kWarpsPerBlock, rem = w % kWarpsPerBlock;
    long long full = nb > rank ? (nb - rank + nranks - 1) / nranks : 0;  // whole blocks owned
    return full * kWarpsPerBlock + ((nb % nranks) == rank ? rem : 0);
  };
  const long long owned = owned_below(w0 + nw) - owned_below(w0);
  const long long last = w0 + nw - 1;
  const bool last_mine = ((last / kWarpsPerBlock) % nranks) == rank;
  return owned * 32 - (last_mine ? (nw * 32 - nq) : 0);
}

// Exact shell-quartet counts of rank `rank` of `nranks` for one task (cached in the table).
static void quartets_of_rank(TaskTable& tt, int rank, int nranks) {
  if (tt.stat_rank == rank && tt.stat_nranks == nranks) return;
  tt.stat_rank = rank; tt.stat_nranks = nranks;
  tt.q_tensor_mine = tt.q_light_mine = tt.q_heavy_mine = 0;
  const int np = (int)tt.h_nq.size();
  if (nranks == 1) {
    tt.q_tensor_mine = tt.nquartets;
    tt.q_light_mine = tt.nquartets_light;
    tt.q_heavy_mine = tt.nquartets - tt.nquartets_light;
    return;
  }
  for (int p = 0; p < np; ++p)
    tt.q_tensor_mine += owned_chunk_quartets(tt.h_prefix[p], tt.h_prefix[p + 1] - tt.h_prefix[p],
                                             tt.h_nq[p], rank, nranks);
  if (tt.nlight > 0) {
    // warp-per-bra-pair kernel: light entry i runs in block i / 4, block b -> rank b % nranks
    for (int i = 0; i < tt.nlight; ++i)
      if ((i / kWarpsPerBlock) % nranks == rank) tt.q_light_mine += tt.h_nq_light[tt.h_lp[i]];
  } else {
    long long w0 = 0;
    for (int p = 0; p < np; ++p) {
      const long long nw = (tt.h_nq_light[p] + 31) / 32;
      tt.q_light_mine += owned_chunk_quartets(w0, nw, tt.h_nq_light[p], rank, nranks);
      w0 += nw;
    }
  }
  for (size_t hh = 0; hh < tt.h_hp.size(); ++hh) {
    // block kernel: heavy entry hh owns blocks [hblk[hh], hblk[hh+1]), each `per` kets
    const int nq = tt.h_nq[tt.h_hp[hh]];
    const long long first = tt.h_hblk[hh];
    const int nb = (int)(tt.h_hblk[hh + 1] - first);
    const int per = (((nq + nb - 1) / nb) + 31) & ~31;
    for (int j = 0; j < nb; ++j)
      if ((first + j) % nranks == rank)
        tt.q_heavy_mine += std::max(0, std::min(nq, (j + 1) * per) - j * per);
  }
}

// Per-batch view for the pack / finalise kernels, which cover EVERY batch in one launch (a build
// used to spend ~0.9 ms of its critical path on 3 x 86 tiny launches and memsets: nothing at
// N = 1, 8 % of the step at 8 GPUs, where every device repeats them).
struct PackDesc {
  const int* idx;
  double* Dp;
  double* Jp;
  int npairs, stride, nb, ncomp;
  int blk0;  // first thread block of this batch (256 pairs per block); desc[nbatch].blk0 = total
};
__device__ __forceinline__ int pack_desc_of_block(const PackDesc* __restrict__ desc, int nbatch, int blk) {
  int lo = 0, hi = nbatch;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (desc[mid].blk0 <= blk) lo = mid; else hi = mid;
  }
  return lo;
}

// D blocks of every shell pair, pair-major packed: Dp[ab][p] = D[bfA+a][bfB+b]
__global__ void pack_d_kernel(const double* __restrict__ D, int N, const PackDesc* __restrict__ desc,
                              int nbatch) {
  const PackDesc d = desc[pack_desc_of_block(desc, nbatch, blockIdx.x)];
  const int p = (blockIdx.x - d.blk0) * blockDim.x + threadIdx.x;
  if (p >= d.npairs) return;
  const int bfA = d.idx[p], bfB = d.idx[d.stride + p];
  for (int c = 0; c < d.ncomp; ++c)
    d.Dp[(size_t)c * d.stride + p] = D[(size_t)(bfA + c / d.nb) * N + bfB + c % d.nb];
}

// J from the packed pair blocks: J[i][j] = J[j][i] += Jp[ab] (+ Jp[ba] on a diagonal shell pair,
// whose block holds both orders) -- i.e. J = Jh + Jh^T of the digestion.  A function pair can
// receive shares from two batches (a fused (sp sp| pair and its segmented twins), so the
// accumulation is atomic; J starts from zero.
__global__ void finalize_j_kernel(const PackDesc* __restrict__ desc, int nbatch, int N,
                                  double* __restrict__ J) {
  const PackDesc d = desc[pack_desc_of_block(desc, nbatch, blockIdx.x)];
  const int p = (blockIdx.x - d.blk0) * blockDim.x + threadIdx.x;
  if (p >= d.npairs) return;
  const int bfA = d.idx[p], bfB = d.idx[d.stride + p], diag = d.idx[2 * d.stride + p];
  for (int c = 0; c < d.ncomp; ++c) {
    const int a = c / d.nb, b = c % d.nb;
    double v = d.Jp[(size_t)c * d.stride + p];
    if (diag) v += d.Jp[(size_t)(b * d.nb + a) * d.stride + p];
    if (v == 0.0) continue;
    atomicAdd(J + (size_t)(bfA + a) * N + bfB + b, v);
    if (!diag) atomicAdd(J + (size_t)(bfB + b) * N + bfA + a, v);
  }
}

// max |D| (the K-row fixed-point scale of the block kernel needs it): non-negative doubles
// order like their bit patterns
__global__ void absmax_kernel(const double* __restrict__ D, size_t n, unsigned long long* out) {
  unsigned long long m = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = max(m, (unsigned long long)__double_as_longlong(fabs(D[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// K = Kh + sign Kh^T: the two transposed halves of the 8-fold digestion.  sign = +1 for a
// symmetric density; for an ANTISYMMETRIC density the transposed half changes sign
// (K[c,a] += (cd|ab) D[d,b] = -(ab|cd) D[b,d]) and the result is ADDED to K (rchem_jk_direct).
__global__ void finalize_k_kernel(const double* __restrict__ Kh, double* __restrict__ K, int N,
                                  double sign, int accumulate) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nn = (size_t)N * N;
  if (idx >= nn) return;
  const size_t i = idx / N, j = idx % N;
  const double v = fma(sign, Kh[j * N + i], Kh[idx]);
  K[idx] = accumulate ? K[idx] + v : v;
}

// Single-process multi-GPU J/K (RCHEM_OPT_NGPUS): device 0 of the group sums the partial [J|K]
// of its peers straight out of THEIR memory over NVLink (peer-to-peer loads, 16 bytes per
// access); every address is read exactly once, so there is nothing to stage.
struct PeerPtrs { const double2* p[15]; int n; };
__global__ void __launch_bounds__(256) peer_reduce_kernel(double2* __restrict__ JK, const PeerPtrs peers,
                                                          size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 s = JK[i];
#pragma unroll 4
    for (int k = 0; k < peers.n; ++k) {
      const double2 v = peers.p[k][i];
      s.x += v.x;
      s.y += v.y;
    }
    JK[i] = s;
  }
}

// {max|D|, max|D - D^T|} as bit patterns (non-negative doubles order like their bits), and the
// split D = S + A into symmetric and antisymmetric parts
__global__ void asym_probe_kernel(const double* __restrict__ D, int N, unsigned long long* out) {
  const size_t nn = (size_t)N * N;
  unsigned long long m = 0, a = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / N, c = i % N;
    const double v = D[i];
    m = max(m, (unsigned long long)__double_as_longlong(fabs(v)));
    if (c < r) a = max(a, (unsigned long long)__double_as_longlong(fabs(v - D[c * N + r])));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (m) atomicMax(out, m);
    if (a) atomicMax(out + 1, a);
  }
}
__global__ void split_density_kernel(const double* __restrict__ D, int N, double* __restrict__ S,
                                     double* __restrict__ A) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * N) return;
  const size_t i = idx / N, j = idx % N;
  const double x = D[idx], y = D[j * N + i];
  S[idx] = 0.5 * (x + y);
  A[idx] = 0.5 * (x - y);
}

// Dense-tensor completion (build_I, basis.rs:451-454 plus the bra<->ket swap): the ERI kernels
// write each unique integral once, at its canonical position (stored orientation of both
// shell pairs, bra pair >= ket pair in kernel order).  This kernel fills every other element
// from its canonical image.
//   fn_shell[i]      : shell of function i
//   pair_key[s*ns+t] : kernel-order rank of shell pair {s,t} (batch << 32 | position)
//   pair_fwd[s*ns+t] : 1 when the pair is stored as (A=s, B=t)
// Tiling (round 2; the first version gave each block one (i,j) and 256 consecutive kl and read
// 4.7x its algorithmic bytes from DRAM): a block owns a 4 x 4 group of (i,j) and a 16 x 16 tile of
// (k,l), one thread per (k,l) walking the 16 (i,j).  The sources with k and l swapped -- stride N
// along l -- use every sector they touch inside the block, and the bra<->ket transposed sources
// -- 8-byte reads at column (i,j) or (j,i) of row (k,l) -- find the other three elements of their
// 32-byte sector in the same thread's next iterations (L1), whichever way the pair is stored.
constexpr int kFillTile = 16;   // (k,l) tile edge: 256 threads
constexpr int kFillGroup = 4;   // (i,j) group edge (8 x 8 measured: L1 thrashes, 80 GB read instead of 47)
__global__ void __launch_bounds__(kFillTile * kFillTile)
tensor_fill_kernel(double* __restrict__ I, int N, int ns, const int* __restrict__ fn_shell,
                   const long long* __restrict__ pair_key,
                   const unsigned char* __restrict__ pair_fwd) {
  const int ng = (N + kFillGroup - 1) / kFillGroup;          // groups per index
  const int i0 = (blockIdx.x / ng) * kFillGroup, j0 = (blockIdx.x % ng) * kFillGroup;
  const int nt = (N + kFillTile - 1) / kFillTile;            // (k,l) tiles per index
  const int k = (blockIdx.y / nt) * kFillTile + threadIdx.x / kFillTile;
  const int l = (blockIdx.y % nt) * kFillTile + threadIdx.x % kFillTile;
  if (k >= N || l >= N) return;
  const int sk = fn_shell[k], sl = fn_shell[l];
  const long long kq = pair_key[sk * ns + sl];
  // stored orientation of each pair; inside a diagonal shell pair (both orders were computed)
  // keep the order with the larger first function so the tensor is bitwise symmetric
  const bool fq = sk == sl ? k >= l : (bool)pair_fwd[sk * ns + sl];
  const size_t ck = fq ? k : l, cl = fq ? l : k;
  const size_t n = (size_t)N;
#pragma unroll 1
  for (int di = 0; di < kFillGroup; ++di) {
    const int i = i0 + di;
    if (i >= N) break;
    const int si = fn_shell[i];
#pragma unroll
    for (int dj = 0; dj < kFillGroup; ++dj) {
      const int j = j0 + dj;
      if (j >= N) break;
      const int sj = fn_shell[j];
      const long long kp = pair_key[si * ns + sj];
      const bool fp = si == sj ? i >= j : (bool)pair_fwd[si * ns + sj];
      const size_t ci = fp ? i : j, cj = fp ? j : i;
      // inside one shell pair (kp == kq) both (ab|cd) and (cd|ab) were computed; keep the one
      // with the larger leading function pair so the tensor is bitwise symmetric under bra <-> ket
      const bool bra_first = kp > kq || (kp == kq && ci * n + cj >= ck * n + cl);
      const size_t src = bra_first ? ((ci * n + cj) * n + ck) * n + cl : ((ck * n + cl) * n + ci) * n + cj;
      const size_t dst = (((size_t)i * n + j) * n + k) * n + l;
      if (src != dst) I[dst] = I[src];
    }
  }
}

// JK_inmem (basis.rs:462-484) in ONE pass over the tensor: element I[i][j][k][l] feeds
// J[i][j] (with D[k][l]) and K[i][k] (with D[j][l]).  One block per (i,j) row of N^2 values;
// one warp per k sums over l.  HBM-bound: 8 N^4 bytes read once.
// U = independent streaming loads in flight per lane (the tensor is read exactly once);
// out-of-range slots are predicated off instead of peeled.
template <int U>
__global__ void __launch_bounds__(256) jk_inmem_kernel(const double* __restrict__ I,
                                                       const double* __restrict__ D,
                                                       double* __restrict__ JK, int N) {
  const size_t nn = (size_t)N * N;
  const int i = blockIdx.x / N, j = blockIdx.x % N;
  const double* row = I + (size_t)blockIdx.x * nn;
  const double* Dj = D + (size_t)j * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double jsum = 0.0;
  for (int k = warp; k < N; k += nwarps) {
    const double* rk = row + (size_t)k * N;
    const double* Dk = D + (size_t)k * N;
    double ks = 0.0;
    for (int l = lane; l < N; l += 32 * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (l + 32 * u < N) ? __ldcs(rk + l + 32 * u) : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (l + 32 * u < N) {
          jsum = fma(v[u], __ldg(Dk + l + 32 * u), jsum);
          ks = fma(v[u], __ldg(Dj + l + 32 * u), ks);
        }
    }
    ks = warp_sum(ks);
    if (lane == 0) atomicAdd(JK + nn + (size_t)i * N + k, ks);
  }
  jsum = warp_sum(jsum);
  if (lane == 0) atomicAdd(JK + (size_t)i * N + j, jsum);
}

// materialises the implicit quartet list of one task: (p, q) for q < nq[p]
__global__ void quartet_list_kernel(const long long* __restrict__ qprefix,
                                    const int* __restrict__ nq, int npairs, long long pair_off_b,
                                    long long pair_off_k, long long* __restrict__ out) {
  const int p = blockIdx.x;
  if (p >= npairs) return;
  const long long base = qprefix[p];
  for (int q = threadIdx.x; q < nq[p]; q += blockDim.x) {
    out[2 * (base + q)] = pair_off_b + p;
    out[2 * (base + q) + 1] = pair_off_k + q;
  }
}

}  // namespace rchem

using namespace rchem;

struct rchem_basis {
  Basis basis;
  ShellSet shells;
  int N = 0;
  int nprim = 0;
  // options
  int boys = kBoysReference;
  double tau = 0.0;
  double prim_eps = kPrimPairEps;
  int device = 0;
  int light_kernel = [] {  // RCHEM_OPT_LIGHT_KERNEL; the environment sets the default (tuning)
    const char* e = std::getenv("RCHEM_LIGHT");
    return e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }();
  double heavy_passes = [] {  // RCHEM_OPT_HEAVY_PASSES; the environment sets the default (tuning)
    const char* e = std::getenv("RCHEM_HEAVY_PASSES");
    return e ? std::max(0.0, atof(e)) : 2.0;
  }();
  int far_sched = [] {  // RCHEM_OPT_FAR_SCHED; the environment sets the default (tuning)
    const char* e = std::getenv("RCHEM_FAR");
    return e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }();
  // device state
  bool ready = false;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  bool external_stream = false;
  std::vector<Batch> batches;      // normal batches first, then the sub batches (Batch::sub)
  std::vector<Shell> seg_shells;   // s and p parts of fused shells, for the sub batches
  rchem_basis* tensor_twin = nullptr;  // segmented clone that builds the dense tensor (build_I)
  const rchem_basis* stats_src = nullptr;  // handle whose events time the last build (this one if null)
  std::vector<TaskTable> tasks;
  double tasks_tau = -1.0;
  double* d_boys = nullptr;      // exact-Boys grids, one per L
  double* d_delta_thr = nullptr; // boys_delta.h tables
  double* d_delta_direct = nullptr;
  double *d_D = nullptr, *d_Kh = nullptr, *d_JK = nullptr;
  double* d_dmax = nullptr;  // max|D| of the current build (device scalar)
  // packed D / J blocks of all batches: two arenas, one descriptor table (pack_d / finalize_j)
  double *d_Dp_all = nullptr, *d_Jp_all = nullptr;
  size_t packed_doubles = 0;
  PackDesc* d_pack_desc = nullptr;
  int pack_blocks = 0;
  // merged light launches (one per class): task descriptors and block prefixes
  EriTask* d_light_tasks = nullptr;
  int* d_light_prefix = nullptr;
  EriTask* h_light_tasks = nullptr;  // pinned staging
  int* h_light_prefix = nullptr;
  cudaEvent_t ev_light = nullptr;
  size_t light_tasks_cap = 0;
  double kbound = 0.0;       // 16 max_X sum_Y n_Y Q_XY (EriTask::kbound)
  // maps for tensor_fill_kernel
  int* d_fn_shell = nullptr;
  long long* d_pair_key = nullptr;
  unsigned char* d_pair_fwd = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_stream = nullptr;  // orders a build after the previous one across rchem_set_stream
  double* d_asym = nullptr;         // {max|D|, max|D - D^T|} of the host-buffer entry point
  double* h_asym = nullptr;         // pinned copy of it (so the device-to-host copy never blocks the host)
  double setup_ms = 0.0;            // host wall time of pair build + Schwarz + task tables
  cudaStream_t last_stream = nullptr;  // stream of the previous build (order_after_previous_build)
  bool last_stream_valid = false;
  int symmetric_only = 0;           // RCHEM_OPT_SYMMETRIC_D_ONLY
  int fuse_sp = [] {                // RCHEM_OPT_FUSE_SP; the environment sets the default (tuning)
    const char* e = std::getenv("RCHEM_FUSE_SP");
    return e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }();
  // single-process multi-GPU (RCHEM_OPT_NGPUS): this handle drives devices device ..
  // device + ngpus - 1; peers[i-1] is a full clone of the handle living on device + i
  int ngpus = 1;
  std::vector<rchem_basis*> peers;
  std::vector<double*> peer_stage;  // on this device: copy of a peer's [J|K] when P2P loads are not possible
  cudaEvent_t ev_D = nullptr;       // D is resident on this device
  cudaEvent_t ev_done = nullptr;    // this handle's share of [J|K] is complete (peer side)
  int group_p2p = -1;               // every device of the group can read every other one's memory (-1 = not probed)
  bool peer_access_on = false;      // cudaDeviceEnablePeerAccess done from this handle's device
  double tasks_heavy = -1.0;        // values the task tables were built with
  int tasks_light = -1;
  // The tasks of one J/K (or tensor) build are independent kernels; they are spread over a
  // few auxiliary streams so the tail of one launch overlaps the head of the next.
  static constexpr int kAuxStreams = 16;  // upper bound; n_aux (6 or 12, RCHEM_STREAMS) are used
  int n_aux = 12;
  cudaStream_t aux[kAuxStreams] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kAuxStreams] = {};
  rchem_stats stats{};
};

namespace {

int upload_batch(rchem_basis* h, Batch& bt, const std::vector<int>& order) {
  // order: permutation of pair slots (Schwarz-descending once Q is known)
  const int np = bt.npairs;
  bt.stride = (np + 31) / 32 * 32;
  const size_t st = bt.stride;
  const int nf = bt.nfields();
  std::vector<double> prim(nf * (size_t)bt.K2 * st, 0.0), geom(kGeomFields * st, 0.0);
  std::vector<int> idx(4 * st, 0);
  std::vector<float4> bnd(st);
  std::vector<float> zminf(st);
  std::vector<PrimPair> pps;
  std::vector<int> shA(np), shB(np);
  std::vector<double> Q(bt.Q.empty() ? 0 : np);
  for (int s = 0; s < np; ++s) {
    const int src = order[s];
    const std::vector<Shell>& pool = bt.sub ? h->seg_shells : h->shells.shells;
    const Shell& A = pool[bt.shA[src]];
    const Shell& B = pool[bt.shB[src]];
    shA[s] = bt.shA[src];
    shB[s] = bt.shB[src];
    if (!bt.Q.empty()) Q[s] = bt.Q[src];
    build_significant_prim_pairs(A, B, h->prim_eps, &pps);  // pps.size() == bt.K2 by construction
    for (int k = 0; k < bt.K2; ++k) {
      const PrimPair& pp = pps[k];
      const double f[kPrimFieldsBase + kMaxPairVariants] = {pp.zeta, pp.rzeta, pp.Px, pp.Py, pp.Pz, pp.fsc,
                                                            pp.pfar, pp.w[0], pp.w[1], pp.w[2], pp.w[3]};
      for (int c = 0; c < nf; ++c) prim[((size_t)c * bt.K2 + k) * st + s] = f[c];
    }
    const PairBound pb = bound_prim_pairs(pps);
    for (int d = 0; d < 3; ++d) {
      geom[d * st + s] = A.ctr[d];
      geom[(3 + d) * st + s] = A.ctr[d] - B.ctr[d];
      geom[(6 + d) * st + s] = pb.M[d];
    }
    geom[9 * st + s] = pb.rad;
    geom[10 * st + s] = pb.zmin;
    const PairBoundF bf = make_pair_bound_f(pb.M[0], pb.M[1], pb.M[2], pb.rad, pb.zmin);
    bnd[s] = make_float4(bf.Mx, bf.My, bf.Mz, bf.rad);
    zminf[s] = bf.zmin;
    geom[11 * st + s] = bt.Q.empty() ? 0.0 : Q[s];
    idx[s] = A.bf0;
    idx[st + s] = B.bf0;
    idx[2 * st + s] = (bt.shA[src] == bt.shB[src]) ? 1 : 0;
    // packed form for the J/K block / light kernels (only read when N < 32768)
    idx[3 * st + s] = (int)(((unsigned)A.bf0 & 0x7fffu) | (((unsigned)B.bf0 & 0x7fffu) << 15) |
                            ((unsigned)idx[2 * st + s] << 30));
  }
  // padding slots replicate pair 0 so stray reads stay finite
  for (int s = np; s < (int)st; ++s) {
    for (size_t c = 0; c < nf * (size_t)bt.K2; ++c) prim[c * st + s] = prim[c * st];
    for (int c = 0; c < kGeomFields; ++c) geom[c * st + s] = geom[c * st];
    bnd[s] = bnd[0];
    zminf[s] = zminf[0];
    for (int c = 0; c < 4; ++c) idx[c * st + s] = idx[c * st];
  }
  bt.shA.swap(shA);
  bt.shB.swap(shB);
  if (!bt.Q.empty()) bt.Q.swap(Q);
  if (!bt.d_prim) {
    CUDA_OK(cudaMalloc(&bt.d_prim, prim.size() * sizeof(double)));
    CUDA_OK(cudaMalloc(&bt.d_geom, geom.size() * sizeof(double)));
    CUDA_OK(cudaMalloc(&bt.d_bnd, bnd.size() * sizeof(float4)));
    CUDA_OK(cudaMalloc(&bt.d_zminf, zminf.size() * sizeof(float)));
    CUDA_OK(cudaMalloc(&bt.d_idx, idx.size() * sizeof(int)));
  }
  CUDA_OK(cudaMemcpyAsync(bt.d_prim, prim.data(), prim.size() * sizeof(double),
                          cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(bt.d_geom, geom.data(), geom.size() * sizeof(double),
                          cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(bt.d_bnd, bnd.data(), bnd.size() * sizeof(float4), cudaMemcpyHostToDevice,
                          h->stream));
  CUDA_OK(cudaMemcpyAsync(bt.d_zminf, zminf.data(), zminf.size() * sizeof(float),
                          cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(bt.d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice,
                          h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));  // host vectors die at scope exit
  return RCHEM_OK;
}

void fill_common(const rchem_basis* h, EriTask* t) {
  std::memset(t, 0, sizeof(*t));
  t->N = h->N;
  t->boys.exact = h->d_boys;
  t->boys.delta.thr = h->d_delta_thr;
  t->boys.delta.direct = h->d_delta_direct;
  t->nranks = 1;
  t->far_sched = h->far_sched;
  for (int l = 0; l < kNumTypes; ++l)
    for (int k = 0; k < 6; ++k)
      t->compscale[l][k] = (l < 3 && l <= h->shells.lmax) ? h->shells.compscale[l][k] : 1.0;
}

void free_tasks(rchem_basis* h);

// Releases every device-side resource of the handle (by non-null pointer, so it is safe after
// a partially failed ensure_ready) except the handle's own stream.
void release_device_state(rchem_basis* h) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return; }
  if (h->device >= 0 && h->device < ndev) cudaSetDevice(h->device);
  free_tasks(h);
  for (Batch& bt : h->batches) {
    cudaFree(bt.d_prim); cudaFree(bt.d_geom); cudaFree(bt.d_bnd); cudaFree(bt.d_zminf); cudaFree(bt.d_idx);
  }
  h->batches.clear();
  auto drop = [](auto*& ptr) { if (ptr) cudaFree(ptr); ptr = nullptr; };
  drop(h->d_boys); drop(h->d_delta_thr); drop(h->d_delta_direct); drop(h->d_D); drop(h->d_Kh);
  drop(h->d_JK); drop(h->d_dmax); drop(h->d_Dp_all); drop(h->d_Jp_all); drop(h->d_pack_desc); drop(h->d_light_tasks); drop(h->d_light_prefix);
  drop(h->d_fn_shell); drop(h->d_pair_key); drop(h->d_pair_fwd); drop(h->d_asym);
  if (h->h_asym) cudaFreeHost(h->h_asym);
  h->h_asym = nullptr;
  if (h->h_light_tasks) cudaFreeHost(h->h_light_tasks);
  if (h->h_light_prefix) cudaFreeHost(h->h_light_prefix);
  h->h_light_tasks = nullptr; h->h_light_prefix = nullptr; h->light_tasks_cap = 0;
  auto drop_ev = [](cudaEvent_t& e) { if (e) cudaEventDestroy(e); e = nullptr; };
  drop_ev(h->ev0); drop_ev(h->ev1); drop_ev(h->ev_fork); drop_ev(h->ev_light); drop_ev(h->ev_stream);
  drop_ev(h->ev_D); drop_ev(h->ev_done);
  for (double*& st : h->peer_stage) drop(st);
  h->peer_stage.clear();
  for (int i = 0; i < rchem_basis::kAuxStreams; ++i) {
    drop_ev(h->ev_join[i]);
    if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
    h->aux[i] = nullptr;
  }
  h->ready = false;
  cudaGetLastError();  // nothing above may leave a latched error behind
}

int ensure_ready(rchem_basis* h) {
  if (h->ready) return RCHEM_OK;
  const auto t_setup0 = std::chrono::steady_clock::now();
  release_device_state(h);  // (a previous attempt may have failed half-way)
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RCHEM_ERR_NO_DEVICE,
                std::string("no CUDA device: librchem_b200 has no CPU path (") +
                    cudaGetErrorString(e) + ")");
  if (h->device < 0 || h->device >= ndev) return fail(RCHEM_ERR_INVALID_ARG, "bad device ordinal");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->shells.lmax > 2)
    return fail(RCHEM_ERR_UNSUPPORTED_AM, "class kernels cover s, p and d shells only");
  if (!h->own_stream) CUDA_OK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  if (!h->external_stream) h->stream = h->own_stream;
  CUDA_OK(cudaEventCreate(&h->ev0));
  CUDA_OK(cudaEventCreate(&h->ev1));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  // measured: s/p-only bases like many streams (short, similar kernels); with d shells the
  // register-heavy kernels compete and 6 is better
  h->n_aux = h->shells.lmax >= 2 ? 6 : 12;
  if (const char* e = std::getenv("RCHEM_STREAMS"))
    h->n_aux = std::min((int)rchem_basis::kAuxStreams, std::max(1, atoi(e)));
  for (int i = 0; i < rchem_basis::kAuxStreams; ++i) {
    CUDA_OK(cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
  }

  // exact-Boys grid
  std::vector<double> table;
  build_boys_tables(&table);
  CUDA_OK(cudaMalloc(&h->d_boys, table.size() * sizeof(double)));
  CUDA_OK(cudaMemcpy(h->d_boys, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice));
  std::vector<double> dthr, ddirect;
  if (!build_boys_delta_tables(&dthr, &ddirect))
    return fail(RCHEM_ERR_CUDA, "internal: reference-Boys table layout exceeded (two steps in one cell)");
  CUDA_OK(cudaMalloc(&h->d_delta_thr, dthr.size() * sizeof(double)));
  CUDA_OK(cudaMemcpy(h->d_delta_thr, dthr.data(), dthr.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&h->d_delta_direct, ddirect.size() * sizeof(double)));
  CUDA_OK(cudaMemcpy(h->d_delta_direct, ddirect.data(), ddirect.size() * sizeof(double), cudaMemcpyHostToDevice));

  // shell pairs -> batches keyed by (la, lb, K2); batch order = pair class, then K2 descending
  const auto& sh = h->shells.shells;
  std::map<std::tuple<int, int, int>, Batch> by_key;
  std::vector<PrimPair> scratch_pps;
  for (int i = 0; i < (int)sh.size(); ++i)
    for (int j = 0; j <= i; ++j) {
      int a = i, b = j;
      if (sh[a].l < sh[b].l) std::swap(a, b);
      // K2 = number of SIGNIFICANT primitive pairs (pair_build.h kPrimPairEps)
      const int K2 = build_significant_prim_pairs(sh[a], sh[b], h->prim_eps, &scratch_pps);
      const int cls = sh[a].l * (sh[a].l + 1) / 2 + sh[b].l;
      Batch& bt = by_key[std::make_tuple(cls, -K2, 0)];
      bt.la = sh[a].l; bt.lb = sh[b].l; bt.K2 = K2;
      bt.shA.push_back(a); bt.shB.push_back(b);
    }
  h->batches.clear();
  for (auto& kv : by_key) {
    kv.second.npairs = (int)kv.second.shA.size();
    h->batches.push_back(std::move(kv.second));
  }

  // Schwarz bounds Q_ab = sqrt(max |(ab|ab)|) with the exact Boys function (so the quartet
  // list does not depend on the Boys option), then sort every batch by Q descending.
  for (Batch& bt : h->batches) {
    std::vector<int> ident(bt.npairs);
    std::iota(ident.begin(), ident.end(), 0);
    int rc = upload_batch(h, bt, ident);
    if (rc) return rc;
    double* dQ = nullptr;
    CUDA_OK(cudaMalloc(&dQ, (size_t)bt.npairs * sizeof(double)));
    EriTask t;
    fill_common(h, &t);
    t.bra = t.ket = bt.view();
    t.boys.exact = h->d_boys + (size_t)(2 * (type_lmax(bt.la) + type_lmax(bt.lb))) * kBoysTableLen;
    t.nwarps = (bt.npairs + 31) / 32;
    t.same = 1;
    t.Qout = dQ;
    EriLaunchFn fn = find_launcher(bt.la, bt.lb, bt.la, bt.lb);
    if (!fn) return fail(RCHEM_ERR_UNSUPPORTED_AM, "no kernel for this class in this build");
    const unsigned grid = (unsigned)((t.nwarps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    CUDA_OK(fn(kBoysExact, kModeSchwarz, t, grid, h->stream));
    bt.Q.resize(bt.npairs);
    CUDA_OK(cudaMemcpyAsync(bt.Q.data(), dQ, (size_t)bt.npairs * sizeof(double),
                            cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaFree(dQ));
    std::vector<int> order(bt.npairs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return bt.Q[x] > bt.Q[y]; });
    rc = upload_batch(h, bt, order);
    if (rc) return rc;
  }

  if (h->shells.fused) {
    // segmented twins of the fused (sp sp| pairs (see Batch::sub)
    h->seg_shells.clear();
    std::vector<int> seg_s(sh.size(), -1), seg_p(sh.size(), -1);
    auto part = [&](int shell, int which) {
      std::vector<int>& memo = which ? seg_p : seg_s;
      if (memo[shell] < 0) {
        Shell r = sh[shell];
        r.l = which;
        r.bf0 = sh[shell].bf0 + which;  // functions of an sp shell: s, px, py, pz
        if (which) r.cn = sh[shell].cn2;
        r.cn2.clear();
        memo[shell] = (int)h->seg_shells.size();
        h->seg_shells.push_back(std::move(r));
      }
      return memo[shell];
    };
    std::map<std::tuple<int, int, int>, Batch> sub_key;
    const size_t n_normal = h->batches.size();
    for (size_t bi = 0; bi < n_normal; ++bi) {
      const Batch& bt = h->batches[bi];
      if (bt.la != kTypeSP || bt.lb != kTypeSP) continue;
      for (int p = 0; p < bt.npairs; ++p) {
        const int A = bt.shA[p], B = bt.shB[p];
        int twins[4][2] = {{part(A, 0), part(B, 0)}, {part(A, 1), part(B, 0)},
                           {part(B, 1), part(A, 0)}, {part(A, 1), part(B, 1)}};
        for (int k = 0; k < 4; ++k) {
          if (k == 2 && A == B) continue;  // (p_A s_A| appears once on a diagonal pair
          const Shell &X = h->seg_shells[twins[k][0]], &Y = h->seg_shells[twins[k][1]];
          const int K2 = build_significant_prim_pairs(X, Y, h->prim_eps, &scratch_pps);
          const int cls = X.l * (X.l + 1) / 2 + Y.l;
          Batch& sb = sub_key[std::make_tuple(cls, -K2, 0)];
          sb.sub = true;
          sb.la = X.l; sb.lb = Y.l; sb.K2 = K2;
          sb.shA.push_back(twins[k][0]); sb.shB.push_back(twins[k][1]);
          sb.Q.push_back(bt.Q[p]);
        }
      }
    }
    for (auto& kv : sub_key) {
      Batch sb = std::move(kv.second);
      sb.npairs = (int)sb.shA.size();
      std::vector<int> order(sb.npairs);
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return sb.Q[x] > sb.Q[y]; });
      h->batches.push_back(std::move(sb));
      int rc = upload_batch(h, h->batches.back(), order);
      if (rc) return rc;
    }
  }

  {  // packed D / J blocks: one arena each, carved per batch (256-byte aligned slices)
    size_t total = 0;
    std::vector<size_t> off;
    for (const Batch& bt : h->batches) {
      off.push_back(total);
      total += ((size_t)bt.ncomp() * bt.stride + 31) & ~(size_t)31;
    }
    h->packed_doubles = total;
    CUDA_OK(cudaMalloc(&h->d_Dp_all, std::max<size_t>(1, total) * sizeof(double)));
    CUDA_OK(cudaMalloc(&h->d_Jp_all, std::max<size_t>(1, total) * sizeof(double)));
    CUDA_OK(cudaMemset(h->d_Dp_all, 0, std::max<size_t>(1, total) * sizeof(double)));
    std::vector<PackDesc> desc;
    int blk = 0;
    for (size_t i = 0; i < h->batches.size(); ++i) {
      Batch& bt = h->batches[i];
      bt.d_Dp = h->d_Dp_all + off[i];
      bt.d_Jp = h->d_Jp_all + off[i];
      desc.push_back(PackDesc{bt.d_idx, bt.d_Dp, bt.d_Jp, bt.npairs, bt.stride, ncart(bt.lb), bt.ncomp(), blk});
      blk += (bt.npairs + 255) / 256;
    }
    desc.push_back(PackDesc{nullptr, nullptr, nullptr, 0, 0, 1, 0, blk});
    h->pack_blocks = blk;
    CUDA_OK(cudaMalloc(&h->d_pack_desc, desc.size() * sizeof(PackDesc)));
    CUDA_OK(cudaMemcpy(h->d_pack_desc, desc.data(), desc.size() * sizeof(PackDesc), cudaMemcpyHostToDevice));
  }
  {  // Schwarz row sums for the K-row fixed-point bound (eri_kernel.cuh krow_add)
    std::vector<double> R(sh.size(), 0.0);
    for (const Batch& bt : h->batches)
      for (int p = 0; p < bt.npairs && !bt.sub; ++p) {
        const int a = bt.shA[p], b = bt.shB[p];
        R[a] += ncart(sh[b].l) * bt.Q[p];
        if (a != b) R[b] += ncart(sh[a].l) * bt.Q[p];
      }
    h->kbound = 16.0 * *std::max_element(R.begin(), R.end());
  }
  CUDA_OK(cudaMalloc(&h->d_dmax, sizeof(double)));
  CUDA_OK(cudaMalloc(&h->d_asym, 2 * sizeof(double)));
  CUDA_OK(cudaMallocHost(&h->h_asym, 2 * sizeof(double)));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_stream, cudaEventDisableTiming));
  h->last_stream_valid = false;
  const size_t nn = (size_t)h->N * h->N;
  CUDA_OK(cudaMalloc(&h->d_D, nn * sizeof(double)));
  CUDA_OK(cudaMalloc(&h->d_Kh, nn * sizeof(double)));
  CUDA_OK(cudaMalloc(&h->d_JK, 2 * nn * sizeof(double)));
  {  // maps for the dense-tensor completion
    const int ns = (int)sh.size();
    std::vector<int> fn_shell(h->N);
    for (int s = 0; s < ns; ++s)
      for (int k = 0; k < ncart(sh[s].l); ++k) fn_shell[sh[s].bf0 + k] = s;
    std::vector<long long> key((size_t)ns * ns, 0);
    std::vector<unsigned char> fwd((size_t)ns * ns, 0);
    for (size_t bi = 0; bi < h->batches.size(); ++bi) {
      const Batch& bt = h->batches[bi];
      for (int p = 0; p < bt.npairs && !bt.sub; ++p) {
        const int a = bt.shA[p], b = bt.shB[p];
        const long long kk = ((long long)bi << 32) | (long long)p;
        key[(size_t)a * ns + b] = key[(size_t)b * ns + a] = kk;
        fwd[(size_t)a * ns + b] = 1;
        if (a != b) fwd[(size_t)b * ns + a] = 0;
      }
    }
    CUDA_OK(cudaMalloc(&h->d_fn_shell, fn_shell.size() * sizeof(int)));
    CUDA_OK(cudaMalloc(&h->d_pair_key, key.size() * sizeof(long long)));
    CUDA_OK(cudaMalloc(&h->d_pair_fwd, fwd.size()));
    CUDA_OK(cudaMemcpy(h->d_fn_shell, fn_shell.data(), fn_shell.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->d_pair_key, key.data(), key.size() * sizeof(long long), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->d_pair_fwd, fwd.data(), fwd.size(), cudaMemcpyHostToDevice));
  }
  h->ready = true;
  h->setup_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_setup0).count();
  return RCHEM_OK;
}

void free_tasks(rchem_basis* h) {
  for (TaskTable& t : h->tasks) {
    if (t.d_prefix) cudaFree(t.d_prefix);
    if (t.d_nq) cudaFree(t.d_nq);
    if (t.d_prefix_light) cudaFree(t.d_prefix_light);
    if (t.d_nq_light) cudaFree(t.d_nq_light);
    if (t.d_hp) cudaFree(t.d_hp);
    if (t.d_lp) cudaFree(t.d_lp);
    if (t.d_hblk_prefix) cudaFree(t.d_hblk_prefix);
  }
  h->tasks.clear();
  h->tasks_tau = -1.0;
}

// The implicit screened quartet list (DESIGN.md "quartet list"): for task (bra batch, ket
// batch <= bra batch) and bra pair p the surviving kets are the prefix
//   q < nq[p] = #{q : Q_bra[p]*Q_ket[q] >= tau}   (Q_ket descending), and q <= p when the
// batches coincide.  Pure integer/compare work on the shared Q arrays.
int ensure_tasks(rchem_basis* h) {
  if (h->tasks_tau == h->tau && h->tasks_heavy == h->heavy_passes &&
      h->tasks_light == h->light_kernel && !h->tasks.empty())
    return RCHEM_OK;
  const auto t_setup0 = std::chrono::steady_clock::now();
  free_tasks(h);
  const double tau = h->tau;
  for (int bi = 0; bi < (int)h->batches.size(); ++bi)
    for (int ki = 0; ki <= bi; ++ki) {
      const Batch &B = h->batches[bi], &K = h->batches[ki];
      if (B.sub != K.sub) continue;  // the segmented twins only meet each other
      TaskTable tt;
      tt.bra = bi; tt.ket = ki;
      tt.sub = B.sub;
      tt.jk_skip = !B.sub && B.la == kTypeSP && B.lb == kTypeSP && K.la == kTypeSP && K.lb == kTypeSP;
      tt.h_nq.resize(B.npairs);
      tt.h_prefix.resize(B.npairs + 1);
      tt.h_prefix[0] = 0;
      for (int p = 0; p < B.npairs; ++p) {
        const double qb = B.Q[p];
        int lo = 0, hi = K.npairs;  // first q with qb*Q[q] < tau
        while (lo < hi) {
          const int mid = (lo + hi) / 2;
          if (qb * K.Q[mid] >= tau) lo = mid + 1; else hi = mid;
        }
        int cut = lo;
        const int full = (bi == ki) ? p + 1 : K.npairs;
        if (bi == ki) cut = std::min(cut, p + 1);
        tt.h_nq[p] = cut;
        tt.h_prefix[p + 1] = tt.h_prefix[p] + (cut + 31) / 32;
        tt.nquartets += cut;
        tt.nquartets_all += full;
      }
      tt.nwarps = tt.h_prefix[B.npairs];
      // J/K split.  Heavy: at least one full pass of the block kernel's threads, and the
      // D/K rows of the bra functions fit in shared memory.
      EriBlockInfo info{0, 0};
      find_block_launcher(B.la, B.lb, K.la, K.lb, &info);
      tt.smem_bytes = (size_t)2 * (ncart(B.la) + ncart(B.lb)) * h->N * sizeof(double) +
                      (size_t)B.K2 * B.prim_bytes() +
                      (size_t)(info.kets_per_block + 96) * sizeof(unsigned short) + 64;
      // (the block and light kernels read the ket pair's functions from a packed 15 + 15 bit word)
      const bool packed_ok = h->N < 32768;
      const bool rows_fit = tt.smem_bytes <= kMaxBlockSmem && info.threads > 0 && packed_ok;
      // a bra pair is "heavy" when its ket prefix fills the block kernel's threads at least
      // kHeavyPasses times; below that the warp-per-bra-pair kernel (no D/K row staging, no
      // block-wide barriers) is the faster home
      // (RCHEM_OPT_HEAVY_PASSES; measured on (H2O)96/6-31G with the light kernel:
      // 0.25: 176 ms, 0.5: 151, 1: 138, 1.5: 135, 2: 135, 3: 136, 4: 146)
      const double kHeavyPasses = h->heavy_passes;
      std::vector<int> nq_light(B.npairs), hp;
      std::vector<long long> prefix_light(B.npairs + 1, 0), hblk(1, 0);
      for (int p = 0; p < B.npairs; ++p) {
        const int cut = tt.h_nq[p];
        const bool heavy = rows_fit && cut >= 1 && (double)cut >= kHeavyPasses * info.threads;
        nq_light[p] = heavy ? 0 : cut;
        prefix_light[p + 1] = prefix_light[p] + (nq_light[p] + 31) / 32;
        if (heavy) {
          hp.push_back(p);
          hblk.push_back(hblk.back() + (cut + info.kets_per_block - 1) / info.kets_per_block);
        } else {
          tt.nquartets_light += cut;
        }
      }
      tt.nwarps_light = prefix_light[B.npairs];
      // The light pairs go to the warp-per-bra-pair kernel when the class has one (the block
      // kernel's classes) and a warp's ket list fits its shared-memory slice; otherwise to
      // the chunk (warp-per-32-kets) kernel.
      {
        std::vector<int> lp;
        int cap = 0;
        for (int p = 0; p < B.npairs; ++p)
          if (nq_light[p] > 0) { lp.push_back(p); cap = std::max(cap, nq_light[p]); }
        const size_t per_warp = (((size_t)B.K2 * B.prim_bytes() +
                                  (size_t)cap * sizeof(int)) + 7) & ~(size_t)7;
        if (h->light_kernel && packed_ok && info.threads > 0 && !lp.empty() && per_warp * kWarpsPerBlock <= 40 * 1024) {
          tt.nlight = (int)lp.size();
          tt.h_lp = lp;
          tt.light_cap = cap;
          tt.light_smem = per_warp * kWarpsPerBlock;
          CUDA_OK(cudaMalloc(&tt.d_lp, lp.size() * sizeof(int)));
          CUDA_OK(cudaMemcpy(tt.d_lp, lp.data(), lp.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
      }
      tt.nheavy = (int)hp.size();
      tt.nblocks_heavy = hblk.back();
      tt.kets_per_block = info.kets_per_block;
      tt.h_nq_light = nq_light;
      tt.h_hp = hp;
      tt.h_hblk = hblk;
      CUDA_OK(cudaMalloc(&tt.d_prefix_light, prefix_light.size() * sizeof(long long)));
      CUDA_OK(cudaMalloc(&tt.d_nq_light, std::max<size_t>(1, nq_light.size()) * sizeof(int)));
      CUDA_OK(cudaMalloc(&tt.d_hp, std::max<size_t>(1, hp.size()) * sizeof(int)));
      CUDA_OK(cudaMalloc(&tt.d_hblk_prefix, hblk.size() * sizeof(long long)));
      CUDA_OK(cudaMemcpy(tt.d_prefix_light, prefix_light.data(), prefix_light.size() * sizeof(long long), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(tt.d_nq_light, nq_light.data(), nq_light.size() * sizeof(int), cudaMemcpyHostToDevice));
      if (!hp.empty()) CUDA_OK(cudaMemcpy(tt.d_hp, hp.data(), hp.size() * sizeof(int), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(tt.d_hblk_prefix, hblk.data(), hblk.size() * sizeof(long long), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMalloc(&tt.d_prefix, tt.h_prefix.size() * sizeof(long long)));
      CUDA_OK(cudaMalloc(&tt.d_nq, std::max<size_t>(1, tt.h_nq.size()) * sizeof(int)));
      CUDA_OK(cudaMemcpy(tt.d_prefix, tt.h_prefix.data(), tt.h_prefix.size() * sizeof(long long),
                         cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(tt.d_nq, tt.h_nq.data(), tt.h_nq.size() * sizeof(int),
                         cudaMemcpyHostToDevice));
      h->tasks.push_back(std::move(tt));
    }
  h->tasks_tau = tau;
  h->tasks_heavy = h->heavy_passes;
  h->tasks_light = h->light_kernel;
  h->setup_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_setup0).count();
  return RCHEM_OK;
}

// launches every task in `mode`; fills stats
int run_tasks(rchem_basis* h, int mode, EriTask proto, int rank, int nranks) {
  h->stats_src = nullptr;
  rchem_stats& st = h->stats;
  st = rchem_stats{};
  st.n_tasks = (int)h->tasks.size();
  const bool split = mode == kModeJK;
  // the share of this rank is the block-interleaved 1/nranks slice
  auto my_blocks = [&](long long nblocks) {
    return nblocks > rank ? (nblocks - rank + nranks - 1) / nranks : 0LL;
  };
  auto make_task = [&](const TaskTable& tt) {
    const Batch &B = h->batches[tt.bra], &K = h->batches[tt.ket];
    EriTask t = proto;
    t.bra = B.view();
    t.ket = K.view();
    t.boys.exact = h->d_boys + (size_t)(type_lmax(B.la) + type_lmax(B.lb) + type_lmax(K.la) + type_lmax(K.lb)) * kBoysTableLen;
    t.nq = tt.d_nq;
    t.same = tt.bra == tt.ket;
    t.rank = rank;
    t.nranks = nranks;
    return t;
  };

  // Light bra pairs (J/K mode): ONE launch per class over all its tasks.  The descriptors go
  // through a pinned staging buffer into device memory before the streams fork.
  struct LightGroup { int la, lb, lc, ld; size_t first, prefix_at; int ntasks; size_t smem; long long grid; };
  std::vector<LightGroup> groups;
  static const bool kMergeLight = [] {
    const char* e = std::getenv("RCHEM_LIGHT_MERGE");
    return e ? atoi(e) != 0 : true;
  }();
  if (split && kMergeLight) {
    std::map<std::tuple<int, int, int, int>, std::vector<const TaskTable*>> by_class;
    size_t total = 0;
    for (const TaskTable& tt : h->tasks) {
      if (tt.nlight <= 0 || tt.jk_skip) continue;
      const Batch &B = h->batches[tt.bra], &K = h->batches[tt.ket];
      by_class[std::make_tuple(B.la, B.lb, K.la, K.lb)].push_back(&tt);
      ++total;
    }
    if (total > 0) {
      if (total > h->light_tasks_cap) {
        if (h->ev_light) CUDA_OK(cudaEventSynchronize(h->ev_light));
        // each buffer is released exactly once and forgotten before the reallocation, so a
        // failed cudaMalloc below cannot leave a stale pointer for rchem_basis_destroy
        if (h->d_light_tasks) CUDA_OK(cudaFree(h->d_light_tasks));
        if (h->d_light_prefix) CUDA_OK(cudaFree(h->d_light_prefix));
        if (h->h_light_tasks) CUDA_OK(cudaFreeHost(h->h_light_tasks));
        if (h->h_light_prefix) CUDA_OK(cudaFreeHost(h->h_light_prefix));
        h->d_light_tasks = nullptr; h->d_light_prefix = nullptr;
        h->h_light_tasks = nullptr; h->h_light_prefix = nullptr;
        h->light_tasks_cap = 0;
        CUDA_OK(cudaMalloc(&h->d_light_tasks, total * sizeof(EriTask)));
        CUDA_OK(cudaMalloc(&h->d_light_prefix, (2 * total + 64) * sizeof(int)));
        CUDA_OK(cudaMallocHost(&h->h_light_tasks, total * sizeof(EriTask)));
        CUDA_OK(cudaMallocHost(&h->h_light_prefix, (2 * total + 64) * sizeof(int)));
        h->light_tasks_cap = total;
        if (!h->ev_light) CUDA_OK(cudaEventCreateWithFlags(&h->ev_light, cudaEventDisableTiming));
      } else {
        CUDA_OK(cudaEventSynchronize(h->ev_light));  // the previous build's copy has left the buffer
      }
      size_t at = 0, pat = 0;
      for (auto& kv : by_class) {
        LightGroup g{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first),
                     std::get<3>(kv.first), at, pat, 0, 0, 0};
        int* prefix = h->h_light_prefix + pat;
        long long blocks = 0;
        for (const TaskTable* tt : kv.second) {
          EriTask t = make_task(*tt);
#ifdef RCHEM_PROBES
          if (std::getenv("RCHEM_PROBE_ALLFAR_LIGHT")) t.far_sched = 2;
#endif
          t.lp = tt->d_lp;
          t.nlight = tt->nlight;
          t.light_cap = tt->light_cap;
          const long long nblocks = ((long long)tt->nlight + kWarpsPerBlock - 1) / kWarpsPerBlock;
          const long long mine = my_blocks(nblocks);
          if (mine <= 0) continue;
          h->h_light_tasks[at + g.ntasks] = t;
          prefix[g.ntasks] = (int)blocks;
          blocks += mine;
          g.ntasks += 1;
          g.smem = std::max(g.smem, tt->light_smem);
        }
        if (blocks > 0x7fffffffLL) return fail(RCHEM_ERR_TOO_LARGE, "light tasks exceed the grid limit");
        prefix[g.ntasks] = (int)blocks;
        g.grid = blocks;
        g.first = at;
        at += g.ntasks;
        pat += g.ntasks + 1;
        if (g.ntasks > 0) groups.push_back(g);
      }
      CUDA_OK(cudaMemcpyAsync(h->d_light_tasks, h->h_light_tasks, at * sizeof(EriTask),
                              cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(cudaMemcpyAsync(h->d_light_prefix, h->h_light_prefix, pat * sizeof(int),
                              cudaMemcpyHostToDevice, h->stream));
      CUDA_OK(cudaEventRecord(h->ev_light, h->stream));
    }
  }

  CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  // fork: the auxiliary streams wait for everything queued so far on the main stream
  CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
  for (int i = 0; i < h->n_aux; ++i) CUDA_OK(cudaStreamWaitEvent(h->aux[i], h->ev_fork, 0));
  int next_stream = 0;
  auto pick_stream = [&]() {
    cudaStream_t s = h->aux[next_stream];
    next_stream = (next_stream + 1) % h->n_aux;
    return s;
  };
  const bool merged = !groups.empty();
  for (const LightGroup& g : groups) {
    EriLightMultiLaunchFn mfn = find_light_multi_launcher(g.la, g.lb, g.lc, g.ld);
    if (!mfn) return fail(RCHEM_ERR_UNSUPPORTED_AM, "no kernel for this class");
    CUDA_OK(mfn(h->boys, h->d_light_tasks + g.first, h->d_light_prefix + g.prefix_at, g.ntasks,
                (unsigned)g.grid, g.smem, pick_stream()));
    st.launches += 1;
  }
  // Launch order: highest classes first.  Their launches are few, small-grid and long-running
  // (one thread walks thousands of recurrence terms), so they should start early and overlap the
  // bulk of the build instead of forming its tail (RCHEM_TASK_ORDER=0: table order, for A/B).
  static const bool kReverseOrder = [] {
    const char* e = std::getenv("RCHEM_TASK_ORDER");
    return e ? atoi(e) != 0 : true;
  }();
  const size_t ntask = h->tasks.size();
  for (size_t ti = 0; ti < ntask; ++ti) {
    TaskTable& tt = h->tasks[kReverseOrder ? ntask - 1 - ti : ti];
    const Batch &B = h->batches[tt.bra], &K = h->batches[tt.ket];
    // J/K: (sp sp|sp sp) runs through its segmented twins; tensor / list modes never see them
    if (split ? tt.jk_skip : tt.sub) continue;
    {
      double P_, H_;
      int ns_ = 1;
      flop_model_types(B.la, B.lb, K.la, K.lb, &P_, &H_, &ns_);
      st.shell_quartets_all += tt.nquartets_all * ns_;
    }
    if (tt.nwarps == 0) continue;
    quartets_of_rank(tt, rank, nranks);
    EriTask t = make_task(tt);
    EriLaunchFn fn = find_launcher(B.la, B.lb, K.la, K.lb);
    if (!fn) return fail(RCHEM_ERR_UNSUPPORTED_AM, "no kernel for this class");
    double fmP = 0.0, fmH = 0.0;
    int nseg = 1;
    const bool fm = flop_model_types(B.la, B.lb, K.la, K.lb, &fmP, &fmH, &nseg);
    const double k4 = (double)B.K2 * K.K2;
    // counts are in SEGMENTED (s/p/d) shell quartets: a quartet of fused sp shells evaluates
    // nseg of them (every s|p-part combination) on one set of primitive quartets
    auto account = [&](long long q) {
      st.shell_quartets += q * nseg;
      st.fused_quartets += q;
      st.prim_quartets += (long long)(q * k4) * nseg;
      st.prim_quartets_evaluated += (long long)(q * k4);
      st.integrals += q * (long long)(ncart(B.la) * ncart(B.lb) * ncart(K.la) * ncart(K.lb));
      if (fm) st.model_flops += q * (k4 * fmP + fmH);
    };
    // --- warp kernels (everything in tensor mode; the light bra pairs in J/K mode) ---
    const long long nwarps = split ? tt.nwarps_light : tt.nwarps;
    if (split && tt.nlight > 0) {
      // --- warp-per-bra-pair kernel (light bra pairs, J/K mode) ---
      const long long nblocks = ((long long)tt.nlight + kWarpsPerBlock - 1) / kWarpsPerBlock;
      const long long mine = my_blocks(nblocks);
      if (!merged) {
        EriLightLaunchFn lfn = find_light_launcher(B.la, B.lb, K.la, K.lb);
        t.lp = tt.d_lp;
        t.nlight = tt.nlight;
        t.light_cap = tt.light_cap;
        CUDA_OK(lfn(h->boys, t, (unsigned)mine, tt.light_smem, pick_stream()));
        if (mine > 0) st.launches += 1;
      }
      account(tt.q_light_mine);
    } else if (nwarps > 0) {
      t.warp_prefix = split ? tt.d_prefix_light : tt.d_prefix;
      t.nq = split ? tt.d_nq_light : tt.d_nq;
      t.nwarps = nwarps;
      const long long nblocks = (nwarps + kWarpsPerBlock - 1) / kWarpsPerBlock;
      const long long mine = my_blocks(nblocks);
      if (mine > 0x7fffffffLL) return fail(RCHEM_ERR_TOO_LARGE, "task exceeds the grid limit");
      CUDA_OK(fn(h->boys, mode, t, (unsigned)mine, pick_stream()));
      if (mine > 0) st.launches += 1;
      account(split ? tt.q_light_mine : tt.q_tensor_mine);
    }
    // --- block kernel (heavy bra pairs, J/K mode) ---
    if (split && tt.nblocks_heavy > 0) {
      EriBlockInfo info{0, 0};
      EriBlockLaunchFn bfn = find_block_launcher(B.la, B.lb, K.la, K.lb, &info);
      t.nq = tt.d_nq;
#ifdef RCHEM_PROBES
      if (std::getenv("RCHEM_PROBE_ALLFAR_BLOCK")) t.far_sched = 2;
#endif
      t.hp = tt.d_hp;
      t.hblk_prefix = tt.d_hblk_prefix;
      t.nheavy = tt.nheavy;
      t.nblocks_heavy = tt.nblocks_heavy;
      const long long mine = my_blocks(tt.nblocks_heavy);
      if (mine > 0x7fffffffLL) return fail(RCHEM_ERR_TOO_LARGE, "task exceeds the grid limit");
      CUDA_OK(bfn(h->boys, t, (unsigned)mine, tt.smem_bytes, pick_stream()));
      if (mine > 0) st.launches += 1;
      account(tt.q_heavy_mine);
    }
  }
  // join: the main stream waits for every auxiliary stream
  for (int i = 0; i < h->n_aux; ++i) {
    CUDA_OK(cudaEventRecord(h->ev_join[i], h->aux[i]));
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_join[i], 0));
  }
  CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  return RCHEM_OK;
}

// A build uses the handle's scratch buffers (Kh, packed J/D, light task tables).  When the
// caller switched streams (rchem_set_stream) since the previous build, the new stream waits
// for the old one's last use of them.
int order_after_previous_build(rchem_basis* h) {
  if (h->last_stream_valid && h->last_stream != h->stream)
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_stream, 0));
  return RCHEM_OK;
}

int make_basis_handle(Basis&& basis, rchem_basis** out) {
  rchem_basis* h = new rchem_basis();
  h->basis = std::move(basis);
  std::string err;
  if (!group_shells(h->basis, &h->shells, &err, h->fuse_sp != 0)) {
    delete h;
    return fail(RCHEM_ERR_UNSUPPORTED_LAYOUT, err);
  }
  h->N = (int)h->basis.cgtos.size();
  for (const CGTO& g : h->basis.cgtos) h->nprim += (int)g.primitives.size();
  *out = h;
  return RCHEM_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* rchem_last_error(void) { return g_err.c_str(); }
int rchem_version(void) { return 100; }

int rchem_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(RCHEM_ERR_NO_DEVICE, cudaGetErrorString(e));
  return n;
}

int rchem_basis_new(int natoms, const uint64_t* atomnos, const double* coords,
                    const char* basis_set_name, rchem_basis** out) {
  if (natoms <= 0 || !atomnos || !coords || !basis_set_name || !out)
    return fail(RCHEM_ERR_INVALID_ARG, "rchem_basis_new: null or empty argument");
  Basis b;
  std::string err;
  std::vector<uint64_t> z(atomnos, atomnos + natoms);
  if (!basis_new(z, coords, basis_set_name, &b, &err)) return fail(RCHEM_ERR_UNKNOWN_BASIS, err);
  return make_basis_handle(std::move(b), out);
}

int rchem_basis_create(int n, const double* origins, const int32_t* powers,
                       const int32_t* prim_offset, const double* exps, const double* coefs,
                       const double* norms, rchem_basis** out) {
  if (n <= 0 || !origins || !powers || !prim_offset || !exps || !coefs || !norms || !out)
    return fail(RCHEM_ERR_INVALID_ARG, "rchem_basis_create: null or empty argument");
  Basis b;
  b.name = "custom";
  for (int i = 0; i < n; ++i) {
    if (prim_offset[i + 1] <= prim_offset[i])
      return fail(RCHEM_ERR_INVALID_ARG, "rchem_basis_create: CGTO without primitives");
    CGTO g;
    for (int d = 0; d < 3; ++d) {
      g.origin[d] = origins[3 * i + d];
      g.powers[d] = powers[3 * i + d];
      if (g.powers[d] < 0) return fail(RCHEM_ERR_INVALID_ARG, "negative Cartesian power");
    }
    for (int p = prim_offset[i]; p < prim_offset[i + 1]; ++p) {
      PGTO pg;
      for (int d = 0; d < 3; ++d) { pg.origin[d] = g.origin[d]; pg.powers[d] = g.powers[d]; }
      pg.exponent = exps[p];
      pg.norm = norms[p];
      if (!(pg.exponent > 0.0)) return fail(RCHEM_ERR_INVALID_ARG, "non-positive exponent");
      g.primitives.push_back(pg);
      g.coefs.push_back(coefs[p]);
    }
    b.cgtos.push_back(std::move(g));
  }
  return make_basis_handle(std::move(b), out);
}

void rchem_basis_destroy(rchem_basis* h) {
  if (!h) return;
  for (rchem_basis* peer : h->peers) rchem_basis_destroy(peer);
  h->peers.clear();
  if (h->tensor_twin) rchem_basis_destroy(h->tensor_twin);
  h->tensor_twin = nullptr;
  release_device_state(h);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int rchem_basis_nbf(const rchem_basis* h) { return h ? h->N : RCHEM_ERR_INVALID_ARG; }
int rchem_basis_nprim(const rchem_basis* h) { return h ? h->nprim : RCHEM_ERR_INVALID_ARG; }
int rchem_basis_nshells(const rchem_basis* h) {
  return h ? (int)h->shells.shells.size() : RCHEM_ERR_INVALID_ARG;
}

int rchem_basis_export(const rchem_basis* h, double* origins, int32_t* powers,
                       int32_t* prim_offset, double* exps, double* coefs, double* norms) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  int p = 0;
  for (int i = 0; i < h->N; ++i) {
    const CGTO& g = h->basis.cgtos[i];
    for (int d = 0; d < 3; ++d) {
      if (origins) origins[3 * i + d] = g.origin[d];
      if (powers) powers[3 * i + d] = g.powers[d];
    }
    if (prim_offset) prim_offset[i] = p;
    for (size_t k = 0; k < g.primitives.size(); ++k, ++p) {
      if (exps) exps[p] = g.primitives[k].exponent;
      if (coefs) coefs[p] = g.coefs[k];
      if (norms) norms[p] = g.primitives[k].norm;
    }
  }
  if (prim_offset) prim_offset[h->N] = p;
  return RCHEM_OK;
}

int rchem_basis_shells(const rchem_basis* h, int32_t* l, int32_t* first_function) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  for (size_t i = 0; i < h->shells.shells.size(); ++i) {
    if (l) l[i] = h->shells.shells[i].l == kTypeSP ? RCHEM_SHELL_SP : h->shells.shells[i].l;
    if (first_function) first_function[i] = h->shells.shells[i].bf0;
  }
  return (int)h->shells.shells.size();
}

double rchem_normalization(const int32_t powers[3], double exponent) {
  const int pw[3] = {powers[0], powers[1], powers[2]};
  return pgto_normalization(pw, exponent);
}

int rchem_get_ijk_list(int m, int32_t* out) {
  if (m < 0) return fail(RCHEM_ERR_INVALID_ARG, "negative angular momentum");
  const auto v = get_ijk_list(m);
  if (out)
    for (size_t i = 0; i < v.size(); ++i)
      for (int d = 0; d < 3; ++d) out[3 * i + d] = v[i][d];
  return (int)v.size();
}

int64_t rchem_ijkl2intindex(int64_t i, int64_t j, int64_t k, int64_t l) {
  if (i < j) std::swap(i, j);
  if (k < l) std::swap(k, l);
  int64_t ij = i * (i + 1) / 2 + j, kl = k * (k + 1) / 2 + l;
  if (ij < kl) std::swap(ij, kl);
  return ij * (ij + 1) / 2 + kl;
}

int rchem_set_option(rchem_basis* h, int key, double value) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  switch (key) {
    case RCHEM_OPT_BOYS:
      if (value != 0.0 && value != 1.0) return fail(RCHEM_ERR_INVALID_ARG, "boys must be 0 or 1");
      h->boys = (int)value;
      return RCHEM_OK;
    case RCHEM_OPT_SCHWARZ_TAU:
      if (!(value >= 0.0)) return fail(RCHEM_ERR_INVALID_ARG, "tau must be >= 0");
      h->tau = value;
      return RCHEM_OK;
    case RCHEM_OPT_DEVICE:
      if (h->ready) return fail(RCHEM_ERR_INVALID_ARG, "device is fixed after the first compute call");
      h->device = (int)value;
      return RCHEM_OK;
    case RCHEM_OPT_PRIM_EPS:
      if (h->ready) return fail(RCHEM_ERR_INVALID_ARG, "prim_eps is fixed after the first compute call");
      if (!(value >= 0.0)) return fail(RCHEM_ERR_INVALID_ARG, "prim_eps must be >= 0");
      h->prim_eps = value;
      return RCHEM_OK;
    case RCHEM_OPT_FAR_SCHED:
      if (value != 0.0 && value != 1.0) return fail(RCHEM_ERR_INVALID_ARG, "far_sched must be 0 or 1");
      h->far_sched = (int)value;
      return RCHEM_OK;
    case RCHEM_OPT_SYMMETRIC_D_ONLY:
      h->symmetric_only = value != 0.0;
      return RCHEM_OK;
    case RCHEM_OPT_LIGHT_KERNEL:
      h->light_kernel = value != 0.0;
      return RCHEM_OK;
    case RCHEM_OPT_HEAVY_PASSES:
      if (!(value >= 0.0)) return fail(RCHEM_ERR_INVALID_ARG, "heavy_passes must be >= 0");
      h->heavy_passes = value;  // (the task tables are rebuilt by the next J/K build)
      return RCHEM_OK;
    case RCHEM_OPT_FUSE_SP: {
      if (h->ready) return fail(RCHEM_ERR_INVALID_ARG, "fuse_sp is fixed after the first compute call");
      if (value != 0.0 && value != 1.0) return fail(RCHEM_ERR_INVALID_ARG, "fuse_sp must be 0 or 1");
      ShellSet regrouped;
      std::string err;
      if (!group_shells(h->basis, &regrouped, &err, value != 0.0))
        return fail(RCHEM_ERR_UNSUPPORTED_LAYOUT, err);
      h->shells = regrouped;
      h->fuse_sp = (int)value;
      return RCHEM_OK;
    }
    case RCHEM_OPT_NGPUS: {
      if (!(value >= 1.0) || value != std::floor(value) || value > 16.0)
        return fail(RCHEM_ERR_INVALID_ARG, "ngpus must be an integer in 1..16");
      int ndev = 0;
      if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RCHEM_ERR_NO_DEVICE, "no CUDA device: librchem_b200 has no CPU path");
      }
      // (RCHEM_MULTI_OVERSUBSCRIBE=1 wraps the group around the visible devices, so the
      // multi-device code path can be exercised on a one-GPU box: tests only)
      if (h->device + (int)value > ndev && !std::getenv("RCHEM_MULTI_OVERSUBSCRIBE"))
        return fail(RCHEM_ERR_INVALID_ARG, "ngpus: devices device .. device+ngpus-1 must exist");
      h->ngpus = (int)value;
      return RCHEM_OK;
    }
  }
  return fail(RCHEM_ERR_INVALID_ARG, "unknown option");
}

double rchem_get_option(const rchem_basis* h, int key) {
  if (!h) return std::numeric_limits<double>::quiet_NaN();
  switch (key) {
    case RCHEM_OPT_BOYS: return h->boys;
    case RCHEM_OPT_SCHWARZ_TAU: return h->tau;
    case RCHEM_OPT_DEVICE: return h->device;
    case RCHEM_OPT_PRIM_EPS: return h->prim_eps;
    case RCHEM_OPT_FAR_SCHED: return h->far_sched;
    case RCHEM_OPT_HEAVY_PASSES: return h->heavy_passes;
    case RCHEM_OPT_LIGHT_KERNEL: return h->light_kernel;
    case RCHEM_OPT_SYMMETRIC_D_ONLY: return h->symmetric_only;
    case RCHEM_OPT_NGPUS: return h->ngpus;
    case RCHEM_OPT_FUSE_SP: return h->fuse_sp;
  }
  return std::numeric_limits<double>::quiet_NaN();
}

int rchem_set_stream(rchem_basis* h, void* s) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  h->stream = (cudaStream_t)s;  // NULL == the legacy default stream
  h->external_stream = true;
  return RCHEM_OK;
}

int rchem_use_own_stream(rchem_basis* h) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  h->external_stream = false;
  h->stream = h->own_stream;  // still null before the first compute call; set in ensure_ready
  return RCHEM_OK;
}

int rchem_get_stats(const rchem_basis* h, rchem_stats* out) {
  if (!h || !out) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  *out = h->stats;
  out->setup_ms = h->setup_ms + (h->tensor_twin ? h->tensor_twin->setup_ms : 0.0);
  const rchem_basis* e = h->stats_src ? h->stats_src : h;
  if (e->ready && e->ev0 && h->stats.launches > 0) {
    if (cudaEventSynchronize(e->ev1) == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) out->kernel_ms = ms;
    }
  }
  return RCHEM_OK;
}

// ---------------- device-buffer entry points ---------------------------------------------
// antisym = 0: J and K of a symmetric D are WRITTEN to JK_dev.
// antisym = 1: D is antisymmetric; J is untouched (it vanishes) and K(D) is ADDED to JK_dev's K.
static int jk_direct_device_impl(rchem_basis* h, const double* D_dev, double* JK_dev, int rank,
                                 int nranks, int antisym) {
  if (!h || !D_dev || !JK_dev) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RCHEM_ERR_INVALID_ARG, "bad rank");
  int rc = ensure_ready(h);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));
  rc = ensure_tasks(h);
  if (rc) return rc;
  rc = order_after_previous_build(h);
  if (rc) return rc;
  const size_t nn = (size_t)h->N * h->N;
  const int N = h->N;
  // J.fill(0); K.fill(0) (basis.rs:389-390): the packed J blocks and the K half-accumulator
  CUDA_OK(cudaMemsetAsync(h->d_Kh, 0, nn * sizeof(double), h->stream));
  const int nbatch = (int)h->batches.size();
  CUDA_OK(cudaMemsetAsync(h->d_Jp_all, 0, h->packed_doubles * sizeof(double), h->stream));
  if (h->pack_blocks > 0)
    pack_d_kernel<<<h->pack_blocks, 256, 0, h->stream>>>(D_dev, N, h->d_pack_desc, nbatch);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemsetAsync(h->d_dmax, 0, sizeof(double), h->stream));
  absmax_kernel<<<148, 256, 0, h->stream>>>(D_dev, nn, reinterpret_cast<unsigned long long*>(h->d_dmax));
  EriTask proto;
  fill_common(h, &proto);
  proto.D = D_dev;
  proto.Kh = h->d_Kh;
  proto.dmax = h->d_dmax;
  proto.kbound = h->kbound;
  rc = run_tasks(h, kModeJK, proto, rank, nranks);
  if (rc) return rc;
  if (!antisym) CUDA_OK(cudaMemsetAsync(JK_dev, 0, nn * sizeof(double), h->stream));
  if (!antisym && h->pack_blocks > 0)
    finalize_j_kernel<<<h->pack_blocks, 256, 0, h->stream>>>(h->d_pack_desc, nbatch, N, JK_dev);
  finalize_k_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(
      h->d_Kh, JK_dev + nn, N, antisym ? -1.0 : 1.0, antisym);
  CUDA_OK(cudaGetLastError());
  // pack_d (+ finalize_j), absmax, finalize_k
  h->stats.launches += (antisym ? 1 : 2) + 2;
  CUDA_OK(cudaEventRecord(h->ev_stream, h->stream));
  h->last_stream = h->stream;
  h->last_stream_valid = true;
  return RCHEM_OK;
}

int rchem_jk_direct_device(rchem_basis* h, const double* D_dev, double* JK_dev, int rank,
                           int nranks) {
  return jk_direct_device_impl(h, D_dev, JK_dev, rank, nranks, 0);
}

int rchem_build_I_device(rchem_basis* h, double* I_dev) {
  if (!h || !I_dev) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  if (h->shells.fused) {
    // The dense tensor is written per canonical element by the segmented classes; a handle with
    // fused sp shells keeps a segmented clone of itself for it (same device, stream, options).
    if (!h->tensor_twin) {
      Basis copy = h->basis;
      rchem_basis* twin = nullptr;
      int rc = make_basis_handle(std::move(copy), &twin);
      if (rc) return rc;
      std::string err;
      if (!group_shells(twin->basis, &twin->shells, &err, false)) {
        rchem_basis_destroy(twin);
        return fail(RCHEM_ERR_UNSUPPORTED_LAYOUT, err);
      }
      twin->fuse_sp = 0;
      twin->device = h->device;
      twin->prim_eps = h->prim_eps;
      h->tensor_twin = twin;
    }
    rchem_basis* t = h->tensor_twin;
    t->boys = h->boys; t->tau = h->tau;
    int rc = ensure_ready(h);  // (gives h its stream)
    if (rc) return rc;
    t->stream = h->stream;
    t->external_stream = true;
    rc = rchem_build_I_device(t, I_dev);
    h->stats = t->stats;
    h->stats_src = t;  // (the timing events of this build live in the clone)
    return rc;
  }
  int rc = ensure_ready(h);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));
  rc = ensure_tasks(h);
  if (rc) return rc;
  const size_t n4 = (size_t)h->N * h->N * h->N * h->N;
  // screened-out quartets must read as zero; without screening every canonical element is
  // written by the ERI kernels and every other element by the fill kernel
  if (h->tau > 0.0) CUDA_OK(cudaMemsetAsync(I_dev, 0, n4 * sizeof(double), h->stream));
  EriTask proto;
  fill_common(h, &proto);
  proto.I = I_dev;
  rc = run_tasks(h, kModeTensor, proto, 0, 1);
  if (rc) return rc;
  const unsigned ngrp = (unsigned)((h->N + kFillGroup - 1) / kFillGroup);
  const unsigned ntile = (unsigned)((h->N + kFillTile - 1) / kFillTile);
  const dim3 grid(ngrp * ngrp, ntile * ntile);
  if (grid.y > 65535) return fail(RCHEM_ERR_TOO_LARGE, "dense tensor: N too large for the fill grid");
  tensor_fill_kernel<<<grid, kFillTile * kFillTile, 0, h->stream>>>(
      I_dev, h->N, (int)h->shells.shells.size(), h->d_fn_shell, h->d_pair_key, h->d_pair_fwd);
  CUDA_OK(cudaGetLastError());
  h->stats.launches += 1;
  return RCHEM_OK;
}

int rchem_jk_inmem_device(int n, const double* I_dev, const double* D_dev, double* JK_dev,
                          void* cuda_stream) {
  if (n <= 0 || !I_dev || !D_dev || !JK_dev) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const size_t nn = (size_t)n * n;
  CUDA_OK(cudaMemsetAsync(JK_dev, 0, 2 * nn * sizeof(double), s));
  // depth of the load pipeline: enough slots to cover a k-row in one or two sweeps
  const int slots = (n + 31) / 32;
  if (slots <= 2) jk_inmem_kernel<2><<<(unsigned)nn, 256, 0, s>>>(I_dev, D_dev, JK_dev, n);
  else if (slots <= 4) jk_inmem_kernel<4><<<(unsigned)nn, 256, 0, s>>>(I_dev, D_dev, JK_dev, n);
  else jk_inmem_kernel<8><<<(unsigned)nn, 256, 0, s>>>(I_dev, D_dev, JK_dev, n);
  CUDA_OK(cudaGetLastError());
  return RCHEM_OK;
}


// ---------------- single-process multi-GPU J/K (RCHEM_OPT_NGPUS) ---------------------------
// One handle, one call, n devices of this node:
//   H2D of D once (device 0 of the group) -> peer copies of D over NVLink -> every device
//   builds its block-interleaved share (one host thread per device enqueues its launches) ->
//   device i sums slice i of the partial [J|K] out of the others' memory (peer_reduce_kernel) and
//   copies that slice to the host buffers over its own PCIe link (n concurrent D2H of 1/n each).
// D must be symmetric on this path (an asymmetric D takes the one-device path).
extern "C++" {
namespace {

int sync_peer_options(rchem_basis* h) {
  while ((int)h->peers.size() < h->ngpus - 1) {
    Basis copy = h->basis;
    rchem_basis* peer = nullptr;
    int rc = make_basis_handle(std::move(copy), &peer);
    if (rc) return rc;
    int ndev = 1;
    cudaGetDeviceCount(&ndev);
    peer->device = (h->device + 1 + (int)h->peers.size()) % std::max(1, ndev);
    peer->prim_eps = h->prim_eps;
    if (peer->fuse_sp != h->fuse_sp) {
      std::string err;
      if (!group_shells(peer->basis, &peer->shells, &err, h->fuse_sp != 0)) {
        rchem_basis_destroy(peer);
        return fail(RCHEM_ERR_UNSUPPORTED_LAYOUT, err);
      }
      peer->fuse_sp = h->fuse_sp;
    }
    h->peers.push_back(peer);
  }
  for (rchem_basis* peer : h->peers) {
    peer->boys = h->boys; peer->tau = h->tau; peer->far_sched = h->far_sched;
    peer->heavy_passes = h->heavy_passes; peer->light_kernel = h->light_kernel;
  }
  return RCHEM_OK;
}

// runs fn(i) for i = 1..n-1 on their own host threads and fn(0) here; first failure wins
template <class F> int for_each_device(int n, F fn) {
  std::vector<int> rc(n, RCHEM_OK);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i)
    th.emplace_back([&, i] { rc[i] = fn(i); if (rc[i]) msg[i] = g_err; });
  rc[0] = fn(0);
  if (rc[0]) msg[0] = g_err;
  for (std::thread& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i]) return fail(rc[i], "device " + std::to_string(i) + " of the group: " + msg[i]);
  return RCHEM_OK;
}

int jk_direct_multi(rchem_basis* h, double* J, double* K) {
  // (D is already in h->d_D on h->stream and was found symmetric)
  const int n = h->ngpus;
  const size_t nn = (size_t)h->N * h->N;
  auto handle = [&](int i) { return i == 0 ? h : h->peers[i - 1]; };
  if (!h->ev_D) CUDA_OK(cudaEventCreateWithFlags(&h->ev_D, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(h->ev_D, h->stream));

  // Can every device of the group read every other one's memory?  (probed once per group; the
  // same physical device appears twice only with RCHEM_MULTI_OVERSUBSCRIBE)
  if (h->group_p2p < 0 || (int)h->peers.size() != n - 1) {
    h->group_p2p = 1;
    for (int i = 0; i < n && h->group_p2p; ++i)
      for (int j = 0; j < n && h->group_p2p; ++j) {
        const int di = handle(i)->device, dj = handle(j)->device;
        if (di == dj) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, di, dj) != cudaSuccess || !can) h->group_p2p = 0;
      }
    cudaGetLastError();
  }
  const bool all_p2p = h->group_p2p == 1;

  // One host thread per device for the whole call: enqueue the device's share of the build,
  // meet the other threads (so every "build done" event exists), then -- with full peer access --
  // the sharded epilogue: device i waits for every build, sums slice i of [J|K] straight out of
  // the other devices' memory (NVLink peer loads) and copies that slice home over ITS OWN PCIe
  // link, so the reduction and the device-to-host copy both shrink with the number of GPUs.
  std::atomic<int> arrived{0}, failed{0};
  int rc = for_each_device(n, [&](int i) -> int {
    rchem_basis* g = handle(i);
    auto build = [&]() -> int {
      int r = ensure_ready(g);
      if (r) return r;
      CUDA_OK(cudaSetDevice(g->device));
      if (!g->ev_done) CUDA_OK(cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming));
      if (i > 0) {
        CUDA_OK(cudaStreamWaitEvent(g->stream, h->ev_D, 0));
        CUDA_OK(cudaMemcpyPeerAsync(g->d_D, g->device, h->d_D, h->device, nn * sizeof(double), g->stream));
      }
      r = jk_direct_device_impl(g, g->d_D, g->d_JK, i, n, 0);
      if (r) return r;
      CUDA_OK(cudaEventRecord(g->ev_done, g->stream));
      return RCHEM_OK;
    };
    const int r = build();
    if (r) failed.fetch_add(1);
    arrived.fetch_add(1);
    while (arrived.load() < n) std::this_thread::yield();  // host-side meeting point
    if (r) return r;
    if (failed.load() || !all_p2p) return RCHEM_OK;
    PeerPtrs pp{};
    const size_t n2 = nn;  // [J|K] = 2 nn doubles = nn double2
    const size_t lo = n2 * (size_t)i / n, hi = n2 * (size_t)(i + 1) / n;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      rchem_basis* o = handle(j);
      CUDA_OK(cudaStreamWaitEvent(g->stream, o->ev_done, 0));
      if (o->device != g->device && !g->peer_access_on) {
        cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(RCHEM_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      pp.p[pp.n++] = reinterpret_cast<const double2*>(o->d_JK) + lo;
    }
    g->peer_access_on = true;
    if (hi > lo) {
      peer_reduce_kernel<<<148, 256, 0, g->stream>>>(reinterpret_cast<double2*>(g->d_JK) + lo, pp, hi - lo);
      CUDA_OK(cudaGetLastError());
      // doubles [2 lo, 2 hi) of [J|K]: the part below nn goes to J, the rest to K
      const size_t a = 2 * lo, b = 2 * hi;
      if (a < nn)
        CUDA_OK(cudaMemcpyAsync(J + a, g->d_JK + a, (std::min(b, nn) - a) * sizeof(double),
                                cudaMemcpyDeviceToHost, g->stream));
      if (b > nn) {
        const size_t k0 = std::max(a, nn);
        CUDA_OK(cudaMemcpyAsync(K + (k0 - nn), g->d_JK + k0, (b - k0) * sizeof(double),
                                cudaMemcpyDeviceToHost, g->stream));
      }
    }
    CUDA_OK(cudaStreamSynchronize(g->stream));
    return RCHEM_OK;
  });
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));

  if (!all_p2p) {
    // No full peer access: device 0 of the group gathers staged copies and sums them.
    PeerPtrs pp{};
    h->peer_stage.resize(n - 1, nullptr);
    for (int i = 1; i < n; ++i) {
      rchem_basis* g = handle(i);
      CUDA_OK(cudaStreamWaitEvent(h->stream, g->ev_done, 0));
      if (!h->peer_stage[i - 1]) CUDA_OK(cudaMalloc(&h->peer_stage[i - 1], 2 * nn * sizeof(double)));
      CUDA_OK(cudaMemcpyPeerAsync(h->peer_stage[i - 1], h->device, g->d_JK, g->device,
                                  2 * nn * sizeof(double), h->stream));
      pp.p[pp.n++] = reinterpret_cast<const double2*>(h->peer_stage[i - 1]);
    }
    peer_reduce_kernel<<<148 * 4, 256, 0, h->stream>>>(reinterpret_cast<double2*>(h->d_JK), pp, nn);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(J, h->d_JK, nn * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaMemcpyAsync(K, h->d_JK + nn, nn * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  // whole-job statistics: counts summed over the group, launches likewise
  for (int i = 1; i < n; ++i) {
    const rchem_stats& s = handle(i)->stats;
    h->stats.shell_quartets += s.shell_quartets;
    h->stats.fused_quartets += s.fused_quartets;
    h->stats.prim_quartets += s.prim_quartets;
    h->stats.prim_quartets_evaluated += s.prim_quartets_evaluated;
    h->stats.integrals += s.integrals;
    h->stats.model_flops += s.model_flops;
    h->stats.launches += s.launches;
  }
  h->stats.launches += all_p2p ? n : 1;  // peer_reduce_kernel
  return RCHEM_OK;
}

}  // namespace
}  // extern "C++"

// ---------------- host-buffer entry points (copies inside) -------------------------------
int rchem_jk_direct(rchem_basis* h, const double* D, double* J, double* K) {
  if (!h || !D || !J || !K) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  const int N = h->N;
  const size_t nn = (size_t)N * N;
  int rc = ensure_ready(h);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));
  rc = order_after_previous_build(h);
  if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(h->d_D, D, nn * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  // The digestion exploits D = D^T (what the reference's caller passes, rchem.rs:101-104).
  // The reference itself accepts any D (basis.rs:383-428 uses no symmetry), so an asymmetric
  // one is split on the device: J(D) = J(S) and K(D) = K(S) + K(A), S/A the symmetric /
  // antisymmetric parts, K(A) antisymmetric -- a second build, only when it is needed.
  CUDA_OK(cudaMemsetAsync(h->d_asym, 0, 2 * sizeof(double), h->stream));
  asym_probe_kernel<<<148, 256, 0, h->stream>>>(h->d_D, N, reinterpret_cast<unsigned long long*>(h->d_asym));
  double* probe = h->h_asym;
  probe[0] = probe[1] = 0.0;
  CUDA_OK(cudaMemcpyAsync(probe, h->d_asym, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (h->ngpus > 1) {
    // Several devices: do not wait for the probe.  The symmetric build is started at once, so the
    // per-device host threads and their launches overlap the upload of D; the verdict is read
    // after the call has finished anyway, and an asymmetric D (which the reference's caller never
    // passes) then takes the one-device split path below.
    rc = sync_peer_options(h);
    if (rc) return rc;
    rc = jk_direct_multi(h, J, K);  // (returns with h->stream drained: the probe has landed)
    if (rc) return rc;
    if (!(probe[1] > 1e-14 * std::max(probe[0], 1e-300))) return RCHEM_OK;
  } else {
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  if (probe[1] > 1e-14 * std::max(probe[0], 1e-300)) {
    if (h->symmetric_only)
      return fail(RCHEM_ERR_ASYMMETRIC_D, "JK_direct: the density matrix is not symmetric "
                                          "(RCHEM_OPT_SYMMETRIC_D_ONLY is set)");
    double *dS = nullptr, *dA = nullptr;
    CUDA_OK(cudaMalloc(&dS, nn * sizeof(double)));
    cudaError_t e = cudaMalloc(&dA, nn * sizeof(double));
    if (e != cudaSuccess) { cudaFree(dS); return fail(RCHEM_ERR_CUDA, cudaGetErrorString(e)); }
    split_density_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(h->d_D, N, dS, dA);
    rc = jk_direct_device_impl(h, dS, h->d_JK, 0, 1, 0);
    rchem_stats first = h->stats;
    if (rc == RCHEM_OK) rc = jk_direct_device_impl(h, dA, h->d_JK, 0, 1, 1);
    if (rc == RCHEM_OK) {
      h->stats.shell_quartets += first.shell_quartets;
      h->stats.fused_quartets += first.fused_quartets;
      h->stats.prim_quartets_evaluated += first.prim_quartets_evaluated;
      h->stats.prim_quartets += first.prim_quartets;
      h->stats.integrals += first.integrals;
      h->stats.model_flops += first.model_flops;
      h->stats.launches += first.launches;
    }
    cudaStreamSynchronize(h->stream);
    cudaFree(dS); cudaFree(dA);
    if (rc) return rc;
  } else {
    rc = rchem_jk_direct_device(h, h->d_D, h->d_JK, 0, 1);
    if (rc) return rc;
  }
  CUDA_OK(cudaMemcpyAsync(J, h->d_JK, nn * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaMemcpyAsync(K, h->d_JK + nn, nn * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return RCHEM_OK;
}

int rchem_build_I(rchem_basis* h, double* I) {
  if (!h || !I) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  int rc = ensure_ready(h);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));
  const size_t n4 = (size_t)h->N * h->N * h->N * h->N;
  size_t free_b = 0, total_b = 0;
  CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
  if (n4 * sizeof(double) > free_b)
    return fail(RCHEM_ERR_TOO_LARGE, "dense N^4 tensor does not fit in device memory");
  double* dI = nullptr;
  CUDA_OK(cudaMalloc(&dI, n4 * sizeof(double)));
  rc = rchem_build_I_device(h, dI);
  if (rc == RCHEM_OK) {
    cudaError_t e = cudaMemcpyAsync(I, dI, n4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(RCHEM_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(dI);
  return rc;
}

int rchem_jk_inmem(int n, const double* I, const double* D, double* J, double* K) {
  if (n <= 0 || !I || !D || !J || !K) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(RCHEM_ERR_NO_DEVICE, "no CUDA device: librchem_b200 has no CPU path");
  const size_t nn = (size_t)n * n, n4 = nn * nn;
  double *dI = nullptr, *dD = nullptr, *dJK = nullptr;
  CUDA_OK(cudaMalloc(&dI, n4 * sizeof(double)));
  cudaError_t e = cudaMalloc(&dD, nn * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dJK, 2 * nn * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(dI, I, n4 * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dD, D, nn * sizeof(double), cudaMemcpyHostToDevice);
  int rc = RCHEM_OK;
  if (e == cudaSuccess) rc = rchem_jk_inmem_device(n, dI, dD, dJK, nullptr);
  if (e == cudaSuccess && rc == RCHEM_OK) e = cudaMemcpy(J, dJK, nn * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && rc == RCHEM_OK) e = cudaMemcpy(K, dJK + nn, nn * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dI); cudaFree(dD); cudaFree(dJK);
  if (e != cudaSuccess) return fail(RCHEM_ERR_CUDA, cudaGetErrorString(e));
  return rc;
}

// ---------------- one-electron matrices -------------------------------------
static int one_electron(rchem_basis* h, int which, int natoms, const double* atomcoords,
                        const uint64_t* atomnos, double* M) {
  if (!h || !M) return fail(RCHEM_ERR_INVALID_ARG, "null argument");
  if (which == 2 && (natoms <= 0 || !atomcoords || !atomnos))
    return fail(RCHEM_ERR_INVALID_ARG, "V needs the atom list");
  const int n = h->N;
  std::vector<double> origins(3 * (size_t)n), exps(h->nprim), coefs(h->nprim), norms(h->nprim), Z;
  std::vector<int> powers(3 * (size_t)n), off(n + 1);
  int rc = rchem_basis_export(h, origins.data(), powers.data(), off.data(), exps.data(),
                              coefs.data(), norms.data());
  if (rc) return rc;
  for (int c = 0; c < natoms && which == 2; ++c) Z.push_back((double)atomnos[c]);
  return one_electron_host(n, origins.data(), powers.data(), off.data(), exps.data(), coefs.data(),
                           norms.data(), which, natoms, atomcoords, Z.data(), M, h->device);
}

int rchem_overlap(rchem_basis* h, double* S) { return one_electron(h, 0, 0, nullptr, nullptr, S); }
int rchem_kinetic(rchem_basis* h, double* T) { return one_electron(h, 1, 0, nullptr, nullptr, T); }
int rchem_nuclear(rchem_basis* h, int natoms, const double* atomcoords, const uint64_t* atomnos,
                  double* V) {
  return one_electron(h, 2, natoms, atomcoords, atomnos, V);
}

// ---------------- screening ---------------------------------------------------------------
int64_t rchem_schwarz(rchem_basis* h, int32_t* shell_a, int32_t* shell_b, int32_t* batch,
                      double* Q) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  int rc = ensure_ready(h);
  if (rc) return rc;
  int64_t n = 0;
  for (size_t bi = 0; bi < h->batches.size(); ++bi) {
    const Batch& bt = h->batches[bi];
    if (bt.sub) continue;  // (segmented twins of fused pairs: an evaluation detail, same Q)
    for (int p = 0; p < bt.npairs; ++p, ++n) {
      if (shell_a) shell_a[n] = bt.shA[p];
      if (shell_b) shell_b[n] = bt.shB[p];
      if (batch) batch[n] = (int32_t)bi;
      if (Q) Q[n] = bt.Q[p];
    }
  }
  return n;
}

int64_t rchem_quartet_list(rchem_basis* h, int64_t* pq, int64_t capacity) {
  if (!h) return fail(RCHEM_ERR_INVALID_ARG, "null handle");
  int rc = ensure_ready(h);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(h->device));
  rc = ensure_tasks(h);
  if (rc) return rc;
  int64_t total = 0;
  for (const TaskTable& tt : h->tasks) total += tt.sub ? 0 : tt.nquartets;
  if (!pq) return total;
  if (capacity < total) return fail(RCHEM_ERR_INVALID_ARG, "quartet list buffer too small");
  std::vector<long long> pair_off(h->batches.size() + 1, 0);
  for (size_t i = 0; i < h->batches.size(); ++i)  // (sub batches come last: offsets of the rest unaffected)
    pair_off[i + 1] = pair_off[i] + h->batches[i].npairs;
  long long* d_out = nullptr;
  CUDA_OK(cudaMalloc(&d_out, std::max<int64_t>(1, total) * 2 * sizeof(long long)));
  int64_t off = 0;
  for (const TaskTable& tt : h->tasks) {
    const Batch& B = h->batches[tt.bra];
    if (tt.nquartets == 0 || tt.sub) continue;
    std::vector<long long> qprefix(B.npairs + 1, 0);
    for (int p = 0; p < B.npairs; ++p) qprefix[p + 1] = qprefix[p] + tt.h_nq[p];
    long long* d_qp = nullptr;
    CUDA_OK(cudaMalloc(&d_qp, qprefix.size() * sizeof(long long)));
    CUDA_OK(cudaMemcpy(d_qp, qprefix.data(), qprefix.size() * sizeof(long long), cudaMemcpyHostToDevice));
    quartet_list_kernel<<<B.npairs, 128, 0, h->stream>>>(d_qp, tt.d_nq, B.npairs, pair_off[tt.bra],
                                                          pair_off[tt.ket], d_out + 2 * off);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaFree(d_qp));
    off += tt.nquartets;
  }
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  CUDA_OK(cudaMemcpy(pq, d_out, (size_t)total * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaFree(d_out));
  return total;
}

}  // extern "C"
