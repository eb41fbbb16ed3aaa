// eri_class_tu.cu -- one translation unit per angular-momentum class AND Boys flavour.
// Compiled 42 times with -DRCHEM_LA=.. -DRCHEM_LB=.. -DRCHEM_LC=.. -DRCHEM_LD=..,
// -DRCHEM_TAG=<abcd>, -DRCHEM_BOYS=<0|1> and -DRCHEM_INC="gen/eri_class_<abcd>.inc" (see
// Makefile), so the classes -- and the two flavours of the big (dd|dd)-type classes, which
// dominate the build time -- compile in parallel.
#include "eri_kernel.cuh"

#define RCHEM_CAT2(a, b) a##b
#define RCHEM_CAT(a, b) RCHEM_CAT2(a, b)
#define RCHEM_SUFFIX RCHEM_CAT(RCHEM_CAT(RCHEM_TAG, _b), RCHEM_BOYS)

namespace rchem {

#include RCHEM_INC

template <int MODE>
static cudaError_t launch_one(const EriTask& task, unsigned grid, cudaStream_t stream) {
  eri_kernel<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD, RCHEM_BOYS, MODE>
      <<<grid, kThreads, 0, stream>>>(task);
  return cudaGetLastError();
}

cudaError_t RCHEM_CAT(launch_eri_, RCHEM_SUFFIX)(int mode, const EriTask& task, unsigned grid,
                                                 cudaStream_t stream) {
  if (grid == 0) return cudaSuccess;
  if (mode == kModeJK) return launch_one<kModeJK>(task, grid, stream);
  if (mode == kModeTensor) return launch_one<kModeTensor>(task, grid, stream);
#if RCHEM_LA == RCHEM_LC && RCHEM_LB == RCHEM_LD
  if (mode == kModeSchwarz) return launch_one<kModeSchwarz>(task, grid, stream);
#endif
  return cudaErrorInvalidValue;
}

// The block kernel is only built for classes with <= 100 VRR targets: beyond that the
// quartet itself dominates (and the register-bound kernel would not profit), so those
// classes digest through the warp kernel.
static constexpr bool kHasBlockKernel =
    EriClass<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD>::kTargets <= 100;

cudaError_t RCHEM_CAT(launch_eri_block_, RCHEM_SUFFIX)(const EriTask& task, unsigned grid,
                                                       size_t smem, cudaStream_t stream) {
  if (grid == 0) return cudaSuccess;
  if constexpr (kHasBlockKernel) {
    auto kern = eri_jk_block_kernel<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD, RCHEM_BOYS>;
    // the opt-in limit is per device; remember what was configured for each
    static size_t configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || smem > configured[dev]) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < 64) configured[dev] = smem;
    }
    kern<<<grid, BlockCfg<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD>::kThreadsBlk, smem, stream>>>(task);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}

// warp-per-bra-pair kernel for the light bra pairs (same classes as the block kernel)
cudaError_t RCHEM_CAT(launch_eri_light_, RCHEM_SUFFIX)(const EriTask& task, unsigned grid,
                                                       size_t smem, cudaStream_t stream) {
  if (grid == 0) return cudaSuccess;
  if constexpr (kHasBlockKernel) {
    if (smem > 48 * 1024) return cudaErrorInvalidValue;  // (the engine keeps it below)
    eri_jk_light_kernel<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD, RCHEM_BOYS>
        <<<grid, kThreads, smem, stream>>>(task);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}

cudaError_t RCHEM_CAT(launch_eri_light_multi_, RCHEM_SUFFIX)(const EriTask* tasks,
                                                             const int* blk_prefix, int ntasks,
                                                             unsigned grid, size_t smem,
                                                             cudaStream_t stream) {
  if (grid == 0) return cudaSuccess;
  if constexpr (kHasBlockKernel) {
    if (smem > 40 * 1024) return cudaErrorInvalidValue;  // (the engine keeps it below)
    eri_jk_light_multi_kernel<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD, RCHEM_BOYS>
        <<<grid, kThreads, smem, stream>>>(tasks, blk_prefix, ntasks);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}

#if RCHEM_BOYS == 0
EriBlockInfo RCHEM_CAT(block_info_, RCHEM_TAG)() {
  using Cfg = BlockCfg<RCHEM_LA, RCHEM_LB, RCHEM_LC, RCHEM_LD>;
  if constexpr (kHasBlockKernel) return EriBlockInfo{Cfg::kThreadsBlk, Cfg::kKetsPerBlock};
  return EriBlockInfo{0, 0};
}
#endif

}  // namespace rchem
