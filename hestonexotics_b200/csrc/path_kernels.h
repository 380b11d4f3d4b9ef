// hestonexotics_b200/csrc/path_kernels.h
//
// Selectors for the instantiations of the path kernel.  Each family is compiled in its own
// translation unit (path_kernels_*.cu) so that the library builds in parallel; hexo_gpu.cu only
// sees function pointers.
#pragma once
#include "path_kernel.cuh"
#include "path_kernel_il.cuh"
#include "path_kernel_ws.cuh"

namespace hexo {

typedef void (*PathKernel)(const PathArgs);
typedef void (*PathKernelWs)(const PathArgs, const uint32_t);

// payoff: hexo_payoff, normal_mode: hexo_normal_mode, segs: kSegsGlobal / kSegsInline / kSegsSingle
PathKernel path_kernel_shishua(int payoff, int normal_mode, int segs);     // the default
PathKernel path_kernel_shishua_cv(int payoff, int normal_mode, int segs);  // + control variate
PathKernel path_kernel_philox(int payoff, int normal_mode, int segs, bool cv);
// HEXO_DRIFT_MARTINGALE (path_kernels_*_mart*.cu)
PathKernel path_kernel_shishua_mart(int payoff, int normal_mode, int segs, bool cv);
PathKernel path_kernel_philox_mart(int payoff, int normal_mode, int segs, bool cv);
// experimental variants (HEXO_WS=1 / HEXO_IL=1)
PathKernelWs path_kernel_ws(int payoff, int normal_mode, bool inline_segs);
PathKernel path_kernel_il(int payoff, int normal_mode, bool inline_segs);

// shared by the selector translation units
template <class Gen, bool CV, bool MART = false>
inline PathKernel select_path_kernel(int payoff, int normal_mode, int segs) {
#define HEXO_PICK(P, N)                                                             \
  (segs == kSegsSingle   ? heston_qe_paths_kernel<P, N, kSegsSingle, Gen, CV, MART> \
   : segs == kSegsInline ? heston_qe_paths_kernel<P, N, kSegsInline, Gen, CV, MART> \
                         : heston_qe_paths_kernel<P, N, kSegsGlobal, Gen, CV, MART>)
  if (payoff == HEXO_PAYOFF_ASIAN)
    return normal_mode == HEXO_NORMAL_F64 ? HEXO_PICK(HEXO_PAYOFF_ASIAN, 1)
                                          : HEXO_PICK(HEXO_PAYOFF_ASIAN, 0);
  return normal_mode == HEXO_NORMAL_F64 ? HEXO_PICK(HEXO_PAYOFF_EUROPEAN, 1)
                                        : HEXO_PICK(HEXO_PAYOFF_EUROPEAN, 0);
#undef HEXO_PICK
}

}  // namespace hexo
