// hestonexotics_b200/csrc/path_kernels.h
//
// Selectors for the instantiations of the path kernel.  Each family is compiled in its own
// translation unit (path_kernels_*.cu) so that the library builds in parallel; hexo_gpu.cu only
// sees function pointers.
#pragma once
#include "path_kernel.cuh"

namespace hexo {

typedef void (*PathKernel)(const PathArgs);

// payoff: hexo_payoff, normal_mode: hexo_normal_mode, segs: kSegsGlobal / kSegsInline / kSegsSingle
PathKernel path_kernel_shishua(int payoff, int normal_mode, int segs);     // the default
PathKernel path_kernel_shishua_ppnd7(int payoff, int segs);  // HEXO_NORMAL_F32_PPND7, plain sums
// HEXO_CV_GEOMETRIC (Asian payoff): shishua generator; mart selects HEXO_DRIFT_MARTINGALE
PathKernel path_kernel_shishua_geo(int normal_mode, int segs, bool mart);
PathKernel path_kernel_shishua_cv(int payoff, int normal_mode, int segs);  // + control variate
PathKernel path_kernel_philox(int payoff, int normal_mode, int segs, bool cv);
// HEXO_DRIFT_MARTINGALE (path_kernels_*_mart*.cu)
PathKernel path_kernel_shishua_mart(int payoff, int normal_mode, int segs, bool cv);
PathKernel path_kernel_philox_mart(int payoff, int normal_mode, int segs, bool cv);

// shared by the selector translation units
// VARIANTS: which of the three sources of per-maturity constants get their own instantiation
// (each costs 4 kernels per family): 3 = single-maturity, parameter-bank and device-memory variants
// (the default family); 2 = single-maturity + device memory; 1 = device memory only (the optional
// Philox families).  In the step loop the parameter-bank and device-memory variants are the
// same code -- both pin the constants in registers.
template <class Gen, int CV, bool MART = false, int VARIANTS = 3>
inline PathKernel select_path_kernel(int payoff, int normal_mode, int segs) {
  constexpr int kOne = VARIANTS >= 2 ? kSegsSingle : kSegsGlobal;
  constexpr int kFew = VARIANTS >= 3 ? kSegsInline : kSegsGlobal;
#define HEXO_PICK(P, N)                                                             \
  (segs == kSegsSingle   ? heston_qe_paths_kernel<P, N, kOne, Gen, CV, MART>        \
   : segs == kSegsInline ? heston_qe_paths_kernel<P, N, kFew, Gen, CV, MART>        \
                         : heston_qe_paths_kernel<P, N, kSegsGlobal, Gen, CV, MART>)
  if (payoff == HEXO_PAYOFF_ASIAN)
    return normal_mode == HEXO_NORMAL_F64 ? HEXO_PICK(HEXO_PAYOFF_ASIAN, 1)
                                          : HEXO_PICK(HEXO_PAYOFF_ASIAN, 0);
  return normal_mode == HEXO_NORMAL_F64 ? HEXO_PICK(HEXO_PAYOFF_EUROPEAN, 1)
                                        : HEXO_PICK(HEXO_PAYOFF_EUROPEAN, 0);
#undef HEXO_PICK
}

}  // namespace hexo
