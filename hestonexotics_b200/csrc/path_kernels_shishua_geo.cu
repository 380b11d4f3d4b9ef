// Path kernels of the geometric-Asian control variate (HEXO_CV_GEOMETRIC): Asian payoff, shishua
// generator, reference or martingale-corrected drift.
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_shishua_geo(int normal_mode, int segs, bool mart) {
#define HEXO_PICKG(N, M)                                                                         \
  (segs == kSegsSingle ? heston_qe_paths_kernel<HEXO_PAYOFF_ASIAN, N, kSegsSingle, Shishua, 2, M> \
                       : heston_qe_paths_kernel<HEXO_PAYOFF_ASIAN, N, kSegsGlobal, Shishua, 2, M>)
  if (mart) return normal_mode == HEXO_NORMAL_F64 ? HEXO_PICKG(1, true) : HEXO_PICKG(0, true);
  return normal_mode == HEXO_NORMAL_F64 ? HEXO_PICKG(1, false) : HEXO_PICKG(0, false);
#undef HEXO_PICKG
}
}  // namespace hexo
