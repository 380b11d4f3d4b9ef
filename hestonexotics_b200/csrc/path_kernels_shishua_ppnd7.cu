// Path kernels of the optional HEXO_NORMAL_F32_PPND7 mode: shishua generator, plain sums.
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_shishua_ppnd7(int payoff, int segs) {
#define HEXO_PICK7(P)                                                                              \
  (segs == kSegsSingle   ? heston_qe_paths_kernel<P, HEXO_NORMAL_F32_PPND7, kSegsSingle, Shishua>  \
   : segs == kSegsInline ? heston_qe_paths_kernel<P, HEXO_NORMAL_F32_PPND7, kSegsInline, Shishua>  \
                         : heston_qe_paths_kernel<P, HEXO_NORMAL_F32_PPND7, kSegsGlobal, Shishua>)
  return payoff == HEXO_PAYOFF_ASIAN ? HEXO_PICK7(HEXO_PAYOFF_ASIAN) : HEXO_PICK7(HEXO_PAYOFF_EUROPEAN);
#undef HEXO_PICK7
}
}  // namespace hexo
