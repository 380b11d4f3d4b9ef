// Experimental variants of the path kernel, selected by HEXO_WS=1 / HEXO_IL=1: producer /
// consumer warp specialisation (path_kernel_ws.cuh) and interleaved look-ahead normals
// (path_kernel_il.cuh).  Measured, parity-tested, not faster than the default.
#include "path_kernels.h"
namespace hexo {
template <bool INL>
static PathKernelWs ws_t(int payoff, int normal_mode) {
  if (payoff == HEXO_PAYOFF_ASIAN)
    return normal_mode == HEXO_NORMAL_F64 ? heston_qe_paths_ws_kernel<HEXO_PAYOFF_ASIAN, 1, INL>
                                          : heston_qe_paths_ws_kernel<HEXO_PAYOFF_ASIAN, 0, INL>;
  return normal_mode == HEXO_NORMAL_F64 ? heston_qe_paths_ws_kernel<HEXO_PAYOFF_EUROPEAN, 1, INL>
                                        : heston_qe_paths_ws_kernel<HEXO_PAYOFF_EUROPEAN, 0, INL>;
}
template <bool INL>
static PathKernel il_t(int payoff, int normal_mode) {
  if (payoff == HEXO_PAYOFF_ASIAN)
    return normal_mode == HEXO_NORMAL_F64 ? heston_qe_paths_il_kernel<HEXO_PAYOFF_ASIAN, 1, INL>
                                          : heston_qe_paths_il_kernel<HEXO_PAYOFF_ASIAN, 0, INL>;
  return normal_mode == HEXO_NORMAL_F64 ? heston_qe_paths_il_kernel<HEXO_PAYOFF_EUROPEAN, 1, INL>
                                        : heston_qe_paths_il_kernel<HEXO_PAYOFF_EUROPEAN, 0, INL>;
}
PathKernelWs path_kernel_ws(int payoff, int normal_mode, bool inline_segs) {
  return inline_segs ? ws_t<true>(payoff, normal_mode) : ws_t<false>(payoff, normal_mode);
}
PathKernel path_kernel_il(int payoff, int normal_mode, bool inline_segs) {
  return inline_segs ? il_t<true>(payoff, normal_mode) : il_t<false>(payoff, normal_mode);
}
}  // namespace hexo
