// Path kernels of HEXO_DRIFT_MARTINGALE with the control-variate sums: shishua generator.
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_shishua_mart_cv(int payoff, int normal_mode, int segs) {
  return select_path_kernel<Shishua, true, true, 2>(payoff, normal_mode, segs);
}
}  // namespace hexo
