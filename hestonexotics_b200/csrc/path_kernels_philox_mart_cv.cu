// Path kernels of HEXO_DRIFT_MARTINGALE with the control-variate sums: Philox4x32-10 generator.
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_philox_mart_cv(int payoff, int normal_mode, int segs) {
  return select_path_kernel<PhiloxGen, true, true, 1>(payoff, normal_mode, segs);
}
}  // namespace hexo
