// Path kernels of HEXO_DRIFT_MARTINGALE: Philox4x32-10 generator, plain sums.
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_philox_mart_cv(int payoff, int normal_mode, int segs);  // ..._mart_cv.cu
PathKernel path_kernel_philox_mart(int payoff, int normal_mode, int segs, bool cv) {
  return cv ? path_kernel_philox_mart_cv(payoff, normal_mode, segs)
            : select_path_kernel<PhiloxGen, false, true, 1>(payoff, normal_mode, segs);
}
}  // namespace hexo
