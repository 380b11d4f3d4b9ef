// Path kernels of HEXO_DRIFT_MARTINGALE: shishua generator, plain sums (12 instantiations).
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_shishua_mart_cv(int payoff, int normal_mode, int segs);  // ..._mart_cv.cu
PathKernel path_kernel_shishua_mart(int payoff, int normal_mode, int segs, bool cv) {
  return cv ? path_kernel_shishua_mart_cv(payoff, normal_mode, segs)
            : select_path_kernel<Shishua, false, true, 2>(payoff, normal_mode, segs);
}
}  // namespace hexo
