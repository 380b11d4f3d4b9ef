// Path kernels of the optional Philox4x32-10 mode with the control-variate sums.
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_philox_plain(int payoff, int normal_mode, int segs);  // path_kernels_philox.cu
PathKernel path_kernel_philox(int payoff, int normal_mode, int segs, bool cv) {
  return cv ? select_path_kernel<PhiloxGen, true, false, 1>(payoff, normal_mode, segs)
            : path_kernel_philox_plain(payoff, normal_mode, segs);
}
}  // namespace hexo
