// Path kernels with the control-variate sums: shishua generator (12 instantiations).
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_shishua_cv(int payoff, int normal_mode, int segs) {
  return select_path_kernel<Shishua, true, false, 2>(payoff, normal_mode, segs);
}
}  // namespace hexo
