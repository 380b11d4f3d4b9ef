// The default path kernels: shishua generator, plain sums (12 instantiations).
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_shishua(int payoff, int normal_mode, int segs) {
  return select_path_kernel<Shishua, false>(payoff, normal_mode, segs);
}
}  // namespace hexo
