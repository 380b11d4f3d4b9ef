// Path kernels of the optional Philox4x32-10 mode, plain sums (12 instantiations).
#include "path_kernels.h"
namespace hexo {
PathKernel path_kernel_philox_plain(int payoff, int normal_mode, int segs) {
  return select_path_kernel<PhiloxGen, false, false, 1>(payoff, normal_mode, segs);
}
}  // namespace hexo
