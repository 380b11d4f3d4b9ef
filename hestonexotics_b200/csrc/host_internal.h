// hestonexotics_b200/csrc/host_internal.h -- host-side pieces shared between translation units
#pragma once
#include <vector>

#include "../../include/hexo_gpu.h"
#include "qe.cuh"

namespace hexo {

// step schedule + per-maturity constants of a request (hexo_gpu.cu)
int build_request_segments(const hexo_price_request* r, std::vector<SegConst>& segs);

// E max(G - K_j, 0) of the geometric-Asian control for every option (geo_asian_host.cu)
int geometric_asian_means(const hexo_price_request* r, const std::vector<SegConst>& segs,
                          std::vector<double>& out);

}  // namespace hexo
