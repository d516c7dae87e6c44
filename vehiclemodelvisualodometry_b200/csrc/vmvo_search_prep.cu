// The MODE 2 instantiations of the fused window search (lean kernels that take preparation records);
// see vmvo_search_kernels.cuh.  A translation unit of its own so that the three modes compile side by side.
#include "vmvo_search_kernels.cuh"

namespace vmvo {

int launch_search_mode2(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st, bool dual, bool imu, bool f64) {
  return launch_search_mode<2>(ctx, p, st, dual, imu, f64);
}

}  // namespace vmvo
