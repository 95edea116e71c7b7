// scan_topk_kernel<MODE = 0, NV = 1..16, SMALLK = true> — see scan_single.cu / scan_single_impl.cuh.
#include "scan_single_impl.cuh"

namespace cqs {
cudaError_t launch_scan_m0_small(const ScanParams& p, int nv, int num_sms, cudaStream_t st) {
  return launch_variant<0, true>(p, nv, num_sms, st);
}
}  // namespace cqs
