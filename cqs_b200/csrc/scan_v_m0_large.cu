// scan_topk_kernel<MODE = 0, NV = 1..16, SMALLK = false> — see scan_single.cu / scan_single_impl.cuh.
#include "scan_single_impl.cuh"

namespace cqs {
cudaError_t launch_scan_m0_large(const ScanParams& p, int nv, int num_sms, cudaStream_t st) {
  return launch_variant<0, false>(p, nv, num_sms, st);
}
}  // namespace cqs
