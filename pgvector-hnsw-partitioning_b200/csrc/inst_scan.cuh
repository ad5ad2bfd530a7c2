// inst_scan.cuh -- instantiates the scan kernels for one (storage type, metric) pair.  Included
// by scan_<type>_<metric>.cu with HBI_T / HBI_IP / HBI_NAME defined, so the four pairs compile in
// parallel.
#include "index.h"
#include "scan_kernel.cuh"

namespace hb {
cudaError_t HB_CAT(scan_fast_, HBI_NAME)(const ScanParams &p, int sms, int mg, cudaStream_t s, ScanLaunchInfo *i)
{
    return launch_scan_t<HBI_T, HBI_IP, false>(p, sms, mg, s, i);
}
cudaError_t HB_CAT(scan_slow_, HBI_NAME)(const ScanParams &p, int sms, int mg, cudaStream_t s, ScanLaunchInfo *i)
{
    return launch_scan_t<HBI_T, HBI_IP, true>(p, sms, mg, s, i);
}
cudaError_t HB_CAT(dist_, HBI_NAME)(const DistBatchParams &p, cudaStream_t s) { return launch_dist_t<HBI_T, HBI_IP>(p, s); }
}   // namespace hb
