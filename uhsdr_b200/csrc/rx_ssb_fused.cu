// rx_ssb_fused.cu -- fused narrow-SSB receiver kernel (placeholder: not yet eligible for any channel).
#include "dsp_device.cuh"
#include "kernels.h"

namespace uhsdr {
bool fused_eligible(const ChanParams &p) { (void)p; return false; }
void fill_fused_coefs(FusedCoefs *fc, const float *d, const float *hi, const float *hq) { (void)fc; (void)d; (void)hi; (void)hq; }
cudaError_t launch_rx_ssb_fused(const RxArgs &a, const FusedCoefs &fc, int sm_count, cudaStream_t stream)
{
    (void)a; (void)fc; (void)sm_count; (void)stream;
    return cudaErrorNotSupported;
}
}  // namespace uhsdr
