// tx_ssb.cu -- SSB transmit modulator (TxProcessor_Run SSB branch, tx_processor.c:891-1078).
#include "dsp_device.cuh"
#include "kernels.h"
#include "host_tables.h"

namespace uhsdr {

int build_tx_params(const HostTables &t, const uhsdr_chan_cfg_t &cfg, TxParams *out, std::string *err)
{
    (void)t; (void)cfg; (void)err;
    memset(out, 0, sizeof(*out));
    return UHSDR_OK;
}

cudaError_t launch_tx_ssb(const TxArgs &a, cudaStream_t stream) { (void)a; (void)stream; return cudaErrorNotSupported; }

}  // namespace uhsdr
