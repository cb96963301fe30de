// host_tables.h -- host-side view of the coefficient-table blob (include/uhsdr_tables.h) and the
// derivation of per-channel kernel parameters from uhsdr_chan_cfg_t.
#pragma once
#include <string>
#include <vector>

#include "uhsdr_b200.h"
#include "uhsdr_dev.h"
#include "uhsdr_tables.h"

namespace uhsdr {

struct HostTables {
    std::vector<uint8_t> blob;
    const uhsdr_tbl_header_t *h = nullptr;
    const uhsdr_tbl_array_t *arr = nullptr;
    const uhsdr_tbl_path_t *path = nullptr;
    const uhsdr_tbl_filter_t *filt = nullptr;
    const uhsdr_tbl_lattice_t *lat = nullptr;
    const uhsdr_tbl_interp_t *interp = nullptr;
    const uhsdr_tbl_extras_t *ex = nullptr;
    // device coefficient pool: every blob array back to back (16-byte aligned starts), then
    // engine-generated tables (FFT twiddles)
    std::vector<float> pool;
    std::vector<int> pool_off;      // array index -> float offset in pool
    int tw256_off = -1, tw512_off = -1;   // cos/sin twiddle tables for the 256/512-point FFTs

    bool load(const void *data, size_t bytes, std::string *err);
    int off(int array_idx) const { return (array_idx >= 0 && array_idx < (int)pool_off.size()) ? pool_off[array_idx] : -1; }
    const float *host_array(int array_idx) const { return pool.data() + pool_off[array_idx]; }
};

// AudioDriver_SetProcessingChain (audio_driver.c:1093-1251) as a pure function of the
// configuration: everything the kernels need, computed with the host libm so that the constants
// are bit-identical to the reference's (SURVEY.md section 7 "libm differences").
int build_chan_params(const HostTables &t, const uhsdr_chan_cfg_t &cfg, ChanParams *out, std::string *err);

int build_tx_params(const HostTables &t, const uhsdr_chan_cfg_t &cfg, TxParams *out, std::string *err);

void default_chan_cfg(uhsdr_chan_cfg_t *cfg);

}  // namespace uhsdr
