// multi.cpp -- one handle over several GPUs of a box (SURVEY.md 8b "(num_channels, device list)", 8e): channels are independent
// (no shared state and no cross-channel reduction anywhere in AudioDriver_RxProcessor, audio_driver.c:2603-2942), so the
// global channel range is cut into contiguous per-device ranges [g C / G, (g + 1) C / G) on the host, every device gets its own
// engine (state arena, streams, staging), and a call is forwarded to the engines from one host thread per device.  No
// collective, no peer traffic.  Plain host C++ on top of the single-device C ABI of engine.cu.
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "uhsdr_b200.h"

struct uhsdr_multi {
    int nch = 0;
    std::vector<uhsdr_engine_t *> eng;
    std::vector<int> first, count, device;
    std::string last_error;
};

extern "C" {

int uhsdr_channel_range(int rank, int world, int total, int *first, int *count)
{
    if (world < 1 || rank < 0 || rank >= world || total < 0 || !first || !count) return UHSDR_ERR_ARG;
    const int base = total / world, rem = total % world;
    *first = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
    return UHSDR_OK;
}

int uhsdr_multi_destroy(uhsdr_multi_t *m)
{
    if (!m) return UHSDR_ERR_ARG;
    for (auto *e : m->eng) if (e) uhsdr_engine_destroy(e);
    delete m;
    return UHSDR_OK;
}

int uhsdr_multi_create(uhsdr_multi_t **out, int num_channels, const int *devices, int num_devices, const void *tables, size_t tables_bytes)
{
    if (!out || !devices || num_devices < 1 || num_channels < num_devices) return UHSDR_ERR_ARG;
    *out = nullptr;
    uhsdr_multi *m = new uhsdr_multi();
    m->nch = num_channels;
    for (int g = 0; g < num_devices; g++) {
        int f = 0, c = 0;
        uhsdr_channel_range(g, num_devices, num_channels, &f, &c);
        uhsdr_engine_t *e = nullptr;
        const int rc = uhsdr_engine_create(&e, c, devices[g], tables, tables_bytes);
        if (rc != UHSDR_OK) { uhsdr_multi_destroy(m); return rc; }       // text in uhsdr_last_error(NULL)
        m->eng.push_back(e); m->first.push_back(f); m->count.push_back(c); m->device.push_back(devices[g]);
    }
    *out = m;
    return UHSDR_OK;
}

int uhsdr_multi_num_devices(const uhsdr_multi_t *m) { return m ? (int)m->eng.size() : UHSDR_ERR_ARG; }
int uhsdr_multi_num_channels(const uhsdr_multi_t *m) { return m ? m->nch : UHSDR_ERR_ARG; }
const char *uhsdr_multi_last_error(const uhsdr_multi_t *m) { return m ? m->last_error.c_str() : uhsdr_last_error(nullptr); }

uhsdr_engine_t *uhsdr_multi_engine(uhsdr_multi_t *m, int index, int *first, int *count)
{
    if (!m || index < 0 || index >= (int)m->eng.size()) return nullptr;
    if (first) *first = m->first[index];
    if (count) *count = m->count[index];
    return m->eng[index];
}

// global channels first, first + stride, ... (count of them): every engine configures its own members of the progression
int uhsdr_multi_configure_channels_strided(uhsdr_multi_t *m, int first, int count, int stride, const uhsdr_chan_cfg_t *cfg, int reset)
{
    if (!m || !cfg || first < 0 || count <= 0 || stride < 1 || (long long)first + (long long)(count - 1) * stride >= m->nch) return UHSDR_ERR_ARG;
    for (size_t g = 0; g < m->eng.size(); g++) {
        const long long lo = m->first[g], hi = lo + m->count[g];
        long long k0 = lo <= first ? 0 : (lo - first + stride - 1) / stride;            // first member of the progression at or above lo
        long long k1 = (hi - 1 - first) / stride;                                        // last member below hi
        if (hi - 1 < first) continue;
        if (k1 > count - 1) k1 = count - 1;
        if (k0 > k1) continue;
        const int rc = uhsdr_configure_channels_strided(m->eng[g], (int)(first + k0 * stride - lo), (int)(k1 - k0 + 1), stride, cfg, reset);
        if (rc != UHSDR_OK) { m->last_error = std::string("device ") + std::to_string(m->device[g]) + ": " + uhsdr_last_error(m->eng[g]); return rc; }
    }
    return UHSDR_OK;
}

int uhsdr_multi_configure_channels(uhsdr_multi_t *m, int first, int count, const uhsdr_chan_cfg_t *cfg, int reset)
{
    return uhsdr_multi_configure_channels_strided(m, first, count, 1, cfg, reset);
}

// AudioDriver_RxProcessor / TxProcessor_Run for all channels of the box: host buffers [num_channels][nblocks*32], channel-major;
// every device works on its contiguous slab of rows from its own host thread (the single-device call overlaps its copies itself).
static int multi_run(uhsdr_multi_t *m, bool tx, const void *in, void *outp, int nblocks, const uint8_t *mute)
{
    if (!m || !in || !outp || nblocks <= 0) return UHSDR_ERR_ARG;
    const size_t row = (size_t)nblocks * UHSDR_BLOCK_SIZE;
    std::vector<int> rcs(m->eng.size(), UHSDR_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < m->eng.size(); g++) {
        th.emplace_back([&, g]() {
            const size_t off = (size_t)m->first[g] * row;
            const uint8_t *mu = mute ? mute + (size_t)m->first[g] * (size_t)nblocks : nullptr;
            if (tx) rcs[g] = uhsdr_tx_process(m->eng[g], (const uhsdr_audio_sample_t *)in + off, (uhsdr_iq_sample_t *)outp + off, nblocks, mu);
            else rcs[g] = uhsdr_rx_process(m->eng[g], (const uhsdr_iq_sample_t *)in + off, (uhsdr_audio_sample_t *)outp + off, nblocks, mu);
        });
    }
    for (auto &t : th) t.join();
    for (size_t g = 0; g < m->eng.size(); g++)
        if (rcs[g] != UHSDR_OK) { m->last_error = std::string("device ") + std::to_string(m->device[g]) + ": " + uhsdr_last_error(m->eng[g]); return rcs[g]; }
    return UHSDR_OK;
}

int uhsdr_multi_rx_process(uhsdr_multi_t *m, const uhsdr_iq_sample_t *iq, uhsdr_audio_sample_t *audio, int nblocks, const uint8_t *mute)
{
    return multi_run(m, false, iq, audio, nblocks, mute);
}

int uhsdr_multi_tx_process(uhsdr_multi_t *m, const uhsdr_audio_sample_t *audio, uhsdr_iq_sample_t *iq, int nblocks, const uint8_t *mute)
{
    return multi_run(m, true, audio, iq, nblocks, mute);
}

int uhsdr_multi_get_status(uhsdr_multi_t *m, int first, int count, uhsdr_chan_status_t *status)
{
    if (!m || !status || first < 0 || count <= 0 || first + count > m->nch) return UHSDR_ERR_ARG;
    for (size_t g = 0; g < m->eng.size(); g++) {
        const int lo = std::max(first, m->first[g]), hi = std::min(first + count, m->first[g] + m->count[g]);
        if (lo >= hi) continue;
        const int rc = uhsdr_get_status(m->eng[g], lo - m->first[g], hi - lo, status + (lo - first));
        if (rc != UHSDR_OK) { m->last_error = uhsdr_last_error(m->eng[g]); return rc; }
    }
    return UHSDR_OK;
}

}  // extern "C"
