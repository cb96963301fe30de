// engine.cu -- the C ABI of include/uhsdr_b200.h: engine lifetime, channel configuration, and the
// host-buffer / device-buffer entry points that stand in for AudioDriver_I2SCallback ->
// AudioDriver_RxProcessor / TxProcessor_Run (mchf-eclipse/drivers/audio/audio_driver.c:2962,2603;
// tx_processor.c:891).  There is no CPU path: every call needs a usable CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "host_tables.h"
#include "kernels.h"
#include "uhsdr_b200.h"

using namespace uhsdr;

struct uhsdr_engine {
    int device = 0;
    int nch = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream[2] = { nullptr, nullptr };
    cudaStream_t aux_stream = nullptr;          // split paths: serial kernels of slice s run beside the FIR kernel of slice s+1
    cudaStream_t aux2_stream = nullptr;         // serial kernels in two phases: phase 2 of slice s beside phase 1 of slice s+1
    static constexpr int kSplitSlices = 4, kSplitSlicesMax = 8;
    cudaEvent_t ev_split[2 * kSplitSlicesMax] = {}, ev_phase[2 * kSplitSlicesMax] = {}, ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    static constexpr int kMaxSlices = 40;
    cudaEvent_t ev_in[kMaxSlices] = {}, ev_k[kMaxSlices] = {};
    cudaEvent_t ev_done = nullptr;
    HostTables tables;
    float *d_pool = nullptr;
    ChanParams *d_params = nullptr;
    ChanState *d_state = nullptr;
    NrState *d_nr = nullptr;
    float *d_spec = nullptr;
    float *d_spec_avg = nullptr, *d_spec_off = nullptr;      // sd.FFT_AVGData [nch][512], sd.display_offset [nch]
    TxState *d_tx = nullptr;
    TxParams *d_txp = nullptr;
    std::vector<ChanParams> h_params;
    std::vector<int> h_tx_enabled;
    // dispatch lists (rebuilt after configure)
    bool lists_dirty = true;
    std::vector<int> h_list_fused, h_list_generic, h_list_split, h_list_split_nr;
    int *d_list_fused = nullptr, *d_list_generic = nullptr, *d_list_split = nullptr, *d_list_split_nr = nullptr;
    // narrow SSB / CW channels with a deferred consumer -- the spectrum ring (h_list_tc_sp) or the spectral NR (h_list_tc_nr, with or
    // without the ring): tensor-core kernel + spectrum tap kernel (+ NR kernel + serial phase 2)
    std::vector<int> h_list_tc_sp, h_list_tc_nr;
    int *d_list_tc_sp = nullptr, *d_list_tc_nr = nullptr;
    float2 *d_iqc = nullptr;             // [2][nch][16] IQ-correction factors of the last 16 blocks, logged by the tensor-core kernel
    float *d_scratch_nr = nullptr;       // AGC output of the h_list_tc_nr channels, [n][nblocks * 8]
    size_t d_scratch_nr_bytes = 0;
    int use_tcx = 1;
    int split_floats_per_block = 0;      // scratch floats per block and channel of the split path
    float *d_scratch = nullptr;
    size_t d_scratch_bytes = 0;
    int use_split = 1;       // general path as front (FIR) kernel + thread-per-channel serial kernel
    int use_front2 = 1;      // register-blocked front kernel (rx_front2.cu) where every split channel's chain fits it
    bool front2_ok = false;
    int use_pipe2 = 1;       // second-generation serial kernel in two phases on two streams (phase 2 of slice s beside phase 1 of slice s+1)
    int split_slices = 0;    // > 0: forces the number of time slices of the split path (experiments)
    int use_serial2 = 1;     // FMA / shared-memory-AGC serial kernel (rx_serial2.cu, shipping build) where every split channel's chain fits it
    bool serial2_ok = false;
    long long tw_blocks_left = 0;    // > 0: some channel's twin-peaks detector may still be active (1050 blocks after a reset / re-arm)
    FusedCoefs fused_coefs;
    bool fused_coefs_valid = false;
    int fused_s1_ci = -1, fused_s2_ci = -1, fused_s2_cq = -1;
    // staging for the host-buffer entry points
    void *d_in = nullptr, *d_out = nullptr;
    uint8_t *d_mute = nullptr;
    size_t d_in_bytes = 0, d_out_bytes = 0, d_mute_bytes = 0;
    std::string last_error;
    int64_t launches = 0;
    int sm_count = 148;
    int use_fused = 1;
    int use_tc = 1;          // tensor-core Hilbert variant of the fused kernel (shipping build)
};

static std::string g_create_error;

#define CK(e, call)                                                                             \
    do {                                                                                        \
        cudaError_t _err = (call);                                                              \
        if (_err != cudaSuccess) {                                                              \
            (e)->last_error = std::string(#call) + ": " + cudaGetErrorString(_err);             \
            return UHSDR_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

extern "C" {

int uhsdr_b200_abi_version(void) { return 2; }

const char *uhsdr_b200_backend(void)
{
#if UHSDR_EXACT
    return "cuda-sm100a-exact";
#else
    return "cuda-sm100a";
#endif
}

const char *uhsdr_strerror(int code)
{
    switch (code) {
    case UHSDR_OK: return "ok";
    case UHSDR_ERR_ARG: return "invalid argument";
    case UHSDR_ERR_NO_DEVICE: return "no usable CUDA device";
    case UHSDR_ERR_CUDA: return "CUDA runtime error";
    case UHSDR_ERR_TABLES: return "bad coefficient-table blob";
    case UHSDR_ERR_UNSUPPORTED: return "configuration not implemented";
    case UHSDR_ERR_STATE: return "channel not configured";
    default: return "unknown error";
    }
}

const char *uhsdr_last_error(const uhsdr_engine_t *e) { return e ? e->last_error.c_str() : g_create_error.c_str(); }

int uhsdr_default_chan_cfg(uhsdr_chan_cfg_t *cfg)
{
    if (!cfg) return UHSDR_ERR_ARG;
    default_chan_cfg(cfg);
    return UHSDR_OK;
}

int uhsdr_tables_validate(const void *tables, size_t tables_bytes)
{
    HostTables t;
    std::string err;
    if (!t.load(tables, tables_bytes, &err)) { g_create_error = err; return UHSDR_ERR_TABLES; }
    return UHSDR_OK;
}

int uhsdr_engine_destroy(uhsdr_engine_t *e)
{
    if (!e) return UHSDR_ERR_ARG;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_pool); cudaFree(e->d_params); cudaFree(e->d_state); cudaFree(e->d_nr); cudaFree(e->d_spec); cudaFree(e->d_spec_avg); cudaFree(e->d_spec_off);
    cudaFree(e->d_tx); cudaFree(e->d_txp); cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->d_mute);
    cudaFree(e->d_list_fused); cudaFree(e->d_list_generic); cudaFree(e->d_list_split); cudaFree(e->d_list_split_nr); cudaFree(e->d_scratch);
    cudaFree(e->d_list_tc_sp); cudaFree(e->d_list_tc_nr); cudaFree(e->d_iqc); cudaFree(e->d_scratch_nr);
    for (auto &s : e->copy_stream) if (s) cudaStreamDestroy(s);
    if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
    if (e->aux2_stream) cudaStreamDestroy(e->aux2_stream);
    for (auto &v : e->ev_split) if (v) cudaEventDestroy(v);
    for (auto &v : e->ev_phase) if (v) cudaEventDestroy(v);
    if (e->ev_join2) cudaEventDestroy(e->ev_join2);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    for (auto &v : e->ev_in) if (v) cudaEventDestroy(v);
    for (auto &v : e->ev_k) if (v) cudaEventDestroy(v);
    if (e->ev_done) cudaEventDestroy(e->ev_done);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return UHSDR_OK;
}

int uhsdr_engine_create(uhsdr_engine_t **out, int num_channels, int device, const void *tables, size_t tables_bytes)
{
    if (!out || num_channels <= 0) { g_create_error = "bad arguments"; return UHSDR_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        g_create_error = std::string("cudaGetDeviceCount: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "no device / bad index");
        return UHSDR_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        g_create_error = "device is not an sm_100-class (Blackwell) GPU; the library carries sm_100a code only";
        return UHSDR_ERR_NO_DEVICE;
    }
    uhsdr_engine *e = new (std::nothrow) uhsdr_engine();
    if (!e) { g_create_error = "out of memory"; return UHSDR_ERR_ARG; }
    e->device = device; e->nch = num_channels; e->sm_count = prop.multiProcessorCount;
    std::string err;
    if (!e->tables.load(tables, tables_bytes, &err)) { g_create_error = err; delete e; return UHSDR_ERR_TABLES; }
    const char *nf = getenv("UHSDR_B200_NO_FUSED");
    if (nf && nf[0] == '1') e->use_fused = 0;
    const char *ns = getenv("UHSDR_B200_NO_SPLIT");
    if (ns && ns[0] == '1') e->use_split = 0;
    const char *nf2 = getenv("UHSDR_B200_NO_FRONT2");
    if (nf2 && nf2[0] == '1') e->use_front2 = 0;
    const char *ns2 = getenv("UHSDR_B200_NO_SERIAL2");
    if (ns2 && ns2[0] == '1') e->use_serial2 = 0;
    const char *ntx = getenv("UHSDR_B200_NO_TCX");
    if (ntx && ntx[0] == '1') e->use_tcx = 0;
    const char *np2 = getenv("UHSDR_B200_NO_PIPE2");
    if (np2 && np2[0] == '1') e->use_pipe2 = 0;
    const char *nss = getenv("UHSDR_B200_SPLIT_SLICES");
    if (nss && atoi(nss) > 0) e->split_slices = std::min(atoi(nss), (int)uhsdr_engine::kSplitSlicesMax);
    const char *nt = getenv("UHSDR_B200_NO_TC");
    if ((nt && nt[0] == '1') || !rx_ssb_tc_available()) e->use_tc = 0;
    auto fail = [&](const char *what, cudaError_t er) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(er);
        uhsdr_engine_destroy(e);
        return UHSDR_ERR_CUDA;
    };
    cudaError_t er;
    if ((er = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", er);
    if ((er = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", er);
    for (auto &s : e->copy_stream) if ((er = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", er);
    for (auto &v : e->ev_in) if ((er = cudaEventCreateWithFlags(&v, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    for (auto &v : e->ev_k) if ((er = cudaEventCreateWithFlags(&v, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    if ((er = cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    {
        // the serial kernels are small and latency-bound: highest priority, so that their few CTAs are placed as soon
        // as the FIR kernel of the next slice frees a slot
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if ((er = cudaStreamCreateWithPriority(&e->aux_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail("cudaStreamCreate", er);
        if ((er = cudaStreamCreateWithPriority(&e->aux2_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail("cudaStreamCreate", er);
    }
    for (auto &v : e->ev_split) if ((er = cudaEventCreateWithFlags(&v, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    for (auto &v : e->ev_phase) if ((er = cudaEventCreateWithFlags(&v, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    if ((er = cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    if ((er = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    if ((er = cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", er);
    const size_t n = (size_t)num_channels;
    if ((er = cudaMalloc(&e->d_pool, e->tables.pool.size() * sizeof(float))) != cudaSuccess) return fail("cudaMalloc pool", er);
    if ((er = cudaMemcpy(e->d_pool, e->tables.pool.data(), e->tables.pool.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy pool", er);
    if ((er = cudaMalloc(&e->d_params, n * sizeof(ChanParams))) != cudaSuccess) return fail("cudaMalloc params", er);
    if ((er = cudaMemset(e->d_params, 0, n * sizeof(ChanParams))) != cudaSuccess) return fail("cudaMemset params", er);
    if ((er = cudaMalloc(&e->d_state, n * sizeof(ChanState))) != cudaSuccess) return fail("cudaMalloc state", er);
    if ((er = cudaMemset(e->d_state, 0, n * sizeof(ChanState))) != cudaSuccess) return fail("cudaMemset state", er);
    if ((er = cudaMalloc(&e->d_list_fused, n * sizeof(int))) != cudaSuccess) return fail("cudaMalloc list", er);
    if ((er = cudaMalloc(&e->d_list_tc_sp, n * sizeof(int))) != cudaSuccess) return fail("cudaMalloc list", er);
    if ((er = cudaMalloc(&e->d_list_tc_nr, n * sizeof(int))) != cudaSuccess) return fail("cudaMalloc list", er);
    if ((er = cudaMalloc(&e->d_iqc, 2 * n * 16 * sizeof(float2))) != cudaSuccess) return fail("cudaMalloc iqc", er);
    if ((er = cudaMalloc(&e->d_list_generic, n * sizeof(int))) != cudaSuccess) return fail("cudaMalloc list", er);
    if ((er = cudaMalloc(&e->d_list_split, n * sizeof(int))) != cudaSuccess) return fail("cudaMalloc list", er);
    if ((er = cudaMalloc(&e->d_list_split_nr, n * sizeof(int))) != cudaSuccess) return fail("cudaMalloc list", er);
    e->h_params.assign(n, ChanParams{});
    e->h_tx_enabled.assign(n, 0);
    *out = e;
    return UHSDR_OK;
}

int uhsdr_engine_num_channels(const uhsdr_engine_t *e) { return e ? e->nch : UHSDR_ERR_ARG; }
void *uhsdr_engine_stream(uhsdr_engine_t *e) { return e ? (void *)e->stream : nullptr; }
int64_t uhsdr_engine_launch_count(const uhsdr_engine_t *e) { return e ? e->launches : 0; }

int uhsdr_engine_sync(uhsdr_engine_t *e)
{
    if (!e) return UHSDR_ERR_ARG;
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    return UHSDR_OK;
}

int uhsdr_configure_channels(uhsdr_engine_t *e, int first, int count, const uhsdr_chan_cfg_t *cfg, int reset)
{
    return uhsdr_configure_channels_strided(e, first, count, 1, cfg, reset);
}

int uhsdr_configure_channels_strided(uhsdr_engine_t *e, int first, int count, int stride, const uhsdr_chan_cfg_t *cfg, int reset)
{
    if (!e || !cfg || first < 0 || count <= 0 || stride < 1 || (long long)first + (long long)(count - 1) * stride >= e->nch) {
        if (e) e->last_error = "configure: bad channel range or NULL cfg";
        return UHSDR_ERR_ARG;
    }
    ChanParams p;
    std::string err;
    int rc = build_chan_params(e->tables, *cfg, &p, &err);
    if (rc != UHSDR_OK) { e->last_error = err; return rc; }
    TxParams tp;
    rc = build_tx_params(e->tables, *cfg, &tp, &err);
    if (rc != UHSDR_OK) { e->last_error = err; return rc; }
    CK(e, cudaSetDevice(e->device));
    const size_t n = (size_t)e->nch;
    if (p.nr_enable && !e->d_nr) {
        CK(e, cudaMalloc(&e->d_nr, n * sizeof(NrState)));
        CK(e, cudaMemsetAsync(e->d_nr, 0, n * sizeof(NrState), e->stream));
        CK(e, launch_nr_boot(e->d_nr, e->nch, e->stream));
        e->launches++;
    }
    if (p.spectrum_enable && !e->d_spec) {
        CK(e, cudaMalloc(&e->d_spec, n * 1024 * sizeof(float)));
        CK(e, cudaMemsetAsync(e->d_spec, 0, n * 1024 * sizeof(float), e->stream));
    }
    if (!e->d_tx) {
        CK(e, cudaMalloc(&e->d_tx, n * sizeof(TxState)));
        CK(e, cudaMemsetAsync(e->d_tx, 0, n * sizeof(TxState), e->stream));
        CK(e, launch_tx_boot(e->d_tx, e->nch, e->stream));
        e->launches++;
        CK(e, cudaMalloc(&e->d_txp, n * sizeof(TxParams)));
        CK(e, cudaMemsetAsync(e->d_txp, 0, n * sizeof(TxParams), e->stream));
    }
    CK(e, launch_configure(e->d_params, e->d_state, e->d_nr, e->d_spec, e->d_tx, e->d_txp, p, tp, first, count, stride, reset, e->stream));
    e->launches++;
    for (int i = 0; i < count; i++) { const int c = first + i * stride; e->h_params[c] = p; e->h_tx_enabled[c] = tp.enabled; }
    e->lists_dirty = true;
    if (reset) e->tw_blocks_left = 1056;      // 1001 blocks of settling + 50 of sampling (audio_driver.c:2194-2225), rounded up
    return UHSDR_OK;
}

int uhsdr_configure_channel(uhsdr_engine_t *e, int channel, const uhsdr_chan_cfg_t *cfg, int reset)
{
    return uhsdr_configure_channels(e, channel, 1, cfg, reset);
}

// Channels whose chain is the narrow-SSB topology with one shared coefficient set run on the
// fused kernel (rx_ssb_fused.cu); everything else runs on the general kernel.
static int rebuild_lists(uhsdr_engine *e)
{
    e->h_list_fused.clear(); e->h_list_generic.clear(); e->h_list_split.clear(); e->h_list_split_nr.clear();
    e->h_list_tc_sp.clear(); e->h_list_tc_nr.clear();
    e->fused_s1_ci = -1;
    e->split_floats_per_block = 0;
    e->front2_ok = e->use_front2 != 0;
    e->serial2_ok = e->use_serial2 != 0;
    for (int c = 0; c < e->nch; c++) {
        const ChanParams &p = e->h_params[c];
        if (!p.configured) { e->last_error = "rx/tx: channel " + std::to_string(c) + " is not configured"; return UHSDR_ERR_STATE; }
        bool fused = e->use_fused && fused_eligible(p);
        bool tcx = !fused && e->use_fused && e->use_tc && e->use_tcx && e->use_serial2 && (p.nr_enable || p.spectrum_enable) && fused_eligible_ext(p);
        if (fused || tcx) {
            if (e->fused_s1_ci < 0) { e->fused_s1_ci = p.s1_ci; e->fused_s2_ci = p.s2_ci; e->fused_s2_cq = p.s2_cq; }
            else if (p.s1_ci != e->fused_s1_ci || p.s2_ci != e->fused_s2_ci || p.s2_cq != e->fused_s2_cq) fused = tcx = false;
        }
        if (fused) e->h_list_fused.push_back(c);
        else if (tcx) (p.nr_enable ? e->h_list_tc_nr : e->h_list_tc_sp).push_back(c);
        else if (e->use_split && rx_split_floats_per_block(p) > 0) {
            // channels with the spectral noise reduction take two serial phases around the warp-cooperative NR kernel
            (p.nr_enable ? e->h_list_split_nr : e->h_list_split).push_back(c);
            if (!rx_front2_eligible(p)) e->front2_ok = false;
            if (!rx_serial2_eligible(p)) e->serial2_ok = false;
            e->split_floats_per_block = std::max(e->split_floats_per_block, rx_split_floats_per_block(p));
        } else {
            if (p.notch_enable) { e->last_error = "rx: the LMS auto-notch runs on the split path only (UHSDR_B200_NO_SPLIT is set, or the chain is not split-eligible)"; return UHSDR_ERR_UNSUPPORTED; }
            e->h_list_generic.push_back(c);
        }
    }
    if (!e->h_list_split.empty())
        CK(e, cudaMemcpyAsync(e->d_list_split, e->h_list_split.data(), e->h_list_split.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (!e->h_list_split_nr.empty())
        CK(e, cudaMemcpyAsync(e->d_list_split_nr, e->h_list_split_nr.data(), e->h_list_split_nr.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (!e->h_list_fused.empty())
        CK(e, cudaMemcpyAsync(e->d_list_fused, e->h_list_fused.data(), e->h_list_fused.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (!e->h_list_tc_sp.empty())
        CK(e, cudaMemcpyAsync(e->d_list_tc_sp, e->h_list_tc_sp.data(), e->h_list_tc_sp.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (!e->h_list_tc_nr.empty())
        CK(e, cudaMemcpyAsync(e->d_list_tc_nr, e->h_list_tc_nr.data(), e->h_list_tc_nr.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (e->fused_s1_ci >= 0) {
        const float *pool = e->tables.pool.data();
        fill_fused_coefs(&e->fused_coefs, pool + e->fused_s1_ci, pool + e->fused_s2_ci, pool + e->fused_s2_cq);
    }
    if (!e->h_list_generic.empty())
        CK(e, cudaMemcpyAsync(e->d_list_generic, e->h_list_generic.data(), e->h_list_generic.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));   // the host vectors may be rebuilt before the copies run
    e->lists_dirty = false;
    return UHSDR_OK;
}

// Launches the receiver kernels for `nblocks` blocks of every channel; rows of the device buffers
// are `chan_stride` samples apart (a time slice of a longer buffer has chan_stride > nblocks*32).
static int rx_launch(uhsdr_engine *e, const uhsdr_iq_sample_t *iq_dev, uhsdr_audio_sample_t *audio_dev, float *audio_f_dev,
                     int nblocks, const uint8_t *mute_dev, long long chan_stride, long long mute_stride, cudaStream_t stream)
{
    RxArgs a;
    a.params = e->d_params; a.state = e->d_state; a.nr = e->d_nr; a.spec_ring = e->d_spec; a.pool = e->d_pool;
    a.iq = iq_dev; a.audio = audio_dev; a.audio_f = audio_f_dev; a.mute = mute_dev; a.nblocks = nblocks;
    a.chan_stride = chan_stride; a.mute_stride = mute_stride; a.scratch = nullptr; a.scratch_stride = 0;
    a.iqc_log = nullptr; a.nr_handoff = 0;
    if (e->tw_blocks_left > 0) {
        CK(e, launch_twinpeaks(e->d_params, e->d_state, iq_dev, e->nch, nblocks, chan_stride, stream));
        e->launches++;
        e->tw_blocks_left -= nblocks;
    }
    if (!e->h_list_fused.empty()) {
        a.chan_list = e->d_list_fused; a.num_items = (int)e->h_list_fused.size();
        // the fused kernel advances in chunks of 4 blocks; other call sizes take the general kernel
        // the tensor-core kernel stores 32-byte vectors; rows are nblocks*256 bytes apart, so only the base matters
        // (and the float copy as 16-byte vectors; the CUDA-core fused kernel as 8-byte ones; the general kernel word by word)
        const bool tc_ok = e->use_tc && ((uintptr_t)audio_dev % 32 == 0) && ((uintptr_t)iq_dev % 32 == 0) && (chan_stride % 4 == 0) &&
                           ((uintptr_t)audio_f_dev % 16 == 0);
        const bool fused_ok = ((uintptr_t)audio_f_dev % 8 == 0) && (chan_stride % 2 == 0) && ((uintptr_t)iq_dev % 16 == 0) &&
                              ((uintptr_t)audio_dev % 16 == 0);
        if (nblocks % 4 == 0 && tc_ok) CK(e, launch_rx_ssb_tc(a, e->fused_s1_ci, e->fused_s2_ci, e->fused_s2_cq, e->sm_count, stream));
        else if (nblocks % 4 == 0 && fused_ok) CK(e, launch_rx_ssb_fused(a, e->fused_coefs, e->sm_count, stream));
        else CK(e, launch_rx_generic(a, stream));
        e->launches++;
    }
    // Narrow SSB / CW with the spectrum ring and / or the spectral NR: the tensor-core kernel up to the output stage (or, with the NR,
    // up to the AGC output), the ring from a tap kernel over the call's last 16 blocks, then the NR frames and the serial kernel's
    // phase 2.  Call sizes / alignments the tensor-core kernel does not take go through the one-kernel general path.
    for (int which = 0; which < 2; which++) {
        const std::vector<int> &lst = which ? e->h_list_tc_nr : e->h_list_tc_sp;
        if (lst.empty()) continue;
        RxArgs x = a;
        x.chan_list = which ? e->d_list_tc_nr : e->d_list_tc_sp; x.num_items = (int)lst.size();
        const bool ok = e->use_tc && nblocks % 4 == 0 && ((uintptr_t)audio_dev % 32 == 0) && ((uintptr_t)iq_dev % 32 == 0) && (chan_stride % 4 == 0) &&
                        ((uintptr_t)audio_f_dev % 16 == 0);
        if (!ok) { CK(e, launch_rx_generic(x, stream)); e->launches++; continue; }
        x.iqc_log = e->d_iqc + (size_t)which * (size_t)e->nch * 16;
        if (which) {
            const size_t need = lst.size() * (size_t)nblocks * 8 * sizeof(float);
            if (need > e->d_scratch_nr_bytes) {
                CK(e, cudaStreamSynchronize(stream));
                cudaFree(e->d_scratch_nr); e->d_scratch_nr = nullptr; e->d_scratch_nr_bytes = 0;
                CK(e, cudaMalloc(&e->d_scratch_nr, need));
                e->d_scratch_nr_bytes = need;
            }
            x.nr_handoff = 1; x.scratch = e->d_scratch_nr; x.scratch_stride = (long long)nblocks * 8;
        }
        CK(e, launch_rx_ssb_tc(x, e->fused_s1_ci, e->fused_s2_ci, e->fused_s2_cq, e->sm_count, stream));
        CK(e, launch_spectrum_tap(x, stream));
        e->launches += 2;
        if (which) {
            CK(e, launch_rx_nr(x, stream));
            CK(e, launch_rx_serial2(x, 2, stream));
            e->launches += 2;
        }
    }
    if (!e->h_list_generic.empty()) {
        a.chan_list = e->d_list_generic; a.num_items = (int)e->h_list_generic.size();
        CK(e, launch_rx_generic(a, stream));
        e->launches++;
    }
    // Split general path.  The call is cut into time slices: the FIR front kernel of slice s+1 (this stream) runs
    // beside the sample-serial kernels of slice s (aux stream); neither fills the GPU on its own.
    // (the serial kernels store 16-byte words and 8-byte float pairs: buffers aligned to less take the one-kernel general path)
    const bool split_ok = ((uintptr_t)audio_dev % 16 == 0) && ((uintptr_t)audio_f_dev % 8 == 0) && (chan_stride % 2 == 0);
    if (!split_ok) {
        for (int with_nr = 0; with_nr < 2; with_nr++) {
            const std::vector<int> &lst = with_nr ? e->h_list_split_nr : e->h_list_split;
            if (lst.empty()) continue;
            a.chan_list = with_nr ? e->d_list_split_nr : e->d_list_split; a.num_items = (int)lst.size();
            CK(e, launch_rx_generic(a, stream));
            e->launches++;
        }
    } else if (!e->h_list_split.empty() || !e->h_list_split_nr.empty()) {
        // the second-generation serial kernel stores 32-byte / 16-byte vectors
        const bool s2 = e->serial2_ok && ((uintptr_t)audio_dev % 32 == 0) && (chan_stride % 4 == 0) && ((uintptr_t)audio_f_dev % 16 == 0);
        // Its chain is cut once more at the AGC output (the hand-off point of the spectral NR): phase 1 (demodulator, lattice, AGC) of
        // slice s+1 runs beside phase 2 (NR frames, biquads, interpolator, output stage) of slice s.  One warp carries 32 channels
        // through a dependent chain, so a serial kernel's time is the chain latency however few SMs it occupies.
        const bool pipe2 = s2 && e->use_pipe2 && nblocks >= 64;
        const int nsl = e->split_slices > 0 ? std::min(e->split_slices, std::max(1, nblocks / 4))
                        : (nblocks >= 64) ? ((pipe2 && nblocks >= 256) ? uhsdr_engine::kSplitSlicesMax : uhsdr_engine::kSplitSlices) : 1;
        const int per = ((nblocks + nsl - 1) / nsl + 3) / 4 * 4;
        const size_t n_all = e->h_list_split.size() + e->h_list_split_nr.size();
        const size_t need = n_all * (size_t)nblocks * (size_t)e->split_floats_per_block * sizeof(float) + 64 * n_all * nsl;
        if (need > e->d_scratch_bytes) {
            CK(e, cudaStreamSynchronize(stream));
            CK(e, cudaStreamSynchronize(e->aux_stream));
            CK(e, cudaStreamSynchronize(e->aux2_stream));
            cudaFree(e->d_scratch); e->d_scratch = nullptr; e->d_scratch_bytes = 0;
            CK(e, cudaMalloc(&e->d_scratch, need));
            e->d_scratch_bytes = need;
        }
        CK(e, cudaEventRecord(e->ev_fork, stream));
        CK(e, cudaStreamWaitEvent(e->aux_stream, e->ev_fork, 0));
        if (pipe2) CK(e, cudaStreamWaitEvent(e->aux2_stream, e->ev_fork, 0));
        float *sc = e->d_scratch;
        int si = 0;
        for (int b0 = 0; b0 < nblocks; b0 += per, si++) {
            const int nb = std::min(per, nblocks - b0);
            RxArgs s = a;
            s.nblocks = nb;
            s.iq = (const char *)iq_dev + (size_t)b0 * BLK * sizeof(uhsdr_iq_sample_t);
            s.audio = (char *)audio_dev + (size_t)b0 * BLK * sizeof(uhsdr_audio_sample_t);
            s.audio_f = audio_f_dev ? audio_f_dev + (size_t)b0 * BLK : nullptr;
            s.mute = mute_dev ? mute_dev + b0 : nullptr;
            s.scratch_stride = (long long)nb * e->split_floats_per_block;
            for (int with_nr = 0; with_nr < 2; with_nr++) {
                const std::vector<int> &lst = with_nr ? e->h_list_split_nr : e->h_list_split;
                if (lst.empty()) continue;
                s.chan_list = with_nr ? e->d_list_split_nr : e->d_list_split; s.num_items = (int)lst.size();
                s.scratch = sc;
                sc += (size_t)s.num_items * (size_t)s.scratch_stride;
                if (e->front2_ok) CK(e, launch_rx_front2(s, stream));
                else CK(e, launch_rx_front(s, stream));
                cudaEvent_t ev = e->ev_split[2 * si + with_nr];
                CK(e, cudaEventRecord(ev, stream));
                CK(e, cudaStreamWaitEvent(e->aux_stream, ev, 0));
                auto serial = [&](int phase, cudaStream_t st) { return s2 ? launch_rx_serial2(s, phase, st) : launch_rx_serial(s, phase, st); };
                if (pipe2) {
                    CK(e, serial(1, e->aux_stream));
                    cudaEvent_t evp = e->ev_phase[2 * si + with_nr];
                    CK(e, cudaEventRecord(evp, e->aux_stream));
                    CK(e, cudaStreamWaitEvent(e->aux2_stream, evp, 0));
                    if (with_nr) CK(e, launch_rx_nr(s, e->aux2_stream));
                    CK(e, serial(2, e->aux2_stream));
                    e->launches += with_nr ? 4 : 3;
                } else if (!with_nr) {
                    CK(e, serial(0, e->aux_stream));
                    e->launches += 2;
                } else {
                    CK(e, serial(1, e->aux_stream));
                    CK(e, launch_rx_nr(s, e->aux_stream));
                    CK(e, serial(2, e->aux_stream));
                    e->launches += 4;
                }
            }
        }
        if (pipe2) {
            CK(e, cudaEventRecord(e->ev_join2, e->aux2_stream));
            CK(e, cudaStreamWaitEvent(stream, e->ev_join2, 0));
        }
        CK(e, cudaEventRecord(e->ev_join, e->aux_stream));
        CK(e, cudaStreamWaitEvent(stream, e->ev_join, 0));
    }
    return UHSDR_OK;
}

int uhsdr_rx_process_device(uhsdr_engine_t *e, const uhsdr_iq_sample_t *iq_dev, uhsdr_audio_sample_t *audio_dev,
                            float *audio_f_dev, int nblocks, const uint8_t *mute_dev)
{
    if (!e || !iq_dev || !audio_dev || nblocks <= 0) { if (e) e->last_error = "rx_process: NULL buffer or nblocks <= 0"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    if (e->lists_dirty) { int rc = rebuild_lists(e); if (rc != UHSDR_OK) return rc; }
    return rx_launch(e, iq_dev, audio_dev, audio_f_dev, nblocks, mute_dev, (long long)nblocks * BLK, nblocks, e->stream);
}

static int ensure_staging(uhsdr_engine *e, size_t in_bytes, size_t out_bytes, size_t mute_bytes)
{
    if (in_bytes > e->d_in_bytes) { cudaFree(e->d_in); e->d_in = nullptr; e->d_in_bytes = 0; CK(e, cudaMalloc(&e->d_in, in_bytes)); e->d_in_bytes = in_bytes; }
    if (out_bytes > e->d_out_bytes) { cudaFree(e->d_out); e->d_out = nullptr; e->d_out_bytes = 0; CK(e, cudaMalloc(&e->d_out, out_bytes)); e->d_out_bytes = out_bytes; }
    if (mute_bytes > e->d_mute_bytes) { cudaFree(e->d_mute); e->d_mute = nullptr; e->d_mute_bytes = 0; CK(e, cudaMalloc(&e->d_mute, mute_bytes)); e->d_mute_bytes = mute_bytes; }
    return UHSDR_OK;
}

// Host-buffer entry point.  The call is cut into time slices (all channels, a range of blocks); the
// H2D copy of slice s+1, the kernels of slice s and the D2H copy of slice s-1 run concurrently on
// three streams, so with pinned host buffers the call is bound by one direction of the PCIe link.
int uhsdr_rx_process(uhsdr_engine_t *e, const uhsdr_iq_sample_t *iq, uhsdr_audio_sample_t *audio, int nblocks, const uint8_t *mute)
{
    if (!e || !iq || !audio || nblocks <= 0) { if (e) e->last_error = "rx_process: NULL buffer or nblocks <= 0"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    if (e->lists_dirty) { int rc = rebuild_lists(e); if (rc != UHSDR_OK) return rc; }
    const size_t row = (size_t)nblocks * BLK * sizeof(uhsdr_iq_sample_t);
    const size_t bytes = (size_t)e->nch * row;
    const size_t mbytes = mute ? (size_t)e->nch * (size_t)nblocks : 0;
    int rc = ensure_staging(e, bytes, bytes, mbytes);
    if (rc != UHSDR_OK) return rc;
    cudaStream_t s_in = e->copy_stream[0], s_out = e->copy_stream[1];
    if (mute) CK(e, cudaMemcpyAsync(e->d_mute, mute, mbytes, cudaMemcpyHostToDevice, e->stream));
    // slices: multiples of 4 blocks (the fused kernel's chunk) unless the call is small
    int nsl = 1;
    if (bytes >= ((size_t)32 << 20) && nblocks >= 64) {
        // the first H2D and the last D2H are not overlapped: more slices = less of that, until a slice is too short to fill the link
        // (measured at 4096 channels x 1500 blocks, 1.57 GB each way: 8 slices 5.25e9, 12: 5.4e9, 20: 5.8e9, 24: 5.8e9 channel-samples/s)
        const char *ov = getenv("UHSDR_B200_SLICES");
        nsl = ov ? atoi(ov) : (int)std::min<size_t>(24, bytes / ((size_t)48 << 20));
        nsl = std::max(1, std::min<int>(uhsdr_engine::kMaxSlices - 1, std::min(nsl, nblocks / 16)));
    }
    int per = ((nblocks + nsl - 1) / nsl + 3) / 4 * 4;
    if (per <= 0) per = nblocks;
    // everything queued earlier on the compute stream (configure kernels, a previous device-side
    // call) must be finished before the copy streams touch the staging buffers
    CK(e, cudaEventRecord(e->ev_done, e->stream));
    CK(e, cudaStreamWaitEvent(s_in, e->ev_done, 0));
    CK(e, cudaStreamWaitEvent(s_out, e->ev_done, 0));
    int si = 0;
    for (int b0 = 0; b0 < nblocks; b0 += per, si++) {
        const int nb = std::min(per, nblocks - b0);
        const size_t off = (size_t)b0 * BLK * sizeof(uhsdr_iq_sample_t);
        const size_t width = (size_t)nb * BLK * sizeof(uhsdr_iq_sample_t);
        CK(e, cudaMemcpy2DAsync((char *)e->d_in + off, row, (const char *)iq + off, row, width, (size_t)e->nch, cudaMemcpyHostToDevice, s_in));
        CK(e, cudaEventRecord(e->ev_in[si], s_in));
        CK(e, cudaStreamWaitEvent(e->stream, e->ev_in[si], 0));
        rc = rx_launch(e, (const uhsdr_iq_sample_t *)((char *)e->d_in + off), (uhsdr_audio_sample_t *)((char *)e->d_out + off), nullptr, nb,
                       mute ? e->d_mute + b0 : nullptr, (long long)nblocks * BLK, nblocks, e->stream);
        if (rc != UHSDR_OK) return rc;
        CK(e, cudaEventRecord(e->ev_k[si], e->stream));
        CK(e, cudaStreamWaitEvent(s_out, e->ev_k[si], 0));
        CK(e, cudaMemcpy2DAsync((char *)audio + off, row, (char *)e->d_out + off, row, width, (size_t)e->nch, cudaMemcpyDeviceToHost, s_out));
    }
    // rejoin: the engine stream is the one callers time and synchronise on
    CK(e, cudaEventRecord(e->ev_done, s_out));
    CK(e, cudaStreamWaitEvent(e->stream, e->ev_done, 0));
    CK(e, cudaStreamSynchronize(e->stream));
    return UHSDR_OK;
}

int uhsdr_tx_process_device(uhsdr_engine_t *e, const uhsdr_audio_sample_t *audio_dev, uhsdr_iq_sample_t *iq_dev,
                            float *iq_f_dev, int nblocks, const uint8_t *mute_dev)
{
    if (!e || !audio_dev || !iq_dev || nblocks <= 0) { if (e) e->last_error = "tx_process: NULL buffer or nblocks <= 0"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    for (int c = 0; c < e->nch; c++)
    {
        if (!e->h_params[c].configured) { e->last_error = "tx: channel " + std::to_string(c) + " is not configured"; return UHSDR_ERR_STATE; }
        if (!e->h_tx_enabled[c]) {
            e->last_error = "tx: channel " + std::to_string(c) + " has no modulator: the SSB (USB/LSB) voice modulator and the AM / FM ones (with a frequency-translate mode) are implemented";
            return UHSDR_ERR_UNSUPPORTED;
        }
    }
    TxArgs a;
    a.params = e->d_params; a.state = e->d_state; a.tx = e->d_tx; a.txp = e->d_txp; a.pool = e->d_pool;
    a.audio = audio_dev; a.iq = iq_dev; a.iq_f = iq_f_dev; a.mute = mute_dev; a.nblocks = nblocks; a.num_items = e->nch;
    a.scratch = nullptr; a.chan_stride = (long long)nblocks * BLK; a.mute_stride = nblocks;
    if (e->use_split && (uintptr_t)audio_dev % 16 == 0) {
        // split modulator, time-sliced: serial kernel of slice s+1 (aux stream) beside the FIR kernel of slice s
        const size_t need = (size_t)e->nch * (size_t)nblocks * BLK * sizeof(float);
        if (need > e->d_scratch_bytes) {
            CK(e, cudaStreamSynchronize(e->stream));
            CK(e, cudaStreamSynchronize(e->aux_stream));
            cudaFree(e->d_scratch); e->d_scratch = nullptr; e->d_scratch_bytes = 0;
            CK(e, cudaMalloc(&e->d_scratch, need));
            e->d_scratch_bytes = need;
        }
        const int nsl = (nblocks >= 64) ? uhsdr_engine::kSplitSlices : 1;
        const int per = ((nblocks + nsl - 1) / nsl + 3) / 4 * 4;
        CK(e, cudaEventRecord(e->ev_fork, e->stream));
        CK(e, cudaStreamWaitEvent(e->aux_stream, e->ev_fork, 0));
        float *sc = e->d_scratch;
        int si = 0;
        for (int b0 = 0; b0 < nblocks; b0 += per, si++) {
            TxArgs s = a;
            s.nblocks = std::min(per, nblocks - b0);
            s.audio = (const char *)audio_dev + (size_t)b0 * BLK * sizeof(uhsdr_audio_sample_t);
            s.iq = (char *)iq_dev + (size_t)b0 * BLK * sizeof(uhsdr_iq_sample_t);
            s.iq_f = iq_f_dev ? iq_f_dev + (size_t)b0 * BLK * 2 : nullptr;
            s.mute = mute_dev ? mute_dev + b0 : nullptr;
            s.scratch = sc;
            sc += (size_t)e->nch * (size_t)s.nblocks * BLK;
            CK(e, launch_tx_serial(s, e->aux_stream));
            CK(e, cudaEventRecord(e->ev_split[si], e->aux_stream));
            CK(e, cudaStreamWaitEvent(e->stream, e->ev_split[si], 0));
            CK(e, launch_tx_ssb(s, e->stream));
            e->launches += 2;
        }
        return UHSDR_OK;
    }
    CK(e, launch_tx_ssb(a, e->stream));
    e->launches++;
    return UHSDR_OK;
}

int uhsdr_tx_process(uhsdr_engine_t *e, const uhsdr_audio_sample_t *audio, uhsdr_iq_sample_t *iq, int nblocks, const uint8_t *mute)
{
    if (!e || !iq || !audio || nblocks <= 0) { if (e) e->last_error = "tx_process: NULL buffer or nblocks <= 0"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    const size_t bytes = (size_t)e->nch * (size_t)nblocks * BLK * sizeof(uhsdr_iq_sample_t);
    const size_t mbytes = mute ? (size_t)e->nch * (size_t)nblocks : 0;
    int rc = ensure_staging(e, bytes, bytes, mbytes);
    if (rc != UHSDR_OK) return rc;
    CK(e, cudaMemcpyAsync(e->d_in, audio, bytes, cudaMemcpyHostToDevice, e->stream));
    if (mute) CK(e, cudaMemcpyAsync(e->d_mute, mute, mbytes, cudaMemcpyHostToDevice, e->stream));
    rc = uhsdr_tx_process_device(e, (const uhsdr_audio_sample_t *)e->d_in, (uhsdr_iq_sample_t *)e->d_out, nullptr, nblocks, mute ? e->d_mute : nullptr);
    if (rc != UHSDR_OK) return rc;
    CK(e, cudaMemcpyAsync(iq, e->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return UHSDR_OK;
}

int uhsdr_get_spectrum_device(uhsdr_engine_t *e, int first, int count, float *mags_dev)
{
    if (!e || !mags_dev || first < 0 || count <= 0 || first + count > e->nch) { if (e) e->last_error = "get_spectrum: bad arguments"; return UHSDR_ERR_ARG; }
    if (!e->d_spec) { e->last_error = "get_spectrum: no channel has spectrum_enable set"; return UHSDR_ERR_STATE; }
    CK(e, cudaSetDevice(e->device));
    CK(e, launch_spectrum(e->d_params, e->d_state, e->d_spec, e->d_pool, e->tables.off(e->tables.ex->spectrum_window_array),
                          e->tables.tw512_off, first, count, mags_dev, e->stream));
    e->launches++;
    return UHSDR_OK;
}

int uhsdr_get_spectrum(uhsdr_engine_t *e, int first, int count, float *mags)
{
    if (!e || !mags || count <= 0) { if (e) e->last_error = "get_spectrum: bad arguments"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    const size_t bytes = (size_t)count * UHSDR_SPECTRUM_FFT_LEN * sizeof(float);
    int rc = ensure_staging(e, 0, bytes, 0);
    if (rc != UHSDR_OK) return rc;
    rc = uhsdr_get_spectrum_device(e, first, count, (float *)e->d_out);
    if (rc != UHSDR_OK) return rc;
    CK(e, cudaMemcpyAsync(mags, e->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return UHSDR_OK;
}

// sd.db_scale / sd.agc_rate / the dBm constant from the reference's settings (UiSpectrum_InitSpectrumDisplayData,
// ui_spectrum.c:955-1083; dB-per-division factors :193-200, :225-237)
static bool spec_disp_params(const uhsdr_spectrum_display_cfg_t *c, SpecDisp *d)
{
    static const float kDbScaling[9] = { 0, 63.2456, 42.1637, 31.6228, 21.0819, 15.8114, 52.7046, 26.3523, 17.5682 };
    if (!c || c->scope_width < 1 || c->scope_width > UHSDR_SPECTRUM_FFT_LEN || c->spectrum_filter < 1 || c->spectrum_filter > 255 ||
        c->spectrum_agc_rate < 0 || c->spectrum_agc_rate > 255) return false;
    int idx = c->spectrum_db_scale;
    if (idx < 0 || idx >= 9) idx = 3;                       // DB_DIV_ADJUST_DEFAULT, :1023-1026
    d->db_scale = kDbScaling[idx];
    if (c->scope_width != UHSDR_SPECTRUM_FFT_LEN) d->db_scale *= (float)c->scope_width / (float)UHSDR_SPECTRUM_FFT_LEN;
    d->agc_rate = ((float)c->spectrum_agc_rate) / 25;       // SPECTRUM_AGC_SCALING
    d->filt_factor = 1 / (float)c->spectrum_filter;
    d->cons = (float)(c->dbm_constant - 225 - 3);
    d->scope_w = c->scope_width;
    return true;
}

int uhsdr_default_spectrum_display_cfg(uhsdr_spectrum_display_cfg_t *cfg)
{
    if (!cfg) return UHSDR_ERR_ARG;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->spectrum_db_scale = 3; cfg->spectrum_agc_rate = 25; cfg->spectrum_filter = 4; cfg->dbm_constant = 0;
    cfg->scope_width = 480;                                 // slayout.scope.w of the 480x320 layout (SPECTRUM_WIDTH_MAX)
    return UHSDR_OK;
}

int uhsdr_spectrum_display_device(uhsdr_engine_t *e, int first, int count, const uhsdr_spectrum_display_cfg_t *cfg, float *disp_dev,
                                  uhsdr_spectrum_level_t *levels_dev, float *avg_dev, float *mags_dev)
{
    SpecDisp dc;
    if (!e || !disp_dev || !levels_dev || first < 0 || count <= 0 || first + count > e->nch || !spec_disp_params(cfg, &dc)) {
        if (e) e->last_error = "spectrum_display: bad arguments";
        return UHSDR_ERR_ARG;
    }
    if (!e->d_spec) { e->last_error = "spectrum_display: no channel has spectrum_enable set"; return UHSDR_ERR_STATE; }
    CK(e, cudaSetDevice(e->device));
    if (!e->d_spec_avg) {
        const size_t n = (size_t)e->nch;
        CK(e, cudaMalloc(&e->d_spec_avg, n * 512 * sizeof(float)));
        CK(e, cudaMemsetAsync(e->d_spec_avg, 0, n * 512 * sizeof(float), e->stream));
        CK(e, cudaMalloc(&e->d_spec_off, n * sizeof(float)));
        CK(e, cudaMemsetAsync(e->d_spec_off, 0, n * sizeof(float), e->stream));
    }
    static_assert(sizeof(uhsdr_spectrum_level_t) == 3 * sizeof(float), "levels are three packed floats");
    CK(e, launch_spectrum_display(e->d_params, e->d_state, e->d_spec, e->d_pool, e->tables.off(e->tables.ex->spectrum_window_array), e->tables.tw512_off,
                                  first, count, dc, e->d_spec_avg, e->d_spec_off, mags_dev, avg_dev, disp_dev, reinterpret_cast<float *>(levels_dev), e->stream));
    e->launches++;
    return UHSDR_OK;
}

int uhsdr_spectrum_display(uhsdr_engine_t *e, int first, int count, const uhsdr_spectrum_display_cfg_t *cfg, float *disp,
                           uhsdr_spectrum_level_t *levels, float *avg)
{
    if (!e || !disp || !levels || !cfg || count <= 0 || cfg->scope_width < 1 || cfg->scope_width > UHSDR_SPECTRUM_FFT_LEN) {
        if (e) e->last_error = "spectrum_display: bad arguments";
        return UHSDR_ERR_ARG;
    }
    CK(e, cudaSetDevice(e->device));
    const size_t nd = (size_t)count * (size_t)cfg->scope_width * sizeof(float), nl = (size_t)count * sizeof(uhsdr_spectrum_level_t);
    const size_t na = avg ? (size_t)count * 512 * sizeof(float) : 0;
    const size_t o_l = (nd + 255) / 256 * 256, o_a = o_l + (nl + 255) / 256 * 256;
    int rc = ensure_staging(e, 0, o_a + na, 0);
    if (rc != UHSDR_OK) return rc;
    char *base = (char *)e->d_out;
    rc = uhsdr_spectrum_display_device(e, first, count, cfg, (float *)base, (uhsdr_spectrum_level_t *)(base + o_l), avg ? (float *)(base + o_a) : nullptr, nullptr);
    if (rc != UHSDR_OK) return rc;
    CK(e, cudaMemcpyAsync(disp, base, nd, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaMemcpyAsync(levels, base + o_l, nl, cudaMemcpyDeviceToHost, e->stream));
    if (avg) CK(e, cudaMemcpyAsync(avg, base + o_a, na, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return UHSDR_OK;
}

int uhsdr_get_status(uhsdr_engine_t *e, int first, int count, uhsdr_chan_status_t *status)
{
    if (!e || !status || first < 0 || count <= 0 || first + count > e->nch) { if (e) e->last_error = "get_status: bad arguments"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    std::vector<ChanState> hs((size_t)count);
    CK(e, cudaMemcpyAsync(hs.data(), e->d_state + first, (size_t)count * sizeof(ChanState), cudaMemcpyDeviceToHost, e->stream));
    std::vector<TxState> ht;
    if (e->d_tx) {
        ht.resize((size_t)count);
        CK(e, cudaMemcpyAsync(ht.data(), e->d_tx + first, (size_t)count * sizeof(TxState), cudaMemcpyDeviceToHost, e->stream));
    }
    CK(e, cudaStreamSynchronize(e->stream));
    for (int i = 0; i < count; i++) {
        const ChanState &s = hs[i];
        uhsdr_chan_status_t &o = status[i];
        memset(&o, 0, sizeof(o));
        o.adc_clip = s.adc_clip; o.adc_half_clip = s.adc_half_clip; o.adc_quarter_clip = s.adc_quarter_clip;
        o.agc_action = s.agc_action; o.agc_hang_action = s.agc_hang_action;
        o.fm_squelched = s.fm_squelched; o.fm_sql_avg = s.fm_sql_avg;
        o.sam_carrier_freq_offset = s.carrier_freq_offset;
        o.iq_corr_c1 = s.M_c1; o.iq_corr_c2 = s.M_c2;
        o.blocks_processed = s.blocks;
        o.twinpeaks_state = s.tw_state; o.twinpeaks_restarts = s.tw_restarts;
        if (!ht.empty()) { o.tx_peak_audio = ht[i].peak_audio; o.tx_alc_val = ht[i].alc_val; }
    }
    return UHSDR_OK;
}

int uhsdr_twinpeaks_rearm(uhsdr_engine_t *e, int first, int count)
{
    if (!e || first < 0 || count <= 0 || first + count > e->nch) { if (e) e->last_error = "twinpeaks_rearm: bad arguments"; return UHSDR_ERR_ARG; }
    CK(e, cudaSetDevice(e->device));
    CK(e, launch_twinpeaks_rearm(e->d_state, first, count, e->stream));
    e->launches++;
    e->tw_blocks_left = 1056;
    return UHSDR_OK;
}

}  // extern "C"
