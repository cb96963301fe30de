// tx_ssb.cu -- SSB voice transmit modulator: TxProcessor_Run, SSB branch
// (mchf-eclipse/drivers/audio/tx_processor.c:891-1078):
//   AudioBufferFill (:339-405)  int32 mic x gain x 2^-16, peak_audio
//   FilterAudio (:416-429)      10-stage lattice band-pass + 3-stage biquad (treble/bass shelves)
//   VoiceCompressor (:173-242)  post-filter gain, per-sample ALC, 32-sample look-ahead delay line
//   TxProcessor_SSB (:467-490)  201-tap Hilbert pair (I/Q filters swapped for LSB), FreqShift
//   IqFinalProcessing (:282-330) x power factor x IQ gain x 1.133 x 65536, phase mix, float -> int32
// One warp per channel: the recurrences (lattice, biquads, ALC) run on lane 0, the two 201-tap
// FIRs, translation and output formatting on all 32 lanes.
#include <type_traits>

#include "fir_device.cuh"
#include "host_tables.h"
#include "kernels.h"

namespace uhsdr {

// ---- host: TxProcessor_Set (:72-119), AudioManagement_CalcTxCompLevel (audio_management.c:240-291) ----
static const float kPiTx = 3.14159265358979f;
static void bq_scale_tx(float c[5], float sa, float sb) { c[3] = c[3] / sa; c[4] = c[4] / sa; c[0] = c[0] / sb; c[1] = c[1] / sb; c[2] = c[2] / sb; }
static void shelf(float c[5], bool high, float f0, float S, float gain, float FS)     // audio_driver.c:906-964
{
    float w0 = 2 * kPiTx * f0 / FS;
    float A = exp10f(gain / 40.0);
    float alpha = sinf(w0) / 2 * sqrtf((A + 1 / A) * (1 / S - 1) + 2);
    float cosw0 = cosf(w0);
    float twoAa = 2 * sqrtf(A) * alpha;
    float scaling;
    if (high) {
        c[0] = A * ((A + 1) + (A - 1) * cosw0 + twoAa);
        c[1] = -2 * A * ((A - 1) + (A + 1) * cosw0);
        c[2] = A * ((A + 1) + (A - 1) * cosw0 - twoAa);
        scaling = (A + 1) - (A - 1) * cosw0 + twoAa;
        c[3] = -2 * ((A - 1) - (A + 1) * cosw0);
        c[4] = twoAa - (A + 1) + (A - 1) * cosw0;
    } else {
        c[0] = A * ((A + 1) - (A - 1) * cosw0 + twoAa);
        c[1] = 2 * A * ((A - 1) - (A + 1) * cosw0);
        c[2] = A * ((A + 1) - (A - 1) * cosw0 - twoAa);
        scaling = (A + 1) + (A - 1) * cosw0 + twoAa;
        c[3] = 2 * ((A - 1) + (A + 1) * cosw0);
        c[4] = twoAa - (A + 1) - (A - 1) * cosw0;
    }
    float DCgain = 1.0 * scaling;
    bq_scale_tx(c, scaling, DCgain);
}

int build_tx_params(const HostTables &t, const uhsdr_chan_cfg_t &cfg, TxParams *tp, std::string *err)
{
    memset(tp, 0, sizeof(*tp));
    const int mode = cfg.dmod_mode;
    tp->am = (mode == UHSDR_DEMOD_AM && cfg.iq_freq_mode != UHSDR_FREQ_IQ_CONV_OFF) ? 1 : 0;      // no AM unless in a translate mode, tx_processor.c:999
    tp->fm = (mode == UHSDR_DEMOD_FM && cfg.iq_freq_mode != UHSDR_FREQ_IQ_CONV_OFF) ? 1 : 0;      // same rule for FM, :1010
    tp->enabled = (mode == UHSDR_DEMOD_USB || mode == UHSDR_DEMOD_LSB || tp->am || tp->fm) ? 1 : 0;    // is_ssb(), uhsdr_board.h:809
    tp->alc_gain_scaling = tp->am ? 0.23 : (tp->fm ? 0.95 : 1.00);
    tp->lsb = mode == UHSDR_DEMOD_LSB;
    {
        float gain_calc = (uint8_t)cfg.tx_mic_gain;      // ts.tx_mic_gain_mult (codec.c:321)
        gain_calc /= 2;                                  // MIC_GAIN_RESCALE
        gain_calc *= (0.0000152587890625);               // AUDIO_BIT_SCALE_DOWN
        tp->gain_calc = gain_calc;
    }
    int li = t.ex->tx_lattice_soprano;
    if (cfg.tx_filter == UHSDR_TX_FILTER_BASS) li = t.ex->tx_lattice_bass;
    else if (cfg.tx_filter == UHSDR_TX_FILTER_TENOR) li = t.ex->tx_lattice_tenor;
    if (mode == UHSDR_DEMOD_FM) li = t.ex->tx_lattice_fm;             // IIR_TX_2k7_FM, tx_processor.c:104-107
    if (li < 0 || li >= (int)t.h->num_lattices || t.lat[li].num_stages > 10) { if (err) *err = "TX lattice missing from the table blob (or longer than the 10 stages the modulator kernels keep in registers)"; return UHSDR_ERR_TABLES; }
    tp->lat.n = t.lat[li].num_stages; tp->lat.k_off = t.off(t.lat[li].k_array); tp->lat.v_off = t.off(t.lat[li].v_array);
    shelf(tp->bq[0], true, 1700, 0.9, cfg.tx_treble_gain, 48000);
    shelf(tp->bq[1], false, 300, 0.7, cfg.tx_bass_gain, 48000);
    tp->bq[2][0] = 1.0f;
    // speech-compressor settings
    static const unsigned alc_params[13][2] = {
        { 1, 15 }, { 2, 12 }, { 4, 10 }, { 6, 9 }, { 7, 8 }, { 8, 7 }, { 10, 6 }, { 12, 5 }, { 15, 4 }, { 17, 3 }, { 20, 2 }, { 25, 1 }, { 25, 0 } };
    unsigned pg, dv;
    const int lvl = (int16_t)cfg.tx_comp_level;
    if (-1 < lvl && lvl < 13) { pg = alc_params[lvl][0]; dv = alc_params[lvl][1]; }
    else if (lvl == 13) { pg = (unsigned)cfg.tx_alc_postfilt_gain; dv = (unsigned)cfg.tx_alc_decay; }
    else { pg = 4; dv = 10; }
    tp->comp_enabled = lvl > -1;
    tp->postfilt_gain = ((float)(float)pg) / 2.0 + 0.5;
    tp->alc_decay = exp10f(-((((float)dv) + 35.0) / 10.0));
    tp->hil_ntaps = t.ex->tx_hilbert_numtaps;
    tp->hil_ci = t.off(t.ex->tx_hilbert_i_array); tp->hil_cq = t.off(t.ex->tx_hilbert_q_array);
    if (tp->hil_ntaps < 1 || tp->hil_ntaps - 1 > H2 || tp->hil_ci < 0 || tp->hil_cq < 0) { if (err) *err = "TX Hilbert pair missing from the table blob"; return UHSDR_ERR_TABLES; }
    int shift = 0;
    switch (cfg.iq_freq_mode) {
    case UHSDR_FREQ_IQ_CONV_P6KHZ: shift = 6000; break;
    case UHSDR_FREQ_IQ_CONV_M6KHZ: shift = -6000; break;
    case UHSDR_FREQ_IQ_CONV_P12KHZ: shift = 12000; break;
    case UHSDR_FREQ_IQ_CONV_M12KHZ: shift = -12000; break;
    default: break;
    }
    tp->shift_freq = abs(shift); tp->shift_down = shift > 0;
    tp->shift_kind = shift == 0 ? 0 : (tp->shift_freq == 12000 ? 1 : 2);
    {
        float nco_freq = (float)tp->shift_freq, sample_rate = 48000.0f;
        double rate = (2 * M_PI * nco_freq) / sample_rate;
        tp->osc_cos = cos(rate); tp->osc_sin = sin(rate);
    }
    {
        float scaling = tp->fm ? 0.875 : 1.133;           // FM_MOD_AMPLITUDE_SCALING / SSB_GAIN_COMP == AM_GAIN_COMP
        scaling *= (1 << 16);                             // IQ_BIT_SCALE_UP
        tp->final_gain_i = cfg.tx_power_factor * cfg.tx_adj_gain_i * scaling;
        tp->final_gain_q = cfg.tx_power_factor * cfg.tx_adj_gain_q * scaling;
    }
    tp->phase_bal = cfg.iq_phase_balance_tx;
    if (tp->fm) {
        tp->fm_word = (65536 * abs(shift)) / 48000;
        tp->fm_swap = shift < 0;
        tp->fm_mult = cfg.fm_dev_5khz ? 2 : 1;
        tp->dds_off = t.off(t.ex->dds_table_array);
        // softdds_setFreqDDS (softdds.c:26-45): step = ((uint64)(freq * 1024) << 22) / 48000
        static const unsigned burst_freq[3] = { 0, 1750, 2135 };          // fm_tone_burst_freq, audio_management.c:328
        const int bm = cfg.fm_tone_burst_mode;
        const float fburst = (uint16_t)((bm >= 0 && bm <= 2) ? burst_freq[bm] : 0);
        const float fsub = (cfg.fm_subaudible_tone_gen_freq > 0.0f && !bm) ? cfg.fm_subaudible_tone_gen_freq : 0.0f;   // no sub-audible tone during a burst
        uint64_t f64 = fsub * 1024; f64 <<= 22; tp->fm_sub_step = (uint32_t)(f64 / 48000);
        f64 = fburst * 1024; f64 <<= 22; tp->fm_burst_step = bm ? (uint32_t)(f64 / 48000) : 0u;
        tp->fm_sub_scale = 0.00045 * tp->fm_mult;
        tp->fm_burst_scale = (16 / 4266.0) * tp->fm_mult;
        if (tp->dds_off < 0) { if (err) *err = "DDS sine table missing from the table blob"; return UHSDR_ERR_TABLES; }
    }
    return UHSDR_OK;
}

// ---- device ---------------------------------------------------------------------------------
// One sample of TxProcessor_FM (tx_processor.c:544-585): pre-emphasis differentiator hpf_b = 0.05 (hpf_b + a - hpf_a) (double
// product), then the 16-bit NCO accumulator `acc += word + a1 * FM_MOD_SCALING * mult; acc %= 65536` -- the += goes through
// float, and a negative sum wraps the way the x86-64 build of the reference does it (conversion to 64 bits, low word).
// Returns the index into the 1024-entry sine table.
__device__ __forceinline__ uint32_t fm_step(float a, float &hpf_a, float &hpf_b, uint32_t &accum, uint32_t &sub_acc, uint32_t &burst_acc,
                                            const TxParams &tp, const float *__restrict__ pool)
{
    hpf_b = (float)(0.05 * (double)__fsub_rn(__fadd_rn(hpf_b, a), hpf_a));
    hpf_a = a;
    float a1 = hpf_b;
    // sub-audible tone / tone burst: softdds_addSingleTone (softdds.c:113-119) on the pre-emphasised audio, :554-564
    if (tp.fm_sub_step) { const uint32_t k = (sub_acc >> 22) & 1023u; sub_acc += tp.fm_sub_step; a1 = __fadd_rn(a1, __fmul_rn(__ldg(pool + tp.dds_off + k), tp.fm_sub_scale)); }
    if (tp.fm_burst_step) { const uint32_t k = (burst_acc >> 22) & 1023u; burst_acc += tp.fm_burst_step; a1 = __fadd_rn(a1, __fmul_rn(__ldg(pool + tp.dds_off + k), tp.fm_burst_scale)); }
    const float inc = __fadd_rn((float)(uint32_t)tp.fm_word, __fmul_rn(__fmul_rn(a1, 16.0f), tp.fm_mult));
    accum = (uint32_t)(long long)__float2ll_rz(__fadd_rn((float)accum, inc));
    accum %= 65536u;
    return accum >> 6;
}

static constexpr int TX_WARPS = 4;

struct TxWork {
    float a[H2 + CHUNK];        // Hilbert input: [200 history | 128 new]
    float scr[2 * BLK];
    TxState st;
};

// SPLIT = false: the whole chain in one kernel (serial stages on lane 0).  SPLIT = true: the sample-serial stages
// (AudioBufferFill, lattice, biquads, compressor) have been run by tx_serial_kernel, one channel per thread, and
// their output sits in a.scratch [channel][nblocks*32]; this kernel does the 201-tap Hilbert pair, the frequency
// translation and the output formatting.
template <bool SPLIT>
__global__ void __launch_bounds__(32 * TX_WARPS)
tx_ssb_kernel(TxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * TX_WARPS + warp;
    if (ch >= a.num_items) return;
    TxWork &w = reinterpret_cast<TxWork *>(smem_raw)[warp];
    const TxParams &tp = a.txp[ch];
    ChanState *rst = a.state + ch;
    const float *__restrict__ pool = a.pool;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.tx + ch);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&w.st);
        for (int i = lane; i < (int)(sizeof(TxState) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    for (int i = lane; i < H2; i += 32) w.a[i] = w.st.hist[i];
    float osc_q = rst->osc_vect_q, osc_i = rst->osc_vect_i;
    int conv = rst->conversion_freq;
    if (tp.shift_kind != 0 && conv != tp.shift_freq) { conv = tp.shift_freq; osc_i = 0.0f; osc_q = 1.0f; }   // freq_shift.c:289-305
    __syncwarp();
    TxState &st = w.st;
    const size_t base = (size_t)ch * (size_t)a.chan_stride;
    const int2 *__restrict__ mic = reinterpret_cast<const int2 *>(a.audio) + base;
    int2 *__restrict__ iq = reinterpret_cast<int2 *>(a.iq) + base;
    float2 *__restrict__ iq_f = a.iq_f ? reinterpret_cast<float2 *>(a.iq_f) + base : nullptr;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;
    const int N = tp.hil_ntaps;
    const float *ci = pool + (tp.lsb ? tp.hil_cq : tp.hil_ci), *cq = pool + (tp.lsb ? tp.hil_ci : tp.hil_cq);

    for (int blk = 0; blk < a.nblocks; blk++) {
        const bool muted = mute && mute[blk];
        float vi = 0.0f, vq = 0.0f;
        if (!muted && tp.enabled) {
            float v;
            if constexpr (SPLIT) {
                v = a.scratch[((size_t)ch * a.nblocks + blk) * BLK + lane];
            } else {
            // AudioBufferFill
            const int2 s = mic[(size_t)blk * BLK + lane];
            float x = (float)s.x;
            if ((double)tp.gain_calc != 1.0) x = __fmul_rn(x, tp.gain_calc);
            float mx = x, mn = x;
            for (int d = 16; d > 0; d >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d)); }
            w.scr[lane] = x;
            __syncwarp();
            if (lane == 0) {
                st.peak_audio = (-mn > mx) ? -mn : mx;
                for (int i = 0; i < BLK; i++) {
                    float v = w.scr[i];
                    v = lattice_step(v, st.lat_s, pool + tp.lat.k_off, pool + tp.lat.v_off, tp.lat.n);
                    w.scr[i] = v;
                }
                // the reference runs the 3 biquad stages block-wise; per-sample order is equivalent
                for (int i = 0; i < BLK; i++) {
                    float v = w.scr[i];
                    for (int sgi = 0; sgi < 3; sgi++) v = biquad_step(v, tp.bq[sgi], st.bq[sgi]);
                    w.scr[i] = v;
                }
                if (tp.comp_enabled) {
                    for (int i = 0; i < BLK; i++) {
                        const float v = __fmul_rn(w.scr[i], tp.postfilt_gain);
                        w.scr[i] = v;
                        // alc_var = fabsf(a*alc_val)/ALC_KNEE - 1.0 (double), :202
                        const float alc_var = (float)((double)__fdiv_rn(fabsf(__fmul_rn(v, st.alc_val)), 30000.0f) - 1.0);
                        if (alc_var < 0.0f) {
                            st.alc_val = __fsub_rn(st.alc_val, __fmul_rn(__fmul_rn(st.alc_val, tp.alc_decay), alc_var));
                        } else {
                            st.alc_val = (float)((double)st.alc_val - (double)st.alc_val * 0.1 * (double)alc_var);
                            if ((double)st.alc_val < 0.001) st.alc_val = (float)0.001;
                        }
                        if (st.alc_val > 1.0f) st.alc_val = 1.0f;
                        w.scr[BLK + i] = __fmul_rn(st.alc_val, tp.alc_gain_scaling);
                    }
                    st.alc_delay_inbuf += BLK;
                }
            }
            __syncwarp();
            v = w.scr[lane];
            if (tp.comp_enabled) {
                // 320-sample delay line: write at inbuf, read at inbuf + 32 (:231-238)
                const uint32_t inb = st.alc_delay_inbuf % 320u, outb = (st.alc_delay_inbuf + BLK) % 320u;
                __syncwarp();
                st.delay[inb + lane] = v;
                __syncwarp();
                v = __fmul_rn(st.delay[outb + lane], w.scr[BLK + lane]);
                if (lane == 0) st.alc_delay_inbuf = inb;
            }
            if (tp.fm) {                     // pre-emphasis + NCO accumulator are sample-serial: lane 0
                __syncwarp();
                w.scr[lane] = v;
                __syncwarp();
                if (lane == 0) for (int i = 0; i < BLK; i++) w.scr[i] = (float)fm_step(w.scr[i], st.fm_hpf_a, st.fm_hpf_b, st.fm_accum, st.fm_dds_sub_acc, st.fm_dds_burst_acc, tp, pool);
                __syncwarp();
                v = w.scr[lane];
                __syncwarp();
            }
            }
            if (tp.fm) {
                // TxProcessor_FM, tx_processor.c:575-585: I = sine table at the accumulator, Q a quarter turn behind
                // (softdds_phase_shift90), buffers swapped for a negative translate frequency; no Hilbert pair, no FreqShift
                const uint32_t idx = (uint32_t)v;
                const float s0 = __ldg(pool + tp.dds_off + idx), s1 = __ldg(pool + tp.dds_off + ((idx + 768u) & 1023u));
                vi = tp.fm_swap ? s1 : s0; vq = tp.fm_swap ? s0 : s1;
            } else {
            w.a[H2 + lane] = v;
            __syncwarp();
            // Hilbert pair, y[n] = sum_k c[k] a[n - (N-1) + k]
            const float *xs = w.a + H2 + lane - (N - 1);
            float yi = 0.0f, yq = 0.0f;
            for (int k = 0; k < N; k++) { yi = mad(xs[k], __ldg(ci + k), yi); yq = mad(xs[k], __ldg(cq + k), yq); }
            __syncwarp();
            // slide the history by one block
            float keep[(H2 + 31) / 32];
            int cnt = 0;
            for (int i = lane; i < H2; i += 32) keep[cnt++] = w.a[BLK + i];
            __syncwarp();
            cnt = 0;
            for (int i = lane; i < H2; i += 32) w.a[i] = keep[cnt++];
            __syncwarp();
            vi = yi; vq = yq;
            if (tp.am) {                     // both AM sidebands and the carrier (2 x AM_CARRIER_LEVEL = 10200), tx_processor.c:783-790
                vi = __fadd_rn(__fsub_rn(yi, yq), 10200.0f);
                vq = __fsub_rn(__fsub_rn(yq, yi), 10200.0f);
            }
            // FreqShift (:483-486)
            if (tp.shift_kind == 1) {
                float ib = tp.shift_down ? vq : vi, qb = tp.shift_down ? vi : vq;
                const int ph = lane & 3;
                float ni = ib, nq = qb;
                if (ph == 1) { ni = qb; nq = -ib; } else if (ph == 2) { ni = -ib; nq = -qb; } else if (ph == 3) { ni = -qb; nq = ib; }
                if (tp.shift_down) { vq = ni; vi = nq; } else { vi = ni; vq = nq; }
            } else if (tp.shift_kind == 2) {
                if (lane == 0) {
                    float q0 = osc_q, i0 = osc_i;
                    for (int n = 0; n < BLK; n++) {
                        const float oq = __fsub_rn(__fmul_rn(q0, tp.osc_cos), __fmul_rn(i0, tp.osc_sin));
                        const float oi = __fadd_rn(__fmul_rn(i0, tp.osc_cos), __fmul_rn(q0, tp.osc_sin));
                        w.scr[n] = oq; w.scr[BLK + n] = oi; q0 = oq; i0 = oi;
                    }
                    const float g = __fdiv_rn(__fsub_rn(3.0f, __fadd_rn(__fmul_rn(q0, q0), __fmul_rn(i0, i0))), 2.0f);
                    osc_q = __fmul_rn(g, q0); osc_i = __fmul_rn(g, i0);
                }
                __syncwarp();
                const float oq = w.scr[lane], oi = w.scr[BLK + lane];
                float ib = tp.shift_down ? vq : vi, qb = tp.shift_down ? vi : vq;
                const float nq = __fsub_rn(__fmul_rn(qb, oq), __fmul_rn(ib, oi));
                const float ni = __fadd_rn(__fmul_rn(ib, oq), __fmul_rn(qb, oi));
                if (tp.shift_down) { vq = ni; vi = nq; } else { vi = ni; vq = nq; }
                __syncwarp();
            }
            }
        }
        // IqFinalProcessing
        vi = __fmul_rn(vi, tp.final_gain_i);
        vq = __fmul_rn(vq, tp.final_gain_q);
        if (tp.phase_bal < 0.0f) vq = __fadd_rn(vq, __fmul_rn(vi, tp.phase_bal));
        else if (tp.phase_bal > 0.0f) vi = __fadd_rn(vi, __fmul_rn(vq, tp.phase_bal));
        const size_t o = (size_t)blk * BLK + lane;
        iq[o] = make_int2(__float2int_rz(vi), __float2int_rz(vq));
        if (iq_f) iq_f[o] = make_float2(vi, vq);
    }
    __syncwarp();
    for (int i = lane; i < H2; i += 32) w.st.hist[i] = w.a[i];
    if (lane == 0) { st.blocks += a.nblocks; rst->osc_vect_q = osc_q; rst->osc_vect_i = osc_i; rst->conversion_freq = conv; }
    __syncwarp();
    if constexpr (SPLIT) {
        for (int i = lane; i < H2; i += 32) a.tx[ch].hist[i] = w.st.hist[i];
        if (lane == 0) a.tx[ch].blocks = st.blocks;
    } else {
        uint32_t *dst = reinterpret_cast<uint32_t *>(a.tx + ch);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&w.st);
        for (int i = lane; i < (int)(sizeof(TxState) / 4); i += 32) dst[i] = src[i];
    }
}

// The sample-serial stages of the SSB modulator, one channel per thread (see rx_serial.cu for the idea):
// AudioBufferFill (tx_processor.c:339-405), IIR_TXFilter lattice + IIR_TX_biquad (:416-429), VoiceCompressor
// with its 320-sample look-ahead delay (:173-242).  Muted blocks advance nothing, as in TxProcessor_Run.
static constexpr int TXS_THREADS = 32;

__global__ void __launch_bounds__(TXS_THREADS)
tx_serial_kernel(TxArgs a)
{
    const int ch = blockIdx.x * TXS_THREADS + threadIdx.x;
    if (ch >= a.num_items) return;
    const TxParams &tp = a.txp[ch];
    if (!tp.enabled) return;
    const float *__restrict__ pool = a.pool;
    TxState &g = a.tx[ch];
    float lat_s[MAX_LAT];
    for (int i = 0; i < MAX_LAT; i++) lat_s[i] = g.lat_s[i];
    BiquadS bq[3];
    float bc[3][5];
    for (int s = 0; s < 3; s++) { bq[s] = g.bq[s]; for (int q = 0; q < 5; q++) bc[s][q] = tp.bq[s][q]; }
    float alc_val = g.alc_val, peak_audio = g.peak_audio;
    float fm_hpf_a = g.fm_hpf_a, fm_hpf_b = g.fm_hpf_b;
    uint32_t fm_accum = g.fm_accum, sub_acc = g.fm_dds_sub_acc, burst_acc = g.fm_dds_burst_acc;
    uint32_t inbuf = g.alc_delay_inbuf;
    float delay[320];
    for (int i = 0; i < 320; i++) delay[i] = g.delay[i];
    const float gain_calc = tp.gain_calc, postfilt_gain = tp.postfilt_gain, alc_decay = tp.alc_decay;
    const bool gain_on = (double)gain_calc != 1.0, comp = tp.comp_enabled != 0;
    const float *lk = pool + tp.lat.k_off, *lv = pool + tp.lat.v_off;
    const int ln = tp.lat.n;
    const int2 *__restrict__ mic = reinterpret_cast<const int2 *>(a.audio) + (size_t)ch * (size_t)a.chan_stride;
    float *__restrict__ out = a.scratch + (size_t)ch * (size_t)a.nblocks * BLK;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

    for (int blk = 0; blk < a.nblocks; blk++) {
        if (mute && mute[blk]) continue;
        float v[BLK], alc[BLK];
        float mx = 0.0f, mn = 0.0f;
        for (int i = 0; i < BLK; i += 2) {
            const int4 s = *reinterpret_cast<const int4 *>(mic + (size_t)blk * BLK + i);
            float x0 = (float)s.x, x1 = (float)s.z;
            if (gain_on) { x0 = __fmul_rn(x0, gain_calc); x1 = __fmul_rn(x1, gain_calc); }
            if (i == 0) { mx = x0; mn = x0; }
            mx = fmaxf(mx, fmaxf(x0, x1)); mn = fminf(mn, fminf(x0, x1));
            v[i] = x0; v[i + 1] = x1;
        }
        peak_audio = (-mn > mx) ? -mn : mx;
        for (int i = 0; i < BLK; i++) {
            float x = lattice_step(v[i], lat_s, lk, lv, ln);
            for (int s = 0; s < 3; s++) x = biquad_step(x, bc[s], bq[s]);
            v[i] = x;
        }
        if (comp) {
            for (int i = 0; i < BLK; i++) {
                const float x = __fmul_rn(v[i], postfilt_gain);
                v[i] = x;
                // alc_var = fabsf(a*alc_val)/ALC_KNEE - 1.0 (double), tx_processor.c:202
                const float alc_var = (float)((double)__fdiv_rn(fabsf(__fmul_rn(x, alc_val)), 30000.0f) - 1.0);
                if (alc_var < 0.0f) {
                    alc_val = __fsub_rn(alc_val, __fmul_rn(__fmul_rn(alc_val, alc_decay), alc_var));
                } else {
                    alc_val = (float)((double)alc_val - (double)alc_val * 0.1 * (double)alc_var);
                    if ((double)alc_val < 0.001) alc_val = (float)0.001;
                }
                if (alc_val > 1.0f) alc_val = 1.0f;
                alc[i] = __fmul_rn(alc_val, tp.alc_gain_scaling);
            }
            inbuf += BLK;
            // 320-sample delay line: write at inbuf, read at inbuf + 32 (:231-238)
            const uint32_t inb = inbuf % 320u, outb = (inbuf + BLK) % 320u;
            for (int i = 0; i < BLK; i++) delay[inb + i] = v[i];
            for (int i = 0; i < BLK; i++) v[i] = __fmul_rn(delay[outb + i], alc[i]);
            inbuf = inb;
        }
        if (tp.fm) {
            for (int i = 0; i < BLK; i++) v[i] = (float)fm_step(v[i], fm_hpf_a, fm_hpf_b, fm_accum, sub_acc, burst_acc, tp, pool);      // table index for the FIR-stage kernel
        }
        for (int i = 0; i < BLK; i += 4) *reinterpret_cast<float4 *>(out + (size_t)blk * BLK + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    g.fm_hpf_a = fm_hpf_a; g.fm_hpf_b = fm_hpf_b; g.fm_accum = fm_accum; g.fm_dds_sub_acc = sub_acc; g.fm_dds_burst_acc = burst_acc;
    for (int i = 0; i < MAX_LAT; i++) g.lat_s[i] = lat_s[i];
    for (int s = 0; s < 3; s++) g.bq[s] = bq[s];
    g.alc_val = alc_val; g.peak_audio = peak_audio; g.alc_delay_inbuf = inbuf;
    for (int i = 0; i < 320; i++) g.delay[i] = delay[i];
}

// Second-generation FIR stage of the split modulator (no mute array): 512-sample chunks, the 201-tap Hilbert pair register-blocked
// (fir_device.cuh fir4_m1_dual: both tap sets on one window, four outputs per lane, taps staged in shared memory), one 32-byte
// store of four {I, Q} samples per lane.  Same arithmetic order as tx_ssb_kernel<true> (taps ascending), so the exact build
// stays bit-identical.
namespace {
constexpr int TX2_CS = 512, TX2_TAPS = 208;
struct Tx2Work {
    alignas(16) float a[H2 + TX2_CS + 8];
    alignas(16) float ti[TX2_TAPS], tq[TX2_TAPS];
    float osc[2 * TX2_CS];
};
}  // namespace

__global__ void __launch_bounds__(32 * TX_WARPS)
tx_fir2_kernel(TxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * TX_WARPS + warp;
    if (ch >= a.num_items) return;
    Tx2Work &w = reinterpret_cast<Tx2Work *>(smem_raw)[warp];
    const TxParams &tp = a.txp[ch];
    ChanState *rst = a.state + ch;
    TxState *tst = a.tx + ch;
    const float *__restrict__ pool = a.pool;
    const int N = tp.hil_ntaps;
    for (int i = lane; i < H2; i += 32) w.a[i] = tst->hist[i];
    for (int i = lane; i < TX2_CS + 8; i += 32) w.a[H2 + i] = 0.0f;
    fir_stage_taps(w.ti, pool + (tp.lsb ? tp.hil_cq : tp.hil_ci), N, lane);      // I/Q filters swapped for LSB (tx_processor.c:477-478)
    fir_stage_taps(w.tq, pool + (tp.lsb ? tp.hil_ci : tp.hil_cq), N, lane);
    const int groups = fir_padded_len(N) / 4, off = N - 1 + fir_pad_front(N);
    float osc_q = rst->osc_vect_q, osc_i = rst->osc_vect_i;
    int conv = rst->conversion_freq;
    if (tp.shift_kind != 0 && conv != tp.shift_freq) { conv = tp.shift_freq; osc_i = 0.0f; osc_q = 1.0f; }   // freq_shift.c:289-305
    __syncwarp();
    const size_t base = (size_t)ch * (size_t)a.chan_stride;
    int2 *__restrict__ iq = reinterpret_cast<int2 *>(a.iq) + base;
    float2 *__restrict__ iq_f = a.iq_f ? reinterpret_cast<float2 *>(a.iq_f) + base : nullptr;
    const float *__restrict__ src = a.scratch + (size_t)ch * (size_t)a.nblocks * BLK;

    for (int s0 = 0; s0 < a.nblocks * BLK; s0 += TX2_CS) {
        const int ns = min(TX2_CS, a.nblocks * BLK - s0);
        if (tp.enabled && !tp.fm) {
            for (int i = lane; i < ns; i += 32) w.a[H2 + i] = src[s0 + i];
            if (tp.shift_kind == 2) {
                // FreqShift_Approx (freq_shift.c:57-108): recursive oscillator, renormalised after every 32-sample block
                if (lane == 0) {
                    float q0 = osc_q, i0 = osc_i;
                    for (int n = 0; n < ns; n++) {
                        const float oq = __fsub_rn(__fmul_rn(q0, tp.osc_cos), __fmul_rn(i0, tp.osc_sin));
                        const float oi = __fadd_rn(__fmul_rn(i0, tp.osc_cos), __fmul_rn(q0, tp.osc_sin));
                        w.osc[n] = oq; w.osc[TX2_CS + n] = oi; q0 = oq; i0 = oi;
                        if ((n & (BLK - 1)) == BLK - 1) {
                            const float g = __fdiv_rn(__fsub_rn(3.0f, __fadd_rn(__fmul_rn(q0, q0), __fmul_rn(i0, i0))), 2.0f);
                            q0 = __fmul_rn(g, q0); i0 = __fmul_rn(g, i0);
                        }
                    }
                    osc_q = q0; osc_i = i0;
                }
                osc_q = __shfl_sync(0xffffffffu, osc_q, 0); osc_i = __shfl_sync(0xffffffffu, osc_i, 0);
            }
            __syncwarp();
        }
        for (int sub = 0; sub < ns; sub += 128) {
            const int n0 = sub + 4 * lane;                 // first of this lane's four samples within the chunk
            float vi[4] = { 0.0f, 0.0f, 0.0f, 0.0f }, vq[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
            if (tp.enabled && n0 < ns) {
                if (tp.fm) {
                    // TxProcessor_FM, tx_processor.c:575-585: I = sine table at the accumulator, Q a quarter turn behind
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const uint32_t idx = (uint32_t)src[s0 + n0 + r];
                        const float t0 = __ldg(pool + tp.dds_off + idx), t1 = __ldg(pool + tp.dds_off + ((idx + 768u) & 1023u));
                        vi[r] = tp.fm_swap ? t1 : t0; vq[r] = tp.fm_swap ? t0 : t1;
                    }
                } else {
                    float yi[4], yq[4];
                    fir4_m1_dual<false>(w.a + H2 + n0 - off, w.ti, w.tq, groups, yi, yq);
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        float i_ = yi[r], q_ = yq[r];
                        if (tp.am) {                         // both AM sidebands and the carrier, tx_processor.c:783-790
                            i_ = __fadd_rn(__fsub_rn(yi[r], yq[r]), 10200.0f);
                            q_ = __fsub_rn(__fsub_rn(yq[r], yi[r]), 10200.0f);
                        }
                        if (tp.shift_kind == 1) {            // FreqShift_QuarterFs: the phase is the sample index in the block, = r
                            float ib = tp.shift_down ? q_ : i_, qb = tp.shift_down ? i_ : q_;
                            float ni = ib, nq = qb;
                            if (r == 1) { ni = qb; nq = -ib; } else if (r == 2) { ni = -ib; nq = -qb; } else if (r == 3) { ni = -qb; nq = ib; }
                            if (tp.shift_down) { q_ = ni; i_ = nq; } else { i_ = ni; q_ = nq; }
                        } else if (tp.shift_kind == 2) {
                            const float oq = w.osc[n0 + r], oi = w.osc[TX2_CS + n0 + r];
                            float ib = tp.shift_down ? q_ : i_, qb = tp.shift_down ? i_ : q_;
                            const float nq = __fsub_rn(__fmul_rn(qb, oq), __fmul_rn(ib, oi));
                            const float ni = __fadd_rn(__fmul_rn(ib, oq), __fmul_rn(qb, oi));
                            if (tp.shift_down) { q_ = ni; i_ = nq; } else { i_ = ni; q_ = nq; }
                        }
                        vi[r] = i_; vq[r] = q_;
                    }
                }
            }
            if (n0 < ns) {
                int wi[4], wq[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {                // IqFinalProcessing, tx_processor.c:282-330
                    float i_ = __fmul_rn(vi[r], tp.final_gain_i), q_ = __fmul_rn(vq[r], tp.final_gain_q);
                    if (tp.phase_bal < 0.0f) q_ = __fadd_rn(q_, __fmul_rn(i_, tp.phase_bal));
                    else if (tp.phase_bal > 0.0f) i_ = __fadd_rn(i_, __fmul_rn(q_, tp.phase_bal));
                    wi[r] = __float2int_rz(i_); wq[r] = __float2int_rz(q_);
                    if (iq_f) iq_f[s0 + n0 + r] = make_float2(i_, q_);
                }
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(iq + s0 + n0), "r"(wi[0]), "r"(wq[0]), "r"(wi[1]), "r"(wq[1]),
                             "r"(wi[2]), "r"(wq[2]), "r"(wi[3]), "r"(wq[3]) : "memory");
            }
        }
        __syncwarp();
        if (tp.enabled && !tp.fm) {
            // keep the newest H2 samples
            float keep[(H2 + 31) / 32];
            int cnt = 0;
            for (int i = lane; i < H2; i += 32) keep[cnt++] = w.a[ns + i];
            __syncwarp();
            cnt = 0;
            for (int i = lane; i < H2; i += 32) w.a[i] = keep[cnt++];
            __syncwarp();
        }
    }
    for (int i = lane; i < H2; i += 32) tst->hist[i] = w.a[i];
    if (lane == 0) { tst->blocks += a.nblocks; rst->osc_vect_q = osc_q; rst->osc_vect_i = osc_i; rst->conversion_freq = conv; }
}

static bool tx_fir2_ok(const TxArgs &a)
{
    static const bool off = [] { const char *v = getenv("UHSDR_B200_NO_FRONT2"); return v && v[0] == '1'; }();
    return !off && a.scratch != nullptr && a.mute == nullptr && ((uintptr_t)a.iq % 32 == 0) && (a.chan_stride % 4 == 0) && (a.nblocks * BLK) % 4 == 0;
}

cudaError_t launch_tx_ssb(const TxArgs &a, cudaStream_t stream)
{
    if (tx_fir2_ok(a)) {
        const size_t smem2 = sizeof(Tx2Work) * TX_WARPS;
        const int grid2 = (a.num_items + TX_WARPS - 1) / TX_WARPS;
        if (grid2 == 0) return cudaSuccess;
        cudaError_t e2 = cudaFuncSetAttribute(tx_fir2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        tx_fir2_kernel<<<grid2, 32 * TX_WARPS, smem2, stream>>>(a);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(TxWork) * TX_WARPS;
    const int grid = (a.num_items + TX_WARPS - 1) / TX_WARPS;
    if (grid == 0) return cudaSuccess;
    if (a.scratch) {
        cudaError_t e = cudaFuncSetAttribute(tx_ssb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tx_ssb_kernel<true><<<grid, 32 * TX_WARPS, smem, stream>>>(a);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(tx_ssb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tx_ssb_kernel<false><<<grid, 32 * TX_WARPS, smem, stream>>>(a);
    return cudaGetLastError();
}

// Second generation of the serial stages: the same operations in the same order in the exact build and for FM modulators (bit-
// identical results), fused forms otherwise (float tolerance); the
// per-sample loop is rolled and touches registers and shared memory only -- lattice and biquad coefficients and states in
// registers (the lattice front-padded to 10 stages: k = v = 0 stages pass the sample through exactly), the block's samples and the
// 320-sample look-ahead delay line of the compressor in shared memory [slot][lane] (the first version kept them in per-thread
// local memory and fetched the lattice coefficients from global memory for every stage of every sample).
namespace {
constexpr int TXS2_DELAY = 320;
}

__global__ void __launch_bounds__(2 * TXS_THREADS)
tx_serial2_kernel(TxArgs a)
{
    // Two warps per 32 channels, one iteration and one __syncthreads per block.  Warp 0 (FILTERS) runs the lattice and the biquads
    // of block t; warp 1 (REST) converts the microphone words of block t + 1, and runs the compressor's detector, the delay line,
    // the FM accumulator and the store of block t - 1.  Each is a dependent chain with nothing else to issue, so splitting the
    // chain over two warps nearly halves the time per block; the buffers between them are double-buffered in shared memory.
    extern __shared__ __align__(16) float txs2_smem[];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int ch = blockIdx.x * TXS_THREADS + lane;
    const bool valid = ch < a.num_items && a.txp[ch < a.num_items ? ch : 0].enabled;
    const TxParams &tp = a.txp[valid ? ch : 0];
    const float *__restrict__ pool = a.pool;
    TxState &g = a.tx[valid ? ch : 0];
    float *dl = txs2_smem + lane;                                    // delay line [320][32]
    float *xin = dl + TXS2_DELAY * TXS_THREADS;                      // converted microphone samples [2][32][32]
    float *xb = xin + 2 * BLK * TXS_THREADS;                         // filter outputs [2][32][32]
    float *sa = xb + 2 * BLK * TXS_THREADS;                          // ALC gains of the block [32][32]
    const float gain_calc = tp.gain_calc, postfilt_gain = tp.postfilt_gain, alc_decay = tp.alc_decay, alc_scale = tp.alc_gain_scaling;
    const bool gain_on = (double)gain_calc != 1.0, comp = tp.comp_enabled != 0, fm = tp.fm != 0;
    const int2 *__restrict__ mic = reinterpret_cast<const int2 *>(a.audio) + (size_t)(valid ? ch : 0) * (size_t)a.chan_stride;
    float *__restrict__ out = a.scratch + (size_t)(valid ? ch : 0) * (size_t)a.nblocks * BLK;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)(valid ? ch : 0) * (size_t)a.mute_stride : nullptr;
    const int nblocks = a.nblocks;
    // one arithmetic per warp (the loops below hold the CTA's barriers: no lane may take another path): the reference's operation
    // order if any channel of the warp is an FM modulator
    const bool exm = UHSDR_EXACT || __any_sync(0xffffffffu, valid && fm);

    if (role == 0) {
        // ======================= FILTERS: lattice + 3 biquads (FilterAudio, tx_processor.c:416-429) =======================
        float lk[10], lv[11], ls[10];
        const int ln = tp.lat.n, lpad = 10 - ln;
#pragma unroll
        for (int j = 0; j < 10; j++) {
            lk[j] = (j >= lpad) ? __ldg(pool + tp.lat.k_off + (j - lpad)) : 0.0f;
            lv[j] = (j >= lpad) ? __ldg(pool + tp.lat.v_off + (j - lpad)) : 0.0f;
            ls[j] = (j >= lpad) ? g.lat_s[j - lpad] : 0.0f;
        }
        lv[10] = __ldg(pool + tp.lat.v_off + ln);
        BiquadS bq[3];
        float bc[3][5];
#pragma unroll
        for (int s = 0; s < 3; s++) { bq[s] = g.bq[s];
#pragma unroll
            for (int q = 0; q < 5; q++) bc[s][q] = tp.bq[s][q]; }
        // Shipping build: fused multiply-adds, biquads with the new sample entering last (one multiply-add on the sample-to-sample
        // path per stage).  Exact build, and the FM modulator in both builds (its audio drives an integer phase accumulator:
        // rounding differences would add up): the reference's operations one by one.
        auto run = [&](auto exm) {
            constexpr bool EXM = decltype(exm)::value;
            for (int t = -1; t <= nblocks; t++) {
                if (valid && t >= 0 && t < nblocks && !(mute && mute[t])) {
                    const float *src = xin + (t & 1) * BLK * TXS_THREADS;
                    float *dst = xb + (t & 1) * BLK * TXS_THREADS;
#pragma unroll 1
                    for (int i = 0; i < BLK; i++) {
                        float f = src[i * TXS_THREADS], acc = 0.0f, fn = 0.0f;
#pragma unroll
                        for (int j = 0; j < 10; j++) {
                            const float gg = ls[j];
                            float gn;
                            if constexpr (EXM) {
                                fn = __fsub_rn(f, __fmul_rn(lk[j], gg));
                                gn = __fadd_rn(__fmul_rn(fn, lk[j]), gg);
                                acc = __fadd_rn(acc, __fmul_rn(gn, lv[j]));
                            } else {
                                fn = fmaf(-lk[j], gg, f);
                                gn = fmaf(fn, lk[j], gg);
                                acc = fmaf(gn, lv[j], acc);
                            }
                            if (j > 0) ls[j - 1] = gn;
                            f = fn;
                        }
                        float x = EXM ? __fadd_rn(acc, __fmul_rn(fn, lv[10])) : fmaf(fn, lv[10], acc);
                        ls[9] = fn;
#pragma unroll
                        for (int s = 0; s < 3; s++) {
                            if constexpr (EXM) { x = biquad_step(x, bc[s], bq[s]); continue; }
                            const float tt = fmaf(bc[s][1], bq[s].x1, fmaf(bc[s][2], bq[s].x2, fmaf(bc[s][3], bq[s].y1, __fmul_rn(bc[s][4], bq[s].y2))));
                            const float y = fmaf(bc[s][0], x, tt);
                            bq[s].x2 = bq[s].x1; bq[s].x1 = x; bq[s].y2 = bq[s].y1; bq[s].y1 = y;
                            x = y;
                        }
                        dst[i * TXS_THREADS] = comp ? __fmul_rn(x, postfilt_gain) : x;
                    }
                }
                __syncthreads();
            }
        };
        if (exm) run(std::true_type{}); else run(std::false_type{});
        if (valid) {
#pragma unroll
            for (int j = 0; j < 10; j++) if (j >= lpad) g.lat_s[j - lpad] = ls[j];
#pragma unroll
            for (int s = 0; s < 3; s++) g.bq[s] = bq[s];
        }
        return;
    }

    // ======================= REST: AudioBufferFill of block t + 1; compressor, delay line, FM accumulator, store of block t - 1 ====
    float alc_val = g.alc_val, peak_audio = g.peak_audio;
    float fm_hpf_a = g.fm_hpf_a, fm_hpf_b = g.fm_hpf_b;
    uint32_t fm_accum = g.fm_accum, sub_acc = g.fm_dds_sub_acc, burst_acc = g.fm_dds_burst_acc;
    uint32_t inbuf = g.alc_delay_inbuf;
    if (valid) for (int i = 0; i < TXS2_DELAY; i++) dl[i * TXS_THREADS] = g.delay[i];
    // microphone words (the .l words of 32 AudioSample_t = every other int32), fetched one block ahead of their conversion
    int4 nxt[BLK / 2];
    auto fetch = [&](int blk) {
#pragma unroll
        for (int i = 0; i < BLK / 2; i++)       // volatile: the compiler must not sink the loads to their first use one block later
            asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(nxt[i].x), "=r"(nxt[i].y), "=r"(nxt[i].z), "=r"(nxt[i].w)
                         : "l"(mic + (size_t)blk * BLK + 2 * i));
    };
    if (valid && nblocks > 0) fetch(0);
    for (int t = -1; t <= nblocks; t++) {
        if (valid && t + 1 < nblocks) {
            // AudioBufferFill (tx_processor.c:339-405) of block t + 1
            const int b1 = t + 1;
            float *dst = xin + (b1 & 1) * BLK * TXS_THREADS;
            float mx = 0.0f, mn = 0.0f;
            const bool muted = mute && mute[b1];
#pragma unroll
            for (int i = 0; i < BLK; i += 2) {
                float x0 = (float)nxt[i / 2].x, x1 = (float)nxt[i / 2].z;
                if (gain_on) { x0 = __fmul_rn(x0, gain_calc); x1 = __fmul_rn(x1, gain_calc); }
                if (i == 0) { mx = x0; mn = x0; }
                mx = fmaxf(mx, fmaxf(x0, x1)); mn = fminf(mn, fminf(x0, x1));
                dst[i * TXS_THREADS] = x0; dst[(i + 1) * TXS_THREADS] = x1;
            }
            if (!muted) peak_audio = (-mn > mx) ? -mn : mx;
            if (b1 + 1 < nblocks) fetch(b1 + 1);
        }
        if (valid && t >= 1 && !(mute && mute[t - 1])) {
            const int b0 = t - 1;
            float *sv = xb + (b0 & 1) * BLK * TXS_THREADS;
            if (comp) {
                // the compressor's detector (:173-242), sample by sample
#pragma unroll 1
                for (int i = 0; i < BLK; i++) {
                    const float xp = sv[i * TXS_THREADS];
                    float alc_var, dn, up;
                    if (exm) {
                        // alc_var = fabsf(a*alc_val)/ALC_KNEE - 1.0 (double), tx_processor.c:202
                        alc_var = (float)((double)__fdiv_rn(fabsf(__fmul_rn(xp, alc_val)), 30000.0f) - 1.0);
                        dn = __fsub_rn(alc_val, __fmul_rn(__fmul_rn(alc_val, alc_decay), alc_var));
                        up = (float)((double)alc_val - (double)alc_val * 0.1 * (double)alc_var);
                        if ((double)up < 0.001) up = (float)0.001;
                    } else {
                        alc_var = fmaf(fabsf(__fmul_rn(xp, alc_val)), 1.0f / 30000.0f, -1.0f);
                        dn = fmaf(-__fmul_rn(alc_val, alc_decay), alc_var, alc_val);
                        up = fmaxf(fmaf(-__fmul_rn(alc_val, 0.1f), alc_var, alc_val), 0.001f);
                    }
                    alc_val = (alc_var < 0.0f) ? dn : up;
                    if (alc_val > 1.0f) alc_val = 1.0f;
                    sa[i * TXS_THREADS] = __fmul_rn(alc_val, alc_scale);
                }
                inbuf += BLK;
                // 320-sample delay line: write at inbuf, read at inbuf + 32 (:231-238)
                const uint32_t inb = inbuf % 320u, outb = (inbuf + BLK) % 320u;
#pragma unroll 4
                for (int i = 0; i < BLK; i++) dl[(inb + i) * TXS_THREADS] = sv[i * TXS_THREADS];
#pragma unroll 4
                for (int i = 0; i < BLK; i++) sv[i * TXS_THREADS] = __fmul_rn(dl[(outb + i) * TXS_THREADS], sa[i * TXS_THREADS]);
                inbuf = inb;
            }
            if (fm) {
#pragma unroll 1
                for (int i = 0; i < BLK; i++)
                    sv[i * TXS_THREADS] = (float)fm_step(sv[i * TXS_THREADS], fm_hpf_a, fm_hpf_b, fm_accum, sub_acc, burst_acc, tp, pool);      // table index for the FIR-stage kernel
            }
#pragma unroll
            for (int i = 0; i < BLK; i += 4)
                *reinterpret_cast<float4 *>(out + (size_t)b0 * BLK + i) = make_float4(sv[i * TXS_THREADS], sv[(i + 1) * TXS_THREADS], sv[(i + 2) * TXS_THREADS], sv[(i + 3) * TXS_THREADS]);
        }
        __syncthreads();
    }
    if (valid) {
        g.fm_hpf_a = fm_hpf_a; g.fm_hpf_b = fm_hpf_b; g.fm_accum = fm_accum; g.fm_dds_sub_acc = sub_acc; g.fm_dds_burst_acc = burst_acc;
        g.alc_val = alc_val; g.peak_audio = peak_audio; g.alc_delay_inbuf = inbuf;
        for (int i = 0; i < TXS2_DELAY; i++) g.delay[i] = dl[i * TXS_THREADS];
    }
}

cudaError_t launch_tx_serial(const TxArgs &a, cudaStream_t stream)
{
    if (a.num_items <= 0) return cudaSuccess;
    if (a.scratch == nullptr) return cudaErrorInvalidValue;
    static const bool no2 = [] { const char *v = getenv("UHSDR_B200_NO_SERIAL2"); return v && v[0] == '1'; }();
    if (!no2 && ((uintptr_t)a.audio % 16 == 0) && (a.chan_stride % 2 == 0)) {
        const size_t smem = (size_t)(TXS2_DELAY + 5 * BLK) * TXS_THREADS * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(tx_serial2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tx_serial2_kernel<<<(a.num_items + TXS_THREADS - 1) / TXS_THREADS, 2 * TXS_THREADS, smem, stream>>>(a);
        return cudaGetLastError();
    }
    tx_serial_kernel<<<(a.num_items + TXS_THREADS - 1) / TXS_THREADS, TXS_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace uhsdr
