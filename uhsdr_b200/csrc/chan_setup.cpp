// chan_setup.cpp -- host-side configuration of one channel: the engine's equivalent of
// AudioDriver_SetProcessingChain (mchf-eclipse/drivers/audio/audio_driver.c:1093-1251),
// AudioFilter_SetRxHilbertAndDecimationFIR (audio_filter.c:1134-1223),
// AudioAgc_SetupAgcWdsp (audio_agc.c:126-339) and AudioDriver_SetSamPllParameters
// (audio_driver.c:709-745).  Runs on the CPU with glibc so that every derived constant carries
// the same float/double promotions as the reference build (no -fsingle-precision-constant).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "host_tables.h"

namespace uhsdr {

static const float kPi = 3.14159265358979f;   // CMSIS arm_math.h PI

bool HostTables::load(const void *data, size_t bytes, std::string *err)
{
    if (data == nullptr || bytes < sizeof(uhsdr_tbl_header_t)) { if (err) *err = "table blob missing or truncated"; return false; }
    const uhsdr_tbl_header_t *hh = static_cast<const uhsdr_tbl_header_t *>(data);
    if (hh->magic != UHSDR_TABLES_MAGIC || hh->version != UHSDR_TABLES_VERSION || hh->total_bytes != bytes) {
        if (err) *err = "table blob: bad magic/version/size";
        return false;
    }
    // every section must lie inside the blob before anything is dereferenced
    auto section_ok = [&](uint32_t off, uint64_t count, size_t elem) {
        return (off % 4) == 0 && (uint64_t)off + count * (uint64_t)elem <= (uint64_t)bytes;
    };
    if (!section_ok(hh->arrays_off, hh->num_arrays, sizeof(uhsdr_tbl_array_t)) || !section_ok(hh->paths_off, hh->num_paths, sizeof(uhsdr_tbl_path_t)) ||
        !section_ok(hh->filters_off, hh->num_filters, sizeof(uhsdr_tbl_filter_t)) || !section_ok(hh->lattices_off, hh->num_lattices, sizeof(uhsdr_tbl_lattice_t)) ||
        !section_ok(hh->interps_off, hh->num_interps, sizeof(uhsdr_tbl_interp_t)) || !section_ok(hh->extras_off, 1, sizeof(uhsdr_tbl_extras_t)) ||
        hh->num_arrays > (1u << 20)) {
        if (err) *err = "table blob: section offset / count out of range";
        return false;
    }
    blob.assign(static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + bytes);
    const uint8_t *b = blob.data();
    h = reinterpret_cast<const uhsdr_tbl_header_t *>(b);
    arr = reinterpret_cast<const uhsdr_tbl_array_t *>(b + h->arrays_off);
    path = reinterpret_cast<const uhsdr_tbl_path_t *>(b + h->paths_off);
    filt = reinterpret_cast<const uhsdr_tbl_filter_t *>(b + h->filters_off);
    lat = reinterpret_cast<const uhsdr_tbl_lattice_t *>(b + h->lattices_off);
    interp = reinterpret_cast<const uhsdr_tbl_interp_t *>(b + h->interps_off);
    ex = reinterpret_cast<const uhsdr_tbl_extras_t *>(b + h->extras_off);
    if (h->num_paths != UHSDR_NUM_FILTER_PATHS) { if (err) *err = "table blob: unexpected path count"; return false; }
    // cross-indices: -1 = NULL where the firmware has a NULL pointer, otherwise inside the table it points into
    {
        const int na = (int)h->num_arrays, nl = (int)h->num_lattices, ni = (int)h->num_interps, nf = (int)h->num_filters;
        auto arr_ok = [&](int a, bool opt) { return (opt && a == -1) || (a >= 0 && a < na); };
        auto cnt_ok = [&](int a, int need) { return a < 0 || (int64_t)arr[a].count >= (int64_t)need; };
        bool ok = true;
        for (uint32_t i = 0; i < h->num_paths && ok; i++) {
            const uhsdr_tbl_path_t &fp = path[i];
            ok = fp.id >= 0 && fp.id < nf && arr_ok(fp.fir_i_array, true) && arr_ok(fp.fir_q_array, true) && arr_ok(fp.dec_array, true) &&
                 (fp.pre_lattice == -1 || (fp.pre_lattice >= 0 && fp.pre_lattice < nl)) && (fp.aa_lattice == -1 || (fp.aa_lattice >= 0 && fp.aa_lattice < nl)) &&
                 (fp.interpolate == -1 || (fp.interpolate >= 0 && fp.interpolate < ni)) && fp.fir_numtaps >= 0 && fp.dec_numtaps >= 0 &&
                 cnt_ok(fp.fir_i_array, fp.fir_numtaps) && cnt_ok(fp.fir_q_array, fp.fir_numtaps) && cnt_ok(fp.dec_array, fp.dec_numtaps);
        }
        for (int i = 0; i < nl && ok; i++)
            ok = lat[i].num_stages >= 0 && lat[i].num_stages <= 64 && arr_ok(lat[i].k_array, false) && arr_ok(lat[i].v_array, false) &&
                 cnt_ok(lat[i].k_array, lat[i].num_stages) && cnt_ok(lat[i].v_array, lat[i].num_stages + 1);
        for (int i = 0; i < ni && ok; i++)
            ok = interp[i].L >= 0 && interp[i].num_coeffs >= 0 && arr_ok(interp[i].coeff_array, false) && cnt_ok(interp[i].coeff_array, interp[i].num_coeffs);
        if (ok) {
            const int ex_arrays[] = { ex->nr_decimate_array, ex->nr_interpolate_array, ex->sqrt_hann_256_array, ex->spectrum_window_array, ex->sam_c0_array,
                                      ex->sam_c1_array, ex->tx_hilbert_i_array, ex->tx_hilbert_q_array, ex->dds_table_array, ex->zoom_biquad_array, ex->zoom_decim_array };
            for (int a : ex_arrays) ok = ok && arr_ok(a, true);
            const int ex_lats[] = { ex->fm_squelch_lattice, ex->tx_lattice_soprano, ex->tx_lattice_tenor, ex->tx_lattice_bass, ex->tx_lattice_fm };
            for (int l : ex_lats) ok = ok && (l == -1 || (l >= 0 && l < nl));
        }
        if (!ok) { if (err) *err = "table blob: cross-index out of range"; return false; }
    }
    pool.clear();
    pool_off.assign(h->num_arrays, 0);
    for (uint32_t i = 0; i < h->num_arrays; i++) {
        if ((arr[i].offset % 4) != 0 || (uint64_t)arr[i].offset + (uint64_t)arr[i].count * 4 > (uint64_t)bytes) { if (err) *err = "table blob: array out of range"; return false; }
        while (pool.size() % 4) pool.push_back(0.0f);
        pool_off[i] = (int)pool.size();
        const float *src = reinterpret_cast<const float *>(b + arr[i].offset);
        pool.insert(pool.end(), src, src + arr[i].count);
    }
    // FFT twiddles e^{-2 pi i k / N}, k < N/2, interleaved (cos, sin), computed in double
    auto add_tw = [&](int N) {
        while (pool.size() % 4) pool.push_back(0.0f);
        int o = (int)pool.size();
        for (int k = 0; k < N / 2; k++) {
            double a = -2.0 * M_PI * (double)k / (double)N;
            pool.push_back((float)cos(a));
            pool.push_back((float)sin(a));
        }
        return o;
    };
    tw256_off = add_tw(256);
    tw512_off = add_tw(512);
    return true;
}

void default_chan_cfg(uhsdr_chan_cfg_t *c)
{
    memset(c, 0, sizeof(*c));
    c->struct_size = sizeof(*c);
    c->dmod_mode = UHSDR_DEMOD_USB;
    c->filter_path = 35;
    c->iq_freq_mode = UHSDR_FREQ_IQ_CONV_M12KHZ;
    c->iq_auto_correction = 1;
    c->rx_adj_gain_i = 1.0f; c->rx_adj_gain_q = 1.0f;
    c->notch_frequency = 800; c->peak_frequency = 750;
    c->bass_gain = 2; c->treble_gain = 0; c->nr_strength = 160;
    c->agc_mode = 2; c->agc_slope = 70; c->agc_hang_enable = 0; c->agc_thresh = 20;
    c->agc_hang_thresh = 45; c->agc_hang_time = 500;
    const int td[6] = { 4000, 2000, 500, 250, 50, 1 };
    for (int i = 0; i < 6; i++) c->agc_tau_decay[i] = td[i];
    c->agc_tau_hang_decay = 500;
    c->sam_sideband = UHSDR_SAM_SIDEBAND_BOTH; c->sam_fade_leveler = 1;
    c->sam_pll_fmax = 2500; c->sam_zeta = 65; c->sam_omegaN = 250;
    c->fm_sql_threshold = 12;
    c->nr_decimation_enable = 1;
    c->codec_gain_calc = 1.0f;
    c->tx_filter = UHSDR_TX_FILTER_SOPRANO; c->tx_bass_gain = 4; c->tx_treble_gain = 4; c->tx_mic_gain = 15;
    c->tx_comp_level = 2; c->tx_alc_decay = 10; c->tx_alc_postfilt_gain = 1;
    c->tx_power_factor = 0.5f; c->tx_adj_gain_i = 1.0f; c->tx_adj_gain_q = 1.0f;
    c->notch_mu = 10;
}

// ---- biquad designers, audio_driver.c:818-964 (a1/a2 stored already negated) ----------------
static void bq_scale(float c[5], float sa, float sb) { c[3] = c[3] / sa; c[4] = c[4] / sa; c[0] = c[0] / sb; c[1] = c[1] / sb; c[2] = c[2] / sb; }

static void design_notch(float c[5], float f0, float FS)       // AudioDriver_CalcBandstop :831
{
    float Q = 10;
    float w0 = 2 * kPi * f0 / FS;
    float alpha = sinf(w0) / (2 * Q);
    c[0] = 1; c[1] = -2 * cosf(w0); c[2] = 1;
    float scaling = 1 + alpha;
    c[3] = 2 * cosf(w0); c[4] = alpha - 1;
    bq_scale(c, scaling, scaling);
}

static void design_peak(float c[5], float f0, float FS)        // AudioDriver_CalcBandpass :850
{
    float Q = 4;
    float BW = 0.03;
    float w0 = 2 * kPi * f0 / FS;
    float alpha = sinf(w0) * sinhf(log(2) / 2 * BW * w0 / sinf(w0));
    c[0] = Q * alpha; c[1] = 0; c[2] = -Q * alpha;
    float scaling = 1 + alpha;
    c[3] = 2 * cosf(w0); c[4] = alpha - 1;
    bq_scale(c, scaling, scaling);
}

static void design_highshelf(float c[5], float f0, float S, float gain, float FS)   // :906
{
    float w0 = 2 * kPi * f0 / FS;
    float A = exp10f(gain / 40.0);
    float alpha = sinf(w0) / 2 * sqrtf((A + 1 / A) * (1 / S - 1) + 2);
    float cosw0 = cosf(w0);
    float twoAa = 2 * sqrtf(A) * alpha;
    c[0] = A * ((A + 1) + (A - 1) * cosw0 + twoAa);
    c[1] = -2 * A * ((A - 1) + (A + 1) * cosw0);
    c[2] = A * ((A + 1) + (A - 1) * cosw0 - twoAa);
    float scaling = (A + 1) - (A - 1) * cosw0 + twoAa;
    c[3] = -2 * ((A - 1) - (A + 1) * cosw0);
    c[4] = twoAa - (A + 1) + (A - 1) * cosw0;
    float DCgain = 1.0 * scaling;
    bq_scale(c, scaling, DCgain);
}

static void design_lowshelf(float c[5], float f0, float S, float gain, float FS)    // :933
{
    float w0 = 2 * kPi * f0 / FS;
    float A = exp10f(gain / 40.0);
    float alpha = sinf(w0) / 2 * sqrtf((A + 1 / A) * (1 / S - 1) + 2);
    float cosw0 = cosf(w0);
    float twoAa = 2 * sqrtf(A) * alpha;
    c[0] = A * ((A + 1) - (A - 1) * cosw0 + twoAa);
    c[1] = 2 * A * ((A - 1) - (A + 1) * cosw0);
    c[2] = A * ((A + 1) - (A - 1) * cosw0 - twoAa);
    float scaling = (A + 1) + (A - 1) * cosw0 + twoAa;
    c[3] = 2 * ((A - 1) + (A + 1) * cosw0);
    c[4] = twoAa - (A + 1) - (A - 1) * cosw0;
    float DCgain = 1.0 * scaling;
    bq_scale(c, scaling, DCgain);
}

static const float kPass[5] = { 1, 0, 0, 0, 0 };

// AudioAgc_SetupAgcWdsp, audio_agc.c:126-339.  The one-time constants (:209-224) are folded in;
// the ring/state re-initialisation rule (:138-142) is applied on the device at configure time.
static void setup_agc(const uhsdr_chan_cfg_t &cfg, float sample_rate, bool remove_dc, AgcP *a)
{
    float tau_attack = 0.001;
    int n_tau = 4;
    float max_input = (float)4096;      // ADC_CLIP_WARN_THRESHOLD
    float out_targ = (float)4096;
    float tau_fast_backaverage = 0.250;
    float tau_fast_decay = 0.005;
    float pop_ratio = 5.0;
    float tau_hang_backmult = 0.500;

    a->mode = cfg.agc_mode;
    a->hang_enable = cfg.agc_hang_enable;
    a->remove_dc = remove_dc ? 1 : 0;
    a->sample_rate = sample_rate;
    a->pop_ratio = pop_ratio;
    float var_gain = exp10f((float)(uint8_t)cfg.agc_slope / 20.0 / 10.0);
    float hangtime = (float)cfg.agc_hang_time / 1000.0;
    switch (cfg.agc_mode) {          // agc_wdsp_conf.switch_mode is re-armed on every configure
    case 1: hangtime = 2.000; break;
    case 2: hangtime = 1.000; break;
    case 3: hangtime = 0.250; break;
    case 4: hangtime = 0.100; break;
    case 0: hangtime = 3.000; tau_hang_backmult = 0.500; tau_fast_decay = 0.05; tau_fast_backaverage = 0.250; break;
    default: break;
    }
    a->hangtime = hangtime;
    float tau_hang_decay = (float)cfg.agc_tau_hang_decay / 1000.0;
    int mode_idx = cfg.agc_mode < 0 ? 0 : (cfg.agc_mode > 5 ? 5 : cfg.agc_mode);
    float tau_decay = (float)cfg.agc_tau_decay[mode_idx] / 1000.0;
    float max_gain = exp10f((float)cfg.agc_thresh / 20.0);
    a->fixed_gain = max_gain / 10.0;
    a->attack_buffsize = ceilf(sample_rate * n_tau * tau_attack);

    a->attack_mult = 1.0 - expf(-1.0 / (sample_rate * tau_attack));
    a->decay_mult = 1.0 - expf(-1.0 / (sample_rate * tau_decay));
    a->fast_decay_mult = 1.0 - expf(-1.0 / (sample_rate * tau_fast_decay));
    a->fast_backmult = 1.0 - expf(-1.0 / (sample_rate * tau_fast_backaverage));
    a->onemfast_backmult = 1.0 - a->fast_backmult;

    a->out_target = out_targ * (1.0 - expf(-(float)n_tau)) * 0.9999;
    a->min_volts = a->out_target / (var_gain * max_gain);
    float tmpA = log10f(a->out_target / (max_input * var_gain * max_gain));
    if (tmpA == 0.0) tmpA = 1e-16;
    a->slope_constant = (a->out_target * (1.0 - 1.0 / var_gain)) / tmpA;
    a->inv_max_input = 1.0 / max_input;
    float hang_thresh;
    if (max_input > a->min_volts) {
        float convert = exp10f((float)cfg.agc_hang_thresh / 20.0);
        float tmpB = (convert - a->min_volts) / (max_input - a->min_volts);
        if (tmpB < 1e-8) tmpB = 1e-8;
        hang_thresh = 1.0 + 0.125 * log10f(tmpB);
    } else {
        hang_thresh = 1.0;
    }
    float tmpC = exp10f((hang_thresh - 1.0) / 0.125);
    a->hang_level = (max_input * tmpC + (a->out_target / (var_gain * max_gain)) * (1.0 - tmpC)) * 0.637;
    a->hang_backmult = 1.0 - expf(-1.0 / (sample_rate * tau_hang_backmult));
    a->onemhang_backmult = 1.0 - a->hang_backmult;
    a->hang_decay_mult = 1.0 - expf(-1.0 / (sample_rate * tau_hang_decay));
}

static int set_lattice(const HostTables &t, int idx, LatticeP *l)
{
    l->n = 0; l->k_off = 0; l->v_off = 0;
    if (idx < 0) return 0;
    if (idx >= (int)t.h->num_lattices) return -1;
    const uhsdr_tbl_lattice_t &r = t.lat[idx];
    if (r.num_stages > MAX_LAT) return -1;
    l->n = r.num_stages; l->k_off = t.off(r.k_array); l->v_off = t.off(r.v_array);
    return 0;
}

int build_chan_params(const HostTables &t, const uhsdr_chan_cfg_t &cfg, ChanParams *p, std::string *err)
{
    auto fail = [&](int code, const char *msg) { if (err) *err = msg; return code; };
    if (cfg.struct_size != sizeof(uhsdr_chan_cfg_t)) return fail(UHSDR_ERR_ARG, "uhsdr_chan_cfg_t.struct_size mismatch");
    if (cfg.filter_path < 1 || cfg.filter_path >= (int)t.h->num_paths) return fail(UHSDR_ERR_ARG, "filter_path out of range");
    if (cfg.dmod_mode < UHSDR_DEMOD_USB || cfg.dmod_mode > UHSDR_DEMOD_DIGI) return fail(UHSDR_ERR_UNSUPPORTED, "dmod_mode not implemented (SSBSTEREO/IQ are stereo-only modes)");
    if (cfg.spectrum_magnify < 0 || cfg.spectrum_magnify > 5) return fail(UHSDR_ERR_ARG, "spectrum_magnify out of range (0..5, MAGNIFY_MAX)");
    if (cfg.iq_freq_mode < 0 || cfg.iq_freq_mode > 4) return fail(UHSDR_ERR_ARG, "iq_freq_mode out of range");
    if (cfg.agc_mode < 0 || cfg.agc_mode > 5) return fail(UHSDR_ERR_ARG, "agc_mode out of range");

    memset(p, 0, sizeof(*p));
    const uhsdr_tbl_path_t &fp = t.path[cfg.filter_path];
    const int mode = cfg.dmod_mode;
    const bool is_am = (mode == UHSDR_DEMOD_AM || mode == UHSDR_DEMOD_SAM);
    // the path must be applicable to the mode (AudioFilter_IsApplicableFilterPath, audio_filter.c:973-1010)
    const int filter_mode = (mode == UHSDR_DEMOD_AM || mode == UHSDR_DEMOD_SAM) ? 2 : (mode == UHSDR_DEMOD_FM ? 3 : (mode == UHSDR_DEMOD_CW ? 0 : 1));
    if ((fp.mode_mask & (1 << filter_mode)) == 0) return fail(UHSDR_ERR_ARG, "filter_path is not applicable to dmod_mode");

    p->configured = 1;
    p->mode = mode;
    p->M = fp.sample_rate_dec;
    if (p->M != 1 && p->M != 2 && p->M != 4) return fail(UHSDR_ERR_TABLES, "unexpected decimation rate");
    p->decimated_freq = 48000 / p->M;
    p->lsb = (mode == UHSDR_DEMOD_LSB) || (mode == UHSDR_DEMOD_CW && cfg.cw_lsb) || (mode == UHSDR_DEMOD_DIGI && cfg.digi_lsb);

    p->iq_auto = cfg.iq_auto_correction ? 1 : 0;
    p->adj_i = cfg.rx_adj_gain_i; p->adj_q = cfg.rx_adj_gain_q; p->phase_bal = cfg.iq_phase_balance_rx;

    // FreqShift, freq_shift.c:275-331; NCO constants :34-47
    int shift = 0;
    switch (cfg.iq_freq_mode) {
    case UHSDR_FREQ_IQ_CONV_P6KHZ: shift = 6000; break;
    case UHSDR_FREQ_IQ_CONV_M6KHZ: shift = -6000; break;
    case UHSDR_FREQ_IQ_CONV_P12KHZ: shift = 12000; break;
    case UHSDR_FREQ_IQ_CONV_M12KHZ: shift = -12000; break;
    default: break;
    }
    p->shift_freq = abs(shift);
    p->shift_down = shift > 0;
    p->shift_kind = shift == 0 ? 0 : (p->shift_freq == 12000 ? 1 : 2);
    {
        float nco_freq = (float)p->shift_freq, sample_rate = 48000.0f;
        double rate = (2 * M_PI * nco_freq) / sample_rate;
        p->osc_cos = cos(rate); p->osc_sin = sin(rate);
    }

    // filters: audio_filter.c:1134-1223, audio_driver.c:2718-2720
    const bool use_dec_iq = (fp.fir_is_new_coeffs && mode != UHSDR_DEMOD_FM) || is_am;
    if (mode == UHSDR_DEMOD_FM) {
        if (p->M != 1) return fail(UHSDR_ERR_UNSUPPORTED, "FM on a decimating filter path");
        p->topo = TOPO_FM;
        p->s1_ntaps = fp.fir_numtaps; p->s1_ci = t.off(fp.fir_i_array); p->s1_cq = t.off(fp.fir_q_array); p->s1_M = 1;
    } else if (is_am) {
        if (fp.fir_numtaps == 0 || p->M == 1) return fail(UHSDR_ERR_UNSUPPORTED, "AM/SAM path without decimator");
        p->topo = TOPO_AM_SAM;
        p->s1_ntaps = fp.fir_numtaps; p->s1_ci = t.off(fp.fir_i_array); p->s1_cq = t.off(fp.fir_q_array); p->s1_M = p->M;
    } else if (use_dec_iq) {
        if (fp.dec_array < 0 || p->M == 1) return fail(UHSDR_ERR_UNSUPPORTED, "SSB path without decimator");
        p->topo = TOPO_SSB_DEC_FIRST;
        p->s1_ntaps = fp.dec_numtaps; p->s1_ci = p->s1_cq = t.off(fp.dec_array); p->s1_M = p->M;
        p->s2_ntaps = fp.fir_numtaps; p->s2_ci = t.off(fp.fir_i_array); p->s2_cq = t.off(fp.fir_q_array); p->s2_M = 1;
    } else {
        if (fp.dec_array < 0 || p->M == 1) return fail(UHSDR_ERR_UNSUPPORTED, "SSB path without decimator");
        p->topo = TOPO_SSB_HIL_FIRST;
        p->s1_ntaps = fp.fir_numtaps; p->s1_ci = t.off(fp.fir_i_array); p->s1_cq = t.off(fp.fir_q_array); p->s1_M = 1;
        p->s2_ntaps = fp.dec_numtaps; p->s2_ci = p->s2_cq = t.off(fp.dec_array); p->s2_M = p->M;
    }
    if (p->s1_ntaps - 1 > H1 || p->s2_ntaps - 1 > H2 || p->s1_ntaps < 1) return fail(UHSDR_ERR_TABLES, "FIR longer than the engine's history slots");

    if (set_lattice(t, fp.pre_lattice, &p->pre) || set_lattice(t, fp.aa_lattice, &p->aa) ||
        set_lattice(t, t.ex->fm_squelch_lattice, &p->sql))
        return fail(UHSDR_ERR_TABLES, "lattice filter with too many stages");

    // interpolator, audio_driver.c:1209-1224 (+ arm_fir_interpolate_init_f32.c:85-100)
    if (fp.interpolate >= 0) {
        const uhsdr_tbl_interp_t &ip = t.interp[fp.interpolate];
        p->interp_L = p->M;
        p->interp_plen = ip.phase_length_field / p->M;
        p->interp_c = t.off(ip.coeff_array);
        if (p->interp_plen < 1 || p->interp_plen - 1 > INTERP_HIST) return fail(UHSDR_ERR_TABLES, "interpolator phase length out of range");
    } else if (mode != UHSDR_DEMOD_FM) {
        return fail(UHSDR_ERR_UNSUPPORTED, "non-FM path without interpolator");
    }

    // biquad EQ, audio_driver.c:994-1047
    float FSdec = 48000 / (fp.sample_rate_dec != 0 ? fp.sample_rate_dec : 1);
    float co[5];
    if (cfg.dsp_active & UHSDR_DSP_MNOTCH_ENABLE) { design_notch(co, (float)(unsigned long)cfg.notch_frequency, FSdec); memcpy(p->bq1[0], co, sizeof(co)); }
    else memcpy(p->bq1[0], kPass, sizeof(kPass));
    if (cfg.dsp_active & UHSDR_DSP_MPEAK_ENABLE) { design_peak(co, (float)(unsigned long)cfg.peak_frequency, FSdec); memcpy(p->bq1[1], co, sizeof(co)); }
    else memcpy(p->bq1[1], kPass, sizeof(kPass));
    design_lowshelf(co, 250, 0.7, cfg.bass_gain, FSdec);
    memcpy(p->bq1[2], co, sizeof(co));
    memcpy(p->bq1[3], kPass, sizeof(kPass));
    design_highshelf(co, 3500, 0.9, cfg.treble_gain, 48000);
    memcpy(p->bq2, co, sizeof(co));

    // fixed post-AGC gain, audio_driver.c:2513-2521
    {
        const float post_agc_gain_scaling = (fp.sample_rate_dec == 4) ? 3.46 : (3.46 * 0.6);
        const float scale_gain = post_agc_gain_scaling * (is_am ? 0.5 : 0.333);
        p->scale_gain = scale_gain;
    }
    setup_agc(cfg, (float)p->decimated_freq, is_am, &p->agc);

    // SAM PLL + fade leveler constants, audio_driver.c:709-745
    {
        const float decimSampleRate = p->decimated_freq;
        const float pll_fmax = cfg.sam_pll_fmax;
        float omegaN = cfg.sam_omegaN;
        float zeta = (float)cfg.sam_zeta / 100.0;
        p->sam_omega_min = -(2.0 * kPi * pll_fmax / decimSampleRate);
        p->sam_omega_max = (2.0 * kPi * pll_fmax / decimSampleRate);
        p->sam_g1 = (1.0 - expf(-2.0 * omegaN * zeta / decimSampleRate));
        p->sam_g2 = (-p->sam_g1 + 2.0 * (1 - expf(-omegaN * zeta / decimSampleRate) * cosf(omegaN / decimSampleRate * sqrtf(1.0 - zeta * zeta))));
        float tauR = 0.02;
        float tauI = 1.4;
        p->sam_mtauR = (expf(-1 / (decimSampleRate * tauR)));
        p->sam_onem_mtauR = (1.0 - p->sam_mtauR);
        p->sam_mtauI = (expf(-1 / (decimSampleRate * tauI)));
        p->sam_onem_mtauI = (1.0 - p->sam_mtauI);
        p->sam_sideband = cfg.sam_sideband;
        p->fade_leveler = cfg.sam_fade_leveler ? 1 : 0;
        p->sam_c0 = t.off(t.ex->sam_c0_array);
        p->sam_c1 = t.off(t.ex->sam_c1_array);
    }
    p->fm_sql_threshold = (uint8_t)cfg.fm_sql_threshold;
    p->fm_tone_det = (mode == UHSDR_DEMOD_FM && cfg.fm_subaudible_tone_det_freq > 0.0f) ? 1 : 0;
    if (p->fm_tone_det) {
        // AudioManagement_CalcSubaudibleDetFreq + AudioFilter_CalcGoertzel, audio_management.c:311-326, audio_filter.c:1281-1288
        const float ratio[3] = { 1.04, 0.95, 1.0 };        // FM_HIGH, FM_LOW, FM_CTR
        const uint32_t size = 400 * 32;                      // FM_SUBAUDIBLE_GOERTZEL_WINDOW * AUDIO_BLOCK_SIZE
        const float freq = cfg.fm_subaudible_tone_det_freq, samplerate = 48000;
        for (int k = 0; k < 3; k++) {
            const float ga = (0.5 + (freq * ratio[k]) * size / samplerate);
            const float gb = (2 * 3.14159265358979f * ga) / size;
            p->fm_gz_sin[k] = sinf(gb); p->fm_gz_cos[k] = cosf(gb); p->fm_gz_r[k] = 2 * p->fm_gz_cos[k];
        }
    }
    p->fm_scaling = cfg.fm_dev_5khz ? (10000 / 2) : 10000;   // FM_RX_SCALING_5K / _2K5, audio_driver.c:1494-1495
    p->fm_translate_on = cfg.iq_freq_mode != UHSDR_FREQ_IQ_CONV_OFF;

    // spectral NR, audio_driver.c:1195, :2355, :2501; audio_nr.c:1857-1867, :2034-2059
    // audio_driver.c:2443-2456 (FM never reaches RxProcessor_DemodAudioPostprocessing's notch: 48 ksps, no decimation)
    p->notch_enable = ((cfg.dsp_active & UHSDR_DSP_NOTCH_ENABLE) && mode != UHSDR_DEMOD_CW && mode != UHSDR_DEMOD_FM &&
                       !(mode == UHSDR_DEMOD_SAM && p->decimated_freq == 24000)) ? 1 : 0;
    p->notch_mu = log10f(((cfg.notch_mu + 1.0) / 1500.0) + 1.0);        // :1170
    const bool nb_active = (cfg.dsp_active & UHSDR_DSP_NB_ENABLE) && cfg.nb_setting > 0;          // is_dsp_nb_active, ui_driver.c:405-408
    const bool nr_active = (cfg.dsp_active & UHSDR_DSP_NR_ENABLE) != 0;
    p->nr_enable = (p->decimated_freq == 12000 && (nr_active || nb_active) && mode != UHSDR_DEMOD_FM) ? 1 : 0;
    p->nr_spectral = p->nr_enable && nr_active;
    p->nb_enable = p->nr_enable && nb_active;
    p->nb_level = 16 - (int)(uint8_t)cfg.nb_setting;
    {
        const int width_i = t.filt[fp.id].width;
        p->nr_decim = (cfg.nr_decimation_enable && width_i < 2701) ? 1 : 0;
        p->nr_alpha = 0.799 + ((float)(uint8_t)cfg.nr_strength / 1000.0);
        const float width = width_i;
        const float offset = fp.offset_hz;
        float NR_sample_rate = p->nr_decim ? 6000.0 : 12000.0;
        const uint16_t NR_FFT_L = 256;
        float lf_freq = (offset - width / 2) / (NR_sample_rate / NR_FFT_L);
        float uf_freq = (offset + width / 2) / (NR_sample_rate / NR_FFT_L);
        uint8_t VAD_low = (int)lf_freq;
        uint8_t VAD_high = (int)uf_freq;
        if (VAD_low == VAD_high) VAD_high++;
        if (VAD_low < 1) VAD_low = 1; else if (VAD_low > NR_FFT_L / 2 - 2) VAD_low = NR_FFT_L / 2 - 2;
        if (VAD_high < 1) VAD_high = 1; else if (VAD_high > NR_FFT_L / 2) VAD_high = NR_FFT_L / 2;
        p->nr_vad_low = VAD_low; p->nr_vad_high = VAD_high;
        p->nr_dec_c = t.off(t.ex->nr_decimate_array);
        p->nr_int_c = t.off(t.ex->nr_interpolate_array);
        p->nr_win_c = t.off(t.ex->sqrt_hann_256_array);
        p->tw256_off = t.tw256_off; p->tw512_off = t.tw512_off;
        p->nr_xih1 = exp10f((float)30 / 10.0);     // NR2.asnr = 30 (audio_nr.c:95, :1886)
    }
    p->spectrum_enable = cfg.spectrum_enable ? 1 : 0;
    p->zoom_m = cfg.spectrum_enable ? cfg.spectrum_magnify : 0;
    if (p->zoom_m) {                                       // AudioDriver_Spectrum_Set, audio_driver.c:1055-1086
        const int zb = t.off(t.ex->zoom_biquad_array), zd = t.off(t.ex->zoom_decim_array);
        if (zb < 0 || zd < 0) return fail(UHSDR_ERR_TABLES, "zoom-FFT filters missing from the table blob");
        p->zoom_bq_off = zb + (p->zoom_m - 1) * 20;
        p->zoom_dec_off = zd + (p->zoom_m - 1) * 4;
    }
    p->codec_gain_calc = cfg.codec_gain_calc;
    {
        // UiSpectrum_CalculateDBm, ui_spectrum.c:2004-2081, for fft_iq_len = 1024 / spec_len = 512 (480x320 layout)
        const int buff_len_int = 1024;
        const float buff_len = buff_len_int;
        const float bin_BW = 48000.0f * 2.0 / (buff_len * (1 << p->zoom_m));
        float width = t.filt[fp.id].width;
        float offset = fp.offset_hz;
        if (offset == 0) offset = width / 2;
        const float lf_freq = offset - width / 2;
        const float uf_freq = offset + width / 2;
        float bw_LOWER = 0.0, bw_UPPER = 0.0;
        const bool both = mode == UHSDR_DEMOD_AM || (mode == UHSDR_DEMOD_SAM && cfg.sam_sideband == UHSDR_SAM_SIDEBAND_BOTH) || mode == UHSDR_DEMOD_FM;
        const bool lsb_active = (mode == UHSDR_DEMOD_SAM) ? (cfg.sam_sideband == UHSDR_SAM_SIDEBAND_LSB) : (p->lsb != 0);
        if (both) { bw_UPPER = uf_freq; bw_LOWER = -uf_freq; }
        else if (lsb_active) { bw_UPPER = -lf_freq; bw_LOWER = -uf_freq; }
        else { bw_UPPER = uf_freq; bw_LOWER = lf_freq; }
        int translate = 0;                                  // AudioDriver_GetTranslateFreq, audio_driver.c:445-464
        switch (cfg.iq_freq_mode) {
        case UHSDR_FREQ_IQ_CONV_P6KHZ: translate = 6000; break;
        case UHSDR_FREQ_IQ_CONV_M6KHZ: translate = -6000; break;
        case UHSDR_FREQ_IQ_CONV_P12KHZ: translate = 12000; break;
        case UHSDR_FREQ_IQ_CONV_M12KHZ: translate = -12000; break;
        default: break;
        }
        const int32_t bin_offset = p->zoom_m != 0 ? 0 : (-(buff_len_int * translate) / (2 * 48000));
        const int32_t posbin = buff_len_int / 4 + bin_offset;
        float Lbin = (float)posbin + roundf(bw_LOWER / bin_BW);
        float Ubin = (float)posbin + roundf(bw_UPPER / bin_BW);
        if (mode == UHSDR_DEMOD_SAM && cfg.sam_sideband == UHSDR_SAM_SIDEBAND_USB) Lbin = Lbin - 1.0;
        if (Lbin < 0) Lbin = 0;
        if (Ubin > (512 - 1)) Ubin = 512 - 1;
        p->dbm_lbin = (int)Lbin; p->dbm_ubin = (int)Ubin;
        p->dbm_span_hz = (float)(((int)Ubin - (int)Lbin) * bin_BW);
    }
    return UHSDR_OK;
}

}  // namespace uhsdr
