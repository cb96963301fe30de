/*
 * uhsdr_port.c -- ORACLE / TEST INFRASTRUCTURE ONLY (see uhsdr_port.h).
 *
 * CPU restatement of the reference's RX/TX block path, one explicit state struct per channel.
 * Every function cites the reference lines it follows (paths under /root/reference/mchf-eclipse;
 * CMSIS = basesw/ovi40/Drivers/CMSIS/DSP_Lib/Source).  Arithmetic keeps the reference's operation
 * order and its float/double promotions (the firmware is built without
 * -fsingle-precision-constant), and this file is compiled with -ffp-contract=off, so that the
 * port is bit-identical to oracle/_ref on the same inputs (tests/test_oracle_pin.py).
 *
 * Parity pin: oracle/_ref (the reference's own object code) + tests/golden/ vectors generated
 * from it by tests/golden/make_golden.py.
 */
#include "uhsdr_port.h"
#include "uhsdr_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

float exp10f(float);

#define BLK 32
#define PI_F 3.14159265358979f /* CMSIS Include/arm_math.h: #define PI 3.14159265358979f */
#define AGC_RB 192             /* AGC_WDSP_RB_SIZE, audio_agc.c:19 */
#define MAX_LAT 12
#define NR_FFT 256
#define NR_HALF 128

/* ------------------------------------------------------------------------------------------ */
/* tables                                                                                     */
/* ------------------------------------------------------------------------------------------ */
struct port_tables {
    uint8_t *blob;
    const uhsdr_tbl_header_t *h;
    const uhsdr_tbl_array_t *arr;
    const uhsdr_tbl_path_t *path;
    const uhsdr_tbl_filter_t *filt;
    const uhsdr_tbl_lattice_t *lat;
    const uhsdr_tbl_interp_t *interp;
    const uhsdr_tbl_extras_t *ex;
};

static const float *tbl_array(const port_tables_t *t, int idx, int *count)
{
    if (idx < 0 || (uint32_t)idx >= t->h->num_arrays) { if (count) *count = 0; return NULL; }
    if (count) *count = (int)t->arr[idx].count;
    return (const float *)(t->blob + t->arr[idx].offset);
}

port_tables_t *port_tables_load(const void *blob, size_t bytes)
{
    if (!blob || bytes < sizeof(uhsdr_tbl_header_t)) return NULL;
    const uhsdr_tbl_header_t *h = (const uhsdr_tbl_header_t *)blob;
    if (h->magic != UHSDR_TABLES_MAGIC || h->version != UHSDR_TABLES_VERSION || h->total_bytes != bytes) return NULL;
    port_tables_t *t = (port_tables_t *)calloc(1, sizeof(*t));
    t->blob = (uint8_t *)malloc(bytes);
    memcpy(t->blob, blob, bytes);
    t->h = (const uhsdr_tbl_header_t *)t->blob;
    t->arr = (const uhsdr_tbl_array_t *)(t->blob + h->arrays_off);
    t->path = (const uhsdr_tbl_path_t *)(t->blob + h->paths_off);
    t->filt = (const uhsdr_tbl_filter_t *)(t->blob + h->filters_off);
    t->lat = (const uhsdr_tbl_lattice_t *)(t->blob + h->lattices_off);
    t->interp = (const uhsdr_tbl_interp_t *)(t->blob + h->interps_off);
    t->ex = (const uhsdr_tbl_extras_t *)(t->blob + h->extras_off);
    return t;
}

void port_tables_free(port_tables_t *t)
{
    if (t) { free(t->blob); free(t); }
}

/* ------------------------------------------------------------------------------------------ */
/* primitive filters (CMSIS portable-C branches)                                              */
/* ------------------------------------------------------------------------------------------ */

/* arm_fir_f32, CMSIS FilteringFunctions/arm_fir_f32.c:482-559: y[n] = sum_k c[k]*x[n-(N-1)+k],
 * k ascending, one float accumulator; hist holds the N-1 previous inputs, oldest first. */
typedef struct { int ntaps; const float *c; float hist[200 + BLK]; } fir_t;

static void fir_reset(fir_t *f, int ntaps, const float *c) { memset(f, 0, sizeof(*f)); f->ntaps = ntaps; f->c = c; }

static void fir_run(fir_t *f, const float *in, float *out, int n)
{
    const int N = f->ntaps;
    float *st = f->hist;
    for (int i = 0; i < n; i++) st[N - 1 + i] = in[i];
    for (int i = 0; i < n; i++) {
        float acc = 0.0f;
        for (int k = 0; k < N; k++) acc += st[i + k] * f->c[k];
        out[i] = acc;
    }
    memmove(st, st + n, sizeof(float) * (size_t)(N - 1));
}

/* arm_fir_decimate_f32, arm_fir_decimate_f32.c:440-516: M new inputs are appended, then one
 * output = sum_k x[k]*c[k] over the newest N samples (k ascending from the oldest). */
typedef struct { int ntaps, M; const float *c; float hist[200 + BLK]; } firdec_t;

static void firdec_reset(firdec_t *f, int ntaps, int M, const float *c) { memset(f, 0, sizeof(*f)); f->ntaps = ntaps; f->M = M; f->c = c; }

static void firdec_run(firdec_t *f, const float *in, float *out, int n_in)
{
    const int N = f->ntaps, M = f->M, nout = n_in / M;
    float *st = f->hist;
    for (int i = 0; i < n_in; i++) st[N - 1 + i] = in[i];
    for (int m = 0; m < nout; m++) {
        const float *x = st + m * M;
        float acc = 0.0f;
        for (int k = 0; k < N; k++) acc += x[k] * f->c[k];
        out[m] = acc;
    }
    memmove(st, st + nout * M, sizeof(float) * (size_t)(N - 1));
}

/* arm_fir_interpolate_f32, arm_fir_interpolate_f32.c:482-573, with phaseLength = numTaps / L
 * (arm_fir_interpolate_init_f32.c:85-100).  Output j of each input uses taps c[(L-1-j) + k*L]. */
typedef struct { int L, plen; const float *c; float hist[64 + BLK]; } firint_t;

static void firint_reset(firint_t *f, int L, int numtaps, const float *c) { memset(f, 0, sizeof(*f)); f->L = L; f->plen = numtaps / L; f->c = c; }

static void firint_run(firint_t *f, const float *in, float *out, int n_in)
{
    const int L = f->L, P = f->plen;
    float *st = f->hist;
    for (int i = 0; i < n_in; i++) {
        st[P - 1 + i] = in[i];
        for (int j = 0; j < L; j++) {
            float sum = 0.0f;
            const float *pc = f->c + (L - 1 - j);
            for (int k = 0; k < P; k++) sum += st[i + k] * pc[k * L];
            *out++ = sum;
        }
    }
    memmove(st, st + n_in, sizeof(float) * (size_t)(P - 1));
}

/* arm_iir_lattice_f32, arm_iir_lattice_f32.c:348-440. s[j] is the g-state of stage j. */
typedef struct { int n; const float *k, *v; float s[MAX_LAT + 1]; } lattice_t;

static void lattice_reset(lattice_t *l, int n, const float *k, const float *v) { memset(l, 0, sizeof(*l)); l->n = n; l->k = k; l->v = v; }

static void lattice_run(lattice_t *l, float *buf, int n)
{
    const int N = l->n;
    for (int i = 0; i < n; i++) {
        float f = buf[i], acc = 0.0f, fn = 0.0f;
        float w[MAX_LAT + 1];
        for (int j = 0; j < N; j++) {
            float g = l->s[j];
            fn = f - (l->k[j] * g);
            float gn = (fn * l->k[j]) + g;
            acc += gn * l->v[j];
            w[j] = gn;
            f = fn;
        }
        acc += fn * l->v[N];
        w[N] = fn;
        for (int j = 0; j < N; j++) l->s[j] = w[j + 1];
        buf[i] = acc;
    }
}

/* arm_biquad_cascade_df1_f32, arm_biquad_cascade_df1_f32.c:349-418 (a1, a2 already negated). */
typedef struct { float c[5]; float x1, x2, y1, y2; } biquad_t;

static void biquad_run(biquad_t *b, float *buf, int n)
{
    for (int i = 0; i < n; i++) {
        float x = buf[i];
        float acc = (b->c[0] * x) + (b->c[1] * b->x1) + (b->c[2] * b->x2) + (b->c[3] * b->y1) + (b->c[4] * b->y2);
        b->x2 = b->x1; b->x1 = x; b->y2 = b->y1; b->y1 = acc;
        buf[i] = acc;
    }
}

static const float BQ_PASS[5] = { 1, 0, 0, 0, 0 };

/* audio_driver.c:818-826 */
static void bq_scale(float c[5], float sa, float sb) { c[3] = c[3] / sa; c[4] = c[4] / sa; c[0] = c[0] / sb; c[1] = c[1] / sb; c[2] = c[2] / sb; }

/* AudioDriver_CalcBandstop, audio_driver.c:831-845 */
static void bq_bandstop(float c[5], float f0, float FS)
{
    float Q = 10;
    float w0 = 2 * PI_F * f0 / FS;
    float alpha = sinf(w0) / (2 * Q);
    c[0] = 1; c[1] = -2 * cosf(w0); c[2] = 1;
    float scaling = 1 + alpha;
    c[3] = 2 * cosf(w0); c[4] = alpha - 1;
    bq_scale(c, scaling, scaling);
}

/* AudioDriver_CalcBandpass, audio_driver.c:850-901 (log(2) is a double expression) */
static void bq_bandpass(float c[5], float f0, float FS)
{
    float Q = 4;
    float BW = 0.03;
    float w0 = 2 * PI_F * f0 / FS;
    float alpha = sinf(w0) * sinhf(log(2) / 2 * BW * w0 / sinf(w0));
    c[0] = Q * alpha; c[1] = 0; c[2] = -Q * alpha;
    float scaling = 1 + alpha;
    c[3] = 2 * cosf(w0); c[4] = alpha - 1;
    bq_scale(c, scaling, scaling);
}

/* AudioDriver_CalcHighShelf, audio_driver.c:906-928 */
static void bq_highshelf(float c[5], float f0, float S, float gain, float FS)
{
    float w0 = 2 * PI_F * f0 / FS;
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

/* AudioDriver_CalcLowShelf, audio_driver.c:933-964 */
static void bq_lowshelf(float c[5], float f0, float S, float gain, float FS)
{
    float w0 = 2 * PI_F * f0 / FS;
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

/* Math_log10f_fast, misc/uhsdr_math.c:27-41 */
static float log10f_fast(float X)
{
    float Y, F;
    int E;
    F = frexpf(fabsf(X), &E);
    Y = 1.23149591368684f;
    Y *= F; Y += -4.11852516267426f;
    Y *= F; Y += 6.02197014179219f;
    Y *= F; Y += -3.13396450166353f;
    Y += E;
    return Y * 0.3010299956639812f;
}

/* Math_sign_new, misc/uhsdr_math.c:65-67 */
static float sign_new(float x) { return (x < 0) ? -1.0 : ((x > 0) ? 1.0 : 0.0); }

/* ------------------------------------------------------------------------------------------ */
/* WDSP AGC, audio_agc.c                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    /* configuration-derived (AudioAgc_SetupAgcWdsp, audio_agc.c:126-339) */
    int mode, hang_enable;
    float tau_attack, tau_decay, max_gain, var_gain, fixed_gain, max_input, out_targ;
    float tau_fast_backaverage, tau_fast_decay, pop_ratio, tau_hang_backmult, hangtime, hang_thresh, tau_hang_decay;
    int n_tau, attack_buffsize, ring_buffsize;
    float attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult, out_target, min_volts;
    float inv_out_target, slope_constant, inv_max_input, hang_level, hang_backmult, onemhang_backmult, hang_decay_mult;
    int remove_dc;
    float sample_rate;
    int initialised;
    /* running state */
    float ring[AGC_RB], abs_ring[AGC_RB];
    int out_index; uint32_t in_index;
    float ring_max, volts, save_volts, fast_backaverage, hang_backaverage;
    int hang_counter, decay_type, state;
    float wold;          /* DC remover, audio_agc.c:579 */
    int action, hang_action;
} agc_t;

static void agc_setup(agc_t *a, const uhsdr_chan_cfg_t *cfg, float sample_rate, int remove_dc)
{
    a->remove_dc = remove_dc;
    a->mode = cfg->agc_mode;
    a->hang_enable = cfg->agc_hang_enable;
    if (a->sample_rate != sample_rate) { a->initialised = 0; a->sample_rate = sample_rate; }
    if (!a->initialised) {
        a->ring_buffsize = AGC_RB;
        a->out_index = -1;
        a->fixed_gain = 1.0;
        a->ring_max = 0.0; a->volts = 0.0; a->save_volts = 0.0;
        a->fast_backaverage = 0.0; a->hang_backaverage = 0.0;
        a->hang_counter = 0; a->decay_type = 0; a->state = 0;
        memset(a->ring, 0, sizeof(a->ring)); memset(a->abs_ring, 0, sizeof(a->abs_ring));
        a->tau_attack = 0.001;
        a->n_tau = 4;
        a->max_input = (float)4096;   /* ADC_CLIP_WARN_THRESHOLD, audio_driver.h:81 */
        a->out_targ = (float)4096;
        a->tau_fast_backaverage = 0.250;
        a->tau_fast_decay = 0.005;
        a->pop_ratio = 5.0;
        a->tau_hang_backmult = 0.500;
        a->initialised = 1;
    }
    a->var_gain = exp10f((float)cfg->agc_slope / 20.0 / 10.0);
    a->hangtime = (float)cfg->agc_hang_time / 1000.0;
    /* switch_mode: the host re-arms it on every configuration (oracle/ref_harness.c) */
    switch (cfg->agc_mode) {
    case 1: a->hangtime = 2.000; break;
    case 2: a->hangtime = 1.000; break;
    case 3: a->hangtime = 0.250; break;
    case 4: a->hangtime = 0.100; break;
    case 0:
        a->hangtime = 3.000; a->tau_hang_backmult = 0.500; a->tau_fast_decay = 0.05; a->tau_fast_backaverage = 0.250;
        break;
    default: break;
    }
    a->tau_hang_decay = (float)cfg->agc_tau_hang_decay / 1000.0;
    a->tau_decay = (float)cfg->agc_tau_decay[cfg->agc_mode] / 1000.0;
    a->max_gain = exp10f((float)cfg->agc_thresh / 20.0);
    a->fixed_gain = a->max_gain / 10.0;
    a->attack_buffsize = ceilf(sample_rate * a->n_tau * a->tau_attack);
    a->in_index = a->attack_buffsize + a->out_index;
    a->in_index %= a->ring_buffsize;

    a->attack_mult = 1.0 - expf(-1.0 / (sample_rate * a->tau_attack));
    a->decay_mult = 1.0 - expf(-1.0 / (sample_rate * a->tau_decay));
    a->fast_decay_mult = 1.0 - expf(-1.0 / (sample_rate * a->tau_fast_decay));
    a->fast_backmult = 1.0 - expf(-1.0 / (sample_rate * a->tau_fast_backaverage));
    a->onemfast_backmult = 1.0 - a->fast_backmult;

    a->out_target = a->out_targ * (1.0 - expf(-(float)a->n_tau)) * 0.9999;
    a->min_volts = a->out_target / (a->var_gain * a->max_gain);
    a->inv_out_target = 1.0 / a->out_target;

    float tmpA = log10f(a->out_target / (a->max_input * a->var_gain * a->max_gain));
    if (tmpA == 0.0) tmpA = 1e-16;
    a->slope_constant = (a->out_target * (1.0 - 1.0 / a->var_gain)) / tmpA;
    a->inv_max_input = 1.0 / a->max_input;

    if (a->max_input > a->min_volts) {
        float convert = exp10f((float)cfg->agc_hang_thresh / 20.0);
        float tmpB = (convert - a->min_volts) / (a->max_input - a->min_volts);
        if (tmpB < 1e-8) tmpB = 1e-8;
        a->hang_thresh = 1.0 + 0.125 * log10f(tmpB);
    } else {
        a->hang_thresh = 1.0;
    }
    float tmpC = exp10f((a->hang_thresh - 1.0) / 0.125);
    a->hang_level = (a->max_input * tmpC + (a->out_target / (a->var_gain * a->max_gain)) * (1.0 - tmpC)) * 0.637;
    a->hang_backmult = 1.0 - expf(-1.0 / (sample_rate * a->tau_hang_backmult));
    a->onemhang_backmult = 1.0 - a->hang_backmult;
    a->hang_decay_mult = 1.0 - expf(-1.0 / (sample_rate * a->tau_hang_decay));
}

/* AudioAgc_RunAgcWdsp (mono), audio_agc.c:349-595 */
static void agc_run(agc_t *a, float *buf, int n)
{
    if (a->mode == 5) {
        for (int i = 0; i < n; i++) buf[i] = buf[i] * a->fixed_gain;
        return;
    }
    for (int i = 0; i < n; i++) {
        if (++a->out_index >= a->ring_buffsize) a->out_index -= a->ring_buffsize;
        if (++a->in_index >= (uint32_t)a->ring_buffsize) a->in_index -= a->ring_buffsize;
        float out_sample = a->ring[a->out_index];
        float abs_out_sample = a->abs_ring[a->out_index];
        a->ring[a->in_index] = buf[i];
        a->abs_ring[a->in_index] = fabsf(buf[i]);

        a->fast_backaverage = a->fast_backmult * abs_out_sample + a->onemfast_backmult * a->fast_backaverage;
        a->hang_backaverage = a->hang_backmult * abs_out_sample + a->onemhang_backmult * a->hang_backaverage;
        a->hang_action = (a->hang_backaverage > a->hang_level) ? 1 : 0;

        if ((abs_out_sample >= a->ring_max) && (abs_out_sample > 0.0)) {
            a->ring_max = 0.0;
            int k = a->out_index;
            for (int j = 0; j < a->attack_buffsize; j++) {
                if (++k == a->ring_buffsize) k = 0;
                if (a->abs_ring[k] > a->ring_max) a->ring_max = a->abs_ring[k];
            }
        }
        if (a->abs_ring[a->in_index] > a->ring_max) a->ring_max = a->abs_ring[a->in_index];
        if (a->hang_counter > 0) --a->hang_counter;

        switch (a->state) {
        case 0:
            if (a->ring_max >= a->volts) {
                a->volts += (a->ring_max - a->volts) * a->attack_mult;
            } else if (a->volts > a->pop_ratio * a->fast_backaverage) {
                a->state = 1;
                a->volts += (a->ring_max - a->volts) * a->fast_decay_mult;
            } else if (a->hang_enable && (a->hang_backaverage > a->hang_level)) {
                a->state = 2;
                a->hang_counter = (int)(a->hangtime * a->sample_rate);
                a->decay_type = 1;
            } else {
                a->state = 3;
                a->volts += (a->ring_max - a->volts) * a->decay_mult;
                a->decay_type = 0;
            }
            break;
        case 1:
            if (a->ring_max >= a->volts) {
                a->state = 0;
                a->volts += (a->ring_max - a->volts) * a->attack_mult;
            } else if (a->volts > a->save_volts) {
                a->volts += (a->ring_max - a->volts) * a->fast_decay_mult;
            } else if (a->hang_counter > 0) {
                a->state = 2;
            } else if (a->decay_type == 0) {
                a->state = 3;
                a->volts += (a->ring_max - a->volts) * a->decay_mult;
            } else {
                a->state = 4;
                a->volts += (a->ring_max - a->volts) * a->hang_decay_mult;
            }
            break;
        case 2:
            if (a->ring_max >= a->volts) {
                a->state = 0;
                a->save_volts = a->volts;
                a->volts += (a->ring_max - a->volts) * a->attack_mult;
            } else if (a->hang_counter == 0) {
                a->state = 4;
                a->volts += (a->ring_max - a->volts) * a->hang_decay_mult;
            }
            break;
        case 3:
            if (a->ring_max >= a->volts) {
                a->state = 0;
                a->save_volts = a->volts;
                a->volts += (a->ring_max - a->volts) * a->attack_mult;
            } else {
                a->volts += (a->ring_max - a->volts) * a->decay_mult;
            }
            break;
        case 4:
            if (a->ring_max >= a->volts) {
                a->state = 0;
                a->save_volts = a->volts;
                a->volts += (a->ring_max - a->volts) * a->attack_mult;
            } else {
                a->volts += (a->ring_max - a->volts) * a->hang_decay_mult;
            }
            break;
        }
        if (a->volts < a->min_volts) { a->volts = a->min_volts; a->action = 0; } else { a->action = 1; }

        float vo = log10f_fast(a->inv_max_input * a->volts);
        if (vo > 0.0) vo = 0.0;
        float mult = (a->out_target - a->slope_constant * vo) / a->volts;
        buf[i] = out_sample * mult;
    }
    if (a->remove_dc) {
        for (int i = 0; i < n; i++) {
            float w = buf[i] + a->wold * 0.9999;
            buf[i] = w - a->wold;
            a->wold = w;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* spectral noise reduction state (audio_nr.c), filled in below                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    /* ISR-side packing, audio_driver.c:2328-2434; mmb.nr_audio_buff[k] (freedv_uhsdr.h:44-55) folded
     * to 256 floats: [0..127] packed input half, [128..255] processed output half */
    float in_buf[4][NR_FFT];
    int trans_count_in, outbuff_count, fill_in_pt;
    int out_buffer;                  /* index of buffer being drained, -1 = none */
    int in_fifo[5], in_head, in_tail;   /* NR_in_buffer FIFO, audio_nr.c:174-236 */
    int out_fifo[5], out_head, out_tail;
    int current_buffer_idx, was_here;   /* NR.current_buffer_idx / NR.was_here, audio_nr.c:314-349 */
    firdec_t dec;                    /* DECIMATE_NR */
    firint_t interp;                 /* INTERPOLATE_NR */
    /* spectral_noise_reduction_3 state, audio_nr.c:1841-2195 */
    float last_sample_buffer_L[NR_HALF];
    float last_iFFT_result[NR_HALF];
    float Hk[NR_HALF], Hk_old[NR_HALF], Nest[NR_HALF][2], xt[NR_HALF], pslp[NR_HALF], X[NR_HALF][2];
    float SNR_prio[NR_HALF], SNR_post[NR_HALF];
    int first_time, nr_first_time_count;
    float alpha;
    int decimation_active;
    float nb_work[128 + 26];                  /* alt_noise_blanking's static working_buffer, audio_nr.c:2282 */
} nr_t;

/* adb.a_buffer[0..1], audio_driver.h:147-157: persists across blocks */
/* LMS automatic notch: AudioDriver_NotchFilter (audio_driver.c:1746-1763) around arm_lms_norm_f32
 * (CMSIS FilteringFunctions/arm_lms_norm_f32.c:372-443, the portable loop); setup audio_driver.c:1165-1187
 * with arm_lms_norm_init_f32.  64 taps, 128-sample decorrelation delay (audio_driver.h:486-492). */
#define NOTCH_TAPS 64
#define NOTCH_DELAY 128
typedef struct {
    float mu, energy, x0;
    float coef[NOTCH_TAPS];
    float state[NOTCH_TAPS + BLK];
    float delay[NOTCH_DELAY];
    unsigned long inbuf, outbuf;              /* function statics lms2_inbuf / lms2_outbuf, never reset */
} notch_t;

static void notch_run(notch_t *n, float *buf, int bs)
{
    memcpy(&n->delay[n->inbuf], buf, (size_t)bs * sizeof(float));
    const float *ref = &n->delay[n->outbuf];
    /* arm_lms_norm_f32(S, pSrc = buf, pRef = ref, pOut = discarded, pErr = buf): the notched audio is the error */
    float *pState = n->state, *cur = &n->state[NOTCH_TAPS - 1];
    float energy = n->energy, x0 = n->x0;
    for (int i = 0; i < bs; i++) {
        *cur++ = buf[i];
        const float in = buf[i];
        energy -= x0 * x0;
        energy += in * in;
        float sum = 0.0f;
        for (int k = 0; k < NOTCH_TAPS; k++) sum += pState[k] * n->coef[k];
        const float d = ref[i];
        const float e = d - sum;
        buf[i] = e;
        const float w = (e * n->mu) / (energy + 0.000000119209289f);
        for (int k = 0; k < NOTCH_TAPS; k++) n->coef[k] += w * pState[k];
        x0 = *pState;
        pState++;
    }
    n->energy = energy; n->x0 = x0;
    memmove(n->state, pState, (NOTCH_TAPS - 1) * sizeof(float));
    n->inbuf += bs;
    n->outbuf = n->inbuf + bs;
    n->inbuf %= NOTCH_DELAY;
    n->outbuf %= NOTCH_DELAY;
}

typedef struct { float a0[BLK], a1[BLK]; } rx_scratch_t;

/* ------------------------------------------------------------------------------------------ */
/* channel                                                                                    */
/* ------------------------------------------------------------------------------------------ */
struct port_chan {
    rx_scratch_t scratch;
    const port_tables_t *t;
    uhsdr_chan_cfg_t cfg;
    const uhsdr_tbl_path_t *path;
    int decimation_rate, decimated_freq;
    int use_decimated_iq;
    /* IQ correction, audio_driver.h:125-135 */
    float teta1, teta2, teta3, teta1_old, teta2_old, teta3_old, M_c1, M_c2;
    int tw_state; uint32_t tw_counter, tw_runs, tw_restarts; float tw_phase;   /* ts.twinpeaks_tested + statics of AudioDriver_RxHandleTwinpeaks */
    /* FreqShift_Approx NCO, freq_shift.c:20-47 */
    int conversion_freq;
    float osc_cos, osc_sin, osc_vect_q, osc_vect_i;
    /* filters */
    firdec_t dec_i, dec_q;
    int dec_active;
    fir_t hil_i, hil_q;
    lattice_t pre, aa, sql_hpf;
    biquad_t bq1[4], bq2;
    firint_t interp;
    int interp_active;
    agc_t agc;
    /* SAM, audio_driver.c:1955-1976, 97-123 */
    float sam_omega_min, sam_omega_max, sam_g1, sam_g2, sam_mtauR, sam_onem_mtauR, sam_mtauI, sam_onem_mtauI;
    float sam_fil_out, sam_lowpass, sam_omega2, sam_phs, sam_dsI, sam_dsQ;
    float sam_a[24], sam_b[24], sam_c[24], sam_d[24];
    int sam_count;
    float fade_dc27, fade_dc_insert;
    const float *sam_c0, *sam_c1;
    int carrier_freq_offset;
    /* FM, audio_driver.c:1516-1531 */
    float fm_i_prev, fm_q_prev, fm_lpf_prev, fm_hpf_prev_a, fm_hpf_prev_b, fm_sql_avg;
    int fm_count, fm_squelched;
    /* spectrum ring, audio_driver.c:1811-1826 */
    float fft_ring[1024];
    biquad_t zoom_bq_i[4], zoom_bq_q[4];      /* IIR_biquad_Zoom_FFT_I/Q, audio_driver.c:165-192 (state survives a reconfiguration) */
    firdec_t zoom_dec_i, zoom_dec_q;          /* DECIMATE_ZOOM_FFT_I/Q, re-initialised by AudioDriver_Spectrum_Set :1073-1086 */
    uint32_t samp_ptr;
    float spec_avg[512], spec_display_offset;     /* sd.FFT_AVGData, sd.display_offset (ui_spectrum.c:1432-1446, :1485) */
    /* NR */
    nr_t nr;
    notch_t notch;
    /* status */
    int adc_clip, adc_half_clip, adc_quarter_clip;
    int64_t blocks;
    /* TX, tx_processor.c */
    lattice_t tx_lat;
    biquad_t tx_bq[3];
    fir_t tx_hil_i, tx_hil_q;
    float alc_val, alc_decay, peak_audio;
    float tx_delay[320];
    uint32_t alc_delay_inbuf;
    float tx_fm_hpf_a, tx_fm_hpf_b; uint32_t tx_fm_accum;   /* TxProcessor_FM statics, tx_processor.c:537-538 */
    uint32_t tx_dds_sub_acc, tx_dds_sub_step, tx_dds_burst_acc, tx_dds_burst_step;   /* soft_dds_t subaudible_tone_dds / tone_burst_dds */
    /* FM subaudible-tone detector, audio_driver.c:1665-1734: Goertzel HIGH / LOW / CTR {r, cos, sin, buf[3]}, smoothed ratio, debounce */
    float gz_r[3], gz_cos[3], gz_sin[3], gz_buf[3][3], fm_subdet; int fm_tdet, fm_tone_detected; unsigned long fm_gcount;
    float tx_postfilt_gain_var;
};

static int lattice_from_table(const port_tables_t *t, int idx, lattice_t *l)
{
    if (idx < 0) { lattice_reset(l, 0, NULL, NULL); return 0; }
    const uhsdr_tbl_lattice_t *r = &t->lat[idx];
    if (r->num_stages > MAX_LAT) return -1;
    lattice_reset(l, r->num_stages, tbl_array(t, r->k_array, NULL), tbl_array(t, r->v_array, NULL));
    return 0;
}

static void nr_init(port_chan_t *c);
static int nb_active(const uhsdr_chan_cfg_t *cfg) { return (cfg->dsp_active & UHSDR_DSP_NB_ENABLE) && cfg->nb_setting > 0; }   /* ui_driver.c:405-408 */
static void nr_isr(port_chan_t *c, int n, float *buf);
static void nr_task(port_chan_t *c);
static void tx_setup(port_chan_t *c);

/* AudioDriver_SetProcessingChain, audio_driver.c:1093-1251 (+ audio_filter.c:1134-1223). */
static int chan_set_chain(port_chan_t *c, const uhsdr_chan_cfg_t *cfg)
{
    const port_tables_t *t = c->t;
    if (cfg->struct_size != sizeof(uhsdr_chan_cfg_t)) return UHSDR_ERR_ARG;
    if (cfg->filter_path < 1 || cfg->filter_path >= (int)t->h->num_paths) return UHSDR_ERR_ARG;
    if (cfg->spectrum_magnify < 0 || cfg->spectrum_magnify > 5) return UHSDR_ERR_ARG;     /* MAGNIFY_MAX */
    c->cfg = *cfg;
    if (cfg->spectrum_magnify != 0) {                 /* AudioDriver_Spectrum_Set, audio_driver.c:1055-1086 */
        const float *zb = tbl_array(t, t->ex->zoom_biquad_array, NULL), *zd = tbl_array(t, t->ex->zoom_decim_array, NULL);
        if (!zb || !zd) return UHSDR_ERR_TABLES;
        const int m = cfg->spectrum_magnify;
        for (int s = 0; s < 4; s++) {
            memcpy(c->zoom_bq_i[s].c, zb + (m - 1) * 20 + 5 * s, 5 * sizeof(float));
            memcpy(c->zoom_bq_q[s].c, zb + (m - 1) * 20 + 5 * s, 5 * sizeof(float));
        }
        firdec_reset(&c->zoom_dec_i, 4, 1 << m, zd + (m - 1) * 4);
        firdec_reset(&c->zoom_dec_q, 4, 1 << m, zd + (m - 1) * 4);
    }
    const uhsdr_tbl_path_t *p = c->path = &t->path[cfg->filter_path];
    const int mode = cfg->dmod_mode;
    const int is_am = (mode == UHSDR_DEMOD_AM || mode == UHSDR_DEMOD_SAM);

    c->decimation_rate = p->sample_rate_dec;
    c->decimated_freq = 48000 / c->decimation_rate;

    if (lattice_from_table(t, p->pre_lattice, &c->pre)) return UHSDR_ERR_TABLES;
    if (lattice_from_table(t, p->aa_lattice, &c->aa)) return UHSDR_ERR_TABLES;

    /* AudioDriver_SetRxTxAudioProcessingAudioFilters, audio_driver.c:994-1050.  Biquad STATES are
     * not touched by the reference here, only coefficients. */
    float FSdec = 48000 / (p->sample_rate_dec != 0 ? p->sample_rate_dec : 1);
    float co[5];
    if (cfg->dsp_active & UHSDR_DSP_MNOTCH_ENABLE) { bq_bandstop(co, (float)(unsigned long)cfg->notch_frequency, FSdec); memcpy(c->bq1[0].c, co, sizeof(co)); }
    else memcpy(c->bq1[0].c, BQ_PASS, sizeof(BQ_PASS));
    memcpy(c->bq1[3].c, BQ_PASS, sizeof(BQ_PASS));
    if (cfg->dsp_active & UHSDR_DSP_MPEAK_ENABLE) { bq_bandpass(co, (float)(unsigned long)cfg->peak_frequency, FSdec); memcpy(c->bq1[1].c, co, sizeof(co)); }
    else memcpy(c->bq1[1].c, BQ_PASS, sizeof(BQ_PASS));
    bq_lowshelf(co, 250, 0.7, cfg->bass_gain, FSdec);
    memcpy(c->bq1[2].c, co, sizeof(co));
    bq_highshelf(co, 3500, 0.9, cfg->treble_gain, 48000);
    memcpy(c->bq2.c, co, sizeof(co));

    tx_setup(c);

    /* NR alpha, audio_driver.c:1195 */
    c->nr.alpha = 0.799 + ((float)(uint8_t)cfg->nr_strength / 1000.0);

    /* interpolator, audio_driver.c:1209-1224 */
    if (p->interpolate >= 0) {
        const uhsdr_tbl_interp_t *ip = &t->interp[p->interpolate];
        firint_reset(&c->interp, c->decimation_rate, ip->phase_length_field, tbl_array(t, ip->coeff_array, NULL));
        c->interp_active = c->interp.plen > 0;
    } else {
        c->interp_active = 0;
    }

    /* AudioDriver_SetSamPllParameters, audio_driver.c:709-745 */
    {
        const float decimSampleRate = c->decimated_freq;
        const float pll_fmax = cfg->sam_pll_fmax;
        float omegaN = cfg->sam_omegaN;
        float zeta = (float)cfg->sam_zeta / 100.0;
        c->sam_omega_min = -(2.0 * PI_F * pll_fmax / decimSampleRate);
        c->sam_omega_max = (2.0 * PI_F * pll_fmax / decimSampleRate);
        c->sam_g1 = (1.0 - expf(-2.0 * omegaN * zeta / decimSampleRate));
        c->sam_g2 = (-c->sam_g1 + 2.0 * (1 - expf(-omegaN * zeta / decimSampleRate) * cosf(omegaN / decimSampleRate * sqrtf(1.0 - zeta * zeta))));
        float tauR = 0.02;
        float tauI = 1.4;
        c->sam_mtauR = (expf(-1 / (decimSampleRate * tauR)));
        c->sam_onem_mtauR = (1.0 - c->sam_mtauR);
        c->sam_mtauI = (expf(-1 / (decimSampleRate * tauI)));
        c->sam_onem_mtauI = (1.0 - c->sam_mtauI);
    }
    /* AudioDriver_SetRxIqCorrection, audio_driver.c:747-755 */
    c->M_c1 = 0.0; c->M_c2 = 1.0; c->teta1_old = 0.0; c->teta2_old = 0.0; c->teta3_old = 0.0;

    /* AudioFilter_SetRxHilbertAndDecimationFIR, audio_filter.c:1134-1223 */
    const float *fir_i = tbl_array(t, p->fir_i_array, NULL), *fir_q = tbl_array(t, p->fir_q_array, NULL);
    fir_reset(&c->hil_i, p->fir_numtaps, fir_i);
    fir_reset(&c->hil_q, p->fir_numtaps, fir_q);
    c->dec_active = 0;
    if (is_am) {
        if (p->fir_numtaps != 0) {
            firdec_reset(&c->dec_i, p->fir_numtaps, c->decimation_rate, fir_i);
            firdec_reset(&c->dec_q, p->fir_numtaps, c->decimation_rate, fir_q);
            c->dec_active = 1;
        }
    } else if (p->dec_array >= 0) {
        const float *dc = tbl_array(t, p->dec_array, NULL);
        firdec_reset(&c->dec_i, p->dec_numtaps, c->decimation_rate, dc);
        firdec_reset(&c->dec_q, p->dec_numtaps, c->decimation_rate, dc);
        c->dec_active = 1;
    }
    /* audio_driver.c:2718-2720 */
    c->use_decimated_iq = (p->fir_is_new_coeffs && mode != UHSDR_DEMOD_FM) || is_am;
    if (mode != UHSDR_DEMOD_FM && !c->dec_active) return UHSDR_ERR_UNSUPPORTED;
    if (mode == UHSDR_DEMOD_FM && c->decimation_rate != 1) return UHSDR_ERR_UNSUPPORTED;

    /* AudioDriver_AgcWdsp_Set, audio_driver.c:628-631 */
    agc_setup(&c->agc, cfg, (float)c->decimated_freq, is_am);
    if (cfg->fm_subaudible_tone_det_freq > 0.0f) {
        /* AudioManagement_CalcSubaudibleDetFreq, audio_management.c:311-326 + AudioFilter_CalcGoertzel, audio_filter.c:1281-1288 */
        const float ratio[3] = { 1.04, 0.95, 1.0 };           /* FM_GOERTZEL_HIGH, FM_GOERTZEL_LOW, centre */
        const uint32_t size = 400 * 32;                         /* FM_SUBAUDIBLE_GOERTZEL_WINDOW * AUDIO_BLOCK_SIZE */
        const float freq = cfg->fm_subaudible_tone_det_freq, samplerate = 48000;
        for (int k = 0; k < 3; k++) {
            float ga = (0.5 + (freq * ratio[k]) * size / samplerate);
            float gb = (2 * 3.14159265358979f * ga) / size;
            c->gz_sin[k] = sinf(gb); c->gz_cos[k] = cosf(gb); c->gz_r[k] = 2 * c->gz_cos[k];
        }
    }
    /* auto-notch init, audio_driver.c:1165-1187: state, energy and the delay buffer are cleared, the coefficients
     * only with reset_dsp_nr (a fresh channel starts from zeros anyway) */
    c->notch.mu = log10f(((cfg->notch_mu + 1.0) / 1500.0) + 1.0);
    memset(c->notch.state, 0, sizeof(c->notch.state));
    memset(c->notch.delay, 0, sizeof(c->notch.delay));
    c->notch.energy = 0.0f; c->notch.x0 = 0.0f;
    return UHSDR_OK;
}

port_chan_t *port_chan_create(const port_tables_t *t, const uhsdr_chan_cfg_t *cfg)
{
    if (!t || !cfg) return NULL;
    port_chan_t *c = (port_chan_t *)calloc(1, sizeof(*c));
    c->t = t;
    /* boot values, SURVEY.md 8a "state inventory" */
    for (int s = 0; s < 4; s++) memcpy(c->bq1[s].c, BQ_PASS, sizeof(BQ_PASS));
    memcpy(c->bq2.c, BQ_PASS, sizeof(BQ_PASS));
    for (int s = 0; s < 3; s++) memcpy(c->tx_bq[s].c, BQ_PASS, sizeof(BQ_PASS));
    c->fm_squelched = 1;                      /* audio_driver.c:475 */
    c->alc_val = 1;                           /* tx_processor.c:137 */
    c->tw_state = UHSDR_TWINPEAKS_WAIT;       /* uhsdr_main.c:339 */
    c->sam_c0 = tbl_array(t, t->ex->sam_c0_array, NULL);
    c->sam_c1 = tbl_array(t, t->ex->sam_c1_array, NULL);
    lattice_from_table(t, t->ex->fm_squelch_lattice, &c->sql_hpf);   /* audio_driver.c:481-485 */
    nr_init(c);
    if (chan_set_chain(c, cfg) != UHSDR_OK) { free(c); return NULL; }
    return c;
}

int port_chan_reconfigure(port_chan_t *c, const uhsdr_chan_cfg_t *cfg) { return chan_set_chain(c, cfg); }
void port_chan_free(port_chan_t *c) { free(c); }

/* ------------------------------------------------------------------------------------------ */
/* RX stages                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* AudioDriver_RxHandleIqCorrection, audio_driver.c:2254-2316 */
/* AudioDriver_RxHandleTwinpeaks, audio_driver.c:2173-2248 (ts.twinpeaks_tested and the function's statics per channel) */
static void rx_twinpeaks(port_chan_t *c)
{
    if (c->tw_state == UHSDR_TWINPEAKS_WAIT) c->tw_counter++;
    if (c->tw_counter > 1000) { c->tw_state = UHSDR_TWINPEAKS_SAMPLING; c->tw_counter = 0; c->tw_phase = 0.0; c->tw_runs = 0; }
    if (c->teta3 != 0.0 && c->tw_state == UHSDR_TWINPEAKS_SAMPLING) {
        float phase_IQ_cur = asinf(c->teta1 / c->teta3);
        if (c->tw_runs == 0) c->tw_phase = phase_IQ_cur;
        else c->tw_phase = 0.05 * phase_IQ_cur + 0.95 * c->tw_phase;
        c->tw_runs++;
        if (c->tw_runs == 50) {
            if (fabsf(c->tw_phase) > (M_PI / 8.0)) {
                c->tw_state = UHSDR_TWINPEAKS_CODEC_RESTART;
                c->tw_restarts++;
                if (c->tw_restarts >= 4) { c->tw_state = UHSDR_TWINPEAKS_UNCORRECTABLE; c->tw_restarts = 0; }
            } else { c->tw_state = UHSDR_TWINPEAKS_DONE; c->tw_restarts = 0; }
        }
    }
}

static void rx_iq_correction(port_chan_t *c, float *ib, float *qb)
{
    if (!c->cfg.iq_auto_correction) {
        for (int i = 0; i < BLK; i++) ib[i] = ib[i] * c->cfg.rx_adj_gain_i;
        for (int i = 0; i < BLK; i++) qb[i] = qb[i] * c->cfg.rx_adj_gain_q;
        /* AudioDriver_IQPhaseAdjust, audio_driver.c:1776-1801 */
        float bal = c->cfg.iq_phase_balance_rx;
        if (bal < 0) { for (int i = 0; i < BLK; i++) qb[i] = qb[i] + ib[i] * bal; }
        else if (bal > 0) { for (int i = 0; i < BLK; i++) ib[i] = ib[i] + qb[i] * bal; }
        return;
    }
    for (int i = 0; i < BLK; i++) {
        c->teta1 += sign_new(ib[i]) * qb[i];
        c->teta2 += sign_new(ib[i]) * ib[i];
        c->teta3 += sign_new(qb[i]) * qb[i];
    }
    c->teta1 = -0.003 * (c->teta1 / BLK) + 0.997 * c->teta1_old;
    c->teta2 = 0.003 * (c->teta2 / BLK) + 0.997 * c->teta2_old;
    c->teta3 = 0.003 * (c->teta3 / BLK) + 0.997 * c->teta3_old;
    c->M_c1 = (c->teta2 != 0.0) ? c->teta1 / c->teta2 : 0.0;
    float help = (c->teta2 * c->teta2);
    if (help > 0.0) help = (c->teta3 * c->teta3 - c->teta1 * c->teta1) / help;
    c->M_c2 = (help > 0.0) ? sqrtf(help) : 1.0;
    rx_twinpeaks(c);
    c->teta1_old = c->teta1; c->teta2_old = c->teta2; c->teta3_old = c->teta3;
    c->teta1 = 0.0; c->teta2 = 0.0; c->teta3 = 0.0;
    for (int i = 0; i < BLK; i++) qb[i] += c->M_c1 * ib[i];
    for (int i = 0; i < BLK; i++) ib[i] = ib[i] * c->M_c2;
}

/* FreqShift, freq_shift.c:275-331 (+ QuarterFs :219-262, Approx :57-108) */
static void freq_shift(port_chan_t *c, float *I, float *Q, int shift)
{
    int newf = abs(shift);
    if (c->conversion_freq != newf) {
        c->conversion_freq = newf;
        double rate = (2 * M_PI * (float)newf) / 48000.0f;
        c->osc_cos = cos(rate); c->osc_sin = sin(rate);
        c->osc_vect_i = 0; c->osc_vect_q = 1;
    }
    int dir_down = shift > 0;
    float *ib = dir_down ? Q : I, *qb = dir_down ? I : Q;
    if (newf == 12000) {
        for (int i = 0; i < BLK; i += 4) {
            float h1 = qb[i + 1], h2 = -ib[i + 1];
            ib[i + 1] = h1; qb[i + 1] = h2;
            h1 = -ib[i + 2]; h2 = -qb[i + 2];
            ib[i + 2] = h1; qb[i + 2] = h2;
            h1 = -qb[i + 3]; h2 = ib[i + 3];
            ib[i + 3] = h1; qb[i + 3] = h2;
        }
        return;
    }
    for (int i = 0; i < BLK; i++) {
        float oq = (c->osc_vect_q * c->osc_cos) - (c->osc_vect_i * c->osc_sin);
        float oi = (c->osc_vect_i * c->osc_cos) + (c->osc_vect_q * c->osc_sin);
        float qt = qb[i], it = ib[i];
        qb[i] = (qt * oq) - (it * oi);
        ib[i] = (it * oq) + (qt * oi);
        c->osc_vect_q = oq; c->osc_vect_i = oi;
    }
    float g = (3 - ((c->osc_vect_q * c->osc_vect_q) + (c->osc_vect_i * c->osc_vect_i))) / 2;
    c->osc_vect_q = g * c->osc_vect_q;
    c->osc_vect_i = g * c->osc_vect_i;
}

static int translate_freq(int iq_freq_mode)
{
    switch (iq_freq_mode) {        /* AudioDriver_GetTranslateFreq, audio_driver.c:445-464 */
    case UHSDR_FREQ_IQ_CONV_P6KHZ: return 6000;
    case UHSDR_FREQ_IQ_CONV_M6KHZ: return -6000;
    case UHSDR_FREQ_IQ_CONV_P12KHZ: return 12000;
    case UHSDR_FREQ_IQ_CONV_M12KHZ: return -12000;
    default: return 0;
    }
}

/* AudioDriver_FadeLeveler, audio_driver.c:1911-1923 */
static float fade_leveler(port_chan_t *c, float audio, float corr)
{
    c->fade_dc27 = c->sam_mtauR * c->fade_dc27 + c->sam_onem_mtauR * audio;
    c->fade_dc_insert = c->sam_mtauI * c->fade_dc_insert + c->sam_onem_mtauI * corr;
    audio = audio + c->fade_dc_insert - c->fade_dc27;
    return audio;
}

/* AudioDriver_DemodSAM, audio_driver.c:1990-2166 */
static void demod_am_sam(port_chan_t *c, const float *ib, const float *qb, float *a, int n, float sampleRate)
{
    if (c->cfg.dmod_mode == UHSDR_DEMOD_AM) {
        for (int i = 0; i < n; i++) {
            float audio = sqrtf(ib[i] * ib[i] + qb[i] * qb[i]);   /* arm_sqrt_f32: sqrtf for in >= 0 */
            if (c->cfg.sam_fade_leveler) audio = fade_leveler(c, audio, 0);
            a[i] = audio;
        }
        return;
    }
    for (int i = 0; i < n; i++) {
        float Sin, Cos;
        sincosf(c->sam_phs, &Sin, &Cos);
        float ai = Cos * ib[i], bi = Sin * ib[i], aq = Cos * qb[i], bq = Sin * qb[i];
        float audio;
        float corr0 = ai + bq, corr1 = -bi + aq;
        if (c->cfg.sam_sideband != UHSDR_SAM_SIDEBAND_BOTH) {
            c->sam_a[0] = c->sam_dsI; c->sam_b[0] = bi; c->sam_c[0] = c->sam_dsQ; c->sam_d[0] = aq;
            c->sam_dsI = ai; c->sam_dsQ = bq;
            for (int j = 0; j < 7; j++) {
                int k = 3 * j;
                c->sam_a[k + 3] = c->sam_c0[j] * (c->sam_a[k] - c->sam_a[k + 5]) + c->sam_a[k + 2];
                c->sam_b[k + 3] = c->sam_c1[j] * (c->sam_b[k] - c->sam_b[k + 5]) + c->sam_b[k + 2];
                c->sam_c[k + 3] = c->sam_c0[j] * (c->sam_c[k] - c->sam_c[k + 5]) + c->sam_c[k + 2];
                c->sam_d[k + 3] = c->sam_c1[j] * (c->sam_d[k] - c->sam_d[k + 5]) + c->sam_d[k + 2];
            }
            float ai_ps = c->sam_a[21], bi_ps = c->sam_b[21], bq_ps = c->sam_c[21], aq_ps = c->sam_d[21];
            for (int j = 23; j > 0; j--) {
                c->sam_a[j] = c->sam_a[j - 1]; c->sam_b[j] = c->sam_b[j - 1];
                c->sam_c[j] = c->sam_c[j - 1]; c->sam_d[j] = c->sam_d[j - 1];
            }
            if (c->cfg.sam_sideband == UHSDR_SAM_SIDEBAND_LSB) audio = (ai_ps + bi_ps) - (aq_ps - bq_ps);
            else audio = (ai_ps - bi_ps) + (aq_ps + bq_ps);
        } else {
            audio = corr0;
        }
        if (c->cfg.sam_fade_leveler) audio = fade_leveler(c, audio, corr0);
        a[i] = audio;

        float phzerror = atan2f(corr1, corr0);
        float del_out = c->sam_fil_out;
        c->sam_omega2 = c->sam_omega2 + c->sam_g2 * phzerror;
        if (c->sam_omega2 < c->sam_omega_min) c->sam_omega2 = c->sam_omega_min;
        else if (c->sam_omega2 > c->sam_omega_max) c->sam_omega2 = c->sam_omega_max;
        c->sam_fil_out = c->sam_g1 * phzerror + c->sam_omega2;
        c->sam_phs = c->sam_phs + del_out;
        while (c->sam_phs >= 2.0 * PI_F) c->sam_phs -= (2.0 * PI_F);
        while (c->sam_phs < 0.0) c->sam_phs += (2.0 * PI_F);
    }
    c->sam_count++;
    if (c->sam_count > 50) {
        float carrier = 0.1 * (c->sam_omega2 * sampleRate) / (2.0 * PI_F);
        carrier = carrier + 0.9 * c->sam_lowpass;
        c->carrier_freq_offset = carrier;
        c->sam_count = 0;
        c->sam_lowpass = carrier;
    }
}

/* AudioDriver_DemodFM, audio_driver.c:1544-1737 (subaudible-tone detection not restated:
 * with the 3 x Goertzel subaudible-tone detector when cfg.fm_subaudible_tone_det_freq > 0). Returns signal_active. */
static int demod_fm(port_chan_t *c, const float *ib, const float *qb, float *a)
{
    float squelch_buf[BLK], goertzel_buf[BLK];
    if (c->cfg.iq_freq_mode != UHSDR_FREQ_IQ_CONV_OFF) {
        const int tone_det_enabled = c->cfg.fm_subaudible_tone_det_freq != 0;
        for (int i = 0; i < BLK; i++) {
            float y = (c->fm_i_prev * qb[i]) - (ib[i] * c->fm_q_prev);
            float x = (c->fm_i_prev * ib[i]) + (qb[i] * c->fm_q_prev);
            float angle = atan2f(y, x);
            squelch_buf[i] = angle;
            float av = c->fm_lpf_prev + (0.05 * (angle - c->fm_lpf_prev));
            c->fm_lpf_prev = av;
            goertzel_buf[i] = av;
            if (((!c->fm_squelched) && (!tone_det_enabled)) || ((c->fm_tone_detected) && (tone_det_enabled)) || ((!c->cfg.fm_sql_threshold))) {
                float b = 0.96 * (c->fm_hpf_prev_b + av - c->fm_hpf_prev_a);
                c->fm_hpf_prev_a = av;
                c->fm_hpf_prev_b = b;
                a[i] = b;
            } else {
                a[i] = 0;
            }
            c->fm_q_prev = qb[i];
            c->fm_i_prev = ib[i];
        }
        lattice_run(&c->sql_hpf, squelch_buf, BLK);
        c->fm_sql_avg = ((1 - 0.005) * c->fm_sql_avg) + (0.005 * sqrtf(fabsf(squelch_buf[0])));
        c->fm_count++;
        c->fm_count %= 200;                    /* FM_SQUELCH_PROC_DECIMATION, uint8 count < 200 */
        if (c->fm_count == 0) {
            if (c->fm_sql_avg > 0.175) c->fm_sql_avg = 0.175;
            float s = c->fm_sql_avg * 172;
            if (s > 24) s = 24;
            s = 22 - s;
            const int thr = (uint8_t)c->cfg.fm_sql_threshold;
            if (thr == 0) c->fm_squelched = 0;
            else if (c->fm_squelched) { if (s >= (float)(thr + 3)) c->fm_squelched = 0; }
            else if (thr > 3) { if (s < (float)(thr - 3)) c->fm_squelched = 1; }
            else { if (s < (float)thr) c->fm_squelched = 1; }
        }
        if (tone_det_enabled) {                               /* :1665-1729 */
            c->fm_gcount++;
            for (int i = 0; i < BLK; i++)
                for (int k = 0; k < 3; k++) {                 /* AudioFilter_GoertzelInput, audio_filter.c:1290-1295 (HIGH, LOW, CTR) */
                    c->gz_buf[k][0] = c->gz_r[k] * c->gz_buf[k][1] - c->gz_buf[k][2] + goertzel_buf[i];
                    c->gz_buf[k][2] = c->gz_buf[k][1];
                    c->gz_buf[k][1] = c->gz_buf[k][0];
                }
            if (c->fm_gcount >= 400) {                        /* FM_SUBAUDIBLE_GOERTZEL_WINDOW */
                float en[3];
                for (int k = 0; k < 3; k++) {                 /* AudioFilter_GoertzelEnergy, :1297-1305 */
                    float ea = (c->gz_buf[k][1] - (c->gz_buf[k][2] * c->gz_cos[k]));
                    float eb = (c->gz_buf[k][2] * c->gz_sin[k]);
                    c->gz_buf[k][0] = 0; c->gz_buf[k][1] = 0; c->gz_buf[k][2] = 0;
                    en[k] = sqrtf(ea * ea + eb * eb);
                }
                float s = en[0] + en[1];
                float r = en[2];
                c->fm_subdet = ((1 - 0.9) * c->fm_subdet) + (r / (s / 2) * 0.9);
                if (c->fm_subdet > 1.75) { c->fm_tdet++; if (c->fm_tdet > 5) c->fm_tdet = 5; }
                else { if (c->fm_tdet) c->fm_tdet--; }
                c->fm_tone_detected = c->fm_tdet >= 2;
                c->fm_gcount = 0;
            }
        } else {
            c->fm_tone_detected = 1;
        }
    }
    return !c->fm_squelched;
}

/* RxProcessor_DemodAudioPostprocessing, audio_driver.c:2436-2592.  a0: decimated audio in,
 * a1: 48 ksps audio out. */
static void rx_postprocess(port_chan_t *c, float *a0, float *a1, int ndec)
{
    const int mode = c->cfg.dmod_mode;
    /* audio_driver.c:2443-2456 */
    if ((c->cfg.dsp_active & UHSDR_DSP_NOTCH_ENABLE) && mode != UHSDR_DEMOD_CW && !(mode == UHSDR_DEMOD_SAM && c->decimated_freq == 24000))
        notch_run(&c->notch, a0, ndec);
    if (c->pre.n > 0) lattice_run(&c->pre, a0, ndec);
    agc_run(&c->agc, a0, ndec);
    /* audio_driver.c:2501: is_dsp_nb_active() || is_dsp_nr() */
    if (c->decimated_freq == 12000 && ((c->cfg.dsp_active & UHSDR_DSP_NR_ENABLE) || nb_active(&c->cfg))) nr_isr(c, ndec, a0);

    const float post_agc_gain_scaling = (c->path->sample_rate_dec == 4) ? 3.46 : (3.46 * 0.6);
    const float scale_gain = post_agc_gain_scaling * ((mode == UHSDR_DEMOD_AM || mode == UHSDR_DEMOD_SAM) ? 0.5 : 0.333);
    for (int i = 0; i < ndec; i++) a0[i] = a0[i] * scale_gain;
    for (int s = 0; s < 4; s++) biquad_run(&c->bq1[s], a0, ndec);
    if (c->interp_active) firint_run(&c->interp, a0, a1, ndec);
    if (c->aa.n > 0) lattice_run(&c->aa, a1, BLK);
}

/* AudioDriver_RxProcessor, audio_driver.c:2603-2942, one block. a0/a1 persist across blocks like
 * adb.a_buffer (a path without interpolator would re-emit stale a_buffer[1] contents). */
static void rx_block(port_chan_t *c, const int32_t *src, int32_t *dst, float *dst_f, int external_mute, rx_scratch_t *sc)
{
    float ib[BLK], qb[BLK];
    float *a0 = sc->a0, *a1 = sc->a1;
    const int mode = c->cfg.dmod_mode;

    for (int i = 0; i < BLK; i++) {
        int32_t l = src[2 * i], r = src[2 * i + 1];
        int32_t level = abs(l) >> 16;
        if (level > 4096 / 4) { c->adc_quarter_clip = 1; if (level > 4096 / 2) { c->adc_half_clip = 1; if (level > 4096) c->adc_clip = 1; } }
        ib[i] = l; qb[i] = r;
    }
    for (int i = 0; i < BLK; i++) ib[i] = ib[i] * (float)(0.0000152587890625);
    for (int i = 0; i < BLK; i++) qb[i] = qb[i] * (float)(0.0000152587890625);
    rx_iq_correction(c, ib, qb);

    /* AudioDriver_SpectrumNoZoomProcessSamples, audio_driver.c:1811-1849 (fft_iq_len 1024) */
    if (c->cfg.spectrum_enable && c->cfg.spectrum_magnify == 0) {
        for (int i = 0; i < BLK; i++) {
            c->fft_ring[c->samp_ptr++] = qb[i];
            c->fft_ring[c->samp_ptr++] = ib[i];
            if (c->samp_ptr >= 1024 - 1) c->samp_ptr = 0;
        }
    }
    if (c->cfg.iq_freq_mode) freq_shift(c, ib, qb, translate_freq(c->cfg.iq_freq_mode));

    /* AudioDriver_SpectrumZoomProcessSamples, audio_driver.c:1860-1909: 4-stage biquad low-pass on I and Q (after the
     * translation), 4-tap FIR decimation by 2^magnify, 32 >> magnify pairs into the ring */
    if (c->cfg.spectrum_enable && c->cfg.spectrum_magnify != 0) {
        float zi[BLK], zq[BLK];
        memcpy(zi, ib, sizeof(zi)); memcpy(zq, qb, sizeof(zq));
        for (int s = 0; s < 4; s++) biquad_run(&c->zoom_bq_i[s], zi, BLK);
        for (int s = 0; s < 4; s++) biquad_run(&c->zoom_bq_q[s], zq, BLK);
        firdec_run(&c->zoom_dec_i, zi, zi, BLK);
        firdec_run(&c->zoom_dec_q, zq, zq, BLK);
        for (int i = 0; i < (BLK >> c->cfg.spectrum_magnify); i++) {
            c->fft_ring[c->samp_ptr++] = zq[i];
            c->fft_ring[c->samp_ptr++] = zi[i];
            if (c->samp_ptr >= 1024 - 1) c->samp_ptr = 0;
        }
    }

    int signal_active = 1;
    const int ndec = BLK / c->decimation_rate;
    const int n_iq = c->use_decimated_iq ? ndec : BLK;
    if (c->use_decimated_iq) {
        firdec_run(&c->dec_i, ib, ib, BLK);
        firdec_run(&c->dec_q, qb, qb, BLK);
    }
    if (mode != UHSDR_DEMOD_SAM && mode != UHSDR_DEMOD_AM) {
        fir_run(&c->hil_i, ib, ib, n_iq);
        fir_run(&c->hil_q, qb, qb, n_iq);
    }
    /* demodulator selection, audio_driver.c:2757-2790 */
    if (mode == UHSDR_DEMOD_AM || mode == UHSDR_DEMOD_SAM) {
        demod_am_sam(c, ib, qb, a0, n_iq, (float)(c->use_decimated_iq ? c->decimated_freq : 48000));
    } else if (mode == UHSDR_DEMOD_FM) {
        signal_active = demod_fm(c, ib, qb, a0);
    } else {
        int lsb = (mode == UHSDR_DEMOD_LSB) || (mode == UHSDR_DEMOD_CW && c->cfg.cw_lsb) || (mode == UHSDR_DEMOD_DIGI && c->cfg.digi_lsb);
        if (lsb) for (int i = 0; i < n_iq; i++) a0[i] = ib[i] - qb[i];
        else for (int i = 0; i < n_iq; i++) a0[i] = ib[i] + qb[i];
    }
    if (mode != UHSDR_DEMOD_FM) {
        if (!c->use_decimated_iq) firdec_run(&c->dec_i, a0, a0, n_iq);
        rx_postprocess(c, a0, a1, ndec);
    } else {
        const float sc_fm = c->cfg.fm_dev_5khz ? (10000 / 2) : 10000;
        for (int i = 0; i < BLK; i++) a1[i] = a0[i] * sc_fm;
        agc_run(&c->agc, a0, BLK);               /* S-meter only, audio_driver.c:2828 */
    }
    biquad_run(&c->bq2, a1, BLK);

    int mute = external_mute || !signal_active;
    if (mute) {
        memset(a0, 0, sizeof(float) * BLK); memset(a1, 0, sizeof(float) * BLK);
    } else {
        for (int i = 0; i < BLK; i++) a1[i] = a1[i] * (float)10;   /* LINE_OUT_SCALING_FACTOR */
        memcpy(a0, a1, sizeof(float) * BLK);
    }
    for (int i = 0; i < BLK; i++) {
        if (mute) { dst[2 * i] = 0; dst[2 * i + 1] = 0; }
        else {
            int32_t l = (int32_t)a1[i], r = (int32_t)a0[i];
            dst[2 * i] = (int32_t)((uint32_t)l << 16);
            dst[2 * i + 1] = (int32_t)((uint32_t)r << 16);
        }
        if (dst_f) dst_f[i] = a1[i];
    }
    c->blocks++;
}

int port_rx(port_chan_t *c, const int32_t *iq, int32_t *audio, float *audio_f, int nblocks, const uint8_t *mute)
{
    if (!c || !iq || !audio || nblocks < 0) return UHSDR_ERR_ARG;
    rx_scratch_t *sc = &c->scratch;
    for (int b = 0; b < nblocks; b++) {
        rx_block(c, iq + (size_t)b * 2 * BLK, audio + (size_t)b * 2 * BLK, audio_f ? audio_f + (size_t)b * BLK : NULL,
                 mute ? mute[b] : 0, sc);
        nr_task(c);   /* the oracle's fixed schedule: deferred NR task once after every block */
    }
    return UHSDR_OK;
}

int port_twinpeaks_rearm(port_chan_t *c) { if (!c) return UHSDR_ERR_ARG; c->tw_state = UHSDR_TWINPEAKS_WAIT; return UHSDR_OK; }

int port_get_status(const port_chan_t *c, uhsdr_chan_status_t *st)
{
    memset(st, 0, sizeof(*st));
    st->adc_clip = c->adc_clip; st->adc_half_clip = c->adc_half_clip; st->adc_quarter_clip = c->adc_quarter_clip;
    st->agc_action = c->agc.action; st->agc_hang_action = c->agc.hang_action;
    st->fm_squelched = c->fm_squelched; st->fm_sql_avg = c->fm_sql_avg;
    st->sam_carrier_freq_offset = c->carrier_freq_offset;
    st->iq_corr_c1 = c->M_c1; st->iq_corr_c2 = c->M_c2;
    st->tx_peak_audio = c->peak_audio; st->tx_alc_val = c->alc_val;
    st->blocks_processed = c->blocks;
    st->twinpeaks_state = c->tw_state; st->twinpeaks_restarts = (int32_t)c->tw_restarts;
    return UHSDR_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* placeholders completed in later sections of this file                                      */
/* ------------------------------------------------------------------------------------------ */
#include "uhsdr_port_nr.inc"
#include "uhsdr_port_tx.inc"
