/*
 * dump_tables.c -- ORACLE / BUILD-TIME TOOL (never linked into the product library).
 *
 * Walks the reference's FilterInfo[] / FilterPathInfo[] and coefficient arrays (linked from
 * /root/reference at build time) and serialises them into the blob format of
 * include/uhsdr_tables.h.  This is the code a UHSDR maintainer would add on the firmware side
 * (INTEGRATION.md); here it produces uhsdr_b200/data/uhsdr_tables.bin.
 *
 * audio_driver.c is #included to reach its file-static NR coefficient arrays
 * (audio_driver.c:195,198).
 */
#include "audio_driver.c"
#include "iq_tx_filter.h"
#include "softdds/dds_table.h"
#include "uhsdr_tables.h"
extern const float32_t SQRT_von_Hann_256[256];   /* audio_nr.c:76 */
#include <stdio.h>

#define MAX_ARRAYS 512
static const float *arr_ptr[MAX_ARRAYS];
static uint32_t arr_cnt[MAX_ARRAYS];
static int n_arr = 0;

static int add_array(const float *p, uint32_t count)
{
    if (p == NULL) return -1;
    for (int i = 0; i < n_arr; i++)
        if (arr_ptr[i] == p && arr_cnt[i] == count) return i;
    arr_ptr[n_arr] = p; arr_cnt[n_arr] = count;
    return n_arr++;
}

#define MAX_LAT 128
static const arm_iir_lattice_instance_f32 *lat_ptr[MAX_LAT];
static uhsdr_tbl_lattice_t lat_row[MAX_LAT];
static int n_lat = 0;
static int add_lattice(const arm_iir_lattice_instance_f32 *l)
{
    if (l == NULL) return -1;
    for (int i = 0; i < n_lat; i++) if (lat_ptr[i] == l) return i;
    lat_ptr[n_lat] = l;
    lat_row[n_lat].num_stages = l->numStages;
    lat_row[n_lat].k_array = add_array(l->pkCoeffs, l->numStages);
    lat_row[n_lat].v_array = add_array(l->pvCoeffs, l->numStages + 1);
    return n_lat++;
}

#define MAX_INT 16
static const arm_fir_interpolate_instance_f32 *int_ptr[MAX_INT];
static uhsdr_tbl_interp_t int_row[MAX_INT];
static int n_int = 0;
static int add_interp(const arm_fir_interpolate_instance_f32 *p)
{
    if (p == NULL) return -1;
    for (int i = 0; i < n_int; i++) if (int_ptr[i] == p) return i;
    int_ptr[n_int] = p;
    int_row[n_int].L = p->L;
    int_row[n_int].phase_length_field = p->phaseLength;
    int_row[n_int].num_coeffs = p->phaseLength;
    int_row[n_int].coeff_array = add_array(p->pCoeffs, p->phaseLength);
    return n_int++;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin [spectrum_window.f32]\n", argv[0]); return 2; }

    static uhsdr_tbl_path_t paths[AUDIO_FILTER_PATH_NUM];
    static uhsdr_tbl_filter_t filters[AUDIO_FILTER_NUM];
    uhsdr_tbl_extras_t ex;
    memset(&ex, 0xff, sizeof(ex));
    memset(paths, 0, sizeof(paths));
    memset(filters, 0, sizeof(filters));

    for (int i = 0; i < AUDIO_FILTER_PATH_NUM; i++) {
        const FilterPathDescriptor *p = &FilterPathInfo[i];
        uhsdr_tbl_path_t *r = &paths[i];
        r->id = p->id; r->mode_mask = p->mode; r->filter_select_id = p->filter_select_id;
        r->fir_numtaps = p->FIR_numTaps;
        r->fir_i_array = add_array(p->FIR_I_coeff_file, p->FIR_numTaps);
        r->fir_q_array = add_array(p->FIR_Q_coeff_file, p->FIR_numTaps);
        r->fir_is_new_coeffs = (p->FIR_I_coeff_file == i_rx_new_coeffs);
        r->dec_array = p->dec ? add_array(p->dec->pCoeffs, p->dec->numTaps) : -1;
        r->dec_numtaps = p->dec ? p->dec->numTaps : 0;
        r->sample_rate_dec = p->sample_rate_dec;
        r->pre_lattice = add_lattice(p->pre_instance);
        r->interpolate = add_interp(p->interpolate);
        r->aa_lattice = add_lattice(p->iir_instance);
        r->offset_hz = p->offset;
        if (p->name) strncpy(r->name, p->name, sizeof(r->name) - 1);
    }
    for (int i = 0; i < AUDIO_FILTER_NUM; i++) {
        filters[i].id = FilterInfo[i].id; filters[i].width = FilterInfo[i].width;
        if (FilterInfo[i].name) strncpy(filters[i].name, FilterInfo[i].name, sizeof(filters[i].name) - 1);
    }
    ex.nr_decimate_array = add_array(NR_decimate_coeffs, 4);
    ex.nr_interpolate_array = add_array(NR_interpolate_coeffs, NR_INTERPOLATE_NO_TAPS);
    ex.sqrt_hann_256_array = add_array(SQRT_von_Hann_256, 256);
    ex.sam_c0_array = add_array(demod_sam_const.c0, SAM_PLL_HILBERT_STAGES);
    ex.sam_c1_array = add_array(demod_sam_const.c1, SAM_PLL_HILBERT_STAGES);
    ex.fm_squelch_lattice = add_lattice(&IIR_15k_hpf);
    ex.tx_hilbert_i_array = add_array(iq_tx_wide.i, iq_tx_wide.num_taps);
    ex.tx_hilbert_q_array = add_array(iq_tx_wide.q, iq_tx_wide.num_taps);
    ex.tx_hilbert_numtaps = iq_tx_wide.num_taps;
    ex.tx_lattice_soprano = add_lattice(&IIR_TX_SOPRANO);
    ex.tx_lattice_tenor = add_lattice(&IIR_TX_WIDE_TREBLE);
    ex.tx_lattice_bass = add_lattice(&IIR_TX_WIDE_BASS);
    ex.tx_lattice_fm = add_lattice(&IIR_TX_2k7_FM);
    {
        static float zb[5 * 20], zd[5 * 4];
        for (int m = 1; m <= 5; m++) {
            for (int i = 0; i < 20; i++) zb[(m - 1) * 20 + i] = mag_coeffs[m][i];
            for (int i = 0; i < 4; i++) zd[(m - 1) * 4 + i] = FirZoomFFTDecimate[m].pCoeffs[i];
        }
        ex.zoom_biquad_array = add_array(zb, 5 * 20);
        ex.zoom_decim_array = add_array(zd, 5 * 4);
    }
    {
        static float dds[DDS_TBL_SIZE];
        for (int i = 0; i < DDS_TBL_SIZE; i++) dds[i] = (float)DDS_TABLE[i];
        ex.dds_table_array = add_array(dds, DDS_TBL_SIZE);
    }

    /* von_Hann_1024 is a function-local constant of ui_spectrum.c (ui_spectrum.c:362); the
     * Makefile extracts it into a raw float file that is appended here. */
    static float win[1024];
    ex.spectrum_window_array = -1;
    if (argc >= 3) {
        FILE *wf = fopen(argv[2], "rb");
        if (wf && fread(win, sizeof(float), 1024, wf) == 1024) ex.spectrum_window_array = add_array(win, 1024);
        if (wf) fclose(wf);
    }

    uhsdr_tbl_header_t h;
    memset(&h, 0, sizeof(h));
    h.magic = UHSDR_TABLES_MAGIC; h.version = UHSDR_TABLES_VERSION;
    uint32_t off = sizeof(h);
    h.num_arrays = n_arr; h.arrays_off = off; off += n_arr * sizeof(uhsdr_tbl_array_t);
    h.num_paths = AUDIO_FILTER_PATH_NUM; h.paths_off = off; off += sizeof(paths);
    h.num_filters = AUDIO_FILTER_NUM; h.filters_off = off; off += sizeof(filters);
    h.num_lattices = n_lat; h.lattices_off = off; off += n_lat * sizeof(uhsdr_tbl_lattice_t);
    h.num_interps = n_int; h.interps_off = off; off += n_int * sizeof(uhsdr_tbl_interp_t);
    h.extras_off = off; off += sizeof(ex);
    static uhsdr_tbl_array_t dir[MAX_ARRAYS];
    for (int i = 0; i < n_arr; i++) { dir[i].offset = off; dir[i].count = arr_cnt[i]; off += arr_cnt[i] * sizeof(float); }
    h.total_bytes = off;

    FILE *f = fopen(argv[1], "wb");
    if (!f) { perror(argv[1]); return 1; }
    fwrite(&h, sizeof(h), 1, f);
    fwrite(dir, sizeof(uhsdr_tbl_array_t), n_arr, f);
    fwrite(paths, sizeof(paths), 1, f);
    fwrite(filters, sizeof(filters), 1, f);
    fwrite(lat_row, sizeof(uhsdr_tbl_lattice_t), n_lat, f);
    fwrite(int_row, sizeof(uhsdr_tbl_interp_t), n_int, f);
    fwrite(&ex, sizeof(ex), 1, f);
    for (int i = 0; i < n_arr; i++) fwrite(arr_ptr[i], sizeof(float), arr_cnt[i], f);
    fclose(f);
    fprintf(stderr, "wrote %s: %u bytes, %d arrays, %d lattices, %d interpolators\n", argv[1], off, n_arr, n_lat, n_int);
    return 0;
}
