/*
 * uhsdr_tables.h -- binary layout of the coefficient-table blob handed to uhsdr_engine_create().
 *
 * The engine keeps the reference's table-driven filter selection but does not embed the tables:
 * the blob is a flat serialisation of what the firmware links from
 *   mchf-eclipse/drivers/audio/audio_filter.c:47-80    FilterInfo[31]       (bandwidth ids/widths)
 *   mchf-eclipse/drivers/audio/audio_filter.c:147-922  FilterPathInfo[87]   (FilterPathDescriptor rows)
 *   mchf-eclipse/drivers/audio/filters/ *.c            FIR / lattice-IIR / interpolator coefficient arrays
 *   mchf-eclipse/drivers/audio/filters/iq_tx_filter.c:744  iq_tx_wide (201-tap TX Hilbert pair)
 *   mchf-eclipse/drivers/audio/audio_driver.c:195,198  NR_decimate_coeffs[4], NR_interpolate_coeffs[40]
 *   mchf-eclipse/drivers/audio/audio_nr.c:76           SQRT_von_Hann_256
 *   mchf-eclipse/drivers/audio/audio_driver.c:1932-1953 demod_sam_const (all-pass Hilbert network)
 * A maintainer produces it on the UHSDR side with ~100 lines of C that walk FilterPathInfo[]
 * (INTEGRATION.md shows the code; oracle/dump_tables.c is that code for this repo's tests).
 *
 * All integers little-endian int32/uint32, all coefficients IEEE float32, offsets in bytes from
 * the start of the blob.  Pointer-valued fields of FilterPathDescriptor become indices (-1 = NULL).
 */
#ifndef UHSDR_TABLES_H
#define UHSDR_TABLES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UHSDR_TABLES_MAGIC   0x42545355u /* "USTB" */
#define UHSDR_TABLES_VERSION 2u

typedef struct {
    uint32_t magic;
    uint32_t version;
    uint32_t total_bytes;
    uint32_t num_arrays;    uint32_t arrays_off;    /* uhsdr_tbl_array_t[num_arrays]     */
    uint32_t num_paths;     uint32_t paths_off;     /* uhsdr_tbl_path_t[num_paths]       */
    uint32_t num_filters;   uint32_t filters_off;   /* uhsdr_tbl_filter_t[num_filters]   */
    uint32_t num_lattices;  uint32_t lattices_off;  /* uhsdr_tbl_lattice_t[num_lattices] */
    uint32_t num_interps;   uint32_t interps_off;   /* uhsdr_tbl_interp_t[num_interps]   */
    uint32_t extras_off;                            /* uhsdr_tbl_extras_t                */
} uhsdr_tbl_header_t;

/* One float32 coefficient array. */
typedef struct { uint32_t offset; uint32_t count; } uhsdr_tbl_array_t;

/* FilterPathDescriptor, audio_filter.h:108-136. */
typedef struct {
    int32_t id;               /* .id: bandwidth id, index into the filter table                  */
    int32_t mode_mask;        /* .mode: bit (1 << FILTER_MODE_x)                                 */
    int32_t filter_select_id; /* .filter_select_id                                               */
    int32_t fir_numtaps;      /* .FIR_numTaps                                                    */
    int32_t fir_i_array;      /* .FIR_I_coeff_file  -> array index                               */
    int32_t fir_q_array;      /* .FIR_Q_coeff_file  -> array index                               */
    int32_t fir_is_new_coeffs;/* FIR_I_coeff_file == i_rx_new_coeffs (audio_driver.c:2719)       */
    int32_t dec_array;        /* .dec->pCoeffs -> array index, -1 when .dec == NULL              */
    int32_t dec_numtaps;      /* .dec->numTaps                                                   */
    int32_t sample_rate_dec;  /* .sample_rate_dec: decimation factor 4 / 2 / 1                   */
    int32_t pre_lattice;      /* .pre_instance   -> lattice index, -1 = none                     */
    int32_t interpolate;      /* .interpolate    -> interp index, -1 = none                      */
    int32_t aa_lattice;       /* .iir_instance   -> lattice index, -1 = none                     */
    int32_t offset_hz;        /* .offset                                                         */
    char    name[24];         /* .name                                                           */
} uhsdr_tbl_path_t;

/* FilterDescriptor, audio_filter.h:96-101. */
typedef struct { int32_t id; int32_t width; char name[12]; } uhsdr_tbl_filter_t;

/* arm_iir_lattice_instance_f32: k has num_stages, v has num_stages+1 entries. */
typedef struct { int32_t num_stages; int32_t k_array; int32_t v_array; } uhsdr_tbl_lattice_t;

/* arm_fir_interpolate_instance_f32 as stored in the tables: the .phaseLength FIELD is what the
 * firmware passes as numTaps (audio_driver.c:1213-1217), so num_coeffs == phase_length_field. */
typedef struct { int32_t L; int32_t phase_length_field; int32_t coeff_array; int32_t num_coeffs; } uhsdr_tbl_interp_t;

typedef struct {
    int32_t nr_decimate_array;     /* 4 taps,  audio_driver.c:195                                */
    int32_t nr_interpolate_array;  /* 40 taps, audio_driver.c:198                                */
    int32_t sqrt_hann_256_array;   /* audio_nr.c:76                                              */
    int32_t spectrum_window_array; /* von_Hann_1024, ui_spectrum.c:362 (1024 floats)             */
    int32_t sam_c0_array;          /* 7 floats, audio_driver.c:1934                              */
    int32_t sam_c1_array;          /* 7 floats, audio_driver.c:1944                              */
    int32_t fm_squelch_lattice;    /* IIR_15k_hpf, audio_driver.c:481-483                        */
    int32_t tx_hilbert_i_array;    /* iq_tx_wide.i (201), audio_filter.c:1239-1252               */
    int32_t tx_hilbert_q_array;    /* iq_tx_wide.q                                               */
    int32_t tx_hilbert_numtaps;
    int32_t tx_lattice_soprano;    /* IIR_TX_SOPRANO,     tx_processor.c:92-102                  */
    int32_t tx_lattice_tenor;      /* IIR_TX_WIDE_TREBLE                                         */
    int32_t tx_lattice_bass;       /* IIR_TX_WIDE_BASS                                           */
    int32_t tx_lattice_fm;         /* IIR_TX_2k7_FM, tx_processor.c:104-107 (-1 in blobs older than this field) */
    int32_t dds_table_array;       /* DDS_TABLE (1024 x int16 as floats), softdds/dds_table.c:19 */
    int32_t zoom_biquad_array;     /* mag_coeffs[1..5], 5 x (4 stages x 5) floats, audio_driver.c:204-363 */
    int32_t zoom_decim_array;      /* FirZoomFFTDecimate[1..5].pCoeffs, 5 x 4 taps, fir_rx_decimate_4.c:108-181 */
    int32_t reserved[4];
} uhsdr_tbl_extras_t;

#ifdef __cplusplus
}
#endif
#endif /* UHSDR_TABLES_H */
