/*
 * examples/host.c -- the "host C code calling CUDA through a thin C-ABI layer" of the north star: a plain C program that stands
 * where the firmware's MchfHw_Codec_HandleBlock -> AudioDriver_I2SCallback (uhsdr_hw_i2s.c:110, audio_driver.c:2962) stands,
 * for many channels at once.  It loads the coefficient-table blob, creates ONE handle over the GPUs named on the command line,
 * configures an alternating USB / LSB plan, pushes synthetic I/Q blocks through uhsdr_multi_rx_process and a two-tone
 * microphone signal through uhsdr_multi_tx_process, and prints a checksum plus the side outputs of the first channels.
 *
 *   cc -O2 -Iinclude examples/host.c -o examples/host -Luhsdr_b200/csrc -luhsdr_b200 -Wl,-rpath,$PWD/uhsdr_b200/csrc -lm
 *   examples/host uhsdr_b200/data/uhsdr_tables.bin 64 40 0 [1 2 ...]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "uhsdr_b200.h"

static void *read_file(const char *path, size_t *bytes)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *buf = malloc((size_t)n);
    if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    *bytes = (size_t)n;
    return buf;
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s tables.bin num_channels nblocks device [device ...]\n", argv[0]); return 2; }
    const int nch = atoi(argv[2]), nblocks = atoi(argv[3]), ndev = argc - 4;
    int devices[16];
    for (int i = 0; i < ndev && i < 16; i++) devices[i] = atoi(argv[4 + i]);
    size_t tbytes = 0;
    void *tables = read_file(argv[1], &tbytes);
    if (!tables || uhsdr_tables_validate(tables, tbytes) != UHSDR_OK) { fprintf(stderr, "bad table blob: %s\n", uhsdr_last_error(NULL)); return 1; }

    uhsdr_multi_t *m = NULL;
    int rc = uhsdr_multi_create(&m, nch, devices, ndev, tables, tbytes);
    if (rc != UHSDR_OK) { fprintf(stderr, "uhsdr_multi_create: %s (%s)\n", uhsdr_strerror(rc), uhsdr_last_error(NULL)); return 1; }

    /* AudioDriver_SetProcessingChain for every channel: even channels USB on FilterPathInfo[35], odd ones LSB on [38] */
    uhsdr_chan_cfg_t usb, lsb;
    uhsdr_default_chan_cfg(&usb);
    uhsdr_default_chan_cfg(&lsb);
    lsb.dmod_mode = UHSDR_DEMOD_LSB; lsb.filter_path = 38;
    if ((rc = uhsdr_multi_configure_channels_strided(m, 0, (nch + 1) / 2, 2, &usb, 1)) != UHSDR_OK ||
        (nch > 1 && (rc = uhsdr_multi_configure_channels_strided(m, 1, nch / 2, 2, &lsb, 1)) != UHSDR_OK)) {
        fprintf(stderr, "configure: %s (%s)\n", uhsdr_strerror(rc), uhsdr_multi_last_error(m)); return 1;
    }

    /* one tone 1 kHz above (USB) / below (LSB) the +12 kHz IF per channel, the firmware's int32 left-justified sample format */
    const size_t ns = (size_t)nblocks * UHSDR_BLOCK_SIZE;
    uhsdr_iq_sample_t *iq = malloc(sizeof(*iq) * nch * ns);
    uhsdr_audio_sample_t *audio = malloc(sizeof(*audio) * nch * ns), *mic = malloc(sizeof(*mic) * nch * ns);
    uhsdr_iq_sample_t *txiq = malloc(sizeof(*txiq) * nch * ns);
    for (int c = 0; c < nch; c++) {
        const double f = 12000.0 + ((c & 1) ? -1.0 : 1.0) * (1000.0 + 10.0 * (c % 32));
        for (size_t n = 0; n < ns; n++) {
            const double ph = 2.0 * M_PI * f * (double)n / UHSDR_SAMPLE_RATE;
            iq[c * ns + n].l = (int32_t)lrint(3000.0 * cos(ph) * 65536.0);
            iq[c * ns + n].r = (int32_t)lrint(3000.0 * sin(ph) * 65536.0);
            mic[c * ns + n].l = (int32_t)lrint(8000.0 * (sin(2.0 * M_PI * 700.0 * n / UHSDR_SAMPLE_RATE) + sin(2.0 * M_PI * 1900.0 * n / UHSDR_SAMPLE_RATE)) * 65536.0);
            mic[c * ns + n].r = 0;
        }
    }
    if ((rc = uhsdr_multi_rx_process(m, iq, audio, nblocks, NULL)) != UHSDR_OK) { fprintf(stderr, "rx: %s (%s)\n", uhsdr_strerror(rc), uhsdr_multi_last_error(m)); return 1; }
    if ((rc = uhsdr_multi_tx_process(m, mic, txiq, nblocks, NULL)) != UHSDR_OK) { fprintf(stderr, "tx: %s (%s)\n", uhsdr_strerror(rc), uhsdr_multi_last_error(m)); return 1; }

    unsigned long long sum_rx = 0, sum_tx = 0;
    for (size_t i = 0; i < nch * ns; i++) {
        sum_rx = sum_rx * 1099511628211ull + (unsigned)audio[i].l;
        sum_tx = sum_tx * 1099511628211ull + (unsigned)txiq[i].l + 31u * (unsigned)txiq[i].r;
    }
    uhsdr_chan_status_t st[2];
    uhsdr_multi_get_status(m, 0, nch > 1 ? 2 : 1, st);
    double p0 = 0.0;
    for (size_t n = ns / 2; n < ns; n++) p0 += (double)(audio[n].l >> 16) * (double)(audio[n].l >> 16);
    printf("backend=%s abi=%d devices=%d channels=%d blocks=%d\n", uhsdr_b200_backend(), uhsdr_b200_abi_version(), uhsdr_multi_num_devices(m), nch, nblocks);
    printf("rx_checksum=%016llx tx_checksum=%016llx ch0_audio_rms=%.1f ch0_agc_action=%d ch0_blocks=%lld\n", sum_rx, sum_tx, sqrt(p0 / (double)(ns - ns / 2)),
           st[0].agc_action, (long long)st[0].blocks_processed);
    uhsdr_multi_destroy(m);
    free(iq); free(audio); free(mic); free(txiq); free(tables);
    return 0;
}
