/* abi_probe.c -- include/srcdsp_b200.h consumed from plain C (C99, -pedantic): the boundary a cgo / JNI / ctypes binding
 * would see.  Without a GPU every constructor must refuse with SRCDSP_E_NOGPU and a message (there is no CPU
 * fallback); with one, a one-channel decimator runs the survey's known-answer test (63 taps of 100,
 * decimate by 8, constant input (1000, -500) -> (1538, -770) after warm-up).  Prints one line per check. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "srcdsp_b200.h"

int main(void)
{
    srcdsp_mixer_t mx = NULL;
    srcdsp_dec_t dec = NULL;
    srcdsp_up_t up = NULL;
    int ndev = -1, st, k;
    int32_t taps[63];
    int16_t *in, *out;
    const size_t n = 8 * 64;

    printf("version %d\n", srcdsp_version());
    st = srcdsp_device_count(&ndev);
    printf("device_count status %d count %d\n", st, ndev);
    st = srcdsp_mixer_create(&mx, 0, 1, 4096);
    if (st == SRCDSP_E_NOGPU) {
        printf("mixer_create refused: %d handle %s message %s\n", st, mx ? "set" : "null", strlen(srcdsp_last_error()) ? "yes" : "no");
        st = srcdsp_dec_create(&dec, 0, 1, 8);
        printf("dec_create refused: %d handle %s\n", st, dec ? "set" : "null");
        st = srcdsp_up_create(&up, 0, 1, 8);
        printf("up_create refused: %d handle %s\n", st, up ? "set" : "null");
        printf("null handle: %d %d\n", srcdsp_dec_reset(NULL), srcdsp_mixer_destroy(NULL));
        return 0;
    }
    if (st != SRCDSP_OK) {
        printf("mixer_create failed: %d %s\n", st, srcdsp_last_error());
        return 1;
    }
    for (k = 0; k < 63; ++k) taps[k] = 100;
    if (srcdsp_dec_create(&dec, 0, 1, 8) != SRCDSP_OK || srcdsp_dec_set_coeffs(dec, taps, 63, 0) != SRCDSP_OK) {
        printf("decimator set-up failed: %s\n", srcdsp_last_error());
        return 1;
    }
    in = (int16_t *)malloc(n * 2 * sizeof(int16_t));
    out = (int16_t *)malloc(n / 8 * 2 * sizeof(int16_t));
    for (k = 0; k < (int)n; ++k) in[2 * k] = 1000, in[2 * k + 1] = -500;
    st = srcdsp_dec_step(dec, in, n, n, out, n / 8);
    printf("dec_step status %d last (%d, %d)\n", st, out[2 * (n / 8 - 1)], out[2 * (n / 8 - 1) + 1]);
    free(in);
    free(out);
    srcdsp_dec_destroy(dec);
    srcdsp_mixer_destroy(mx);
    return st == SRCDSP_OK ? 0 : 1;
}
