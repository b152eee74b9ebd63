/*
 * srcdsp_oracle.h -- CPU restatement of the SrcDsp DDC hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain-C, single-threaded restatement of the reference's
 * integer arithmetic.  It is NOT part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, link or call it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path (SURVEY.md
 * section 4), so the oracle is pinned against the reference itself: oracle/_ref/ is the
 * unmodified reference headers compiled from /root/reference (see oracle/Makefile and
 * oracle/ref_harness.cpp), tests/test_oracle.py checks this restatement bit-for-bit against
 * it whenever it is present, and against tests/golden/ *.npz fixtures which were generated
 * from it by tests/golden/make_golden.py.
 *
 * All sample buffers are interleaved I/Q int16 (byte-identical to
 * std::vector<std::complex<int16_t>>::data()).
 */
#ifndef SRCDSP_ORACLE_H
#define SRCDSP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic input (counter based, identical on device: csrc/synth.cuh) ---- */
uint32_t orc_hash32(uint32_t seed, uint32_t channel, uint64_t n);
void orc_synth_fill(int16_t *iq, uint32_t seed, uint32_t channel, uint64_t n0, size_t n,
                    int amp_shift);

/* ---- mixer: mixers.h:51-67 (setFrequency), :91-98 (adjustFrequency), :149-160 (table),
 *      :168-188 (step) ---- */
void orc_mixer_table(int16_t *table, unsigned n_table);
int orc_mixer_set_frequency(float lo_freq, unsigned n_table);
float orc_mixer_adjust_nominal(float nominal, float adjust);
void orc_mixer_step(const int16_t *table, unsigned n_table, int *phi, int freq,
                    const int16_t *in_iq, int16_t *out_iq, size_t n);

/* ---- decimator: dsptl_dnsampling_filters.h:114-134 (setCoeffs), :172-220 (step)
 *      == dnsampling_filters.h:83-97, :128-172 ---- */
int orc_dec_coeff_scaling(const int32_t *taps, int ntaps);
/* history: (ntaps-1) complex samples, oldest first; updated in place.  n_in % M == 0. */
unsigned orc_decf_coeff_scaling(const float *taps, int ntaps);
void orc_decf_step(const float *taps, int ntaps, int M, unsigned shift, float *history_iq, const float *in_iq, size_t n_in,
                   float *out_iq);
void orc_dec_step(const int32_t *taps, int ntaps, int M, unsigned shift, int16_t *history_iq,
                  const int16_t *in_iq, size_t n_in, int16_t *out_iq);

/* ---- non-decimating FIR: filters.h:85-97, :130-169 ---- */
void orc_fir_step(const int32_t *taps, int ntaps, unsigned shift, int16_t *history_iq,
                  const int16_t *in_iq, size_t n_in, int16_t *out_iq);

/* ---- upsampler: upsampling_filters.h:107-126 (setCoefficients), :149-233, :240-323 ---- */
int orc_up_left_shift_factor(int L);             /* round(log2 L) */
int orc_up_length(const int32_t *taps, int ntaps); /* ntaps minus trailing zero taps */
/* history: (ntaps/L - 1) complex samples, oldest first; updated in place.
 * Produces L*(n_in + n_flush) outputs; the n_flush trailing inputs are zeros
 * (flush=true  <=>  n_flush = length / L). */
void orc_up_step(const int32_t *taps, int ntaps, int L, unsigned shift, int16_t *history_iq,
                 const int16_t *in_iq, size_t n_in, size_t n_flush, int16_t *out_iq);

/* ---- correlator: correlators.h:143-155, :167-194, :209-303 ---- */
typedef struct {
    int N, S, top, coeff_scaling;
    uint32_t coeffs_energy, corr_value[3], energy_value[3];
    double threshold_factor;
    int32_t *history; /* [N*S][2] circular */
    int32_t *coeffs;  /* [N][2] conjugated pattern */
    int16_t *bits;    /* [N][2] */
} orc_corr_t;
int orc_corr_init(orc_corr_t *c, int N, int S);
void orc_corr_free(orc_corr_t *c);
void orc_corr_reset(orc_corr_t *c);
int orc_corr_set_pattern(orc_corr_t *c, const int32_t *pattern_iq, double threshold_coeff); /* -1: energy assert */
/* returns 1 and *corr_index when a peak was found (the rest of the block is not consumed) */
int orc_corr_step(orc_corr_t *c, const int16_t *in_iq, size_t n, int *corr_index);

/* ---- shared scalar helpers: dsp_complex.cpp:63-73, dsp_complex.h:83-108 ---- */
int16_t orc_limit_scale16(int32_t v, unsigned shift);     /* symmetric  +-32767        */
int16_t orc_limit_scale_asym(int32_t v, unsigned shift);  /* [-32768, 32767]           */

#ifdef __cplusplus
}
#endif
#endif
