/*
 * srcdsp_oracle.c -- CPU restatement of the SrcDsp DDC hot path.  TEST INFRASTRUCTURE ONLY;
 * see srcdsp_oracle.h for the rules and the parity-pinning statement.
 *
 * Every function cites the reference file:line it restates.  Integer accumulation is done in
 * uint32_t so that overflow wraps exactly as the reference's int32_t does on x86 (the
 * reference documents overflow as the caller's problem: dsptl_dnsampling_filters.h:36-38).
 */
#include "srcdsp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* constants.h:21 */
static const double ORC_PI = 3.141592653589793238462643383279502884197169399375105820974944592307816406286;

/* ------------------------------------------------------------------------------------------ */
/* synthetic input: "lowbias32" finaliser over (seed, channel, n).  Same function on device. */
uint32_t orc_hash32(uint32_t seed, uint32_t channel, uint64_t n)
{
    uint32_t x = seed ^ (channel * 0x9E3779B1u) ^ ((uint32_t)n * 0x85EBCA6Bu) ^
                 ((uint32_t)(n >> 32) * 0xC2B2AE35u);
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

void orc_synth_fill(int16_t *iq, uint32_t seed, uint32_t channel, uint64_t n0, size_t n,
                    int amp_shift)
{
    for (size_t k = 0; k < n; ++k) {
        uint32_t h = orc_hash32(seed, channel, n0 + k);
        int16_t re = (int16_t)(h & 0xFFFFu);
        int16_t im = (int16_t)(h >> 16);
        iq[2 * k] = (int16_t)(re >> amp_shift);
        iq[2 * k + 1] = (int16_t)(im >> amp_shift);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* dsp_complex.cpp:63-73 -- arithmetic shift, then symmetric clamp to +-INT16_MAX */
int16_t orc_limit_scale16(int32_t v, unsigned shift)
{
    int32_t a = v >> shift;
    if (a > 32767) a = 32767;
    if (a < -32767) a = -32767;
    return (int16_t)a;
}

/* dsp_complex.h:83-108 -- arithmetic shift, then clamp to [lowest, max] of int16 */
int16_t orc_limit_scale_asym(int32_t v, unsigned shift)
{
    int32_t a = v >> shift;
    if (a > 32767) a = 32767;
    else if (a < -32768) a = -32768;
    return (int16_t)a;
}

/* ------------------------------------------------------------------------------------------ */
/* mixers.h:149-160 -- sine table, amplitude INT16_MAX >> 1, truncating cast from double */
void orc_mixer_table(int16_t *table, unsigned n_table)
{
    const int16_t max = (int16_t)(INT16_MAX >> 1);
    for (unsigned k = 0; k < n_table; ++k)
        table[k] = (int16_t)(max * sin(2 * ORC_PI * (double)k / n_table));
}

/* mixers.h:51-67 -- float arithmetic, round half away from zero */
int orc_mixer_set_frequency(float lo_freq, unsigned n_table)
{
    int16_t freq;
    if (lo_freq >= 0)
        freq = (int16_t)roundf(lo_freq * (float)n_table / 2);
    else {
        freq = (int16_t)roundf((float)n_table - roundf(-lo_freq * (float)n_table / 2));
        if (freq == (int16_t)n_table) freq = 0;
    }
    return freq;
}

/* mixers.h:91-98 -- wrap of the nominal frequency (the caller then re-quantises it) */
float orc_mixer_adjust_nominal(float nominal, float adjust)
{
    nominal += adjust;
    if (nominal > 1) nominal -= 2;
    if (nominal < -1) nominal += 2;
    return nominal;
}

/* mixers.h:168-188 with dsp_complex.cpp:31-37 (product) and :63-73 (limitScale16(., 14)) */
void orc_mixer_step(const int16_t *table, unsigned n_table, int *phi, int freq,
                    const int16_t *in_iq, int16_t *out_iq, size_t n)
{
    unsigned p = (unsigned)*phi;
    for (size_t k = 0; k < n; ++k) {
        int32_t c = table[(p + n_table / 4) % n_table];
        int32_t s = table[p];
        int32_t xr = in_iq[2 * k], xi = in_iq[2 * k + 1];
        int32_t r = xr * c - xi * s;
        int32_t i = xi * c + s * xr;
        out_iq[2 * k] = orc_limit_scale16(r, 14);
        out_iq[2 * k + 1] = orc_limit_scale16(i, 14);
        p = (p + (unsigned)freq) % n_table;
    }
    *phi = (int)p;
}

/* ------------------------------------------------------------------------------------------ */
/* dsptl_dnsampling_filters.h:126-132 == dnsampling_filters.h:90-96 */
int orc_dec_coeff_scaling(const int32_t *taps, int ntaps)
{
    double sum = 0;
    for (int k = 0; k < ntaps; ++k) sum += abs(taps[k]);
    return (int)floor(log2(sum));
}

/* out[i] = limitScale16( sum_k c[k] * xx[i*M - k], shift ),  xx = history ++ input.
 * dsptl_dnsampling_filters.h:188-219.  The reference needs n_in >= ntaps-1 to save its
 * history (:218-219); for shorter blocks this restatement keeps the last ntaps-1 samples of
 * history ++ input, which is what a longer call split in two would have produced. */
static void fir_core(const int32_t *taps, int ntaps, int M, unsigned shift, int16_t *hist,
                     const int16_t *in, size_t n_in, int16_t *out)
{
    const size_t H = (size_t)(ntaps - 1);
    int16_t *xx = (int16_t *)malloc((H + n_in + 1) * 2 * sizeof(int16_t));
    memcpy(xx, hist, H * 2 * sizeof(int16_t));
    memcpy(xx + 2 * H, in, n_in * 2 * sizeof(int16_t));
    for (size_t j = 0; j < n_in; j += (size_t)M) {
        uint32_t ar = 0, ai = 0;
        const int16_t *x = xx + 2 * (H + j);
        for (int k = 0; k < ntaps; ++k) {
            ar += (uint32_t)taps[k] * (uint32_t)(int32_t)x[-2 * k];
            ai += (uint32_t)taps[k] * (uint32_t)(int32_t)x[-2 * k + 1];
        }
        out[2 * (j / M)] = orc_limit_scale16((int32_t)ar, shift);
        out[2 * (j / M) + 1] = orc_limit_scale16((int32_t)ai, shift);
    }
    memmove(hist, xx + 2 * n_in, H * 2 * sizeof(int16_t));
    free(xx);
}

void orc_dec_step(const int32_t *taps, int ntaps, int M, unsigned shift, int16_t *history_iq,
                  const int16_t *in_iq, size_t n_in, int16_t *out_iq)
{
    fir_core(taps, ntaps, M, shift, history_iq, in_iq, n_in, out_iq);
}

/* ------------------------------------------------------------------------------------------ */
/* The float instantiation FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M>.
 * dsptl_dnsampling_filters.h:128-132: `abs(coeff[index])` inside namespace dsptl is ::abs(int) -- the float tap is
 * truncated to int first -- and coeffScaling is an unsigned that receives int(floor(log2(sum))).  A sum of 0 makes the
 * reference's value undefined; as compiled by g++ for x86-64 it is INT_MIN, returned here as 0x80000000. */
/* complex<int32_t>(y) of a float sum as the reference's x86-64 build does it (cvttss2si): truncation, and the
   "integer indefinite" value 0x80000000 for NaN and for every |y| >= 2^31 -- positive overflow included, which then
   clamps to -32767.  (A C cast is undefined there; written out so that the restatement does not depend on it.) */
static int32_t orc_cvttss2si(float y)
{
    return (fabsf(y) < 2147483648.0f) ? (int32_t)y : INT32_MIN;
}

unsigned orc_decf_coeff_scaling(const float *taps, int ntaps)
{
    double sum = 0;
    for (int k = 0; k < ntaps; ++k) sum += abs((int)taps[k]);
    return sum >= 1 ? (unsigned)(int)floor(log2(sum)) : 0x80000000u;
}

/* dsptl_dnsampling_filters.h:188-219 with float types: y += coeff[k] * xx[i*M - k] is one rounded float multiply and
 * one rounded float add per component, in tap order (:195-210; no FMA contraction: see the pragma / build flags);
 * :214 converts y to complex<int32_t> (truncation), applies limitScale16 (dsp_complex.cpp:63-73) with the shift
 * count taken mod 32 (x86 sar) and converts the int16 pair back to float. */
#pragma STDC FP_CONTRACT OFF
void orc_decf_step(const float *taps, int ntaps, int M, unsigned shift, float *history_iq, const float *in_iq, size_t n_in,
                   float *out_iq)
{
    const size_t H = (size_t)(ntaps - 1);
    float *xx = (float *)malloc((H + n_in + 1) * 2 * sizeof(float));
    memcpy(xx, history_iq, H * 2 * sizeof(float));
    memcpy(xx + 2 * H, in_iq, n_in * 2 * sizeof(float));
    for (size_t j = 0; j < n_in; j += (size_t)M) {
        volatile float ar = 0.f, ai = 0.f; /* volatile: every partial sum is rounded to float */
        const float *x = xx + 2 * (H + j);
        for (int k = 0; k < ntaps; ++k) {
            volatile float pr = taps[k] * x[-2 * k], pi = taps[k] * x[-2 * k + 1];
            ar = ar + pr;
            ai = ai + pi;
        }
        out_iq[2 * (j / M)] = (float)orc_limit_scale16(orc_cvttss2si(ar), shift & 31u);
        out_iq[2 * (j / M) + 1] = (float)orc_limit_scale16(orc_cvttss2si(ai), shift & 31u);
    }
    memmove(history_iq, xx + 2 * n_in, H * 2 * sizeof(float));
    free(xx);
}

/* filters.h:130-169 -- circular-buffer FIR; in age order it is the M = 1 decimator */
void orc_fir_step(const int32_t *taps, int ntaps, unsigned shift, int16_t *history_iq,
                  const int16_t *in_iq, size_t n_in, int16_t *out_iq)
{
    fir_core(taps, ntaps, 1, shift, history_iq, in_iq, n_in, out_iq);
}

/* ------------------------------------------------------------------------------------------ */
/* upsampling_filters.h:120 */
int orc_up_left_shift_factor(int L) { return (int)round(log2((double)L)); }

/* upsampling_filters.h:122-123 */
int orc_up_length(const int32_t *taps, int ntaps)
{
    int length = ntaps;
    while (length > 0 && taps[length - 1] == 0) --length;
    return length;
}

/* out[j*L + p] = limitScale<cs16>( sum_{i<H} c[p + i*L] * xx[j - i], shift ), H = ntaps/L.
 * upsampling_filters.h:163-194 (and the flush loop :196-231 with zero inputs).  The circular
 * buffer + top of the reference is restated in age order (top itself is unobservable). */
void orc_up_step(const int32_t *taps, int ntaps, int L, unsigned shift, int16_t *history_iq,
                 const int16_t *in_iq, size_t n_in, size_t n_flush, int16_t *out_iq)
{
    const int H = ntaps / L;
    const size_t Hm1 = (size_t)(H - 1);
    const size_t n_tot = n_in + n_flush;
    int16_t *xx = (int16_t *)calloc((Hm1 + n_tot + 1) * 2, sizeof(int16_t));
    memcpy(xx, history_iq, Hm1 * 2 * sizeof(int16_t));
    memcpy(xx + 2 * Hm1, in_iq, n_in * 2 * sizeof(int16_t));
    for (size_t j = 0; j < n_tot; ++j) {
        const int16_t *x = xx + 2 * (Hm1 + j);
        for (int p = 0; p < L; ++p) {
            uint32_t ar = 0, ai = 0;
            for (int i = 0; i < H; ++i) {
                uint32_t c = (uint32_t)taps[p + i * L];
                ar += c * (uint32_t)(int32_t)x[-2 * i];
                ai += c * (uint32_t)(int32_t)x[-2 * i + 1];
            }
            out_iq[2 * (j * (size_t)L + (size_t)p)] = orc_limit_scale_asym((int32_t)ar, shift);
            out_iq[2 * (j * (size_t)L + (size_t)p) + 1] = orc_limit_scale_asym((int32_t)ai, shift);
        }
    }
    memmove(history_iq, xx + 2 * n_tot, Hm1 * 2 * sizeof(int16_t));
    free(xx);
}

/* ---- FixedPatternCorrelator<int16_t, int32_t, N, S> --------------------------
 * Sequential restatement with the reference's own state: circular history of N*S slots + insertion
 * index `top`, the two 3-deep registers, the bit samples.  correlators.h:143-155 (reset), :167-194
 * (setPattern), :209-303 (step).  All sums are mod 2^32 (the reference's int32_t / uint32_t). */
void orc_corr_reset(orc_corr_t *c)
{
    c->top = 0;
    memset(c->corr_value, 0, sizeof c->corr_value);
    memset(c->energy_value, 0, sizeof c->energy_value);
    memset(c->history, 0, sizeof(int32_t) * 2 * (size_t)c->N * (size_t)c->S);
    memset(c->bits, 0, sizeof(int16_t) * 2 * (size_t)c->N);
}

int orc_corr_init(orc_corr_t *c, int N, int S)
{
    memset(c, 0, sizeof *c);
    c->N = N;
    c->S = S;
    c->history = (int32_t *)calloc((size_t)2 * N * S, sizeof(int32_t));
    c->coeffs = (int32_t *)calloc((size_t)2 * N, sizeof(int32_t));
    c->bits = (int16_t *)calloc((size_t)2 * N, sizeof(int16_t));
    if (!c->history || !c->coeffs || !c->bits) return -1;
    orc_corr_reset(c);
    return 0;
}

void orc_corr_free(orc_corr_t *c)
{
    free(c->history);
    free(c->coeffs);
    free(c->bits);
    memset(c, 0, sizeof *c);
}

int orc_corr_set_pattern(orc_corr_t *c, const int32_t *pattern_iq, double threshold_coeff)
{
    double tmp = 0;
    for (int k = 0; k < c->N; ++k) {
        c->coeffs[2 * k] = pattern_iq[2 * k];
        c->coeffs[2 * k + 1] = -pattern_iq[2 * k + 1]; /* conjugate, :172-176 */
        const uint32_t e = (uint32_t)c->coeffs[2 * k] * (uint32_t)c->coeffs[2 * k] +
                           (uint32_t)c->coeffs[2 * k + 1] * (uint32_t)c->coeffs[2 * k + 1];
        tmp += (double)(int32_t)e;
    }
    if (!(tmp <= 1073217600.0)) return -1; /* assert at :183 */
    c->coeffs_energy = (uint32_t)tmp;
    c->threshold_factor = threshold_coeff * sqrt((double)c->coeffs_energy);
    c->coeff_scaling = (int)floor(log2(sqrt((double)c->coeffs_energy)));
    return 0;
}

int orc_corr_step(orc_corr_t *c, const int16_t *in_iq, size_t n, int *corr_index)
{
    const int N = c->N, S = c->S, H = N * S;
    for (size_t index = 0; index < n; ++index) {
        c->history[2 * c->top] = in_iq[2 * index];
        c->history[2 * c->top + 1] = in_iq[2 * index + 1];
        uint32_t re = 0, im = 0, en = 0;
        c->energy_value[2] = c->energy_value[1];
        c->energy_value[1] = c->energy_value[0];
        /* every S-th slot: downwards from top (newest -> coeffs[N-1-k]), then the wrapped part (:226-235) */
        int k, h;
        for (k = 0; (h = c->top - k * S) >= 0; ++k) {
            const uint32_t xr = (uint32_t)c->history[2 * h], xi = (uint32_t)c->history[2 * h + 1];
            const uint32_t cr = (uint32_t)c->coeffs[2 * (N - 1 - k)], ci = (uint32_t)c->coeffs[2 * (N - 1 - k) + 1];
            re += xr * cr - xi * ci;
            im += xr * ci + xi * cr;
            en += xr * xr + xi * xi;
        }
        for (k = 0; (h = c->top + (k + 1) * S) < H; ++k) {
            const uint32_t xr = (uint32_t)c->history[2 * h], xi = (uint32_t)c->history[2 * h + 1];
            const uint32_t cr = (uint32_t)c->coeffs[2 * k], ci = (uint32_t)c->coeffs[2 * k + 1];
            re += xr * cr - xi * ci;
            im += xr * ci + xi * cr;
            en += xr * xr + xi * xi;
        }
        const int32_t sr = (int32_t)re >> c->coeff_scaling, si = (int32_t)im >> c->coeff_scaling; /* scale32 */
        c->energy_value[0] = en >> (c->coeff_scaling / 2);
        c->corr_value[2] = c->corr_value[1];
        c->corr_value[1] = c->corr_value[0];
        c->corr_value[0] = (uint32_t)(sr >> 2) * (uint32_t)(sr >> 2) + (uint32_t)(si >> 2) * (uint32_t)(si >> 2);
        if (c->corr_value[1] > c->corr_value[2] && c->corr_value[1] > c->corr_value[0]) {
            const double corr = sqrt((double)c->corr_value[1]), energy = sqrt((double)c->energy_value[1]);
            if (corr > energy * 2.7 && energy > 300) {
                *corr_index = (int)index - 1;
                const int new_top = c->top > 0 ? c->top - 1 : H - 1;
                for (k = 0; (h = new_top - k * S) >= 0; ++k) {
                    c->bits[2 * (N - 1 - k)] = (int16_t)c->history[2 * h];
                    c->bits[2 * (N - 1 - k) + 1] = (int16_t)c->history[2 * h + 1];
                }
                for (k = 0; (h = new_top + (k + 1) * S) < H; ++k) {
                    c->bits[2 * k] = (int16_t)c->history[2 * h];
                    c->bits[2 * k + 1] = (int16_t)c->history[2 * h + 1];
                }
                return 1; /* the break at :292: top is not advanced */
            }
        }
        c->top = (c->top + 1) % H;
    }
    return 0;
}
