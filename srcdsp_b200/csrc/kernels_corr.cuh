// kernels_corr.cuh -- FixedPatternCorrelator<int16_t, int32_t, N, S> (reference correlators.h:54-303), the
// stage right behind the DDC in a receiver, as data-parallel kernels.
//
// The reference walks the input sample by sample: correlation of the last N samples taken every S-th with the
// conjugated pattern, their energy, a 3-point peak test on the squared magnitudes, and it STOPS at the first
// peak above the threshold.  Here every sample's (corrValue, energyValue) pair is computed independently
// (the sliding dot product is a function of the time-ordered past only), the peak test runs on all positions
// at once and an atomicMin keeps the first hit; a small second kernel then advances the state exactly as the
// sequential loop would have left it at that sample (history, the two 3-deep registers, the bit samples).
// All arithmetic is mod 2^32 like the reference's int32_t / uint32_t; the threshold test is evaluated in double
// like the reference (sqrt is correctly rounded on both sides, so the decisions are identical).
#pragma once

#include "common.cuh"

namespace srcdsp {

constexpr int CORR_THREADS = 256;
constexpr int CORR_MAX_N = 256;

struct CorrParams {
    const uint32_t *in;       // [C][in_stride] packed cs16
    size_t in_stride;
    int n;                    // samples per channel in this call
    int N, S, H;              // pattern length, stride, H = N * S - 1 carried samples (time order, oldest first)
    const int *coef;          // [N][2] conjugated pattern (re, im)
    const uint32_t *hist_in;  // [C][H]
    uint32_t *hist_out;       // [C][H]
    const uint32_t *reg_in;   // [C][6]: corrValue[0..2], energyValue[0..2]
    uint32_t *reg_out;
    int *found;               // [C]: index of the sample at which the reference returns true, INT_MAX if none
    uint32_t *bits;           // [C][N] bit samples of the last detection
    unsigned long long *cnt;  // [C] cntProcessedSamples
    int coeff_scaling;
};

// sample t of the logical stream (t < 0: carried history)
__device__ __forceinline__ uint32_t corr_sample(const uint32_t *x, const uint32_t *hist, int H, int t)
{
    return t >= 0 ? __ldg(x + t) : (t >= -H ? __ldg(hist + H + t) : 0u);
}

// correlators.h:226-247 for the sample at time t: squared magnitude of the scaled correlation and the scaled energy
template <typename F>
__device__ __forceinline__ void corr_point(F sample_at, const int *coef, int N, int S, int t, int coeff_scaling, uint32_t &cv,
                                           uint32_t &ev)
{
    int re = 0, im = 0;
    uint32_t e = 0;
    for (int j = 0; j < N; ++j) {
        const uint32_t w = sample_at(t - j * S);
        const int xr = sx_lo(w), xi = sx_hi(w);
        const int cr = coef[2 * (N - 1 - j)], ci = coef[2 * (N - 1 - j) + 1];
        re += xr * cr - xi * ci;  // std::complex<int32_t> product, wraps like the reference
        im += xr * ci + xi * cr;
        e += (uint32_t)(xr * xr + xi * xi);
    }
    re >>= coeff_scaling;  // scale32, dsp_complex.cpp:43-46
    im >>= coeff_scaling;
    ev = e >> (coeff_scaling / 2);
    cv = (uint32_t)((re >> 2) * (re >> 2) + (im >> 2) * (im >> 2));
}

// correlators.h:259-266
__device__ __forceinline__ bool corr_is_peak(uint32_t c_prev2, uint32_t c_mid, uint32_t c_now, uint32_t e_mid)
{
    if (!(c_mid > c_prev2 && c_mid > c_now)) return false;
    const double corr = sqrt((double)c_mid), energy = sqrt((double)e_mid);
    return corr > energy * 2.7 && energy > 300;
}

// grid (tiles, C): every thread one sample.  The block stages its samples (+ halo) in shared memory already
// unpacked to (re, im) int32 pairs together with |x|^2, so that the N-tap loop is one LDS.64 + one LDS.32 +
// four IMAD + one IADD per tap (the unpacking and the squares would otherwise be redone N times per sample).
__global__ void __launch_bounds__(CORR_THREADS) corr_scan_kernel(const CorrParams P)
{
    extern __shared__ uint32_t corr_smem[];
    const int N = P.N, S = P.S, H = P.H;
    const int halo = (N - 1) * S + 2;  // the two preceding points of the 3-point test are computed here as well
    const int W = halo + CORR_THREADS;
    int2 *xs = reinterpret_cast<int2 *>(corr_smem);                 // [W] (re, im)
    uint32_t *es = reinterpret_cast<uint32_t *>(xs + W);            // [W] re^2 + im^2 (mod 2^32)
    int *cf = reinterpret_cast<int *>(es + W);                      // [2N], stored newest-first: cf[2j] = coef of x[t - jS]
    uint32_t *cvs = reinterpret_cast<uint32_t *>(cf + 2 * N);       // [CORR_THREADS + 2]
    uint32_t *evs = cvs + CORR_THREADS + 2;
    const int ch = blockIdx.y;
    const int t0 = blockIdx.x * CORR_THREADS;
    const uint32_t *x = P.in + (size_t)ch * P.in_stride;
    const uint32_t *hist = P.hist_in + (size_t)ch * H;
    for (int i = threadIdx.x; i < W; i += CORR_THREADS) {
        const int t = t0 - halo + i;
        const uint32_t w = t < P.n ? corr_sample(x, hist, H, t) : 0u;
        const int xr = sx_lo(w), xi = sx_hi(w);
        xs[i] = make_int2(xr, xi);
        es[i] = (uint32_t)(xr * xr + xi * xi);
    }
    for (int j = threadIdx.x; j < N; j += CORR_THREADS) {
        cf[2 * j] = P.coef[2 * (N - 1 - j)];
        cf[2 * j + 1] = P.coef[2 * (N - 1 - j) + 1];
    }
    __syncthreads();
    // points t0-2 .. t0+255 (the first two by threads 0, 1 as extra work); points before the block start come
    // from the carried registers: t = -1 -> corrValue[0], t = -2 -> corrValue[1]
    for (int i = threadIdx.x; i < CORR_THREADS + 2; i += CORR_THREADS) {
        const int t = t0 - 2 + i;
        uint32_t cv = 0, ev = 0;
        if (t >= 0 && t < P.n) {
            int re = 0, im = 0;
            uint32_t e = 0;
            const int base = t - (t0 - halo);  // index of sample t in the staged window
#pragma unroll 4
            for (int j = 0; j < N; ++j) {
                const int2 v = xs[base - j * S];
                const int cr = cf[2 * j], ci = cf[2 * j + 1];
                re += v.x * cr - v.y * ci;  // std::complex<int32_t> product, wraps like the reference
                im += v.x * ci + v.y * cr;
                e += es[base - j * S];
            }
            re >>= P.coeff_scaling;  // scale32, dsp_complex.cpp:43-46
            im >>= P.coeff_scaling;
            ev = e >> (P.coeff_scaling / 2);
            cv = (uint32_t)((re >> 2) * (re >> 2) + (im >> 2) * (im >> 2));
        } else if (t < 0) {
            cv = P.reg_in[ch * 6 + (-1 - t)];
            ev = P.reg_in[ch * 6 + 3 + (-1 - t)];
        }
        cvs[i] = cv;
        evs[i] = ev;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t < P.n && corr_is_peak(cvs[threadIdx.x], cvs[threadIdx.x + 1], cvs[threadIdx.x + 2], evs[threadIdx.x + 1]))
        atomicMin(P.found + ch, t);
}

// one block per channel: leave the state as the sequential loop would at the sample it stopped at
__global__ void __launch_bounds__(CORR_THREADS) corr_finish_kernel(const CorrParams P)
{
    const int ch = blockIdx.x;
    const int N = P.N, S = P.S, H = P.H;
    const uint32_t *x = P.in + (size_t)ch * P.in_stride;
    const uint32_t *hist = P.hist_in + (size_t)ch * H;
    const int f = P.found[ch];
    const bool hit = f < P.n;
    const int last = hit ? f : P.n - 1;  // last sample the reference touched
    auto at = [&](int t) { return corr_sample(x, hist, H, t); };
    // the two registers: [0] = point `last`, [1] = last - 1, [2] = last - 2 (correlators.h:221-223, 243-245)
    if (threadIdx.x < 3) {
        const int t = last - threadIdx.x;
        uint32_t cv, ev;
        if (t >= 0)
            corr_point(at, P.coef, N, S, t, P.coeff_scaling, cv, ev);
        else {
            cv = P.reg_in[ch * 6 + (-1 - t)];
            ev = P.reg_in[ch * 6 + 3 + (-1 - t)];
        }
        P.reg_out[ch * 6 + threadIdx.x] = cv;
        P.reg_out[ch * 6 + 3 + threadIdx.x] = ev;
    }
    // history: on a hit `top` is not advanced (the break at :292 skips :297), so the next call overwrites the slot
    // of sample f: the carried stream ends at f - 1.  Otherwise it ends at n - 1.
    const int end = hit ? f : P.n;  // exclusive
    for (int i = threadIdx.x; i < H; i += CORR_THREADS) P.hist_out[(size_t)ch * H + i] = at(end - H + i);
    // bit samples: every S-th sample back from the peak (the point before f), correlators.h:276-288
    // (with S == 1 every slot is in the peak's stride class, also the one sample f has just overwritten: the
    // oldest bit sample is then sample f itself, as in the reference)
    if (hit)
        for (int j = threadIdx.x; j < N; j += CORR_THREADS)
            P.bits[(size_t)ch * N + (N - 1 - j)] = at((S == 1 && j == N - 1) ? f : f - 1 - j * S);
    if (threadIdx.x == 0) P.cnt[ch] += (unsigned long long)(last + 1);  // ++cntProcessedSamples per started sample (:216)
}

}  // namespace srcdsp
