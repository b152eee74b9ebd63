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

// Register-blocked scan for power-of-two strides S <= 256.  Outputs S apart share their stride-S window, so one
// thread owns R = 8 outputs t, t + S, ..., t + 7S: per block of 8 taps it loads the 15 samples of the sliding
// window once (LDS.64, already unpacked) and the 8 coefficient pairs once (broadcast LDS.128) for 256 IMADs --
// 1.3 issued instructions per multiply instead of 2.2.  The energies come from the same registers (|x|^2 of the
// first output's window) and slide for the others: E[t + S] = E[t] + e[t + S] - e[t - (N-1)S] (mod 2^32, exact).
// A CTA covers 256 * 8 consecutive outputs; thread (a, s) = (tid / S, tid % S) owns the outputs
// t0 + (a * 8 + r) * S + s.  Lanes with the same a read consecutive samples, lanes of different a are R * S apart,
// so the staged window is skewed by PAD elements per R * S with (R * S + PAD) = S (mod 16): all window loads of a
// warp are bank-conflict free, and the offsets inside a tap block are compile-time multiples of S.
// The 3-point peak test runs its double-precision square roots (as the reference does) only where an exact integer
// pre-test cannot already reject the point: energy > 300 <=> e > 90000, and corr > 2.7 * energy needs c > 7.2 * e.
constexpr int CORR_R = 8;
struct CorrSkew {
    int S, logS, PAD, shiftR;
    __host__ __device__ int operator()(int p) const { return p + PAD * ((((p >> logS) + shiftR)) / CORR_R); }
};
__host__ __device__ inline CorrSkew corr_skew(int N, int S)
{
    CorrSkew k;
    k.S = S;
    k.logS = 0;
    while ((1 << k.logS) < S) ++k.logS;
    const int Np = (N + CORR_R - 1) / CORR_R * CORR_R, e = (2 + S - 1) / S;
    k.PAD = (((S - CORR_R * S) % 16) + 16) % 16;  // (R * S + PAD) = S (mod 16): conflict-free LDS.64 half-warps
    k.shiftR = (CORR_R - (Np - 1 + e) % CORR_R) % CORR_R;
    return k;
}
// dynamic shared memory of the blocked kernel, bytes
inline size_t corr_blocked_smem(int N, int S)
{
    const int Np = (N + CORR_R - 1) / CORR_R * CORR_R, e = (2 + S - 1) / S;
    const int W = (Np - 1 + e) * S + CORR_THREADS * CORR_R;
    const int Wp = corr_skew(N, S)(W - 1) + 1;
    return ((size_t)2 * Wp + 2 * Np + 2 * (CORR_THREADS * CORR_R + 2)) * 4 + 16;
}

// correlators.h:259-266 with the cheap exact rejections first (see above); the accepted points take the reference's path
__device__ __forceinline__ bool corr_is_peak_fast(uint32_t c_prev2, uint32_t c_mid, uint32_t c_now, uint32_t e_mid)
{
    if (!(c_mid > c_prev2 && c_mid > c_now)) return false;
    if (e_mid <= 90000u) return false;                                           // sqrt(e) > 300 <=> e > 90000
    if ((unsigned long long)c_mid * 10ull <= (unsigned long long)e_mid * 72ull) return false;  // 2.7^2 = 7.29
    return corr_is_peak(c_prev2, c_mid, c_now, e_mid);
}

__device__ __forceinline__ uint32_t corr_sq(const int2 v) { return (uint32_t)(v.x * v.x + v.y * v.y); }

__global__ void __launch_bounds__(CORR_THREADS) corr_scan_blocked_kernel(const CorrParams P, int vec_in)
{
    extern __shared__ __align__(16) uint32_t corr_smem[];
    constexpr int R = CORR_R, TT = CORR_THREADS * R;
    const int N = P.N, S = P.S, H = P.H;
    const int Np = (N + R - 1) / R * R;
    const int e = (2 + S - 1) / S;
    const int halo = (Np - 1 + e) * S;  // staged in front of the tile: a multiple of S that covers the 2 extra points
    const CorrSkew skew = corr_skew(N, S);
    const int PAD = skew.PAD;
    const int W = halo + TT;
    const int Wp = skew(W - 1) + 1;
    int *cf = reinterpret_cast<int *>(corr_smem);                 // [2 * Np] newest-first, (re, im), zero padded
    int2 *xs = reinterpret_cast<int2 *>(cf + 2 * Np);             // [Wp] (re, im), skewed
    uint32_t *cvs = reinterpret_cast<uint32_t *>(xs + Wp);        // [TT + 2]
    uint32_t *evs = cvs + TT + 2;
    const int ch = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const uint32_t *x = P.in + (size_t)ch * P.in_stride;
    const uint32_t *hist = P.hist_in + (size_t)ch * H;
    // staging in groups of 4 samples aligned in the channel row (one LDG.128 where the row allows it)
    {
        const int tw = t0 - halo;                  // stream index of staged position 0
        const int extra = tw & 3;                  // (two's complement: also right for tw < 0)
        const int tstart = tw - extra;
        for (int g = threadIdx.x; 4 * g < W + extra; g += CORR_THREADS) {
            const int t = tstart + 4 * g;
            uint32_t w[4];
            if (vec_in && t >= 0 && t + 4 <= P.n) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(x + t));
                w[0] = q.x, w[1] = q.y, w[2] = q.z, w[3] = q.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = t + i < P.n ? corr_sample(x, hist, H, t + i) : 0u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = 4 * g - extra + i;
                if (p >= 0 && p < W) xs[skew(p)] = make_int2(sx_lo(w[i]), sx_hi(w[i]));
            }
        }
    }
    for (int j = threadIdx.x; j < Np; j += CORR_THREADS) {
        cf[2 * j] = j < N ? P.coef[2 * (N - 1 - j)] : 0;
        cf[2 * j + 1] = j < N ? P.coef[2 * (N - 1 - j) + 1] : 0;
    }
    __syncthreads();

    const int s = threadIdx.x & (S - 1), a = threadIdx.x >> skew.logS;
    const int q0 = a * R + Np - 1 + e;                 // stride-class index of (r = 0, tap 0)
    const int base = skew(q0 * S + s);                 // (q0 + shiftR) is a multiple of R: offsets below are static
    const int blk = R * S + PAD;
    int re[R], im[R];
#pragma unroll
    for (int r = 0; r < R; ++r) re[r] = im[r] = 0;
    uint32_t e0 = 0;  // energy of output r = 0: |x|^2 over its N taps
    auto tap_block = [&](int jb, bool check) {
        const int2 *xp = xs + (base - jb * blk);
        int2 v[2 * R - 1];
#pragma unroll
        for (int k = -(R - 1); k <= R - 1; ++k) v[k + R - 1] = xp[k * S - (k < 0 ? PAD : 0)];
        int cr[R], ci[R];
#pragma unroll
        for (int u = 0; u < R; u += 2) {
            const int4 c = *reinterpret_cast<const int4 *>(cf + 2 * (jb * R + u));
            cr[u] = c.x, ci[u] = c.y, cr[u + 1] = c.z, ci[u + 1] = c.w;
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const uint32_t sq = corr_sq(v[R - 1 - u]);  // tap jb * R + u of output r = 0
            e0 += (!check || jb * R + u < N) ? sq : 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int2 xv = v[r - u + R - 1];
                re[r] += xv.x * cr[u] - xv.y * ci[u];  // std::complex<int32_t> product, wraps like the reference
                im[r] += xv.x * ci[u] + xv.y * cr[u];
            }
        }
    };
#pragma unroll 1
    for (int jb = 0; jb < N / R; ++jb) tap_block(jb, false);
    if (N % R) tap_block(N / R, true);  // zero-padded coefficients; the energy leaves the padding taps out
    // energies of the other outputs slide along the stride class
    uint32_t en[R];
    en[0] = e0;
#pragma unroll
    for (int r = 1; r < R; ++r) en[r] = en[r - 1] + corr_sq(xs[base + r * S]) - corr_sq(xs[skew((q0 + r - N) * S + s)]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = (a * R + r) * S + s;  // output t0 + i
        uint32_t cv = 0, ev = 0;
        if (t0 + i < P.n) {
            const int sr = re[r] >> P.coeff_scaling, si = im[r] >> P.coeff_scaling;  // scale32, dsp_complex.cpp:43-46
            ev = en[r] >> (P.coeff_scaling / 2);
            cv = (uint32_t)((sr >> 2) * (sr >> 2) + (si >> 2) * (si >> 2));
        }
        cvs[i + 2] = cv;
        evs[i + 2] = ev;
    }
    // the two points in front of the tile (3-point test of its first outputs): t = -1 -> corrValue[0], t = -2 -> [1]
    if (threadIdx.x < 2) {
        const int t = t0 - 2 + threadIdx.x;
        uint32_t cv = 0, ev = 0;
        if (t >= 0 && t < P.n) {
            int sr = 0, si = 0;
            uint32_t en1 = 0;
            const int pb = halo - 2 + threadIdx.x;  // staged position of sample t
            for (int j = 0; j < N; ++j) {
                const int2 xv = xs[skew(pb - j * S)];
                sr += xv.x * cf[2 * j] - xv.y * cf[2 * j + 1];
                si += xv.x * cf[2 * j + 1] + xv.y * cf[2 * j];
                en1 += corr_sq(xv);
            }
            sr >>= P.coeff_scaling;
            si >>= P.coeff_scaling;
            ev = en1 >> (P.coeff_scaling / 2);
            cv = (uint32_t)((sr >> 2) * (sr >> 2) + (si >> 2) * (si >> 2));
        } else if (t < 0) {
            cv = P.reg_in[ch * 6 + (-1 - t)];
            ev = P.reg_in[ch * 6 + 3 + (-1 - t)];
        }
        cvs[threadIdx.x] = cv;
        evs[threadIdx.x] = ev;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int i = threadIdx.x + k * CORR_THREADS, t = t0 + i;
        if (t < P.n && corr_is_peak_fast(cvs[i], cvs[i + 1], cvs[i + 2], evs[i + 1])) atomicMin(P.found + ch, t);
    }
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
