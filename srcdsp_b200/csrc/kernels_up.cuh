// kernels_up.cuh -- batched polyphase interpolating FIR (K4) and the stand-alone NCO mixer.
//
// Upsampler: replaces the loop nest of FilterUpsamplingFir::step (upsampling_filters.h:163-194,
// flush :196-231, iterator overload :252-284):
//
//   out[j*L + p] = limitScale<cs16>( sum_{i<H} c[p + i*L] * xx[j - i], shift ),  H = Nt / L,
//   xx = history ++ x ++ (n_flush zeros)
//
// The reference's circular buffer + `top` is kept in AGE ORDER on the device (a[0..H-1],
// a[H-1] newest, a[0] the stale slot the next insert overwrites); only age order is observable.
//
// One thread owns one output phase p and 8 consecutive input positions j (sliding-window reuse
// of the input in registers, the H taps of its phase in registers); lanes of a warp run over
// p fastest so that each store instruction writes runs of L consecutive output samples.
#pragma once

#include "common.cuh"

namespace srcdsp {

constexpr int UP_NT = 128;
constexpr int UP_R = 8;   // consecutive inputs per thread per iteration
constexpr int UP_NI = 8;  // iterations per CTA (CTA span = (UP_NT / L) * UP_R * UP_NI inputs)
constexpr int UP_HC = 8;  // taps per phase per inner iteration

struct UpParams {
    const uint32_t *in;
    uint32_t *out;
    size_t in_stride, out_stride;
    long long n_in;   // real input samples per channel
    long long n_tot;  // n_in + n_flush
    int L;
    int H;            // taps per phase = Nt / L
    int Hp;           // H rounded up to a multiple of UP_HC
    int G;            // j-groups per CTA = UP_NT / L
    const int32_t *taps_poly;  // [L][Hp + 4]: tp[p][i] = c[p + i*L], zero padded
    const uint32_t *hist_in;   // [C][H] age order
    unsigned shift;
    int tiles_per_ch;
    int vec_in;
};

__global__ void __launch_bounds__(UP_NT) up_fir_kernel(const UpParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int L = P.L, Hp = P.Hp, HP = Hp + 4;
    const int span = P.G * UP_R * UP_NI;     // inputs per CTA
    int32_t *tp = reinterpret_cast<int32_t *>(smem);  // [L][HP]
    uint32_t *xs = smem + L * HP;                      // [span + Hp + 8]

    const int tid = threadIdx.x;
    const int ch = blockIdx.x / P.tiles_per_ch;
    const int tile = blockIdx.x - ch * P.tiles_per_ch;
    const long long J0 = (long long)tile * span;
    const uint32_t *x = P.in + (size_t)ch * P.in_stride;
    const uint32_t *hist = P.hist_in + (size_t)ch * P.H;

    for (int i = tid; i < L * HP; i += UP_NT) tp[i] = P.taps_poly[i];

    // xs[col] = xx[J0 - Hp + col], col in [0, span + Hp)
    const int n_cols = span + Hp;
    const long long n_lo = J0 - Hp;  // multiple of 8: 16-byte aligned when vec_in
    for (int g = tid; g < (n_cols >> 2); g += UP_NT) {
        const long long n0 = n_lo + 4ll * g;
        uint4 q;
        if (P.vec_in && n0 >= 0 && n0 + 4 <= P.n_in) {
            q = __ldg(reinterpret_cast<const uint4 *>(x + n0));
        } else {
            uint32_t v[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const long long n = n0 + s;
                uint32_t w = 0;
                if (n >= 0) {
                    if (n < P.n_in) w = __ldg(x + n);
                } else if (n >= -(long long)P.H) {
                    w = __ldg(hist + (P.H + n));
                }
                v[s] = w;
            }
            q = make_uint4(v[0], v[1], v[2], v[3]);
        }
        reinterpret_cast<uint4 *>(xs)[g] = q;
    }
    __syncthreads();

    const int p = tid % L;
    const int jg = tid / L;
    if (jg >= P.G) return;
    const int32_t *tpr = tp + p * HP;
    uint32_t *o = P.out + (size_t)ch * P.out_stride;

    for (int it = 0; it < UP_NI; ++it) {
        const int jr = (it * P.G + jg) * UP_R;  // first input (relative to J0) of this thread
        if (J0 + jr >= P.n_tot) break;
        int ar[UP_R], ai[UP_R];
#pragma unroll
        for (int r = 0; r < UP_R; ++r) ar[r] = ai[r] = 0;
        for (int i0 = 0; i0 < Hp; i0 += UP_HC) {
            // sample for (r, u): xx[J0 + jr + r - i0 - u] = xs[jr - i0 - 8 + Hp + (8 + r - u)]
            const int cb = jr - i0 - 8 + Hp;
            uint32_t w[16];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const uint4 q = *reinterpret_cast<const uint4 *>(xs + cb + 4 * c4);
                w[4 * c4 + 0] = q.x, w[4 * c4 + 1] = q.y, w[4 * c4 + 2] = q.z, w[4 * c4 + 3] = q.w;
            }
            int c[UP_HC];
            {
                const int4 c0 = *reinterpret_cast<const int4 *>(tpr + i0);
                const int4 c1 = *reinterpret_cast<const int4 *>(tpr + i0 + 4);
                c[0] = c0.x, c[1] = c0.y, c[2] = c0.z, c[3] = c0.w;
                c[4] = c1.x, c[5] = c1.y, c[6] = c1.z, c[7] = c1.w;
            }
            int re[16], im[16];
#pragma unroll
            for (int e = 1; e < 16; ++e) {
                re[e] = sx_lo(w[e]);
                im[e] = sx_hi(w[e]);
            }
#pragma unroll
            for (int u = 0; u < UP_HC; ++u) {
#pragma unroll
                for (int r = 0; r < UP_R; ++r) {
                    ar[r] += c[u] * re[8 + r - u];
                    ai[r] += c[u] * im[8 + r - u];
                }
            }
        }
        // outputs of (j, p) for the thread's UP_R consecutive inputs: one pointer, stride L
        uint32_t *op = o + (J0 + jr) * L + p;
        if (J0 + jr + UP_R <= P.n_tot) {
#pragma unroll
            for (int r = 0; r < UP_R; ++r) op[r * L] = scale_pack_asym_sat(ar[r], ai[r], P.shift);
        } else {
#pragma unroll
            for (int r = 0; r < UP_R; ++r)
                if (J0 + jr + r < P.n_tot) op[r * L] = scale_pack_asym_sat(ar[r], ai[r], P.shift);
        }
    }
}

// Register-blocked variant for L in {4, 8, 16} and 16-byte aligned output rows: one thread owns 4 ADJACENT phases
// x 4 consecutive inputs.  The sliding window (11 samples per 8 taps) is unpacked once for 16 outputs instead of
// 15 samples for 8, the 4 x 8 taps stay in registers for the whole CTA when the filter has at most 8 taps per
// phase (SINGLE), the 4 phases of an input leave as one STG.128 with an immediate offset, and all index math is
// 32-bit: ~21 issued instructions per output (16 of them the IMADs) instead of ~36, so the multiply pipe -- the
// roofline of this path, 16 IMAD per output at 64 lanes/clk/SM -- is what binds, not the issue slots.
constexpr int UP4_R = 4;    // consecutive inputs per thread per iteration
constexpr int UP4_PH = 4;   // adjacent phases per thread
constexpr int UP4_NI = 8;   // iterations per CTA
constexpr int up4_span(int L) { return (UP_NT / (L / UP4_PH)) * UP4_R * UP4_NI; }

template <int L, bool SINGLE>
__global__ void __launch_bounds__(UP_NT) up_fir4_kernel(const UpParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int PG = L / UP4_PH;        // phase groups per input
    constexpr int G = UP_NT / PG;         // j-groups per CTA
    constexpr int SPAN_IT = G * UP4_R;    // inputs per CTA iteration
    constexpr int span = SPAN_IT * UP4_NI;
    const int Hp = SINGLE ? UP_HC : P.Hp, HP = Hp + 4;
    int32_t *tp = reinterpret_cast<int32_t *>(smem);  // [L][HP]
    uint32_t *xs = smem + L * HP;                      // [span + Hp + 8]

    const int tid = threadIdx.x;
    const int ch = blockIdx.x / P.tiles_per_ch;
    const int tile = blockIdx.x - ch * P.tiles_per_ch;
    const long long J0 = (long long)tile * span;
    const uint32_t *x = P.in + (size_t)ch * P.in_stride;
    const uint32_t *hist = P.hist_in + (size_t)ch * P.H;

    for (int i = tid; i < L * HP; i += UP_NT) tp[i] = P.taps_poly[i];

    // xs[col] = xx[J0 - Hp + col], col in [0, span + Hp)
    const int n_cols = span + Hp;
    const long long n_lo = J0 - Hp;  // multiple of 8: 16-byte aligned when vec_in
    const bool interior = P.vec_in && n_lo >= 0 && n_lo + n_cols <= P.n_in;
    if (interior) {
        const uint4 *src = reinterpret_cast<const uint4 *>(x + n_lo);
        for (int g = tid; g < (n_cols >> 2); g += UP_NT) reinterpret_cast<uint4 *>(xs)[g] = __ldg(src + g);
    } else {
        for (int g = tid; g < (n_cols >> 2); g += UP_NT) {
            const long long n0 = n_lo + 4ll * g;
            uint4 q;
            if (P.vec_in && n0 >= 0 && n0 + 4 <= P.n_in) {
                q = __ldg(reinterpret_cast<const uint4 *>(x + n0));
            } else {
                uint32_t v[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const long long n = n0 + s;
                    uint32_t w = 0;
                    if (n >= 0) {
                        if (n < P.n_in) w = __ldg(x + n);
                    } else if (n >= -(long long)P.H) {
                        w = __ldg(hist + (P.H + n));
                    }
                    v[s] = w;
                }
                q = make_uint4(v[0], v[1], v[2], v[3]);
            }
            reinterpret_cast<uint4 *>(xs)[g] = q;
        }
    }
    __syncthreads();

    const int phg = tid % PG, jg = tid / PG;
    const int32_t *tpr = tp + (UP4_PH * phg) * HP;
    const long long left = P.n_tot - J0;
    const int rem = left < (long long)span ? (int)left : span;  // inputs of this CTA that exist
    uint32_t *o = P.out + (size_t)ch * P.out_stride + (size_t)J0 * L + UP4_PH * phg;
    const unsigned shift = P.shift;

    int c[UP4_PH][UP_HC];
    auto load_taps = [&](int i0) {
#pragma unroll
        for (int ph = 0; ph < UP4_PH; ++ph) {
            const int4 c0 = *reinterpret_cast<const int4 *>(tpr + ph * HP + i0);
            const int4 c1 = *reinterpret_cast<const int4 *>(tpr + ph * HP + i0 + 4);
            c[ph][0] = c0.x, c[ph][1] = c0.y, c[ph][2] = c0.z, c[ph][3] = c0.w;
            c[ph][4] = c1.x, c[ph][5] = c1.y, c[ph][6] = c1.z, c[ph][7] = c1.w;
        }
    };
    if (SINGLE) load_taps(0);

#pragma unroll 1
    for (int it = 0; it < UP4_NI; ++it) {
        const int jr = it * SPAN_IT + jg * UP4_R;  // first input (relative to J0) of this thread
        if (jr >= rem) break;
        int ar[UP4_PH][UP4_R], ai[UP4_PH][UP4_R];
#pragma unroll
        for (int ph = 0; ph < UP4_PH; ++ph)
#pragma unroll
            for (int r = 0; r < UP4_R; ++r) ar[ph][r] = ai[ph][r] = 0;
#pragma unroll 1
        for (int i0 = 0; i0 < Hp; i0 += UP_HC) {
            // sample for (r, u): xx[J0 + jr + r - i0 - u] = xs[jr - i0 - 8 + Hp + (8 + r - u)], 8 + r - u in [1, 11]
            const uint4 *wp = reinterpret_cast<const uint4 *>(xs + (jr - i0 - 8 + Hp));
            const uint4 q0 = wp[0], q1 = wp[1], q2 = wp[2];
            if (!SINGLE) load_taps(i0);
            const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            int re[12], im[12];
#pragma unroll
            for (int e = 1; e < 12; ++e) {
                re[e] = sx_lo(w[e]);
                im[e] = sx_hi(w[e]);
            }
#pragma unroll
            for (int u = 0; u < UP_HC; ++u) {
#pragma unroll
                for (int r = 0; r < UP4_R; ++r) {
#pragma unroll
                    for (int ph = 0; ph < UP4_PH; ++ph) {
                        ar[ph][r] += c[ph][u] * re[8 + r - u];
                        ai[ph][r] += c[ph][u] * im[8 + r - u];
                    }
                }
            }
            if (SINGLE) break;
        }
        // the 4 phases of input jr + r are 4 consecutive output samples: one 16-byte store each
        uint4 *op = reinterpret_cast<uint4 *>(o + (size_t)jr * L);
#pragma unroll
        for (int r = 0; r < UP4_R; ++r) {
            if (jr + r < rem)
                op[r * (L / 4)] = make_uint4(scale_pack_asym_sat(ar[0][r], ai[0][r], shift), scale_pack_asym_sat(ar[1][r], ai[1][r], shift),
                                             scale_pack_asym_sat(ar[2][r], ai[2][r], shift), scale_pack_asym_sat(ar[3][r], ai[3][r], shift));
        }
    }
}

// New age-ordered history a'[k] = xx[n_tot - H + k], k < H  (xx includes the flush zeros).
__global__ void up_history_kernel(const uint32_t *__restrict__ in, size_t in_stride, long long n_in,
                                  long long n_tot, const uint32_t *__restrict__ hist_in,
                                  uint32_t *__restrict__ hist_out, int H)
{
    const int ch = blockIdx.y;
    const uint32_t *x = in + (size_t)ch * in_stride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < H; k += gridDim.x * blockDim.x) {
        const long long n = n_tot - H + k;
        uint32_t w = 0;
        if (n >= 0) {
            if (n < n_in) w = __ldg(x + n);
        } else {
            w = hist_in[(size_t)ch * H + (H + n)];
        }
        hist_out[(size_t)ch * H + k] = w;
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone mixer: Mixer::step (mixers.h:168-188).  HBM bound: 4 B in + 4 B out per sample.
// out may alias in (each thread reads its samples before writing them).
// ---------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) mixer_kernel(const uint32_t *in, size_t in_stride, uint32_t *out,
                                                    size_t out_stride, long long n,
                                                    const uint32_t *__restrict__ cs_table,
                                                    const int *__restrict__ phi, int *__restrict__ phi_out,
                                                    const int *__restrict__ freq, PhaseMod pm)
{
    const int ch = blockIdx.y;
    const uint32_t *x = in + (size_t)ch * in_stride;
    uint32_t *y = out + (size_t)ch * out_stride;
    const unsigned ph0 = (unsigned)phi[ch], fr = (unsigned)freq[ch];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const long long n4 = n >> 2;
        for (long long g = t0; g < n4; g += stride) {
            const uint4 q = reinterpret_cast<const uint4 *>(x)[g];
            uint32_t v[4] = {q.x, q.y, q.z, q.w};
            const unsigned nm = pm.mask ? ((unsigned)(4 * g) & pm.mask)
                                        : (unsigned)((4 * g) % (long long)pm.n_table);
            const unsigned pb = ph0 + nm * fr;  // < 2^32: nm, fr < n_table <= 32768
#pragma unroll
            for (int s = 0; s < 4; ++s) v[s] = mix_sample(v[s], __ldg(cs_table + pm(pb + s * fr)));
            reinterpret_cast<uint4 *>(y)[g] = make_uint4(v[0], v[1], v[2], v[3]);
        }
        for (long long k = (n4 << 2) + t0; k < n; k += stride) {
            const unsigned nm = pm.mask ? ((unsigned)k & pm.mask) : (unsigned)(k % (long long)pm.n_table);
            y[k] = mix_sample(x[k], __ldg(cs_table + pm(ph0 + nm * fr)));
        }
    } else {
        for (long long k = t0; k < n; k += stride) {
            const unsigned nm = pm.mask ? ((unsigned)k & pm.mask) : (unsigned)(k % (long long)pm.n_table);
            y[k] = mix_sample(x[k], __ldg(cs_table + pm(ph0 + nm * fr)));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned nm = (unsigned)(n % (long long)pm.n_table);
        phi_out[ch] = (int)pm(ph0 + nm * fr);
    }
}

// Same with the oscillator as a per-channel sequence in TIME order (power-of-two tables, 16-byte aligned
// rows): the CTA builds seq[n] = digits of T[(phi0 + n * freq) mod N], n < N, in shared memory once and then
// streams its share of the channel -- LDG.128, two conflict-free LDS.128 for the 4 oscillator values, dp2a mix,
// STG.128.  The table gather of mixer_kernel has a lane stride equal to the channel frequency: with evenly
// spaced channel frequencies most lanes collide on a few lines / banks (2.1 TB/s); this form is HBM bound.
constexpr int MIXSEQ_THREADS = 256;
__global__ void __launch_bounds__(MIXSEQ_THREADS) mixer_seq_kernel(const uint32_t *in, size_t in_stride, uint32_t *out,
                                                                   size_t out_stride, long long n,
                                                                   const uint32_t *__restrict__ cs_table,
                                                                   const int *__restrict__ phi, int *__restrict__ phi_out,
                                                                   const int *__restrict__ freq, unsigned mask)
{
    extern __shared__ __align__(16) uint32_t seq[];  // Bre[N] ++ Bim[N]
    const int ch = blockIdx.y;
    const unsigned ph0 = (unsigned)phi[ch], fr = (unsigned)freq[ch];
    const unsigned N = mask + 1;
    for (unsigned i = threadIdx.x; i < N; i += MIXSEQ_THREADS) {
        uint32_t bre, bim;
        mix_digits(__ldg(cs_table + ((ph0 + i * fr) & mask)), bre, bim);
        seq[i] = bre;
        seq[N + i] = bim;
    }
    __syncthreads();
    const uint4 *x = reinterpret_cast<const uint4 *>(in + (size_t)ch * in_stride);
    uint4 *y = reinterpret_cast<uint4 *>(out + (size_t)ch * out_stride);
    const long long n4 = n >> 2;
    const long long per = (n4 + gridDim.x - 1) / gridDim.x;  // groups of 4 samples per CTA: a contiguous run
    const long long g0 = (long long)blockIdx.x * per, g1 = g0 + per < n4 ? g0 + per : n4;
    auto mix4 = [&](uint4 q, long long g) {
        const unsigned idx = (unsigned)(4 * g) & mask;
        const uint4 a = *reinterpret_cast<const uint4 *>(seq + idx);
        const uint4 b = *reinterpret_cast<const uint4 *>(seq + N + idx);
        q.x = mix_sample_dp2a(q.x, a.x, b.x);
        q.y = mix_sample_dp2a(q.y, a.y, b.y);
        q.z = mix_sample_dp2a(q.z, a.z, b.z);
        q.w = mix_sample_dp2a(q.w, a.w, b.w);
        __stcs(y + g, q);
    };
    constexpr int U = 4;  // 16-byte loads in flight per thread
    long long g = g0 + threadIdx.x;
    for (; g + (U - 1) * MIXSEQ_THREADS < g1; g += U * MIXSEQ_THREADS) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = __ldcs(x + g + u * MIXSEQ_THREADS);
#pragma unroll
        for (int u = 0; u < U; ++u) mix4(q[u], g + u * MIXSEQ_THREADS);
    }
    for (; g < g1; g += MIXSEQ_THREADS) mix4(__ldcs(x + g), g);
    if (blockIdx.x == 0) {
        const uint32_t *xs = in + (size_t)ch * in_stride;
        uint32_t *ys = out + (size_t)ch * out_stride;
        for (long long k = (n4 << 2) + threadIdx.x; k < n; k += MIXSEQ_THREADS) {
            const unsigned i = (unsigned)k & mask;
            ys[k] = mix_sample_dp2a(xs[k], seq[i], seq[N + i]);
        }
        if (threadIdx.x == 0) phi_out[ch] = (int)((ph0 + ((unsigned)n & mask) * fr) & mask);
    }
}

// synthetic complex baseband (host twin: oracle/srcdsp_oracle.c:orc_synth_fill)
__global__ void synth_kernel(uint32_t *out, size_t stride, long long n, uint32_t seed, uint32_t ch0,
                             unsigned long long n0, int amp_shift)
{
    const int ch = blockIdx.y;
    uint32_t *y = out + (size_t)ch * stride;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
        const uint32_t h = hash32(seed, ch0 + ch, n0 + (unsigned long long)k);
        const int re = ((int)(short)(h & 0xFFFFu)) >> amp_shift;
        const int im = ((int)(short)(h >> 16)) >> amp_shift;
        y[k] = pack_iq(re, im);
    }
}

}  // namespace srcdsp
