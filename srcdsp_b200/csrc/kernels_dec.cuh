// kernels_dec.cuh -- batched polyphase decimating FIR on the INT32 multiply-add pipe (K1), with
// the NCO mix optionally fused into the load stage (K2).
//
// Replaces the loop nest of FilterDnsamplingFir::step (dsptl_dnsampling_filters.h:188-219 ==
// dnsampling_filters.h:140-171) and, when MIX, Mixer::step (mixers.h:172-177) in front of it:
//
//   out[i] = limitScale16( sum_{k<Nt} c[k] * xx[i*M - k], shift ),   xx = history ++ mix(x)
//
// Data layout
//   global : channel-major [C][stride] interleaved I/Q int16 (one 32-bit word per sample)
//   shared : the input tile stored POLYPHASE-TRANSPOSED, xp[p][j] = xx[j*M - p], so that the
//            8 consecutive outputs of one thread read consecutive words of one row (sliding
//            window reuse in registers) and a warp reads consecutive 32-byte pieces; 16-byte
//            chunks are XOR-swizzled in pairs so that the 32-byte lane stride is conflict free.
//            Taps are stored polyphase too, tp[p][q] = c[q*M + p], zero padded to Qp = 8*ceil.
//   regs   : per thread 8 complex int32 accumulators, a 16-sample window, 8 taps.
//
// Work per output: 2*Nt IMAD (real tap x complex sample).  One CTA = NT threads = NT*8 outputs
// of one channel; the (Nt-1)-sample halo in front of the tile comes from the input itself or,
// for the first samples of a block, from the bank's carried history -- exactly the reference's
// internal delay line.
#pragma once

#include "common.cuh"

namespace srcdsp {

constexpr int DEC_NT = 128;  // max threads per CTA (blockDim.x is 128, 64 or 32: large M shrinks the tile)
constexpr int DEC_R = 8;     // consecutive outputs per thread
constexpr int DEC_QC = 8;    // taps per polyphase branch handled per inner iteration
constexpr int DEC_FB = 8;    // independent 16-byte loads in flight per thread while staging a tile

struct DecParams {
    const uint32_t *in;      // [C][in_stride] packed cs16
    uint32_t *out;           // [C][out_stride]
    size_t in_stride, out_stride;
    long long n_in;          // input samples per channel in this step (multiple of M)
    long long n_out;         // n_in / M
    int M;                   // decimation ratio
    int ntaps;               // Nt
    int Qp;                  // ceil(Nt / M) rounded up to a multiple of DEC_QC
    int JP;                  // row pitch of xp in words
    const int32_t *taps_poly;  // [M][Qp] polyphase, zero padded
    const uint32_t *hist_in;   // [C][H] previous block's last H = Nt-1 (mixed) samples
    int H;
    unsigned shift;          // coeffScaling - leftShift
    int tiles_per_ch;
    int vec_in, vec_out;     // 16-byte aligned fast paths usable
    // fused mixer (MIX only)
    const uint32_t *cs_table;  // [n_table] packed (cos, sin)
    const int *phi;            // [C] phase at sample 0 of this step
    const int *freq;           // [C]
    PhaseMod pm;
};

__device__ __forceinline__ int swz_col(int j)
{
    // swap neighbouring 16-byte chunks in every other group of 8 chunks
    return j ^ ((j >> 3) & 4);
}

template <int MT /* compile-time M, 0 = runtime */, bool MIX>
__global__ void __launch_bounds__(DEC_NT) dec_fir_kernel(const DecParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int M = MT ? MT : P.M;
    const int Qp = P.Qp;
    const int JP = P.JP;
    int32_t *tp = reinterpret_cast<int32_t *>(smem);  // [M][Qp]
    uint32_t *xp = smem + M * Qp;                     // [M][JP]

    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int TB = NT * DEC_R;  // outputs per tile
    const int ch = blockIdx.x / P.tiles_per_ch;
    const int tile = blockIdx.x - ch * P.tiles_per_ch;
    const long long I0 = (long long)tile * TB;  // first output of the tile
    const uint32_t *x = P.in + (size_t)ch * P.in_stride;
    const uint32_t *hist = P.hist_in + (size_t)ch * P.H;

    // ---- stage taps --------------------------------------------------------------------------
    for (int i = tid; i < M * Qp; i += NT) tp[i] = P.taps_poly[i];

    // ---- stage the input tile, polyphase transposed (+ fused NCO mix) -------------------------
    // column j_rel of row p holds sample n = (j_lo + j_rel)*M - p, j_lo = I0 - Qp.
    // The tile covers u = n - n_base in [0, Jlen*M), n_base = j_lo*M - (M-1).
    const int Jlen = TB + Qp;
    const long long n_base = (I0 - Qp) * M - (M - 1);
    // groups of 4 consecutive samples; g0 keeps the global address 16-byte aligned
    const int mis = (int)(((n_base % 4) + 4) % 4);  // n_base - mis is a multiple of 4
    const int n_groups = (Jlen * M + mis + 3) >> 2;

    unsigned phi_base = 0, fr = 0;
    if (MIX) {
        // phase of sample n_base - mis (may be "negative time": only used for n >= 0)
        fr = (unsigned)P.freq[ch];
        const unsigned nt = P.pm.n_table;
        long long nb = n_base - mis;
        long long nbm = nb % (long long)nt;
        if (nbm < 0) nbm += nt;
        phi_base = (unsigned)(((unsigned long long)P.phi[ch] + (unsigned long long)nbm * fr) % nt);
    }

    // Loads are issued in batches of DEC_FB independent 16-byte requests per thread before any of
    // them is consumed, so that one DRAM round trip is paid per batch, not per group.
    for (int gb = tid; gb < n_groups; gb += NT * DEC_FB) {
        uint4 q[DEC_FB];
        bool fast[DEC_FB];
#pragma unroll
        for (int k = 0; k < DEC_FB; ++k) {
            const int g = gb + k * NT;
            const long long n0 = n_base - mis + 4ll * g;
            fast[k] = P.vec_in && g < n_groups && n0 >= 0 && n0 + 4 <= P.n_in;
            if (fast[k]) q[k] = __ldg(reinterpret_cast<const uint4 *>(x + n0));
        }
#pragma unroll
        for (int k = 0; k < DEC_FB; ++k) {
            const int g = gb + k * NT;
            if (g >= n_groups) break;
            const long long n0 = n_base - mis + 4ll * g;
            uint32_t v[4];
            if (fast[k]) {
                v[0] = q[k].x, v[1] = q[k].y, v[2] = q[k].z, v[3] = q[k].w;
                if (MIX) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const unsigned ph = P.pm(phi_base + (unsigned)(4 * g + s) * fr);
                        v[s] = mix_sample(v[s], __ldg(P.cs_table + ph));
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const long long n = n0 + s;
                    uint32_t w = 0;
                    if (n >= 0) {
                        if (n < P.n_in) {
                            w = __ldg(x + n);
                            if (MIX) {
                                const unsigned ph = P.pm(phi_base + (unsigned)(4 * g + s) * fr);
                                w = mix_sample(w, __ldg(P.cs_table + ph));
                            }
                        }
                    } else if (n >= -(long long)P.H) {
                        w = __ldg(hist + (P.H + n));  // already mixed when it was input
                    }
                    v[s] = w;
                }
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int u = 4 * g + s - mis;
                if (u >= 0 && u < Jlen * M) {
                    const int j = u / M;
                    const int p = M - 1 - (u - j * M);
                    xp[p * JP + swz_col(j)] = v[s];
                }
            }
        }
    }
    __syncthreads();

    // ---- compute: 8 consecutive outputs per thread ----------------------------------------------
    int ar[DEC_R], ai[DEC_R];
#pragma unroll
    for (int r = 0; r < DEC_R; ++r) ar[r] = ai[r] = 0;

    // window for (r, u) of tap chunk q0 is column 8*tid - q0 - 8 + Qp + (8 + r - u)
    for (int p = 0; p < M; ++p) {
        const uint32_t *row = xp + p * JP;
        const int32_t *tpr = tp + p * Qp;
        for (int q0 = 0; q0 < Qp; q0 += DEC_QC) {
            const int cb = DEC_R * tid - q0 - 8 + Qp;  // multiple of 8
            uint32_t w[16];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const uint4 q = *reinterpret_cast<const uint4 *>(row + swz_col(cb + 4 * c4));
                w[4 * c4 + 0] = q.x, w[4 * c4 + 1] = q.y, w[4 * c4 + 2] = q.z, w[4 * c4 + 3] = q.w;
            }
            int c[DEC_QC];
            {
                const int4 c0 = *reinterpret_cast<const int4 *>(tpr + q0);
                const int4 c1 = *reinterpret_cast<const int4 *>(tpr + q0 + 4);
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
            for (int u = 0; u < DEC_QC; ++u) {
#pragma unroll
                for (int r = 0; r < DEC_R; ++r) {
                    ar[r] += c[u] * re[8 + r - u];
                    ai[r] += c[u] * im[8 + r - u];
                }
            }
        }
    }

    // ---- epilogue: limitScale16, pack, store -----------------------------------------------------
    const long long i0 = I0 + (long long)DEC_R * tid;
    uint32_t *o = P.out + (size_t)ch * P.out_stride + i0;
    uint32_t y[DEC_R];
#pragma unroll
    for (int r = 0; r < DEC_R; ++r) y[r] = scale_pack<true>(ar[r], ai[r], P.shift);
    if (P.vec_out && i0 + DEC_R <= P.n_out) {
        reinterpret_cast<uint4 *>(o)[0] = make_uint4(y[0], y[1], y[2], y[3]);
        reinterpret_cast<uint4 *>(o)[1] = make_uint4(y[4], y[5], y[6], y[7]);
    } else {
#pragma unroll
        for (int r = 0; r < DEC_R; ++r)
            if (i0 + r < P.n_out) o[r] = y[r];
    }
}

// New history = last H samples of (old history ++ mix(x)); dsptl_dnsampling_filters.h:218-219.
// Runs on the same stream as the FIR kernel; reads hist_in, writes hist_out (ping-pong), and,
// for the fused mixer, advances the per-channel phase: phi' = (phi + n_in*freq) mod N
// (mixers.h:177 applied n_in times).
template <bool MIX>
__global__ void dec_history_kernel(const uint32_t *__restrict__ in, size_t in_stride, long long n_in,
                                   const uint32_t *__restrict__ hist_in, uint32_t *__restrict__ hist_out,
                                   int H, const uint32_t *__restrict__ cs_table, const int *__restrict__ phi,
                                   int *__restrict__ phi_out, const int *__restrict__ freq, PhaseMod pm)
{
    const int ch = blockIdx.y;
    const uint32_t *x = in + (size_t)ch * in_stride;
    unsigned ph0 = 0, fr = 0;
    if (MIX) {
        ph0 = (unsigned)phi[ch];
        fr = (unsigned)freq[ch];
    }
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
        const long long n = n_in - H + h;
        uint32_t w;
        if (n >= 0) {
            w = __ldg(x + n);
            if (MIX) {
                const unsigned nm = (unsigned)(n % (long long)pm.n_table);
                const unsigned ph = (unsigned)((ph0 + (unsigned long long)nm * fr) % pm.n_table);
                w = mix_sample(w, __ldg(cs_table + ph));
            }
        } else {
            w = hist_in[(size_t)ch * H + (h + n_in)];
        }
        hist_out[(size_t)ch * H + h] = w;
    }
    if (MIX && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned nm = (unsigned)(n_in % (long long)pm.n_table);
        phi_out[ch] = (int)((ph0 + (unsigned long long)nm * fr) % pm.n_table);
    }
}

}  // namespace srcdsp
