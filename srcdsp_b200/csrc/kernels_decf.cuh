// kernels_decf.cuh -- the float instantiation of the decimating FIR,
//   dsptl::FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M>::step
//   (dsptl_dnsampling_filters.h:172-220), bit-exact on the FP32 pipe.
//
// What the reference computes for that instantiation (probed on the compiled reference, oracle/ref_harness.cpp):
//
//   y = complex<float>(0, 0)
//   for k = 0 .. N-1 (ascending):  y.re += c[k] * xx[i*M - k].re ;  y.im += c[k] * xx[i*M - k].im      (:195-210)
//   out[i] = complex<float>( limitScale16( complex<int32_t>(y), coeffScaling - leftShift ) )             (:214)
//
// i.e. one rounded multiply and one rounded add per component and tap (g++ for x86-64 does not contract them),
// summed in tap order; the sum is truncated toward zero to int32, shifted, clamped to +-32767 and converted back.
// The result depends on the order of the additions, so this kernel keeps it: every output is ONE sequential chain
// of __fmul_rn / __fadd_rn per component (never an FMA, never a tree), and parallelism comes from the outputs.
//
// Shape.  4 FP32 instructions per tap and output make the FP32 pipe (128 lanes/clk/SM) the roofline; 8 bytes of
// sample per 4 lane-instructions are twice what shared memory delivers (128 B/clk/SM), so samples are reused in
// registers: a thread owns the output PAIR (2q, 2q+1) and walks the samples downwards from xx[(2q+1)*M]; sample
// (2q+1)*M - e is tap k = e of the upper output and tap k = e - M of the lower one.  Per 4 samples: 2 LDS.128
// (samples) + 2 LDS.128 (broadcast: 4 taps of each output, from two pre-shifted zero-padded copies of the taps) feed
// 32 FP32 instructions; a thread carries PAIRS such pairs (1 with 256 threads per CTA by default: more warps hide the
// LDS latency and overlap another CTA's staging better than more chains per thread do).  The lane stride is 2*M samples
// = M 16-byte units; for even M every block of 2*M samples is followed by one unit of padding so that the stride is
// odd and the LDS.128 are conflict free.  Zero taps in front of / behind a filter add +-0 to its chain, which leaves
// every finite sum unchanged (samples must be finite -- the reference's int conversion of NaN/Inf is undefined).
//
// Measured steps (256 channels, /16 x 255 taps, 1 Mi samples per channel; tools/decfbench.py):
//   plain staging loads 3.03 ms -> 8-byte cp.async (all copies of a thread in flight) 1.14 -> bulk L2 prefetch of the
//   successor tile 1.00 -> whole blocks with compile-time offsets (BC > 0) 0.92 -> taps through the parameter space
//   (CT: uniform constant loads, no tap LDS) 0.79 ms = 21.3 G out/s = 0.58 of the FP32 pipe.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace srcdsp {


struct DecfParams {
    const float2 *in;
    float2 *out;
    size_t in_stride, out_stride;  // complex samples per channel row
    long long n_in, n_out;         // per channel
    const float2 *hist;            // [C][N - 1] age order: hist[N - 2] = xx[-1]
    const float *taps2;            // [2][E]: [1][e] = c[e], [0][e] = c[e - M], zero outside the filter
    int M, N, E;                   // ratio, taps, padded walk length (multiple of 4, >= N + M)
    int lead;                      // local sample index of stream sample tile_out * M * tile
    int blk, padw;                 // samples per block (2 * M) and padding samples behind each (2 if M is even)
    unsigned blk_magic;            // floor(2^32 / blk) + 1: idx / blk = umulhi(idx, magic) for the staged range
    int blk_chunks;                // 4-sample chunks per block (M / 2), or INT_MAX without padding
    int n_local;                   // staged samples per tile
    int tile_out;                  // outputs per tile = 2 * PAIRS * blockDim.x
    int tiles_per_ch;
    int prefetch_dist;             // CTAs resident on the device at once (0: no L2 prefetch)
    unsigned stage_step;           // bytes the padded staging position advances per round of blockDim.x samples (0: blockDim.x % blk != 0)
    unsigned shift;                // (coeffScaling - leftShift) & 31, as x86 `sar` applies it
};

// taps of short filters travel in the kernel's parameter space (constant bank): 2 * E floats <= 3584 bytes
constexpr int DECF_CT_MAX4 = 224;
struct DecfTaps {
    float4 t[DECF_CT_MAX4];  // [0, E/4): lower output (c[e - M]), [E/4, E/2): upper output (c[e])
};

__device__ __forceinline__ float decf_limit(float y, unsigned shift)
{
    // complex<int32_t>(y): truncation as the reference's x86-64 build does it (cvttss2si) -- 0x80000000 for NaN and for every
    // |y| >= 2^31, positive overflow included (cvt.rzi would saturate to INT_MAX there); then limitScale16 (dsp_complex.cpp:63-73)
    int v = (fabsf(y) < 2147483648.0f ? __float2int_rz(y) : (int)0x80000000) >> shift;
    v = max(-32767, min(32767, v));
    return (float)v;
}

// CT: the taps are read from the parameter space (uniform constant loads: no LDS, no shared-memory wavefronts for them)
template <int PAIRS, int BC, bool CT>
__global__ void __launch_bounds__(256) decf_fir_kernel(const __grid_constant__ DecfParams P, const __grid_constant__ DecfTaps K)
{
    extern __shared__ __align__(16) uint8_t decf_smem[];
    float *ts = reinterpret_cast<float *>(decf_smem);              // [2][E]
    float2 *xs = reinterpret_cast<float2 *>(ts + 2 * P.E);         // padded samples (2 * E * 4 bytes is a multiple of 16)
    const int tid = threadIdx.x, T = blockDim.x;
    const unsigned ch = blockIdx.x / (unsigned)P.tiles_per_ch;
    const int tile = (int)(blockIdx.x - ch * (unsigned)P.tiles_per_ch);

    // L2 prefetch of the tile that the CTA taking this one's place will stage (prefetch_dist = resident CTAs of the
    // grid): one bulk prefetch by one thread, so that the staging below waits for L2, not for HBM
    if (tid == 0 && P.prefetch_dist > 0 && blockIdx.x + (unsigned)P.prefetch_dist < gridDim.x) {
        const unsigned nb = blockIdx.x + (unsigned)P.prefetch_dist, nch = nb / (unsigned)P.tiles_per_ch;
        const long long a = max((long long)(nb - nch * (unsigned)P.tiles_per_ch) * P.tile_out * P.M - P.lead, 0ll);
        const long long b = min(a + P.n_local, P.n_in);
        if (b > a) {
            const uintptr_t p0 = reinterpret_cast<uintptr_t>(P.in + (size_t)nch * P.in_stride + a);
            const uintptr_t lo = (p0 + 15) & ~(uintptr_t)15, hi = (p0 + (uintptr_t)(b - a) * 8) & ~(uintptr_t)15;
            if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
        }
    }
    if (!CT)
        for (int i = tid; i < 2 * P.E; i += T) ts[i] = __ldg(P.taps2 + i);
    // stage: local index idx <-> stream sample s = tile * tile_out * M + idx - lead
    const float2 *x = P.in + (size_t)ch * P.in_stride;
    const float2 *hist = P.hist + (size_t)ch * (P.N - 1);
    const long long s0 = (long long)tile * P.tile_out * P.M - P.lead;
    if (s0 >= 0 && s0 + P.n_local <= P.n_in) {
        // interior tile: every sample comes from the input row -- 8-byte cp.async (rows are only 8-byte aligned against
        // the block grid), all of a thread's copies in flight at once, no registers
        const uint32_t xs_u32 = (uint32_t)__cvta_generic_to_shared(xs);
        const float2 *src = x + s0;
        if (P.stage_step) {
            // the CTA is a whole number of blocks wide: the padded position advances by a constant per round
            uint32_t dst = xs_u32 + 8u * (uint32_t)(tid + P.padw * (int)__umulhi((unsigned)tid, P.blk_magic));
            const float2 *sp = src + tid;
#pragma unroll 4
            for (int idx = tid; idx < P.n_local; idx += T, dst += P.stage_step, sp += T)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(sp) : "memory");
        } else {
            for (int idx = tid; idx < P.n_local; idx += T) {
                const uint32_t dst = xs_u32 + 8u * (uint32_t)(idx + P.padw * (int)__umulhi((unsigned)idx, P.blk_magic));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + idx) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    } else {
        // edge tile: carried history in front of the block, zeros behind it
        for (int idx = tid; idx < P.n_local; idx += T) {
            const long long s = s0 + idx;
            float2 v = make_float2(0.f, 0.f);
            if (s >= 0) {
                if (s < P.n_in) v = __ldg(x + s);
            } else if (s >= -(long long)(P.N - 1)) {
                v = __ldg(hist + (P.N - 1 + s));
            }
            xs[idx + P.padw * (int)__umulhi((unsigned)idx, P.blk_magic)] = v;
        }
    }
    __syncthreads();

    // pair q = tid + j * T: upper output 2q + 1 sits at local index (2q + 1) * M + lead = the LAST sample of block
    // q + j0 (even M) -- chunks of 4 never straddle a block
    float2 acc[PAIRS][2];
    int pos[PAIRS];
#pragma unroll
    for (int j = 0; j < PAIRS; ++j) {
        acc[j][0] = acc[j][1] = make_float2(0.f, 0.f);
        const int top = (2 * (tid + j * T) + 1) * P.M + P.lead;
        pos[j] = top + P.padw * (int)__umulhi((unsigned)top, P.blk_magic);
    }
    int in_blk = 0;
    const float4 *t0 = reinterpret_cast<const float4 *>(ts), *t1 = reinterpret_cast<const float4 *>(ts + P.E);
    const int e4 = P.E / 4;
    auto tap0 = [&](int c) { return CT ? K.t[c] : t0[c]; };
    auto tap1 = [&](int c) { return CT ? K.t[e4 + c] : t1[c]; };
    // one chunk = 4 samples (e .. e + 3) for the upper (UP) and / or lower (LO) output of every pair
    auto chunk = [&](int c, auto up_tag, auto lo_tag) {
        constexpr bool UP = decltype(up_tag)::value, LO = decltype(lo_tag)::value;
        float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0;  // taps e .. e + 3 of the lower / upper output (broadcast)
        if (LO) k0 = tap0(c);
        if (UP) k1 = tap1(c);
#pragma unroll
        for (int j = 0; j < PAIRS; ++j) {
            // samples pos-3 .. pos (ascending address); e ascends as the address descends
            const float4 hi = *reinterpret_cast<const float4 *>(xs + pos[j] - 1);  // {x[pos-1], x[pos]}
            const float4 lo = *reinterpret_cast<const float4 *>(xs + pos[j] - 3);  // {x[pos-3], x[pos-2]}
            float2 &a0 = acc[j][0], &a1 = acc[j][1];
            if (UP) a1.x = __fadd_rn(a1.x, __fmul_rn(k1.x, hi.z)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.x, hi.w));
            if (LO) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.x, hi.z)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.x, hi.w));
            if (UP) a1.x = __fadd_rn(a1.x, __fmul_rn(k1.y, hi.x)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.y, hi.y));
            if (LO) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.y, hi.x)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.y, hi.y));
            if (UP) a1.x = __fadd_rn(a1.x, __fmul_rn(k1.z, lo.z)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.z, lo.w));
            if (LO) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.z, lo.z)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.z, lo.w));
            if (UP) a1.x = __fadd_rn(a1.x, __fmul_rn(k1.w, lo.x)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.w, lo.y));
            if (LO) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.w, lo.x)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.w, lo.y));
            pos[j] -= 4;
        }
        if (++in_blk == P.blk_chunks) {  // uniform: step over the padding in front of the block just finished
            in_blk = 0;
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) pos[j] -= P.padw;
        }
    };
    // chunks below M / 4 hold only zero taps of the lower output (its tap index e - M is negative), chunks from
    // ceil(N / 4) on only zero taps of the upper one: skipping a zero tap leaves its chain unchanged
    const int c1 = min(P.M / 4, (P.N + 3) / 4), c2 = (P.N + 3) / 4, c3 = P.E / 4;
    int c = 0;
    if (BC > 0) {
        // M = 2 * BC (a power of two up to 32) and N >= 2 * M: whole blocks of BC chunks with compile-time offsets --
        // no per-chunk position / padding bookkeeping, every LDS address an immediate off one register per pair.
        // The first block is upper-only below chunk BC / 2 (= M / 4), all later whole blocks feed both outputs.
        auto block = [&](int cb, auto first_tag) {
            constexpr bool FIRST = decltype(first_tag)::value;
#pragma unroll
            for (int i = 0; i < BC; ++i) {
                const bool lo_on = !FIRST || i >= BC / 2;
                float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lo_on) k0 = tap0(cb + i);
                const float4 k1 = tap1(cb + i);
#pragma unroll
                for (int j = 0; j < PAIRS; ++j) {
                    const float4 hi = *reinterpret_cast<const float4 *>(xs + pos[j] - 4 * i - 1);
                    const float4 lo = *reinterpret_cast<const float4 *>(xs + pos[j] - 4 * i - 3);
                    float2 &a0 = acc[j][0], &a1 = acc[j][1];
                    a1.x = __fadd_rn(a1.x, __fmul_rn(k1.x, hi.z)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.x, hi.w));
                    if (lo_on) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.x, hi.z)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.x, hi.w));
                    a1.x = __fadd_rn(a1.x, __fmul_rn(k1.y, hi.x)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.y, hi.y));
                    if (lo_on) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.y, hi.x)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.y, hi.y));
                    a1.x = __fadd_rn(a1.x, __fmul_rn(k1.z, lo.z)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.z, lo.w));
                    if (lo_on) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.z, lo.z)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.z, lo.w));
                    a1.x = __fadd_rn(a1.x, __fmul_rn(k1.w, lo.x)), a1.y = __fadd_rn(a1.y, __fmul_rn(k1.w, lo.y));
                    if (lo_on) a0.x = __fadd_rn(a0.x, __fmul_rn(k0.w, lo.x)), a0.y = __fadd_rn(a0.y, __fmul_rn(k0.w, lo.y));
                }
            }
#pragma unroll
            for (int j = 0; j < PAIRS; ++j) pos[j] -= 4 * BC + 2;  // the block and the padding in front of it
        };
        block(0, std::true_type{});
        c = BC;
#pragma unroll 1
        for (; c + BC <= c2; c += BC) block(c, std::false_type{});
        // in_blk is 0 here: the per-chunk code below finishes the last (partial) block and the lower-only tail
    }
    for (; c < c1; ++c) chunk(c, std::true_type{}, std::false_type{});
#pragma unroll 2
    for (; c < c2; ++c) chunk(c, std::true_type{}, std::true_type{});
    for (; c < c3; ++c) chunk(c, std::false_type{}, std::true_type{});
    float2 *o = P.out + (size_t)ch * P.out_stride;
    const long long o0 = (long long)tile * P.tile_out;
#pragma unroll
    for (int j = 0; j < PAIRS; ++j) {
        const long long i = o0 + 2 * (tid + j * T);
        const float2 y0 = make_float2(decf_limit(acc[j][0].x, P.shift), decf_limit(acc[j][0].y, P.shift));
        const float2 y1 = make_float2(decf_limit(acc[j][1].x, P.shift), decf_limit(acc[j][1].y, P.shift));
        if (i + 1 < P.n_out && ((reinterpret_cast<uintptr_t>(o + i) & 15) == 0)) {
            *reinterpret_cast<float4 *>(o + i) = make_float4(y0.x, y0.y, y1.x, y1.y);
        } else {
            if (i < P.n_out) o[i] = y0;
            if (i + 1 < P.n_out) o[i + 1] = y1;
        }
    }
}

// history <- the last N - 1 samples of (history ++ input)   (dsptl_dnsampling_filters.h:218-219; shorter blocks than
// N - 1 behave like one long call, as in the integer bank)
__global__ void decf_history_kernel(const float2 *in, size_t in_stride, long long n_in, const float2 *hist_in, float2 *hist_out, int H)
{
    const unsigned ch = blockIdx.y;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < H; j += gridDim.x * blockDim.x) {
        const long long s = n_in - H + j;
        hist_out[(size_t)ch * H + j] = s >= 0 ? __ldg(in + (size_t)ch * in_stride + s) : __ldg(hist_in + (size_t)ch * H + (H + s));
    }
}

}  // namespace srcdsp
