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
// The result depends on the order of the additions, so the kernels keep it: every output is ONE sequential chain of a
// rounded multiply and a rounded add per component (never a fused multiply-add, never a tree), and parallelism comes
// from the outputs.  The two components of a sample share a packed instruction (decf_mac below: FFMA2 with a -0.0
// addend + FADD2), which halves the issue slots without changing a single rounding.
//
// Shape.  4 FP32 lane-operations per tap and output make the FP32 pipe (128 lanes/clk/SM) the roofline; 8 bytes of
// sample per 4 lane-operations are twice what shared memory delivers (128 B/clk/SM), so samples are reused in
// registers.  decf_fir_kernel: a thread owns the output PAIR (2q, 2q+1) and walks the samples downwards from
// xx[(2q+1)*M]; sample (2q+1)*M - e is tap k = e of the upper output and tap k = e - M of the lower one.  Per 4
// samples: 2 LDS.128 (samples) + the taps (uniform constant loads from the parameter space for short filters, 2
// broadcast LDS.128 from two pre-shifted zero-padded copies otherwise) feed 16 packed instructions; a thread carries
// PAIRS such pairs.  The lane stride is 2*M samples = M 16-byte units; for even M every block of 2*M samples is followed
// by one unit of padding so that the stride is odd and the LDS.128 are conflict free.  Zero taps in front of / behind a
// filter add +-0 to its chain, which leaves every finite sum unchanged (samples must be finite -- the reference's int
// conversion of NaN/Inf is undefined).  decf_quad_kernel (further down): FOUR outputs per thread, half the
// shared-memory reads per MAC; the default from /16 on and for long filters at /8.
//
// Measured steps (256 channels, /16 x 255 taps, 1 Mi samples per channel; tools/decfbench.py):
//   plain staging loads 3.03 ms -> 8-byte cp.async (all copies of a thread in flight) 1.14 -> bulk L2 prefetch of the
//   successor tile 1.00 -> whole blocks with compile-time offsets (BC > 0) 0.92 -> taps through the parameter space
//   (CT: uniform constant loads, no tap LDS) 0.79 -> packed FFMA2 + FADD2 0.71 -> four outputs per thread 0.67 ms =
//   25.1 G out/s = 0.69 of the FP32 pipe (0.73 on bench.py's 4 Mi-sample cfg2f; profiles/r2_decf_packed_quad_ab.txt).
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
    const float *taps2;            // pair kernel [2][E]: [1][e] = c[e], [0][e] = c[e - M], zero outside the filter; quad kernel: c[0 .. 4 * rows * MC)
    int tap_floats;                // floats of taps2 (shared memory in front of the samples; 2 * E or 4 * rows * M / 4)
    int rows;                      // quad kernel: rows of M samples per output (ceil(N / M), >= 3)
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
    unsigned nz_bits, nz_mask;     // 0x80000000 and 0: -0.0f = nz_bits | (tid & nz_mask), a value the compiler can neither see nor
                                   // prove uniform (decf_mac: the uniform operand slot of the FFMA2 belongs to the tap)
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

// Packed FP32 (sm_100: FFMA2 / FADD2, two lanes of one 64-bit register pair per instruction): (re, im) of a sample are such
// a pair, so a tap costs one multiply and one add INSTRUCTION per output instead of two each -- the same FP32 pipe time
// (a packed instruction occupies the pipe for two issue cycles) at half the issue slots, which is what bound this kernel.
// Each lane rounds exactly like the scalar instruction.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (a single
// rounding: not the reference's result), explicit .rn or -fmad=false notwithstanding, so the multiply is written as
// fma(x, k, -0.0) with a -0.0 the compiler cannot see (DecfParams::nz_bits): x * k + (-0) is the rounded product with its own sign
// (+0 + -0 = +0, -0 + -0 = -0), and an fma followed by an add has no fused form.  cuobjdump -sass: FFMA2 = FADD2 count.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f32x2_bcast(float v)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float2 f32x2_unpack(f32x2_t v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// acc.re += k * x.re ; acc.im += k * x.im   (dsptl_dnsampling_filters.h:209-210 with float types: two roundings per component)
__device__ __forceinline__ void decf_mac(f32x2_t &acc, float k, f32x2_t x, f32x2_t nz)
{
    f32x2_t p;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(x), "l"(f32x2_bcast(k)), "l"(nz));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc) : "l"(acc), "l"(p));
}

// Head of both kernels: L2 prefetch for the successor CTA, taps into shared memory (unless they travel in the parameter
// space), the tile's samples into the padded staging area.  Ends with __syncthreads().
__device__ __forceinline__ void decf_stage_tile(const DecfParams &P, float *ts, float2 *xs, unsigned ch, int tile, bool taps_to_smem)
{
    const int tid = threadIdx.x, T = blockDim.x;
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
    if (taps_to_smem)
        for (int i = tid; i < P.tap_floats; i += T) ts[i] = __ldg(P.taps2 + i);
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
}

// CT: the taps are read from the parameter space (uniform constant loads: no LDS, no shared-memory wavefronts for them)
template <int PAIRS, int BC, bool CT>
__global__ void __launch_bounds__(256) decf_fir_kernel(const __grid_constant__ DecfParams P, const __grid_constant__ DecfTaps K)
{
    extern __shared__ __align__(16) uint8_t decf_smem[];
    float *ts = reinterpret_cast<float *>(decf_smem);              // [2][E]
    float2 *xs = reinterpret_cast<float2 *>(ts + P.tap_floats);    // padded samples (tap_floats * 4 bytes is a multiple of 16)
    const int tid = threadIdx.x, T = blockDim.x;
    const unsigned ch = blockIdx.x / (unsigned)P.tiles_per_ch;
    const int tile = (int)(blockIdx.x - ch * (unsigned)P.tiles_per_ch);
    decf_stage_tile(P, ts, xs, ch, tile, !CT);

    // pair q = tid + j * T: upper output 2q + 1 sits at local index (2q + 1) * M + lead = the LAST sample of block
    // q + j0 (even M) -- chunks of 4 never straddle a block
    f32x2_t acc[PAIRS][2];  // (re, im) of the lower / upper output, packed
    const f32x2_t nz = f32x2_bcast(__uint_as_float(P.nz_bits | (threadIdx.x & P.nz_mask)));
    int pos[PAIRS];
#pragma unroll
    for (int j = 0; j < PAIRS; ++j) {
        acc[j][0] = acc[j][1] = 0ull;
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
            const ulonglong2 hi = *reinterpret_cast<const ulonglong2 *>(xs + pos[j] - 1);  // {x[pos-1], x[pos]}
            const ulonglong2 lo = *reinterpret_cast<const ulonglong2 *>(xs + pos[j] - 3);  // {x[pos-3], x[pos-2]}
            f32x2_t &a0 = acc[j][0], &a1 = acc[j][1];
            if (UP) decf_mac(a1, k1.x, hi.y, nz);
            if (LO) decf_mac(a0, k0.x, hi.y, nz);
            if (UP) decf_mac(a1, k1.y, hi.x, nz);
            if (LO) decf_mac(a0, k0.y, hi.x, nz);
            if (UP) decf_mac(a1, k1.z, lo.y, nz);
            if (LO) decf_mac(a0, k0.z, lo.y, nz);
            if (UP) decf_mac(a1, k1.w, lo.x, nz);
            if (LO) decf_mac(a0, k0.w, lo.x, nz);
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
                    const ulonglong2 hi = *reinterpret_cast<const ulonglong2 *>(xs + pos[j] - 4 * i - 1);
                    const ulonglong2 lo = *reinterpret_cast<const ulonglong2 *>(xs + pos[j] - 4 * i - 3);
                    f32x2_t &a0 = acc[j][0], &a1 = acc[j][1];
                    decf_mac(a1, k1.x, hi.y, nz);
                    if (lo_on) decf_mac(a0, k0.x, hi.y, nz);
                    decf_mac(a1, k1.y, hi.x, nz);
                    if (lo_on) decf_mac(a0, k0.y, hi.x, nz);
                    decf_mac(a1, k1.z, lo.y, nz);
                    if (lo_on) decf_mac(a0, k0.z, lo.y, nz);
                    decf_mac(a1, k1.w, lo.x, nz);
                    if (lo_on) decf_mac(a0, k0.w, lo.x, nz);
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
        const float2 sum0 = f32x2_unpack(acc[j][0]), sum1 = f32x2_unpack(acc[j][1]);
        const float2 y0 = make_float2(decf_limit(sum0.x, P.shift), decf_limit(sum0.y, P.shift));
        const float2 y1 = make_float2(decf_limit(sum1.x, P.shift), decf_limit(sum1.y, P.shift));
        if (i + 1 < P.n_out && ((reinterpret_cast<uintptr_t>(o + i) & 15) == 0)) {
            *reinterpret_cast<float4 *>(o + i) = make_float4(y0.x, y0.y, y1.x, y1.y);
        } else {
            if (i < P.n_out) o[i] = y0;
            if (i + 1 < P.n_out) o[i + 1] = y1;
        }
    }
}

// The same computation with FOUR consecutive outputs per thread (M = 4 * MC in {4, 8, 16, 32}, filters of at least 3 M taps).
// With the packed MAC the pair kernel above is bound by shared memory, not by issue slots: 2 LDS.128 (8 wavefronts) feed
// 16 packed instructions = 8 SM-clocks of the FP32 pipe (ncu: shared memory 88 % busy at 0.65 of the pipe).  A quad of
// outputs reuses every sample four times from registers -- half the wavefronts per MAC.  Thread t owns outputs
// 4t .. 4t + 3 (acc[j] = output 4t + 3 - j) and walks the samples down from xx[(4t + 3) M]; sample (4t + 3) M - e is tap
// e - j M of output j.  The walk is cut into ROWS of M samples = MC chunks: output j is active in rows [j, j + rows), so
// the active set is a compile-time range per row (ramp up {0}, {0,1}, {0,1,2}; all four; ramp down {1,2,3}, {2,3}, {3}),
// every output uses the SAME tap array shifted by whole rows (chunk (r - j) MC + i: one copy of the taps, zero padded
// to whole rows -- a zero tap adds +-0), and all LDS offsets inside a row are immediates.  The lane stride is 4 M samples +
// one 16-byte unit of padding (odd in units of 16 bytes: conflict-free LDS.128); a block of 4 rows ends at a thread's top
// sample, so rows never straddle the padding.  Each output is still ONE chain in tap order (k ascends with e).
template <int MC, bool CT>
__global__ void __launch_bounds__(256) decf_quad_kernel(const __grid_constant__ DecfParams P, const __grid_constant__ DecfTaps K)
{
    extern __shared__ __align__(16) uint8_t decf_smem[];
    float *ts = reinterpret_cast<float *>(decf_smem);
    float2 *xs = reinterpret_cast<float2 *>(ts + P.tap_floats);
    const int tid = threadIdx.x;
    const unsigned ch = blockIdx.x / (unsigned)P.tiles_per_ch;
    const int tile = (int)(blockIdx.x - ch * (unsigned)P.tiles_per_ch);
    decf_stage_tile(P, ts, xs, ch, tile, !CT);

    f32x2_t acc[4] = {0ull, 0ull, 0ull, 0ull};
    const f32x2_t nz = f32x2_bcast(__uint_as_float(P.nz_bits | (threadIdx.x & P.nz_mask)));
    const int top = (4 * tid + 3) * P.M + P.lead;  // the last sample of a block
    int pos = top + P.padw * (int)__umulhi((unsigned)top, P.blk_magic);
    const float4 *tsm = reinterpret_cast<const float4 *>(ts);
    auto row = [&](int r, auto jlo_tag, auto jhi_tag) {
        constexpr int JLO = decltype(jlo_tag)::value, JHI = decltype(jhi_tag)::value;
#pragma unroll
        for (int i = 0; i < MC; ++i) {
            const ulonglong2 hi = *reinterpret_cast<const ulonglong2 *>(xs + pos - 4 * i - 1);  // {x[pos-1], x[pos]}
            const ulonglong2 lo = *reinterpret_cast<const ulonglong2 *>(xs + pos - 4 * i - 3);  // {x[pos-3], x[pos-2]}
            float4 k[4];
#pragma unroll
            for (int j = JLO; j <= JHI; ++j) k[j] = CT ? K.t[(r - j) * MC + i] : tsm[(r - j) * MC + i];
#pragma unroll
            for (int j = JLO; j <= JHI; ++j) decf_mac(acc[j], k[j].x, hi.y, nz);
#pragma unroll
            for (int j = JLO; j <= JHI; ++j) decf_mac(acc[j], k[j].y, hi.x, nz);
#pragma unroll
            for (int j = JLO; j <= JHI; ++j) decf_mac(acc[j], k[j].z, lo.y, nz);
#pragma unroll
            for (int j = JLO; j <= JHI; ++j) decf_mac(acc[j], k[j].w, lo.x, nz);
        }
        pos -= 4 * MC + ((r & 3) == 3 ? P.padw : 0);  // the row, and the padding in front of the block after its fourth row
    };
    typedef std::integral_constant<int, 0> I0;
    typedef std::integral_constant<int, 1> I1;
    typedef std::integral_constant<int, 2> I2;
    typedef std::integral_constant<int, 3> I3;
    const int R = P.rows;
    row(0, I0{}, I0{});
    row(1, I0{}, I1{});
    row(2, I0{}, I2{});
#pragma unroll 1
    for (int r = 3; r < R; ++r) row(r, I0{}, I3{});
    row(R, I1{}, I3{});
    row(R + 1, I2{}, I3{});
    row(R + 2, I3{}, I3{});

    float2 *o = P.out + (size_t)ch * P.out_stride;
    const long long i0 = (long long)tile * P.tile_out + 4 * tid;
    float2 y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 sum = f32x2_unpack(acc[3 - j]);
        y[j] = make_float2(decf_limit(sum.x, P.shift), decf_limit(sum.y, P.shift));
    }
    if (i0 + 3 < P.n_out && ((reinterpret_cast<uintptr_t>(o + i0) & 15) == 0)) {
        *reinterpret_cast<float4 *>(o + i0) = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
        *reinterpret_cast<float4 *>(o + i0 + 2) = make_float4(y[2].x, y[2].y, y[3].x, y[3].y);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j < P.n_out) o[i0 + j] = y[j];
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
