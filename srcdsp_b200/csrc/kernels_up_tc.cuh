// kernels_up_tc.cuh -- the polyphase interpolating FIR (FilterUpsamplingFir::step, upsampling_filters.h:149-233 /
// :240-323) as ONE exact int8 tcgen05 MMA per 4096 outputs.
//
//   out[j*L + p] = limitScale<cs16>( sum_{i<H} c[p + i*L] * xx[j - i], shift ),  H = Nt / L
//
// The CUDA-core kernel (kernels_up.cuh) spends 16 IMADs per output and is bound by the multiply pipe at 85 % of
// its peak.  Here a GROUP is JJ = 32 / L consecutive inputs = 32 consecutive outputs g = jj * L + p, and
//
//   D[(g, w), (J, re/im)] = sum_k A[(g, w), k] * B[(J, re/im), k]            (M = 128, N = 256, K = 32, kind::i8)
//
//   B[(J, c), k]      = byte planes of the group's sample window xx[JJ*J - (H-1) + m], m < KW = JJ + H - 1 <= 15:
//                       k = m      : low byte - 128  (as s8: low ^ 0x80)          } x = 256 * hi + lo' + 128
//                       k = 16 + m : high byte (s8)                               }
//                       k = 15     : the constant 1
//   A[(g, w), k]      = signed base-256 digits d_0..d_2 of the tap c[p + (jj - m + H - 1) * L] (0 outside the filter):
//                       slot w holds d_w on the low-plane columns and d_{w-1} on the high-plane columns, so that
//                       sum_w 256^w D_w = sum c * (256 hi + lo') = sum c * (x - 128);
//                       column 15 holds the digits of 128 * sum_i c[p + i*L], which puts the bias back.
//
// All of it mod 2^32 = the reference's int32 accumulator.  The sample operand is built by converter threads straight
// from a cp.async ring of raw input tiles (16-byte aligned channel rows; 5 tiles in flight per CTA -- a tile is only
// 2 KB and HBM latency is about two tile periods); the epilogue (8 warps, 16x256b TMEM loads: the 4 slots of an
// output sit 8 lanes apart) recombines, shifts, saturates and stores; per-tile overhead is kept off the issue slots
// (incremental channel / tile walks, no divisions), because with one MMA per tile everything else IS the kernel.
//
// Measured (tools/upbench.py, 256 ch x 1 Mi in): 0.82-0.90 T out/s for EVERY filter length it accepts -- the bound
// is the epilogue's TMEM read: 4 slots x (re, im) x 4 B = 32 B of accumulator per output at the 64-90 B/clk/SM that
// tcgen05.ld delivers (timing ablation: without the stores 2.40 instead of 2.59 ms on cfg4).  The CUDA-core kernel
// does 0.99-1.03 T out/s up to 8 taps per phase and 0.51 T out/s from 9 to 16, so this kernel is chosen for more
// than 8 taps per phase (x8 with 80 / 96 taps: 4.24 -> 2.6 ms; x16 with 224 taps: 8.4 -> 4.8 ms; x32 with 480 taps:
// 19.5 -> 9.6 ms).
#pragma once

#include "kernels_dec_tc.cuh"
#include "kernels_up.cuh"

namespace srcdsp {

constexpr int UPT_EPI_WARPS = 8;    // 2 per TMEM lane quadrant, 128 accumulator columns each
constexpr int UPT_MMA_WARP = 8;
constexpr int UPT_CONV_WARP0 = 9;
constexpr int UPT_CONV_WARPS = 4;   // 128 threads = one group each
constexpr int UPT_THREADS = 32 * (UPT_CONV_WARP0 + UPT_CONV_WARPS);
constexpr int UPT_STAGES = 4;       // sample-operand stages (converter -> MMA)
constexpr int UPT_RAW = 6;          // raw input tiles in flight (cp.async ring): HBM latency is ~2 tile periods
constexpr int UPT_GROUPS = 128;     // groups per tile (N = 2 * groups)
constexpr int UPT_A_BYTES = 128 * 32;
constexpr int UPT_STAGE_BYTES = 2 * UPT_GROUPS * 32;
inline int upt_raw_words(int JJ, int H) { return ((H - 1 + 3) & ~3) + UPT_GROUPS * JJ; }  // samples per raw tile (aligned halo + tile)
inline size_t upt_smem_bytes(int JJ, int H) { return UPT_A_BYTES + UPT_STAGES * UPT_STAGE_BYTES + (size_t)UPT_RAW * upt_raw_words(JJ, H) * 4 + 256; }

struct UpTcParams {
    const uint32_t *in;
    uint32_t *out;
    size_t in_stride, out_stride;
    long long n_in;    // real input samples per channel
    long long n_tot;   // n_in + n_flush
    int L, JJ, H, KW;  // ratio, inputs per group (32 / L), taps per phase, window samples per group (JJ + H - 1)
    int halo;          // H - 1 rounded up to a multiple of 4: a raw tile starts 16-byte aligned in the channel row
    int raw_words;     // samples per raw tile = halo + 128 * JJ
    const uint8_t *a_image;   // [2][128][16] tap operand (see above)
    const uint8_t *n_image;   // [2][32 * S][16] the same digits as the N-side operand of up_tc2_kernel: row w * 32 + g
    const uint32_t *hist_in;  // [C][H] age order
    unsigned shift;
    int tiles_per_ch;
    long long total_tiles;
    int *error_flag;
    int debug;                     // 8 = wait-cycle accounting (timing experiments)
    unsigned long long *counters;  // debug & 8: [0] conv total [1] conv wait empty [2] conv wait data [3] mma total [4] mma wait full [5] mma wait tempty [6] epi total [7] epi wait tfull
};

// sample n of the logical stream history ++ x ++ zeros (hist is age ordered: hist[H - 1] = xx[-1])
__device__ __forceinline__ uint32_t upt_sample(const uint32_t *x, const uint32_t *hist, int H, long long n_in, long long n)
{
    if (n >= 0) return n < n_in ? __ldg(x + n) : 0u;
    return n >= -(long long)H ? __ldg(hist + (H + n)) : 0u;
}
__device__ __forceinline__ void upt_cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// converter role (shared by both kernel forms): cp.async raw ring -> byte-plane sample operand stages
__device__ __forceinline__ void upt_converter_role(const UpTcParams &P, uint8_t *stages, uint32_t *raw, uint32_t bar_full, uint32_t bar_empty,
                                                   int tid, int lane, int n_tiles, unsigned ch0, int tt0, bool acct)
{
    // ================= converters: cp.async raw ring -> byte-plane sample operand =================
    // All 128 threads copy a raw tile (halo + 128 * JJ samples, 16-byte chunks) UPT_RAW - 1 tiles ahead; thread J then
    // builds the two operand rows (re, im) of group J from its KW-sample window in the raw tile.
    const int J = tid - 32 * UPT_CONV_WARP0;
    const int KW = P.KW, RW = P.raw_words, tile_span = UPT_GROUPS * P.JJ;
    const uint32_t raw_u32 = smem_u32(raw);
    // tile k of this CTA (k = 0 .. n_tiles): channel / tile-in-channel walk incrementally
    unsigned ich = ch0;   // of the next tile to ISSUE
    int itt = tt0, issued = 0;
    auto issue_tile = [&]() {
        if (issued < n_tiles) {
            const uint32_t *x = P.in + (size_t)ich * P.in_stride;
            const long long s0 = (long long)itt * tile_span - P.halo;  // stream index of raw sample 0 (multiple of 4)
            const uint32_t dst = raw_u32 + (uint32_t)((issued % UPT_RAW) * RW) * 4;
            if (s0 >= 0 && s0 + RW <= P.n_in) {
                for (int c = J; 4 * c < RW; c += 32 * UPT_CONV_WARPS) upt_cp_async16(dst + 16 * c, x + s0 + 4 * c);
            } else {  // carried history in front of the block / zeros behind it: plain stores
                const uint32_t *hist = P.hist_in + (size_t)ich * P.H;
                uint32_t *d = raw + (size_t)(issued % UPT_RAW) * RW;
                for (int i = J; i < RW; i += 32 * UPT_CONV_WARPS) d[i] = upt_sample(x, hist, P.H, P.n_in, s0 + i);
            }
            ++issued;
            if (++itt == P.tiles_per_ch) {
                itt = 0;
                ++ich;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per call, empty past the end
    };
    long long c_t0 = clock64(), c_wait = 0, c_data = 0;
    for (int d = 0; d < UPT_RAW - 1; ++d) issue_tile();
    const int woff = P.halo - (P.H - 1) + J * P.JJ;  // raw index of window sample 0 of group J
    int stage = 0;
    uint32_t par = 1;  // first wait on a fresh "empty" barrier passes
    for (int k = 0; k < n_tiles; ++k) {
        const long long c_a = acct ? clock64() : 0;
        asm volatile("cp.async.wait_group %0;" ::"n"(UPT_RAW - 2) : "memory");  // this thread's chunks of tile k have landed
        asm volatile("bar.sync 1, %0;" ::"n"(32 * UPT_CONV_WARPS) : "memory");   // everybody's; and tile k - 1 is converted
        issue_tile();                                                            // tile k + RAW - 1 into the slot of tile k - 1
        const long long c_b = acct ? clock64() : 0;
        const uint32_t *rw = raw + (size_t)(k % UPT_RAW) * RW + woff;
        uint32_t w[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) w[m] = m < KW ? rw[m] : 0u;
        uint32_t lo_re[4], hi_re[4], lo_im[4], hi_im[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            split4(make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]), lo_re[i], hi_re[i], lo_im[i], hi_im[i]);
            lo_re[i] ^= 0x80808080u;  // low byte - 128 as s8
            lo_im[i] ^= 0x80808080u;
        }
        // k = 15 of the low plane: the constant 1 that multiplies the bias column of the tap operand
        lo_re[3] = (lo_re[3] & 0x00FFFFFFu) | 0x01000000u;
        lo_im[3] = (lo_im[3] & 0x00FFFFFFu) | 0x01000000u;
        mbar_wait(bar_empty + 8 * stage, par, P.error_flag);
        const long long c_c = acct ? clock64() : 0;
        // rows n = 2 * J (re), 2 * J + 1 (im); [chunk][row][16 B]
        uint8_t *st = stages + stage * UPT_STAGE_BYTES;
        uint4 *r0 = reinterpret_cast<uint4 *>(st + (2 * J) * 16);
        uint4 *r1 = reinterpret_cast<uint4 *>(st + 2 * UPT_GROUPS * 16 + (2 * J) * 16);
        r0[0] = make_uint4(lo_re[0], lo_re[1], lo_re[2], lo_re[3]);
        r0[1] = make_uint4(lo_im[0], lo_im[1], lo_im[2], lo_im[3]);
        r1[0] = make_uint4(hi_re[0], hi_re[1], hi_re[2], hi_re[3]);
        r1[1] = make_uint4(hi_im[0], hi_im[1], hi_im[2], hi_im[3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
        if (++stage == UPT_STAGES) {
            stage = 0;
            par ^= 1;
        }
        if (acct) {
            c_data += c_b - c_a;
            c_wait += c_c - c_b;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acct && lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)(clock64() - c_t0));
        atomicAdd(P.counters + 1, (unsigned long long)c_wait);
        atomicAdd(P.counters + 2, (unsigned long long)c_data);
    }
}
// DIG3: the fourth weight slot is in use (3-digit taps, or the bias column reaches it)
template <bool DIG3>
__global__ void __launch_bounds__(UPT_THREADS, 1) up_tc_kernel(const __grid_constant__ UpTcParams P)
{
    extern __shared__ __align__(128) uint8_t upt_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *a_smem = upt_smem;
    uint8_t *stages = upt_smem + UPT_A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stages + UPT_STAGES * UPT_STAGE_BYTES);
    uint32_t *raw = reinterpret_cast<uint32_t *>(stages + UPT_STAGES * UPT_STAGE_BYTES + 256);  // [UPT_RAW][raw_words]
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * UPT_STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * UPT_STAGES, bar_tempty = bar_tfull + 16;
    __shared__ uint32_t tmem_base_s;

    for (int i = tid; i < UPT_A_BYTES / 16; i += UPT_THREADS)
        reinterpret_cast<uint4 *>(a_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(P.a_image) + i);
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < UPT_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, UPT_CONV_WARPS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, UPT_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UPT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const long long first_tile = P.total_tiles * blockIdx.x / gridDim.x, tile_end = P.total_tiles * (blockIdx.x + 1) / gridDim.x;
    const int n_tiles = (int)(tile_end - first_tile);
    const unsigned ch0 = (unsigned)(first_tile / P.tiles_per_ch);
    const int tt0 = (int)(first_tile - (long long)ch0 * P.tiles_per_ch);
    const bool acct = SRCDSP_EXP(P, 8);

    if (warp >= UPT_CONV_WARP0) {
        upt_converter_role(P, stages, raw, bar_full, bar_empty, tid, lane, n_tiles, ch0, tt0, acct);
    } else if (warp == UPT_MMA_WARP) {
        // ================= MMA issuer: one MMA per tile =================
        const uint32_t idesc = umma_idesc_i8(1, 1, 128, 2 * UPT_GROUPS);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO 128 B, descriptor version 1
        auto desc = [&](uint32_t addr, uint32_t lbo) { return ((uint64_t)desc_hi << 32) | ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); };
        const uint64_t da = desc(smem_u32(a_smem), 128 * 16);
        int stage = 0, acc = 0;
        uint32_t par = 0, acc_phases = 0;
        long long m_wt = 0, m_wf = 0;
        const long long m_t0 = clock64();
        for (int k = 0; k < n_tiles; ++k) {
            const long long m_a = acct ? clock64() : 0;
            mbar_wait(bar_tempty + 8 * acc, ((acc_phases >> acc) & 1) ^ 1, P.error_flag);
            const long long m_b = acct ? clock64() : 0;
            mbar_wait(bar_full + 8 * stage, par, P.error_flag);
            if (acct) {
                m_wt += m_b - m_a;
                m_wf += clock64() - m_b;
            }
            tc_fence_after();
            if (elect_one()) {
                umma_i8(tmem_base + acc * (2 * UPT_GROUPS), da, desc(smem_u32(stages + stage * UPT_STAGE_BYTES), 2 * UPT_GROUPS * 16), idesc, 0);
                tc_commit(bar_empty + 8 * stage);
                tc_commit(bar_tfull + 8 * acc);
            }
            __syncwarp();
            if (++stage == UPT_STAGES) {
                stage = 0;
                par ^= 1;
            }
            acc_phases ^= 1u << acc;
            acc ^= 1;
        }
        if (acct && lane == 0) {
            atomicAdd(P.counters + 3, (unsigned long long)(clock64() - m_t0));
            atomicAdd(P.counters + 4, (unsigned long long)m_wf);
            atomicAdd(P.counters + 5, (unsigned long long)m_wt);
        }
    } else {
        // ================= epilogue: warp (q, hc) = TMEM lanes 32q.., accumulator columns 128 hc .. =================
        // Two passes of 64 columns; each requests its four 16x256b loads before the first value is used.
        const int q = warp & 3, hc = warp >> 2;
        const int g = 8 * q + (lane >> 2);  // output within the group
        int acc = 0;
        uint32_t acc_phases = 0;
        const long long n_out = P.n_tot * P.L;
        long long e_w = 0;
        const long long e_t0 = clock64();
        unsigned ch = ch0;
        int tt = tt0;
        uint32_t *o = P.out + (size_t)ch * P.out_stride;
        const int lane_off = (64 * hc + (lane & 3)) * 32 + g;  // (group 64 hc + lane & 3, output g) inside the tile
#define UPT_LD16(dst, addr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 "                                                               \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                        \
                 : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]), "=r"(dst[4]), "=r"(dst[5]), "=r"(dst[6]),      \
                   "=r"(dst[7]), "=r"(dst[8]), "=r"(dst[9]), "=r"(dst[10]), "=r"(dst[11]), "=r"(dst[12]), "=r"(dst[13]),  \
                   "=r"(dst[14]), "=r"(dst[15])                                                                            \
                 : "r"(addr))
// ties the loaded registers to the wait: nothing that uses them may be scheduled above it
#define UPT_TOUCH16(r)                                                                                                   \
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),     \
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]))
        for (int kt = 0; kt < n_tiles; ++kt) {
            const long long tile_out0 = (long long)tt * (UPT_GROUPS * 32);
            const bool full_tile = tile_out0 + UPT_GROUPS * 32 <= n_out;
            const int left = full_tile ? 0x7fffffff : (int)(n_out - tile_out0);  // outputs of this tile that exist
            uint32_t *ot = o + tile_out0 + lane_off;
            const long long e_a = acct ? clock64() : 0;
            mbar_wait(bar_tfull + 8 * acc, (acc_phases >> acc) & 1, P.error_flag);
            if (acct) e_w += clock64() - e_a;
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * (2 * UPT_GROUPS) + 128 * hc;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                uint32_t a0[16], h0[16], a1[16], h1[16];
                UPT_LD16(a0, t_addr + 64 * pass);
                UPT_LD16(h0, t_addr + 64 * pass + (16u << 16));
                UPT_LD16(a1, t_addr + 64 * pass + 32);
                UPT_LD16(h1, t_addr + 64 * pass + 32 + (16u << 16));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                UPT_TOUCH16(a0);
                UPT_TOUCH16(h0);
                UPT_TOUCH16(a1);
                UPT_TOUCH16(h1);
                if (pass == 1) {  // the accumulator is in registers: hand it back before the arithmetic and the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                }
                auto finish = [&](const uint32_t *a, const uint32_t *h, int grp) {  // groups grp + 4 * k of this lane
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // sum_w 256^w * D_w (mod 2^32, exactly the reference's int32 wrap)
                        uint32_t re = a[4 * k] + (a[4 * k + 2] << 8) + (h[4 * k] << 16);
                        uint32_t im = a[4 * k + 1] + (a[4 * k + 3] << 8) + (h[4 * k + 1] << 16);
                        if (DIG3) {
                            re += h[4 * k + 2] << 24;
                            im += h[4 * k + 3] << 24;
                        }
                        const int off = (grp + 4 * k) * 32;
                        const uint32_t word = scale_pack_asym_sat((int)re, (int)im, P.shift);
                        if (SRCDSP_EXP(P, 1) ? word == 0x12345678u : (full_tile || lane_off + off < left)) ot[off] = word;
                    }
                };
                finish(a0, h0, 32 * pass);
                finish(a1, h1, 32 * pass + 16);
            }
            acc_phases ^= 1u << acc;
            acc ^= 1;
            if (++tt == P.tiles_per_ch) {
                tt = 0;
                ++ch;
                o = P.out + (size_t)ch * P.out_stride;
            }
        }
        if (acct && lane == 0) {
            atomicAdd(P.counters + 6, (unsigned long long)(clock64() - e_t0));
            atomicAdd(P.counters + 7, (unsigned long long)e_w);
        }
#undef UPT_LD16
#undef UPT_TOUCH16
    }

    tc_fence_before();
    __syncthreads();
    if (warp == UPT_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second form: operand roles swapped.  A = the sample operand (M = 128 rows = 64 groups x (re, im)), B = the taps
// (N = 32 * S rows: slot-major, row w * 32 + g), two MMAs per tile (groups 0-63 and 64-127).  An accumulator lane then
// holds ONE component of ONE group and all S slots of its 32 outputs sit in that lane's columns: the epilogue reads
// exactly S slots (24 bytes of TMEM per output for 2-digit taps instead of 32 -- the TMEM read rate is what bounds this
// kernel), needs no cross-lane slot reduction, and after one exchange with the neighbouring lane (re <-> im) every
// thread owns 8 complete consecutive outputs = two STG.128.
// MEASURED (tools/upbench.py, 256 ch x 1 Mi in, same box): bit-identical, but 3-12 % slower than up_tc_kernel for
// every shape (x8/96 taps 2.93 vs 2.61 ms, x16/224 taps 5.44 vs 4.78 ms, x32/480 taps 9.88 vs 9.62 ms): 98 KB of
// accumulator per tile through 32x32b.x16 loads come out at ~60 B/clk/SM, where the 16x256b loads of the first form
// deliver ~90 B/clk/SM for its 131 KB.  The bytes were not the bound -- the load shape is.  Kept as an opt-in
// (SRCDSP_UP_TC_FORM=2) and as a parity-tested record of the experiment.
// ---------------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(UPT_THREADS, 1) up_tc2_kernel(const __grid_constant__ UpTcParams P)
{
    extern __shared__ __align__(128) uint8_t upt_smem[];
    constexpr int N = 32 * S;                 // accumulator columns per MMA
    constexpr int B_BYTES = N * 32;           // taps operand: [2][N][16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *b_smem = upt_smem;               // taps (4 KB reserved: UPT_A_BYTES)
    uint8_t *stages = upt_smem + UPT_A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stages + UPT_STAGES * UPT_STAGE_BYTES);
    uint32_t *raw = reinterpret_cast<uint32_t *>(stages + UPT_STAGES * UPT_STAGE_BYTES + 256);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * UPT_STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * UPT_STAGES, bar_tempty = bar_tfull + 16;
    __shared__ uint32_t tmem_base_s;

    for (int i = tid; i < B_BYTES / 16; i += UPT_THREADS)
        reinterpret_cast<uint4 *>(b_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(P.n_image) + i);
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < UPT_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, UPT_CONV_WARPS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, UPT_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UPT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const long long first_tile = P.total_tiles * blockIdx.x / gridDim.x, tile_end = P.total_tiles * (blockIdx.x + 1) / gridDim.x;
    const int n_tiles = (int)(tile_end - first_tile);
    const unsigned ch0 = (unsigned)(first_tile / P.tiles_per_ch);
    const int tt0 = (int)(first_tile - (long long)ch0 * P.tiles_per_ch);
    const bool acct = SRCDSP_EXP(P, 8);

    if (warp >= UPT_CONV_WARP0) {
        upt_converter_role(P, stages, raw, bar_full, bar_empty, tid, lane, n_tiles, ch0, tt0, acct);
    } else if (warp == UPT_MMA_WARP) {
        // ================= MMA issuer: two MMAs per tile (groups 0-63, 64-127) =================
        const uint32_t idesc = umma_idesc_i8(1, 1, 128, N);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO 128 B, descriptor version 1
        auto desc = [&](uint32_t addr, uint32_t lbo) { return ((uint64_t)desc_hi << 32) | ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); };
        const uint64_t db = desc(smem_u32(b_smem), N * 16);
        int stage = 0, acc = 0;
        uint32_t par = 0, acc_phases = 0;
        long long m_wt = 0, m_wf = 0;
        const long long m_t0 = clock64();
        for (int k = 0; k < n_tiles; ++k) {
            const long long m_a = acct ? clock64() : 0;
            mbar_wait(bar_tempty + 8 * acc, ((acc_phases >> acc) & 1) ^ 1, P.error_flag);
            const long long m_b = acct ? clock64() : 0;
            mbar_wait(bar_full + 8 * stage, par, P.error_flag);
            if (acct) {
                m_wt += m_b - m_a;
                m_wf += clock64() - m_b;
            }
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = smem_u32(stages + stage * UPT_STAGE_BYTES);
                // sample rows 0-127 / 128-255 of the stage ([chunk][256 rows][16 B]: the chunk stride stays 4096)
                umma_i8(tmem_base + acc * (2 * N), desc(st, 2 * UPT_GROUPS * 16), db, idesc, 0);
                umma_i8(tmem_base + acc * (2 * N) + N, desc(st + 128 * 16, 2 * UPT_GROUPS * 16), db, idesc, 0);
                tc_commit(bar_empty + 8 * stage);
                tc_commit(bar_tfull + 8 * acc);
            }
            __syncwarp();
            if (++stage == UPT_STAGES) {
                stage = 0;
                par ^= 1;
            }
            acc_phases ^= 1u << acc;
            acc ^= 1;
        }
        if (acct && lane == 0) {
            atomicAdd(P.counters + 3, (unsigned long long)(clock64() - m_t0));
            atomicAdd(P.counters + 4, (unsigned long long)m_wf);
            atomicAdd(P.counters + 5, (unsigned long long)m_wt);
        }
    } else {
        // ================= epilogue: warp (q, sub) = TMEM lanes 32q.. of the tile's MMA `sub` =================
        // lane = (group 16q + lane / 2 of the sub-tile, component lane & 1); columns w * 32 + g
        const int q = warp & 3, sub = warp >> 2;
        const int cmp = lane & 1;
        const int grp = 64 * sub + 16 * q + (lane >> 1);  // group within the tile
        int acc = 0;
        uint32_t acc_phases = 0;
        const long long n_out = P.n_tot * P.L;
        long long e_w = 0;
        const long long e_t0 = clock64();
        unsigned ch = ch0;
        int tt = tt0;
        uint32_t *o = P.out + (size_t)ch * P.out_stride;
        // after the exchange this lane owns outputs 8 * cmp .. 8 * cmp + 7 of each half (16 outputs) of its group
        const int lane_off = grp * 32 + 8 * cmp;
        const uint32_t swap_sel = cmp ? 0x1032u : 0x3210u;  // odd lanes keep im and receive re: halves swapped after the pack
#define UPT_LD32x16(dst, addr)                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                               \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                        \
                 : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]), "=r"(dst[4]), "=r"(dst[5]), "=r"(dst[6]),      \
                   "=r"(dst[7]), "=r"(dst[8]), "=r"(dst[9]), "=r"(dst[10]), "=r"(dst[11]), "=r"(dst[12]), "=r"(dst[13]),  \
                   "=r"(dst[14]), "=r"(dst[15])                                                                            \
                 : "r"(addr))
#define UPT_TOUCH16(r)                                                                                                   \
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),     \
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]))
        for (int kt = 0; kt < n_tiles; ++kt) {
            const long long tile_out0 = (long long)tt * (UPT_GROUPS * 32);
            const bool full_tile = tile_out0 + UPT_GROUPS * 32 <= n_out;
            const int left = full_tile ? 0x7fffffff : (int)(n_out - tile_out0);  // outputs of this tile that exist
            uint32_t *ot = o + tile_out0 + lane_off;
            const long long e_a = acct ? clock64() : 0;
            mbar_wait(bar_tfull + 8 * acc, (acc_phases >> acc) & 1, P.error_flag);
            if (acct) e_w += clock64() - e_a;
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * (2 * N) + sub * N;
#pragma unroll
            for (int half = 0; half < 2; ++half) {  // outputs 16 * half .. + 15 of the group
                uint32_t d0[16], d1[16], d2[16], d3[16];
                UPT_LD32x16(d0, t_addr + 16 * half);
                UPT_LD32x16(d1, t_addr + 32 + 16 * half);
                UPT_LD32x16(d2, t_addr + 64 + 16 * half);
                if (S == 4) UPT_LD32x16(d3, t_addr + 96 + 16 * half);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                UPT_TOUCH16(d0);
                UPT_TOUCH16(d1);
                UPT_TOUCH16(d2);
                if (S == 4) UPT_TOUCH16(d3);
                if (half == 1) {  // the accumulator is in registers: hand it back before the arithmetic and the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                }
                // this lane's component of 16 outputs: sum_w 256^w * D_w (mod 2^32), >> shift, saturated pairs
                uint32_t pr[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t v0 = d0[2 * i] + (d1[2 * i] << 8) + (d2[2 * i] << 16);
                    uint32_t v1 = d0[2 * i + 1] + (d1[2 * i + 1] << 8) + (d2[2 * i + 1] << 16);
                    if (S == 4) {
                        v0 += d3[2 * i] << 24;
                        v1 += d3[2 * i + 1] << 24;
                    }
                    pr[i] = scale_pack_asym_sat((int)v0, (int)v1, P.shift);  // {hi = output 2i+1, lo = output 2i} of ONE component
                }
                // exchange with the other component's lane: re lanes keep outputs 0-7 (pairs 0-3), im lanes 8-15 (pairs 4-7)
                uint32_t w4[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t keep = cmp ? pr[4 + i] : pr[i];
                    const uint32_t send = cmp ? pr[i] : pr[4 + i];
                    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    // keep = {a1, a0} of my component, recv = {b1, b0} of the other: words {b0, a0}, {b1, a1}, halves swapped on im lanes
                    w4[2 * i] = prmt(prmt(keep, recv, 0x5410), 0, swap_sel);
                    w4[2 * i + 1] = prmt(prmt(keep, recv, 0x7632), 0, swap_sel);
                }
                uint32_t *op = ot + 16 * half;
                if (full_tile || lane_off + 16 * half + 8 <= left) {
                    *reinterpret_cast<uint4 *>(op) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    *reinterpret_cast<uint4 *>(op + 4) = make_uint4(w4[4], w4[5], w4[6], w4[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (lane_off + 16 * half + i < left) op[i] = w4[i];
                }
            }
            acc_phases ^= 1u << acc;
            acc ^= 1;
            if (++tt == P.tiles_per_ch) {
                tt = 0;
                ++ch;
                o = P.out + (size_t)ch * P.out_stride;
            }
        }
        if (acct && lane == 0) {
            atomicAdd(P.counters + 6, (unsigned long long)(clock64() - e_t0));
            atomicAdd(P.counters + 7, (unsigned long long)e_w);
        }
#undef UPT_LD32x16
#undef UPT_TOUCH16
    }

    tc_fence_before();
    __syncthreads();
    if (warp == UPT_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace srcdsp
