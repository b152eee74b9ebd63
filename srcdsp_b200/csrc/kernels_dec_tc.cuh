// kernels_dec_tc.cuh -- polyphase decimating FIR as an EXACT int8 Toeplitz GEMM on the 5th-gen
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).  Same contract as dec_fir_kernel:
//
//   out[i] = limitScale16( sum_{k<Nt} c[k] * xx[i*M - k], shift )      (bit-exact, mod 2^32)
//
// Why: the INT32 multiply pipe caps the CUDA-core kernel at 2*Nt IMAD per output (36.5 G out/s for
// /16, 255 taps); the same sum split into byte planes is an int8 GEMM that the
// tensor cores finish fast enough for HBM to become the limit.
//
// Decomposition (all exact):
//   sample component v (int16)  = 256 * hi + lo,   hi = v >> 8 (s8),  lo = v & 255 (u8)
//   tap c[k]                    = sum_{pl<P} 256^pl * d_pl[k],  d_pl in [-128,127] (signed digits)
//   sum_k c[k] v[.]             = sum_w 256^w * D_w,   D_w = sum_{pl+sp=w} sum_k d_pl[k] * plane_sp[.]
//
// GEMM shape per tile (one channel, 128 row-blocks of G = 32*M samples, 4096 outputs):
//   D[rho, n] += A[rho, t] * B[n, t]          M=128 rows, N=256 columns, K=32 bytes per MMA
//   B (samples): row n = (row-block m, component re/im), K = 32 consecutive samples of one byte
//       plane -> SWIZZLE_NONE K-major, rows linear at 16-byte pitch, so the operand for "lag j"
//       (previous row-blocks) is the same shared-memory tile with the start address moved back
//       by 2*j rows.
//   A (taps):   row rho = 4*b + w (output b of the row-block, weight slot w), one resident
//       "master" Toeplitz matrix T[4u + w, t] = d_w[M*u - t - r]; every (K-step, lag) operand is
//       the master with the start address moved by 4*(32*j - a) rows, and the operand for the hi
//       byte plane is the same moved back by ONE row (weight slot w -> w+1).
//   (the linear-address behaviour of SWIZZLE_NONE descriptors is checked by tools/umma_probe.cu)
//
// Warp roles (416 threads, 1 CTA / SM, persistent over tiles):
//   warps 0-3  epilogue: tcgen05.ld 32x32b, combine the 4 weight slots with 2 shuffles,
//              >> shift, symmetric clamp, pack, store
//   warp  4    TMEM alloc + single-thread MMA issue, tcgen05.commit -> mbarriers
//   warps 5-12 producers: warp k owns every 8th K-step and shared-memory stage k, so that no
//              warp sits on the critical path of every K-step: LDG.128 (8 lanes per 128-byte
//              line, two batches of 8 loads in flight per lane) -> PRMT byte-plane split -> STS.32
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace srcdsp {

constexpr int TC_NRB = 128;              // row-blocks per tile (MMA N = 2 * NRB)
constexpr int TC_BOUT = 32;              // outputs per row-block
constexpr int TC_NPW = 16;               // producer warps: HBM streaming scales with warps, not with loads per warp (tools/ldbench.cu)
constexpr int TC_SPLIT = 2;              // warps sharing one K-step (half of the row-blocks each)
constexpr int TC_OWNERS = TC_NPW / TC_SPLIT;  // K-steps being staged concurrently
constexpr int TC_PROD_WARP0 = 5;         // first producer warp
constexpr int TC_PF_WARP = TC_PROD_WARP0 + TC_NPW;  // L2 prefetch issuer (one lane)
constexpr int TC_THREADS = 32 * (TC_PF_WARP + 1);
constexpr int TC_MAX_STAGES = 12;         // stage ring is decoupled from the producer warps: step gs -> stage gs % n_stages
constexpr int TC_BATCH = 4;              // 16-byte loads per lane per batch (4 row-blocks per iteration); two batches in flight
constexpr int TC_MAX_KSTEPS = 64;        // M <= 64
constexpr int TC_MAX_J = 16;

struct TcKstep {
    int a_row;        // master row of (b = 0, w = 0) for lag 0:  4 * (32 - a) + 4 guard rows
    int res_off;      // byte offset of the residue's master inside the A region
    unsigned jmask;   // bit j set: lag j contributes
};

struct TcParams {
    const uint32_t *in;
    uint32_t *out;
    size_t in_stride, out_stride;
    long long n_in, n_out;
    int nrb;          // row-blocks per tile: TC_NRB (128), or 64 in "p2" mode
    int p2;           // p2 mode (taps of at most 2 digits): 64 outputs x 2 digit slots per row-block, the two byte planes
                      // accumulate into separate halves of the accumulator (128 columns each) and meet in the epilogue
    int ksteps;       // K-steps per tile: M, or 2 * M in p2 mode (row-blocks of 64 * M samples)
    int M;            // decimation ratio
    unsigned epi_sleep_ns;  // epilogue back-off between polls of the accumulator barrier
    unsigned conv_sleep_ns; // TMA variant: converter back-off between polls of the raw-stage barrier (0: plain try_wait loop)
    unsigned mma_sleep_ns;  // MMA warp's back-off between polls of the byte-plane ring (0: plain try_wait loop).  The warp
                            // shares a scheduler with converter warps; its polling takes their issue slots
    int G;            // 32 * M samples per row-block
    int J;            // lags: 1 + ceil((Nt-1)/G)
    int tiles_per_ch;
    long long total_tiles;
    const uint8_t *master;   // device image of the A region (+ the MMA plan behind it)
    int master_bytes;
    int plan_hdr_off, plan_ent_off;  // byte offsets of the MMA plan inside the image: per K-step {first entry, count},
                                     // per (K-step, lag) entry the 4 operand addresses >> 4 (see tc_mma_role)
    int a_rows;              // rows per (residue, kc) chunk  -> LBO_A = a_rows * 16
    int rbp;                 // padded rows per (plane, kc) chunk of a stage (odd) -> LBO_B = rbp * 16
    int front_pad;           // unused rows in front of every chunk: 2 * (4*ceil((J-1)/4) - (J-1))
    int grouped;             // A-row layout: 0: rho = 4*b + w (shuffle epilogue); 1: rho = 32*(b>>3) + 8*w + (b&7)
                             // (weight slots 8 rows apart -> one thread holds all slots of an output, 16x256b loads)
    int n_stages;            // shared-memory stages (TC_NPW < n_stages <= TC_MAX_STAGES)
    const uint32_t *hist_in;
    int H;
    unsigned shift;
    int vec_in;
    int vec_out;             // dec_band_kernel: output rows are 16-byte aligned (STG.128)
    int *error_flag;         // mapped host memory: [0] set by a timed-out barrier wait, [1..4] which one
    int *counters;           // device memory: wait-cycle counters of the timing variants
    // fused NCO mix (MIX instantiation): packed (cos, sin) table of n_table = mix_mask + 1 entries (power of two)
    const uint32_t *cs_table;
    const int *phi, *freq;   // [C] phase at sample 0 of this step, frequency
    unsigned mix_mask;
    unsigned seq_mask;       // dec_tma_kernel: length - 1 of the oscillator sequence kept in shared memory (a period of every channel's oscillator)
    int table_bytes;         // bytes of the shared-memory copy of the table (0 without MIX)
    int rb_stride, kc_stride;  // samples between row-blocks / K-steps (G and 32; timing experiments permute them)
    int pf_dist;  // dec_tc_kernel: K-steps the L2 prefetcher runs ahead of the MMAs (0: no prefetch, no tensor map)
    int debug;  // timing experiments only (results become wrong): 1 = skip the MMAs, 4 = skip the epilogue math
    TcKstep ks[TC_MAX_KSTEPS];
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.  The bound is WALL-CLOCK time
// (%globaltimer, checked every 1024 failed attempts): a legitimate stall -- time-slicing under MPS, compute
// preemption, a debugger or sanitizer, throttled clocks under a persistent CTA -- is far shorter than
// TC_WAIT_TIMEOUT_NS, a lost arrival is not.  On a timeout the waiter records which barrier in mapped host memory
// (the record survives) and traps: the trap ends the kernel but poisons the CUDA context, so the next call on the
// handle reports SRCDSP_E_CUDA and the process has to tear the context down.  It is a bug tripwire, not a
// recoverable condition.
constexpr unsigned long long TC_WAIT_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000u)  // suspend-time hint (ns)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int *error_flag)
{
    if (error_flag && atomicExch(error_flag, 1) == 0) {
        error_flag[1] = (int)bar;
        error_flag[2] = (int)parity;
        error_flag[3] = (int)blockIdx.x;
        error_flag[4] = (int)threadIdx.x;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// The loop is a handful of instructions while the phase is incomplete (try_wait suspends the thread for up to the
// hint, so waiting warps leave the issue slots to the working ones).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *error_flag)
{
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t n = 1;; ++n) {
        if (mbar_try_wait(bar, parity)) return;
        if ((n & 1023u) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0)
                t0 = now;
            else if (now - t0 > TC_WAIT_TIMEOUT_NS)
                mbar_timeout(bar, parity, error_flag);
        }
    }
}
// Same for a role that waits long and is not on the critical path (the epilogue warps wait most of a tile for the
// accumulator): try_wait's own suspend ends at every mbarrier event of the CTA (~40 ns apart here), so a failed
// attempt is followed by a plain sleep -- ~10x fewer polling instructions taken from the working warps' issue slots.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, int *error_flag, uint32_t sleep_ns)
{
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t n = 1;; ++n) {
        __nanosleep(sleep_ns);
        if (mbar_try_wait(bar, parity)) return;
        if ((n & 1023u) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0)
                t0 = now;
            else if (now - t0 > TC_WAIT_TIMEOUT_NS)
                mbar_timeout(bar, parity, error_flag);
        }
    }
}
// DBG & 16: per-role wait-cycle accounting into P.error_flag[1..] (timing experiments)
template <int DBG>
__device__ __forceinline__ void mbar_wait_acc(uint32_t bar, uint32_t parity, int *error_flag, long long &acc)
{
    if (DBG & 16) {
        const long long t0 = clock64();
        mbar_wait(bar, parity, error_flag);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity, error_flag);
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes)
{
    // SWIZZLE_NONE, K-major: start >> 4 | LBO >> 4 << 16 | SBO(128 B) >> 4 << 32 | version 1 << 46
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_i8(int a_signed, int b_signed, int m, int n)
{
    // c_format S32 (2) @4, a_format @7, b_format @10 (0 = u8, 1 = s8), K-major both, N>>3 @17, M>>4 @24
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// generic sample fetch: history for n < 0, zero outside [−H, n_in)
__device__ __forceinline__ uint32_t tc_sample(const uint32_t *x, const uint32_t *hist, int H, long long n_in, long long n)
{
    if (n >= 0) return n < n_in ? __ldg(x + n) : 0u;
    return n >= -(long long)H ? __ldg(hist + (H + n)) : 0u;
}
// same with the NCO mix applied to block samples (history is already mixed); ph0/fr: channel phase at
// sample 0 and frequency, table in global memory (edge path only)
__device__ __forceinline__ uint32_t tc_sample_mix(const TcParams &P, const uint32_t *x, const uint32_t *hist, long long n,
                                                  unsigned ph0, unsigned fr)
{
    const uint32_t w = tc_sample(x, hist, P.H, P.n_in, n);
    if (n < 0 || n >= P.n_in) return w;
    const unsigned ph = (ph0 + ((unsigned)n & P.mix_mask) * fr) & P.mix_mask;
    return mix_sample_packed(w, __ldg(P.cs_table + ph));
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// 4 consecutive samples -> one 32-bit word per byte plane
__device__ __forceinline__ void split4(const uint4 q, uint32_t &re_lo, uint32_t &re_hi, uint32_t &im_lo, uint32_t &im_hi)
{
    const uint32_t a = prmt(q.x, q.y, 0x5140);  // {w0.b0, w1.b0, w0.b1, w1.b1}
    const uint32_t b = prmt(q.z, q.w, 0x5140);
    const uint32_t c = prmt(q.x, q.y, 0x7362);  // {w0.b2, w1.b2, w0.b3, w1.b3}
    const uint32_t d = prmt(q.z, q.w, 0x7362);
    re_lo = prmt(a, b, 0x5410);
    re_hi = prmt(a, b, 0x7632);
    im_lo = prmt(c, d, 0x5410);
    im_hi = prmt(c, d, 0x7632);
}

// A batch = up to TC_BATCH iterations; iteration `it` covers row-blocks rb + 4*it (rb includes
// the lane's group lane >> 3), each lane one 16-byte piece (4 samples): 8 lanes read one whole
// 128-byte line.
struct TcBatch {
    uint4 q[TC_BATCH];
    bool fast;  // q[] holds prefetched data; otherwise the batch is done by tc_batch_generic at store time
};

// Interior batch (warp-uniform decision): straight-line, no per-iteration predicates.
template <int DBG>
__device__ __forceinline__ void tc_batch_load_fast(TcBatch &t, const uint32_t *p, size_t stride_words, int nit)
{
    t.fast = true;
#pragma unroll
    for (int it = 0; it < TC_BATCH; ++it) {
        if (it < nit) {
            if (DBG & 2)
                t.q[it] = make_uint4(it, 2, 3, 4);  // timing experiment: no global loads
            else
                t.q[it] = ldg_stream(reinterpret_cast<const uint4 *>(p + it * stride_words));
        }
    }
}

// MIX: tab = shared-memory copy of the packed (cos, sin) table, ph = phase of this lane's first
// sample of iteration 0, fr = channel frequency, dph = phase step between iterations (4 row-blocks)
template <int DBG, bool MIX>
__device__ __forceinline__ void tc_batch_store_fast(const TcBatch &t, uint8_t *lo, uint8_t *hi, int nit, const uint32_t *tab,
                                                    unsigned ph, unsigned fr, unsigned dph, unsigned mask)
{
    uint32_t sink = 0;
#pragma unroll
    for (int it = 0; it < TC_BATCH; ++it) {
        if (it < nit) {
            uint32_t re_lo, re_hi, im_lo, im_hi;
            uint4 q = t.q[it];
            if (MIX) {  // mixers.h:172-177 on the 4 samples of this piece, then the same byte-plane split
                const unsigned p0 = (ph + it * dph) & mask;
                q.x = mix_sample_packed(q.x, tab[p0]);
                q.y = mix_sample_packed(q.y, tab[(p0 + fr) & mask]);
                q.z = mix_sample_packed(q.z, tab[(p0 + 2 * fr) & mask]);
                q.w = mix_sample_packed(q.w, tab[(p0 + 3 * fr) & mask]);
            }
            split4(q, re_lo, re_hi, im_lo, im_hi);
            if (DBG & 8) {  // timing experiment: no shared-memory stores
                sink ^= re_lo ^ re_hi ^ im_lo ^ im_hi;
                continue;
            }
            *reinterpret_cast<uint32_t *>(lo + it * 128) = re_lo;  // 4 row-blocks further = 8 rows = 128 bytes
            *reinterpret_cast<uint32_t *>(lo + it * 128 + 16) = im_lo;
            *reinterpret_cast<uint32_t *>(hi + it * 128) = re_hi;
            *reinterpret_cast<uint32_t *>(hi + it * 128 + 16) = im_hi;
        }
    }
    if ((DBG & 8) && sink == 0x12345678u) *reinterpret_cast<uint32_t *>(lo) = sink;
}

// Edge batches (history in front of the block, ragged end, unaligned buffers): scalar, checked,
// not prefetched.  Out of line to keep the hot loop small.
// n0: sample index of (row-block rb, this lane's piece, this K-step); dst: its (plane lo, re) word.
__device__ __noinline__ void tc_batch_generic(const TcParams &P, const uint32_t *x, const uint32_t *hist, long long n0,
                                              int rb, int rb_lo, int rb_hi, int nit, uint8_t *dst, bool mix, unsigned ph0,
                                              unsigned fr)
{
    const int chunk = P.rbp * 16;  // bytes per (plane, kc) chunk
    for (int it = 0; it < nit; ++it) {
        const int r = rb + 4 * it;
        if (r < rb_lo || r >= rb_hi) continue;
        const long long n = n0 + (long long)it * 4 * P.rb_stride;
        uint4 q;
        if (mix) {
            q.x = tc_sample_mix(P, x, hist, n, ph0, fr);
            q.y = tc_sample_mix(P, x, hist, n + 1, ph0, fr);
            q.z = tc_sample_mix(P, x, hist, n + 2, ph0, fr);
            q.w = tc_sample_mix(P, x, hist, n + 3, ph0, fr);
        } else {
            q.x = tc_sample(x, hist, P.H, P.n_in, n);
            q.y = tc_sample(x, hist, P.H, P.n_in, n + 1);
            q.z = tc_sample(x, hist, P.H, P.n_in, n + 2);
            q.w = tc_sample(x, hist, P.H, P.n_in, n + 3);
        }
        uint32_t re_lo, re_hi, im_lo, im_hi;
        split4(q, re_lo, re_hi, im_lo, im_hi);
        uint8_t *lo = dst + it * 128;
        uint8_t *hi = lo + 2 * chunk;
        *reinterpret_cast<uint32_t *>(lo) = re_lo;
        *reinterpret_cast<uint32_t *>(lo + 16) = im_lo;
        *reinterpret_cast<uint32_t *>(hi) = re_hi;
        *reinterpret_cast<uint32_t *>(hi + 16) = im_hi;
    }
}

// ---------------------------------------------------------------------------------------------
// Roles shared by the two tensor-core kernels (dec_tc_kernel: LDG producers; dec_tma_kernel in
// kernels_dec_tma.cuh: TMA-fed raw ring + converter warps).  "stages" is the ring of byte-plane
// sample stages the MMAs read; full / empty are its mbarriers.
// ---------------------------------------------------------------------------------------------
struct TcRole {
    uint8_t *a_smem, *stages;
    int stage_bytes, NS, J, KS, warp, lane;
    uint32_t bar_full, bar_empty, bar_tfull, bar_tempty, tmem_base;
    long long first_tile, tile_step, tile_end;  // this CTA's tiles: first_tile, first_tile + tile_step, ... < tile_end
    uint32_t progress;  // != 0: shared-memory word that receives the number of K-steps issued so far (L2 prefetch pacing)
};

// MMA issuer (one warp, one elected lane issues)
template <int DBG>
__device__ __forceinline__ void tc_mma_role(const TcParams &P, const TcRole &R)
{
    uint8_t *const a_smem = R.a_smem, *const stages = R.stages;
    const int stage_bytes = R.stage_bytes, NS = R.NS, J = R.J, KS = R.KS, warp = R.warp, lane = R.lane;
    const uint32_t bar_full = R.bar_full, bar_empty = R.bar_empty, bar_tfull = R.bar_tfull, bar_tempty = R.bar_tempty;
    const uint32_t tmem_base = R.tmem_base;
    const long long first_tile = R.first_tile, tile_step = R.tile_step;
    (void)a_smem, (void)stages, (void)stage_bytes, (void)NS, (void)J, (void)KS, (void)warp, (void)lane;
    (void)bar_full, (void)bar_empty, (void)bar_tfull, (void)bar_tempty, (void)tmem_base;
    // The whole warp walks the loop (warp-uniform control flow and operands, so descriptors
    // stay in uniform registers); one elected lane issues the MMAs and the commits.
    //
    // The issuing thread is a serial resource: everything address-like is precomputed on the host
    // into the "MMA plan" that sits behind the master image in shared memory (per K-step a header
    // {first entry, count}, per contributing lag one entry {A lo, A hi, B lo, B hi} of operand
    // addresses >> 4, A relative to the master, B relative to the sample stage), and the plan of
    // the next K-step is fetched before this one's barrier wait.  Per K-step that leaves: wait,
    // fence, a handful of adds, the MMAs, the commit.
    {
        const uint32_t idesc_lo = umma_idesc_i8(1, 0, 128, 2 * P.nrb);  // taps s8 x lo plane u8
        const uint32_t idesc_hi = umma_idesc_i8(1, 1, 128, 2 * P.nrb);  // taps s8 x hi plane s8
        const uint32_t hi_cols = P.p2 ? 2 * P.nrb : 0;  // p2: the hi plane has its own accumulator columns
        const uint32_t a_base = smem_u32(a_smem), s_base = smem_u32(stages);
        // descriptor = {high word: SBO 128 B, version 1} {low word: LBO >> 4 << 16 | address >> 4}
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_const = (a_base >> 4) + (((uint32_t)P.a_rows * 16 >> 4) << 16);
        const uint32_t b_const = (s_base >> 4) + (((uint32_t)P.rbp * 16 >> 4) << 16);
        const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
        const uint2 *plan_hdr = reinterpret_cast<const uint2 *>(a_smem + P.plan_hdr_off);
        const uint4 *plan_ent = reinterpret_cast<const uint4 *>(a_smem + P.plan_ent_off);
        const bool no_mma = SRCDSP_EXP(P, 1), no_hi = SRCDSP_EXP(P, 64);
        auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
        int stage = 0;
        uint32_t sb16 = 0;  // (stage * stage_bytes) >> 4
        uint32_t phase = 0;
        uint32_t acc_phases = 0;  // bit a: parity of accumulator buffer a
        int acc = 0;
        long long m_full = 0, m_tempty = 0;
        uint32_t steps_done = 0;
        const long long m_t0 = clock64();
        // plan of the first K-step (the entry table is padded, so the second load is always in bounds)
        uint2 hdr = plan_hdr[0];
        uint4 e0 = plan_ent[hdr.x], e1 = plan_ent[hdr.x + 1];
        for (long long tile = first_tile; tile < R.tile_end; tile += tile_step) {
            mbar_wait_acc<DBG>(bar_tempty + 8 * acc, ((acc_phases >> acc) & 1) ^ 1, P.error_flag, m_tempty);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (2 * TC_NRB);  // (p2: 128 + 128 columns of the same 256)
            uint32_t accumulate = 0;
            for (int kc = 0; kc < KS; ++kc) {
                if (!(DBG & 16) && P.mma_sleep_ns)
                    mbar_wait_backoff(bar_full + 8 * stage, phase, P.error_flag, P.mma_sleep_ns);
                else
                    mbar_wait_acc<DBG>(bar_full + 8 * stage, phase, P.error_flag, m_full);
                tc_fence_after();
                const uint32_t cnt = no_mma ? 0u : hdr.y;
                const uint32_t bs = b_const + sb16;
                // the hi plane accumulates into the same columns (weight slot + 1 of the master) or, in p2 mode, into
                // its own columns, where its first MMA of the tile must not accumulate either
                const uint32_t acc_hi0 = P.p2 ? accumulate : 1u;
#ifdef SRCDSP_TIMING_EXPERIMENTS
                if (SRCDSP_EXP(P, 256)) {
                    // timing experiment (wrong results): the MMA cost structure of the operand-swapped form -- samples on the
                    // M side (two halves of 128 rows), a band of N = (debug >> 12) tap rows on the N side, per lag and plane
                    const int nn = (P.debug >> 12) ? (P.debug >> 12) : 56;
                    const uint32_t it_lo = umma_idesc_i8(0, 1, 128, nn), it_hi = umma_idesc_i8(1, 1, 128, nn);
                    if (elect_one()) {
                        const uint32_t cnt_t = SRCDSP_EXP(P, 512) ? (cnt ? 1u : 0u) : cnt;  // 512: lags merged into one band
                        for (uint32_t i = 0; i < cnt_t; ++i) {
                            const uint4 e = plan_ent[hdr.x + i];
                            for (uint32_t half = 0; half < 2; ++half) {
                                umma_i8(d_tmem + half * 128 + 8 * i, desc(bs + e.z + half * 128), desc(a_const + e.x), it_lo, 1);
                                umma_i8(d_tmem + half * 128 + 8 * i, desc(bs + e.w + half * 128), desc(a_const + e.y), it_hi, 1);
                            }
                        }
                        tc_commit(bar_empty + 8 * stage);
                        if (kc == KS - 1) tc_commit(bar_tfull + 8 * acc);
                    }
                } else
#endif
                if (elect_one()) {
                    if (cnt > 0) {
                        umma_i8(d_tmem, desc(a_const + e0.x), desc(bs + e0.z), idesc_lo, accumulate);
                        if (!no_hi) umma_i8(d_tmem + hi_cols, desc(a_const + e0.y), desc(bs + e0.w), idesc_hi, acc_hi0);
                    }
                    if (cnt > 1) {
                        umma_i8(d_tmem, desc(a_const + e1.x), desc(bs + e1.z), idesc_lo, 1);
                        if (!no_hi) umma_i8(d_tmem + hi_cols, desc(a_const + e1.y), desc(bs + e1.w), idesc_hi, 1);
                    }
                    for (uint32_t i = 2; i < cnt; ++i) {
                        const uint4 e = plan_ent[hdr.x + i];
                        umma_i8(d_tmem, desc(a_const + e.x), desc(bs + e.z), idesc_lo, 1);
                        if (!no_hi) umma_i8(d_tmem + hi_cols, desc(a_const + e.y), desc(bs + e.w), idesc_hi, 1);
                    }
                    tc_commit(bar_empty + 8 * stage);  // frees the stage when these MMAs have read it
                    if (kc == KS - 1) tc_commit(bar_tfull + 8 * acc);
                    if (R.progress) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(R.progress), "r"(++steps_done) : "memory");
                }
                accumulate |= (cnt != 0);
                __syncwarp();
                // next K-step's plan: in flight while the tensor pipe works and the next barrier is awaited
                hdr = plan_hdr[kc + 1 == KS ? 0 : kc + 1];
                e0 = plan_ent[hdr.x];
                e1 = plan_ent[hdr.x + 1];
                sb16 += stage16;
                if (++stage == NS) {
                    stage = 0;
                    sb16 = 0;
                    phase ^= 1;
                }
            }
            acc_phases ^= 1u << acc;
            acc ^= 1;
        }
        if ((DBG & 16) && lane == 0) {
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(P.counters);
            atomicAdd(cnt + 3, (unsigned long long)(clock64() - m_t0));  // MMA warp total
            atomicAdd(cnt + 4, (unsigned long long)m_full);               // MMA waiting for a full stage
            atomicAdd(cnt + 5, (unsigned long long)m_tempty);             // MMA waiting for a free accumulator
        }
    }
}

// epilogue warps 0..3: TMEM lanes 32*warp .. 32*warp+31
template <int DBG>
__device__ __forceinline__ void tc_epilogue_role(const TcParams &P, const TcRole &R)
{
    uint8_t *const a_smem = R.a_smem, *const stages = R.stages;
    const int stage_bytes = R.stage_bytes, NS = R.NS, J = R.J, KS = R.KS, warp = R.warp, lane = R.lane;
    const uint32_t bar_full = R.bar_full, bar_empty = R.bar_empty, bar_tfull = R.bar_tfull, bar_tempty = R.bar_tempty;
    const uint32_t tmem_base = R.tmem_base;
    const long long first_tile = R.first_tile, tile_step = R.tile_step;
    (void)a_smem, (void)stages, (void)stage_bytes, (void)NS, (void)J, (void)KS, (void)warp, (void)lane;
    (void)bar_full, (void)bar_empty, (void)bar_tfull, (void)bar_tempty, (void)tmem_base;
    const int w = lane & 3;
    const int b = 8 * warp + (lane >> 2);
    long long e_wait = 0;
    const long long e_t0 = clock64();
    uint32_t acc_phases = 0;
    int acc = 0;
    for (long long tile = first_tile; tile < R.tile_end; tile += tile_step) {
        const unsigned ch = (unsigned)tile / (unsigned)P.tiles_per_ch;  // total_tiles < 2^31 (host check)
        const long long tt = (long long)((unsigned)tile - ch * (unsigned)P.tiles_per_ch);
        uint32_t *o = P.out + (size_t)ch * P.out_stride;
        if ((DBG & 16) || P.epi_sleep_ns == 0) {
            mbar_wait_acc<DBG>(bar_tfull + 8 * acc, (acc_phases >> acc) & 1, P.error_flag, e_wait);
        } else {
            mbar_wait_backoff(bar_tfull + 8 * acc, (acc_phases >> acc) & 1, P.error_flag, P.epi_sleep_ns);
        }
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(32 * warp) << 16) + acc * (2 * TC_NRB);
        if (P.p2) {
            // p2 mode: lane = 16*(b>>3) + 8*p + (b&7) for output b (64 per row-block) and digit slot p; columns 0..127
            // hold the lo byte plane's sums, 128..255 the hi plane's.  A 16x256b load of lanes 16h..16h+15 hands one
            // thread both digit slots of output b = 16*warp + 8*h + lane/4 for the (re, im) pairs of 4 row-blocks.
            for (int h = 0; h < 2; ++h) {
                const int b2 = 16 * warp + 8 * h + (lane >> 2);
                const long long idx_lane = (tt * P.nrb + (lane & 3)) * 64 + b2;
                uint32_t *o_lane = o + idx_lane;
                const bool full_tile = (tt + 1) * (long long)(TC_NRB * TC_BOUT) <= P.n_out;
#pragma unroll 1
                for (int c0 = 0; c0 < (SRCDSP_EXP(P, 4) ? 0 : 2 * P.nrb); c0 += 32) {
                    uint32_t a[16], hh[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
                        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                          "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]),
                          "=r"(a[15])
                        : "r"(t_addr + ((uint32_t)(16 * h) << 16) + c0));
                    asm volatile(
                        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
                        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(hh[0]), "=r"(hh[1]), "=r"(hh[2]), "=r"(hh[3]), "=r"(hh[4]), "=r"(hh[5]), "=r"(hh[6]), "=r"(hh[7]),
                          "=r"(hh[8]), "=r"(hh[9]), "=r"(hh[10]), "=r"(hh[11]), "=r"(hh[12]), "=r"(hh[13]), "=r"(hh[14]),
                          "=r"(hh[15])
                        : "r"(t_addr + ((uint32_t)(16 * h) << 16) + 2 * P.nrb + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint32_t *op = o_lane + c0 * 32;  // row-block c0 / 2, 64 outputs each
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // lo*d0 + 256 * (lo*d1 + hi*d0) + 65536 * hi*d1 (mod 2^32, exactly the reference's int32 wrap)
                        const uint32_t re = a[4 * k] + ((a[4 * k + 2] + hh[4 * k]) << 8) + (hh[4 * k + 2] << 16);
                        const uint32_t im = a[4 * k + 1] + ((a[4 * k + 3] + hh[4 * k + 1]) << 8) + (hh[4 * k + 3] << 16);
                        const uint32_t word = scale_pack_sym_sat((int)re, (int)im, P.shift);
                        if (full_tile || idx_lane + (c0 / 2 + 4 * k) * 64 < P.n_out) op[4 * k * 64] = word;
                    }
                }
            }
        } else if (P.grouped) {
            // weight slots of output b sit 8 TMEM lanes apart: a 16x256b load hands one thread the
            // slots {0,1} (lanes 0-15 of the warp's quadrant) or {2,3} (lanes 16-31) of output
            // b = 8*warp + lane/4 for the column pairs (re, im) of 4 row-blocks -> no shuffles
            const long long idx_lane = (tt * TC_NRB + (lane & 3)) * TC_BOUT + b;  // output index for c0 = 0, k = 0
            uint32_t *o_lane = o + idx_lane;
            const bool full_tile = (tt + 1) * (long long)(TC_NRB * TC_BOUT) <= P.n_out;
#pragma unroll 1
            for (int c0 = 0; c0 < (SRCDSP_EXP(P, 4) ? 0 : 2 * TC_NRB); c0 += 32) {
                uint32_t a[16], h[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                      "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]),
                      "=r"(a[15])
                    : "r"(t_addr + c0));
                asm volatile(
                    "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]), "=r"(h[4]), "=r"(h[5]), "=r"(h[6]), "=r"(h[7]),
                      "=r"(h[8]), "=r"(h[9]), "=r"(h[10]), "=r"(h[11]), "=r"(h[12]), "=r"(h[13]), "=r"(h[14]),
                      "=r"(h[15])
                    : "r"(t_addr + (16u << 16) + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // this lane's outputs of the chunk: row-blocks m = c0/2 + 4*k + (lane & 3), output b
                uint32_t *op = o_lane + c0 * (TC_BOUT / 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // sum_w 256^w * D_w (mod 2^32, exactly the reference's int32 wrap)
                    const uint32_t re = a[4 * k] + (a[4 * k + 2] << 8) + (h[4 * k] << 16) + (h[4 * k + 2] << 24);
                    const uint32_t im = a[4 * k + 1] + (a[4 * k + 3] << 8) + (h[4 * k + 1] << 16) + (h[4 * k + 3] << 24);
                    const uint32_t word = scale_pack_sym_sat((int)re, (int)im, P.shift);
                    if (full_tile || idx_lane + (c0 / 2 + 4 * k) * TC_BOUT < P.n_out) op[4 * k * TC_BOUT] = word;
                }
            }
        } else {
        const bool full_tile = (tt + 1) * (long long)(TC_NRB * TC_BOUT) <= P.n_out;
#pragma unroll 1
        for (int c0 = 0; c0 < (SRCDSP_EXP(P, 4) ? 0 : 2 * TC_NRB); c0 += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                  "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                  "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                  "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(t_addr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // sum_w 256^w * D_w over the 4 lanes of the quad (mod 2^32, exactly the int32 wrap), as a transposing
            // reduction: every lane shifts its slot's sums into place, then two exchange rounds halve the columns a
            // lane keeps, so that lane w ends with the complete sums of the row-blocks m = c0/2 + 4*i + w -- 24
            // shuffles per 32 columns instead of 64, and no register selection afterwards
            const int sh8 = 8 * w;
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] <<= sh8;
            const bool hi2 = (w & 2) != 0, hi1 = (w & 1) != 0;
            uint32_t k1[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {  // lanes with w & 2 keep the row-blocks 4*i + 2, 4*i + 3 (columns 8*i + 4 ..)
                    const uint32_t keep = hi2 ? v[8 * i + 4 + c] : v[8 * i + c];
                    const uint32_t send = hi2 ? v[8 * i + c] : v[8 * i + 4 + c];
                    k1[4 * i + c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
            }
            const long long idx0 = (tt * TC_NRB + (c0 >> 1) + w) * TC_BOUT + b;  // row-block c0/2 + w, output b
            uint32_t *op = o + idx0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // lanes with w & 1 keep the odd row-block of the pair
                const uint32_t kr = hi1 ? k1[4 * i + 2] : k1[4 * i], sr = hi1 ? k1[4 * i] : k1[4 * i + 2];
                const uint32_t ki = hi1 ? k1[4 * i + 3] : k1[4 * i + 1], si = hi1 ? k1[4 * i + 1] : k1[4 * i + 3];
                const uint32_t re = kr + __shfl_xor_sync(0xffffffffu, sr, 1);
                const uint32_t im = ki + __shfl_xor_sync(0xffffffffu, si, 1);
                if (full_tile || idx0 + 4 * i * TC_BOUT < P.n_out) op[4 * i * TC_BOUT] = scale_pack_sym_sat((int)re, (int)im, P.shift);
            }
        }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        acc_phases ^= 1u << acc;
        acc ^= 1;
    }
    if ((DBG & 16) && lane == 0) {
        unsigned long long *cnt = reinterpret_cast<unsigned long long *>(P.counters);
        atomicAdd(cnt + 6, (unsigned long long)(clock64() - e_t0));  // epilogue warp total
        atomicAdd(cnt + 7, (unsigned long long)e_wait);               // epilogue waiting for accumulators
    }
}

template <int DBG, bool MIX>
__global__ void __launch_bounds__(TC_THREADS, 1) dec_tc_kernel(const __grid_constant__ TcParams P, const __grid_constant__ CUtensorMap in_map)
{
    extern __shared__ __align__(128) uint8_t tc_smem_raw[];
    uint8_t *smem = tc_smem_raw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int J = P.J;
    const int stage_bytes = 4 * P.rbp * 16;  // 2 planes x 2 kc chunks
    uint8_t *a_smem = smem;
    uint32_t *tab_smem = reinterpret_cast<uint32_t *>(smem + ((P.master_bytes + 127) & ~127));
    uint8_t *stages = reinterpret_cast<uint8_t *>(tab_smem) + P.table_bytes;
    const int NS = P.n_stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stages + NS * stage_bytes);
    // bars: full[NS], empty[NS], tmem_full[2], tmem_empty[2]
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * TC_MAX_STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * TC_MAX_STAGES, bar_tempty = bar_tfull + 16;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t progress_s;
    if (tid == 0) progress_s = 0;

    // ---- setup ------------------------------------------------------------------------------
    for (int i = tid; i < P.master_bytes / 16; i += TC_THREADS)
        reinterpret_cast<uint4 *>(a_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(P.master) + i);
    if (MIX)
        for (int i = tid; i < P.table_bytes / 4; i += TC_THREADS) tab_smem[i] = __ldg(P.cs_table + i);
    // rows of the stages that no producer ever writes (the padding row) must be defined: zero all
    for (int i = tid; i < NS * stage_bytes / 16; i += TC_THREADS)
        reinterpret_cast<uint4 *>(stages)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + 8 * s, TC_SPLIT);  // the producer warps sharing the K-step
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // a CTA walks a contiguous run of tiles (consecutive tiles of a channel: the halo rows of a tile are the
    // previous tile's last rows, still in L2; the fused mixer's per-channel state changes rarely)
    const long long first_tile = P.total_tiles * blockIdx.x / gridDim.x, tile_end = P.total_tiles * (blockIdx.x + 1) / gridDim.x;
    const long long tile_step = 1;
    const int KS = P.ksteps;  // K-steps per tile
    const TcRole role{a_smem, stages, stage_bytes, NS, J, KS, warp, lane, bar_full, bar_empty, bar_tfull, bar_tempty, tmem_base,
                      first_tile, tile_step, tile_end, P.pf_dist > 0 ? smem_u32(&progress_s) : 0u};

    if (warp == TC_PF_WARP) {
        // =====================================================================================
        // L2 prefetcher: the producers' register-staged loads cannot keep enough bytes in flight to
        // cover HBM latency (tools/ldbench.cu); one TMA prefetch per K-step -- the same box the
        // producers are going to read, pf_dist K-steps ahead of the MMAs -- turns their loads into
        // L2 hits.  No shared memory, no barrier: pacing follows the MMA warp's progress word.
        // =====================================================================================
        if (lane == 0 && P.pf_dist > 0) {
            const uint32_t prog = smem_u32(&progress_s);
            uint32_t gs = 0;
            for (long long tile = first_tile; tile < tile_end; tile += tile_step) {
                const unsigned tl = (unsigned)tile;
                const unsigned ch = tl / (unsigned)P.tiles_per_ch;
                const unsigned tt = tl - ch * (unsigned)P.tiles_per_ch;
                const int row0 = (int)tt * TC_NRB - (J - 1);
                for (int kc = 0; kc < KS; ++kc, ++gs) {
                    for (;;) {
                        uint32_t done;
                        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(prog) : "memory");
                        if ((int)(gs - done) <= P.pf_dist) break;
                        __nanosleep(100);
                    }
                    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&in_map), "r"(32 * kc),
                                 "r"(row0), "r"((int)ch)
                                 : "memory");
                }
            }
        }
    } else if (warp >= TC_PROD_WARP0) {
        // =====================================================================================
        // producers.  Owner o = pw / TC_SPLIT stages K-steps o, o + TC_OWNERS, ... of the flattened
        // (tile, K-step) sequence (step gs -> stage gs % NS); its TC_SPLIT warps take half of the
        // 128 row-blocks each (warp 0 of the owner also the J-1 halo row-blocks in front of the
        // tile).  A warp has few loads in flight (hardware limit, tools/ldbench.cu), so HBM is
        // covered by the NUMBER of producer warps; the code per warp is a plain load -> split ->
        // store loop with everything non-trivial hoisted to once per K-step.
        // =====================================================================================
        const int pw = warp - TC_PROD_WARP0;
        const int owner = pw / TC_SPLIT, part = pw % TC_SPLIT;
        long long w_wait = 0, w_fence = 0;
        const long long w_t0 = clock64();
        const int piece = lane & 7, grp = lane >> 3;
        const int halo_it = (part == 0) ? (J - 1 + 3) >> 2 : 0;  // iterations of the halo batch of this warp
        constexpr int ROWS_PER_WARP = TC_NRB / TC_SPLIT;
        constexpr int MAIN_BATCHES = ROWS_PER_WARP / (4 * TC_BATCH);
        const int chunk = P.rbp * 16;
        const size_t it_stride = (size_t)4 * P.rb_stride;  // words between iterations of a batch
        const int row_first = part * ROWS_PER_WARP;        // first main row-block of this warp

        long long tile = first_tile;
        int kc = owner;
        while (kc >= KS && tile < tile_end) {
            kc -= KS;
            tile += tile_step;
        }
        long long cur_tile = -1;
        const uint32_t *x = nullptr, *hist = nullptr;
        long long tile0 = 0;
        unsigned ph0 = 0, fr = 0, dph = 0;  // MIX: channel phase at sample 0, frequency, phase step per iteration
        int stage = owner % NS;  // step gs uses stage gs % NS
        uint32_t parity = 1;     // first wait on a fresh "empty" barrier passes
        while (tile < tile_end) {
            if (tile != cur_tile) {  // one division per tile
                const unsigned tl = (unsigned)tile;
                const unsigned ch = tl / (unsigned)P.tiles_per_ch;
                const unsigned tt = tl - ch * (unsigned)P.tiles_per_ch;
                x = P.in + (size_t)ch * P.in_stride;
                hist = P.hist_in + (size_t)ch * P.H;
                tile0 = (long long)tt * TC_NRB * (long long)P.G;
                cur_tile = tile;
                if (MIX) {
                    ph0 = (unsigned)P.phi[ch];
                    fr = (unsigned)P.freq[ch];
                    dph = (((unsigned)(4 * P.rb_stride) & P.mix_mask) * fr) & P.mix_mask;
                }
            }
            const long long col = tile0 + (long long)P.kc_stride * kc + 4 * piece;  // sample of (row 0, piece)
            const bool fast_main = P.vec_in && col - 4 * piece + (long long)(TC_NRB - 1) * P.rb_stride + 32 <= P.n_in;
            const bool fast_halo = P.vec_in && tile0 > 0;  // the previous tile of the same block exists

            // (plane lo, this lane's kc chunk, re row of row-block rb, this lane's word) for rb = grp;
            // rows are shifted by front_pad so that the unused lanes of a prefetched halo batch
            // (row-blocks below -(J-1)) land in padding instead of out of bounds
            uint8_t *dst0 = stages + stage * stage_bytes + (piece >> 2) * chunk +
                            (P.front_pad + 2 * (grp + (J - 1))) * 16 + (piece & 3) * 4;
            // This warp's batches of the K-step: b = -1 (the halo rows, first warp of the owner only),
            // 0 .. MAIN_BATCHES-1.  Two register slots: the loads of batch b + 1 are in flight while
            // batch b is mixed / split / stored, and the first loads are issued before the wait for the
            // stage, so only one load latency per K-step is exposed.
            auto issue = [&](TcBatch &t, int b) {
                const int rb = b < 0 ? -4 * halo_it + grp : row_first + 4 * TC_BATCH * b + grp;
                t.fast = b < 0 ? fast_halo : fast_main;
                if (t.fast) tc_batch_load_fast<DBG>(t, x + (col + (long long)rb * P.rb_stride), it_stride, b < 0 ? halo_it : TC_BATCH);
            };
            auto finish = [&](const TcBatch &t, int b) {
                const int rb = b < 0 ? -4 * halo_it + grp : row_first + 4 * TC_BATCH * b + grp;
                const int nit = b < 0 ? halo_it : TC_BATCH;
                uint8_t *dst = dst0 + (rb - grp) * 32;
                const long long n0 = col + (long long)rb * P.rb_stride;
                if (t.fast)
                    tc_batch_store_fast<DBG, MIX>(t, dst, dst + 2 * chunk, nit, tab_smem,
                                                  (ph0 + ((unsigned)n0 & P.mix_mask) * fr) & P.mix_mask, fr, dph, P.mix_mask);
                else
                    tc_batch_generic(P, x, hist, n0, rb, b < 0 ? -(J - 1) : 0, b < 0 ? 0 : TC_NRB, nit, dst, MIX, ph0, fr);
            };
            TcBatch slot[2];
            const int b_first = halo_it > 0 ? -1 : 0;
            issue(slot[0], b_first);
            mbar_wait_acc<DBG>(bar_empty + 8 * stage, parity, P.error_flag, w_wait);
#pragma unroll
            for (int i = 0; i < 1 + MAIN_BATCHES; ++i) {
                const int b = b_first + i;
                if (b < MAIN_BATCHES) {
                    if (b + 1 < MAIN_BATCHES) issue(slot[(i + 1) & 1], b + 1);
                    finish(slot[i & 1], b);
                }
            }
            // every writer fences its generic-proxy stores towards the async proxy (the MMA reads
            // shared memory through it); one arrival per warp
            if (DBG & 16) {
                const long long f0 = clock64();
                fence_async_smem();
                w_fence += clock64() - f0;
            } else {
                fence_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * stage);
            stage += TC_OWNERS;
            while (stage >= NS) {
                stage -= NS;
                parity ^= 1;
            }
            kc += TC_OWNERS;
            while (kc >= KS) {
                kc -= KS;
                tile += tile_step;
            }
        }
        if ((DBG & 16) && lane == 0) {
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(P.counters);
            atomicAdd(cnt + 0, (unsigned long long)(clock64() - w_t0));  // producer total
            atomicAdd(cnt + 1, (unsigned long long)w_wait);               // producer waiting for an empty stage
            atomicAdd(cnt + 2, (unsigned long long)w_fence);              // producer in fence.proxy.async
        }
    } else if (warp == 4) {
        // =====================================================================================
        // MMA issuer (one thread)
        // =====================================================================================
        tc_mma_role<DBG>(P, role);
    } else {
        // =====================================================================================
        // epilogue warps 0..3: TMEM lanes 32*warp .. 32*warp+31  ->  rho = 4*b + w
        // =====================================================================================
        tc_epilogue_role<DBG>(P, role);
    }

    // ---- teardown ---------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace srcdsp
