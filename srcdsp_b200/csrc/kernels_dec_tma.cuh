// kernels_dec_tma.cuh -- the tcgen05 int8 Toeplitz decimator (kernels_dec_tc.cuh) fed by TMA.
//
// Same GEMM, same resident tap "master", same epilogue as dec_tc_kernel; what changes is how the
// samples reach the byte-plane stages.  dec_tc_kernel's producers load global memory into
// registers, so HBM latency is covered only by the number of producer warps (6.1 TB/s ceiling for
// the 128-bytes-per-2-KB K-step pattern, tools/ldbench.cu).  Here one thread issues a 3-D TMA box
// per K-step -- (32 samples) x (J-1 halo + 128 row-blocks) x (1 channel) of the raw interleaved
// int16 stream -- into a ring of RAW stages (the bytes in flight no longer cost registers or
// warps: 7.4 TB/s with the same pattern, tools/tmabench.cu), and converter warps turn a raw
// stage into a byte-plane stage: LDS.128 -> (NCO mix) -> 8 PRMT -> 4 STS.32, all on-chip.
//
//   warps 0-3   epilogue            (tc_epilogue_role)
//   warp  4     MMA issuer + TMEM   (tc_mma_role)
//   warp  5     TMA issuer (one lane)
//   warps 6..   converters in groups of W warps: a group takes every (n_conv / W)-th K-step, each of
//               its warps every W-th group of 4 rows of it
//
// Edges (rare): rows in front of the block come from the carried history and the ragged last
// row-block is not part of the TMA tensor; the TMA box zero-fills both and the converters patch
// those rows straight from global memory (tc_sample, the same code dec_tc_kernel uses).
#pragma once

#include <cuda.h>

#include "kernels_dec_tc.cuh"

namespace srcdsp {

constexpr int TMA_CONV_WARP0 = 6;
constexpr int TMA_MAX_CONV = 16;
constexpr int TMA_MAX_THREADS = 32 * (TMA_CONV_WARP0 + TMA_MAX_CONV);
constexpr int TMA_MAX_RAW = 12;
constexpr int TMA_MAX_CONV_MIX = 12;  // fused mixer: fewer warps, more registers each (the mix keeps many values live)

struct TmaExtra {
    int n_raw;        // raw stages
    int n_conv;       // converter warps = W * groups (W = the kernel's template parameter; groups <= min(n_raw, n_stages - 1))
    int shared_raw;   // n_raw % groups != 0: a raw stage is converted by different groups in turn
    int raw_rows;     // rows of a raw stage: 4 * ceil((J-1)/4) + 128 (the box lands at row raw_rows - box_rows)
    int box_rows;     // J - 1 + 128
    long long rows_full;  // floor(n_in / G): row-blocks that are part of the TMA tensor
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// shared-window (32-bit) accesses with compile-time offsets: the converters' hot loop is nothing but
// these, and generic pointers cost a window-base recomputation per access
template <int OFF>
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(OFF) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// (mix4_planes_dp2a / mix_digits: common.cuh)
// One group of 4 rows: the warp's 32 lanes hold, per byte plane, 8 rows x 16 bytes of the sample operand -- lane
// (grp, piece) has word (piece & 3) of row (grp, piece >> 2).  That is exactly the fragment layout of an 8 x 8 b16 matrix
// (thread t: row t / 4, columns 2 (t % 4) and 2 (t % 4) + 1), so the four plane words of a piece leave as ONE
// stmatrix.x4 instead of four STS.32: matrix 0 = re lo, 1 = im lo, 2 = re hi, 3 = im hi, and lane L supplies the
// shared address of row L % 8 of matrix L / 8 (tma_stm_lane).  Same bytes, same 4 conflict-free wavefronts, a
// quarter of the store instructions.
__device__ __forceinline__ int tma_stm_lane(int lane, int chunk)
{
    const int m = lane >> 3, j = lane & 7;  // matrix, row: (grp, kc chunk) = (j >> 1, j & 1)
    return (m >> 1) * 2 * chunk + (m & 1) * 16 + (j & 1) * chunk + (j >> 1) * 32;
}
template <int DST_OFF>
__device__ __forceinline__ void tma_store_planes(uint32_t re_lo, uint32_t re_hi, uint32_t im_lo, uint32_t im_hi, uint32_t dst)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0+%5], {%1, %2, %3, %4};" ::"r"(dst), "r"(re_lo), "r"(im_lo), "r"(re_hi),
                 "r"(im_hi), "n"(DST_OFF)
                 : "memory");
}
template <int DST_OFF>
__device__ __forceinline__ void tma_split_store(const uint4 q, uint32_t dst)
{
    uint32_t re_lo, re_hi, im_lo, im_hi;
    split4(q, re_lo, re_hi, im_lo, im_hi);
    tma_store_planes<DST_OFF>(re_lo, re_hi, im_lo, im_hi, dst);
}
// The oscillator values of the 4 samples of a piece: Bre / Bim digit words (mix_digits).  They come from the
// channel's oscillator sequence in TIME order in shared memory (two arrays Bre[n], Bim[n] for phase
// (phi0 + n * freq) mod N; the sequence has period N): a piece takes two conflict-free LDS.128 instead of 4
// gathers from the sine table, whose stride (the channel frequency) makes the lanes collide on a few banks.
struct MixPiece {
    uint4 re, im;
};
// lo = shared address of the sequence, idx4 = (n mod N) * 4 of the piece's first sample, im_off = bytes from Bre to Bim
__device__ __forceinline__ MixPiece tma_mix_piece(uint32_t lo, unsigned idx4, unsigned im_off)
{
    MixPiece m;
    m.re = lds128<0>(lo + idx4);
    m.im = lds128<0>(lo + im_off + idx4);
    return m;
}
template <int DST_OFF>
__device__ __forceinline__ void tma_mix_store(const uint4 q, const MixPiece &m, uint32_t dst)
{
    uint32_t re_lo, re_hi, im_lo, im_hi;
    mix4_planes_dp2a(q, m.re, m.im, re_lo, re_hi, im_lo, im_hi);
    tma_store_planes<DST_OFF>(re_lo, re_hi, im_lo, im_hi, dst);
}

// 4 row groups W apart (this warp's share of 4 * W groups): all loads first, then mix / split / store
// SRC_GS: bytes between consecutive groups of 4 raw rows (512 for rows of 32 samples, 1024 for dec_band_kernel's rows of 64)
template <bool MIX, int W, int SRC_GS = 512>
__device__ __forceinline__ void tma_convert4(uint32_t src, uint32_t dst, uint32_t lo, unsigned idx4, unsigned didx4, unsigned mask4)
{
    const uint4 v0 = lds128<0>(src), v1 = lds128<W * SRC_GS>(src), v2 = lds128<2 * W * SRC_GS>(src), v3 = lds128<3 * W * SRC_GS>(src);
    if (MIX) {
        tma_mix_store<0>(v0, tma_mix_piece(lo, idx4, mask4 + 4), dst);
        tma_mix_store<W * 128>(v1, tma_mix_piece(lo, (idx4 + didx4) & mask4, mask4 + 4), dst);
        tma_mix_store<2 * W * 128>(v2, tma_mix_piece(lo, (idx4 + 2 * didx4) & mask4, mask4 + 4), dst);
        tma_mix_store<3 * W * 128>(v3, tma_mix_piece(lo, (idx4 + 3 * didx4) & mask4, mask4 + 4), dst);
    } else {
        tma_split_store<0>(v0, dst);
        tma_split_store<W * 128>(v1, dst);
        tma_split_store<2 * W * 128>(v2, dst);
        tma_split_store<3 * W * 128>(v3, dst);
    }
}

// Same with ONE oscillator piece for all 4 row groups: when the distance between a warp's row groups
// (4 * W * G samples) is a multiple of the sequence's period -- the usual case, e.g. N = 4096 with W * M a
// multiple of 32 -- the groups see the same oscillator values, which are then fetched once per K-step
// instead of once per piece.
template <int W, int SRC_GS = 512>
__device__ __forceinline__ void tma_convert4_same(uint32_t src, uint32_t dst, const MixPiece &m)
{
    const uint4 v0 = lds128<0>(src), v1 = lds128<W * SRC_GS>(src), v2 = lds128<2 * W * SRC_GS>(src), v3 = lds128<3 * W * SRC_GS>(src);
    tma_mix_store<0>(v0, m, dst);
    tma_mix_store<W * 128>(v1, m, dst);
    tma_mix_store<2 * W * 128>(v2, m, dst);
    tma_mix_store<3 * W * 128>(v3, m, dst);
}

// generic-pointer variant for the edge path (tab = the oscillator sequence Bre[N] ++ Bim[N], p0 = n mod N)
template <bool MIX>
__device__ __forceinline__ void tma_convert_store(uint4 q, uint8_t *dst, int hi_off, const uint32_t *tab, unsigned p0, unsigned mask)
{
    if (MIX) {  // mixers.h:172-177 on the 4 samples of this piece
        const uint32_t *im = tab + mask + 1;
        q.x = mix_sample_dp2a(q.x, tab[p0], im[p0]);
        q.y = mix_sample_dp2a(q.y, tab[(p0 + 1) & mask], im[(p0 + 1) & mask]);
        q.z = mix_sample_dp2a(q.z, tab[(p0 + 2) & mask], im[(p0 + 2) & mask]);
        q.w = mix_sample_dp2a(q.w, tab[(p0 + 3) & mask], im[(p0 + 3) & mask]);
    }
    uint32_t re_lo, re_hi, im_lo, im_hi;
    split4(q, re_lo, re_hi, im_lo, im_hi);
    *reinterpret_cast<uint32_t *>(dst) = re_lo;
    *reinterpret_cast<uint32_t *>(dst + 16) = im_lo;
    *reinterpret_cast<uint32_t *>(dst + hi_off) = re_hi;
    *reinterpret_cast<uint32_t *>(dst + hi_off + 16) = im_hi;
}

// W = converter warps per group (4 or 8): each takes 32 / W of the 32 main row groups of a K-step
template <int DBG, bool MIX, int W>
__global__ void __launch_bounds__(MIX ? 32 * (TMA_CONV_WARP0 + TMA_MAX_CONV_MIX) : TMA_MAX_THREADS, 1)
    dec_tma_kernel(const __grid_constant__ TcParams P, const __grid_constant__ TmaExtra X, const __grid_constant__ CUtensorMap in_map)
{
    extern __shared__ __align__(128) uint8_t tc_smem_raw[];
    uint8_t *smem = tc_smem_raw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int J = P.J;
    const int stage_bytes = 4 * P.rbp * 16;  // byte-plane stage: 2 planes x 2 kc chunks
    const int raw_bytes = X.raw_rows * 128;
    const int NS = P.n_stages, NR = X.n_raw, NCW = X.n_conv;
    uint8_t *a_smem = smem;
    uint32_t *tab_smem = reinterpret_cast<uint32_t *>(smem + ((P.master_bytes + 127) & ~127));
    uint8_t *raw = reinterpret_cast<uint8_t *>(tab_smem) + P.table_bytes;
    uint8_t *stages = raw + NR * raw_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stages + NS * stage_bytes);
    // bars: full[TC_MAX_STAGES], empty[TC_MAX_STAGES], tmem_full[2], tmem_empty[2], raw_full[TMA_MAX_RAW], raw_empty[TMA_MAX_RAW]
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * TC_MAX_STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * TC_MAX_STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t bar_rfull = bar_tempty + 16, bar_rempty = bar_rfull + 8 * TMA_MAX_RAW;
    __shared__ uint32_t tmem_base_s;

    // ---- setup ------------------------------------------------------------------------------
    for (int i = tid; i < P.master_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(a_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(P.master) + i);
    // (MIX: tab_smem holds the current channel's oscillator sequence; the converters build it)
    // padding rows of both rings must be defined: zero everything once
    for (int i = tid; i < (NR * raw_bytes + NS * stage_bytes) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(raw)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + 8 * s, W);  // one arrival per converter warp of the K-step's group
            mbar_init(bar_empty + 8 * s, 1);   // tcgen05.commit
        }
        for (int s = 0; s < NR; ++s) {
            mbar_init(bar_rfull + 8 * s, 1);   // expect_tx arrival + the box's bytes
            mbar_init(bar_rempty + 8 * s, W);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
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
    // previous tile's last rows, still in L2; the fused mixer's per-channel sequence changes rarely)
    const long long first_tile = P.total_tiles * blockIdx.x / gridDim.x, tile_end = P.total_tiles * (blockIdx.x + 1) / gridDim.x;
    const long long tile_step = 1;
    const int KS = P.ksteps;
    const TcRole role{a_smem, stages, stage_bytes, NS, J, KS, warp, lane, bar_full, bar_empty, bar_tfull, bar_tempty, tmem_base,
                      first_tile, tile_step, tile_end, 0u};

    if (warp >= TMA_CONV_WARP0) {
        // =====================================================================================
        // converters: raw stage (rows of 32 interleaved samples) -> byte-plane stage
        // =====================================================================================
        // The warps form NG = n_conv / W groups; group g converts K-steps g, g + NG, ... of the CTA's
        // flattened (tile, K-step) sequence, its W warps share the row groups of such a K-step.  A warp's
        // per-K-step overhead (two barrier waits, proxy fence, two arrivals) is then paid once per
        // NG K-steps, and the groups' K-steps overlap in time.
        const int cw = warp - TMA_CONV_WARP0;
        const int NG = NCW / W;
        const int g = cw / W, wi = cw - g * W;
        const int piece = lane & 7, grp = lane >> 3;
        const int chunk = P.rbp * 16;
        const int halo_rows4 = X.raw_rows - P.nrb;    // 4 * ceil((J-1)/4)
        const int HQ = halo_rows4 / 4;                // row groups in front of the tile (halo + padding rows)
        const int NQ = X.raw_rows / 4;                // row groups per stage
        // raw row i (0 .. raw_rows) holds row-block rb = i - halo_rows4; its byte-plane rows are
        // front_pad + 2 * (rb + J - 1) = 2 * i (+1 for im)  [front_pad = 2 * (halo_rows4 - (J - 1))]
        const int src_lane = grp * 128 + piece * 16;
        const int dst_lane = (piece >> 2) * chunk + grp * 32 + (piece & 3) * 4;
        const int hi_off = 2 * chunk;
        // main row groups of this warp: HQ + wi + W * i, i < 32 / W; the HQ groups in front go to one warp
        // of the group in turn
        const uint32_t src_main = smem_u32(raw) + (HQ + wi) * 512 + src_lane;
        const uint32_t dst_main = smem_u32(stages) + (HQ + wi) * 128 + tma_stm_lane(lane, chunk);  // stmatrix row address of this lane
        const uint32_t tab_u32 = smem_u32(tab_smem);
        const unsigned mask4 = P.seq_mask << 2;
        const bool two_batches = P.nrb / (4 * W) == 8;  // 8 main row groups per warp (W = 4, 128 row-blocks); else 4
        int rs = g % NR, ss = g % NS;
        uint32_t rpar = 0, spar = 1;  // first wait on a fresh "empty" barrier passes
        int halo_turn = 0;
        long long w_wait_raw = 0, w_wait_split = 0, w_fence = 0, w_arrive = 0;
        const long long w_t0 = clock64();
        long long tile = first_tile, cur_tile = -1, tile0 = 0;
        int kc = g;
        while (kc >= KS && tile < tile_end) {
            kc -= KS;
            tile += tile_step;
        }
        const uint32_t *x = nullptr, *hist = nullptr;
        unsigned tt = 0, ph0 = 0, fr = 0, didx4 = 0, n_lane = 0;
        int cur_ch = -1;
        bool edge = false;
        while (tile < tile_end) {
            if (tile != cur_tile) {  // one division per tile
                const unsigned tl = (unsigned)tile;
                const unsigned ch = tl / (unsigned)P.tiles_per_ch;
                tt = tl - ch * (unsigned)P.tiles_per_ch;
                x = P.in + (size_t)ch * P.in_stride;
                hist = P.hist_in + (size_t)ch * P.H;
                tile0 = (long long)tt * P.nrb * (long long)P.G;  // sample of (row-block 0, K-step 0)
                // rows the TMA tensor does not hold: history in front of the block, everything from the ragged row-block on
                edge = (tt == 0 && J > 1) || ((long long)(tt + 1) * P.nrb > X.rows_full);
                if (MIX) {
                    if ((int)ch != cur_ch) {
                        // new channel: all converter warps rebuild the oscillator sequence lo[n] = T[(phi0 + n * freq) mod N]
                        // (rare: a CTA walks consecutive tiles).  Every group passes here once per tile, so the counts
                        // match; the first barrier waits until no warp reads the previous channel's sequence any more.
                        ph0 = (unsigned)P.phi[ch];
                        fr = (unsigned)P.freq[ch];
                        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32) : "memory");
                        for (unsigned i = cw * 32 + lane; i <= P.seq_mask; i += NCW * 32) {
                            uint32_t bre, bim;
                            mix_digits(__ldg(P.cs_table + ((ph0 + i * fr) & P.mix_mask)), bre, bim);
                            tab_smem[i] = bre;
                            tab_smem[P.seq_mask + 1 + i] = bim;
                        }
                        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32) : "memory");
                        cur_ch = (int)ch;
                    }
                    didx4 = ((unsigned)(4 * W * P.G) & P.seq_mask) << 2;  // between a warp's row groups
                    // sample index mod N of (row-block 4 * wi + grp, K-step 0, this lane's piece)
                    n_lane = (unsigned)(tile0 + (long long)(4 * wi + grp) * P.G + 4 * piece) & P.seq_mask;
                }
                cur_tile = tile;
            }
            // A raw stage shared by several groups (n_raw not a multiple of the group count): TMA boxes land out
            // of order, so another group's previous use of this stage may not even have landed yet, and the
            // parity wait below would then pass a whole phase early.  Waiting first until that use has been
            // released (the phase before ours of the "empty" barrier; passes at once on a fresh barrier) closes it.
            if (X.shared_raw) mbar_wait_acc<DBG>(bar_rempty + 8 * rs, rpar ^ 1, P.error_flag, w_wait_raw);
            if (!(DBG & 16) && P.conv_sleep_ns)
                mbar_wait_backoff(bar_rfull + 8 * rs, rpar, P.error_flag, P.conv_sleep_ns);
            else
                mbar_wait_acc<DBG>(bar_rfull + 8 * rs, rpar, P.error_flag, w_wait_raw);
            mbar_wait_acc<DBG>(bar_empty + 8 * ss, spar, P.error_flag, w_wait_split);
            if (SRCDSP_EXP(P, 8)) {
                // timing experiment: barriers only
            } else if (!edge) {
                const uint32_t src = src_main + rs * raw_bytes;
                const uint32_t dst = dst_main + ss * stage_bytes;
                const unsigned idx4 = MIX ? ((n_lane + 32 * kc) & P.seq_mask) << 2 : 0u;
                if (MIX && didx4 == 0) {
                    const MixPiece m = tma_mix_piece(tab_u32, idx4, mask4 + 4);
                    tma_convert4_same<W>(src, dst, m);
                    if (two_batches) tma_convert4_same<W>(src + 16 * 512, dst + 16 * 128, m);
                } else {
                    tma_convert4<MIX, W>(src, dst, tab_u32, idx4, didx4, mask4);
                    if (two_batches) tma_convert4<MIX, W>(src + 16 * 512, dst + 16 * 128, tab_u32, (idx4 + 4 * didx4) & mask4, didx4, mask4);
                }
                if (HQ > 0 && wi == halo_turn) {
                    // the row groups in front of the tile (the previous tile's last row-blocks): same shared-space
                    // code as the main groups -- this warp is the K-step's straggler
                    uint32_t hsrc = src - (HQ + wi) * 512, hdst = dst - (HQ + wi) * 128;
                    // (n mod N) * 4 of (raw row grp, this lane's piece): halo_rows4 row-blocks before row-block grp
                    unsigned hidx4 = 0;
                    if (MIX)
                        hidx4 = ((unsigned)(tile0 + (long long)(grp - halo_rows4) * P.G + 32 * kc + 4 * piece) & P.seq_mask) << 2;
                    const unsigned hstep4 = ((unsigned)(4 * P.G) & P.seq_mask) << 2;
                    for (int q = 0; q < HQ; ++q) {
                        const uint4 v = lds128<0>(hsrc);
                        if (MIX)
                            tma_mix_store<0>(v, tma_mix_piece(tab_u32, hidx4, mask4 + 4), hdst);
                        else
                            tma_split_store<0>(v, hdst);
                        hsrc += 512;
                        hdst += 128;
                        hidx4 = (hidx4 + hstep4) & mask4;
                    }
                }
            } else {
                const uint8_t *src = raw + rs * raw_bytes + src_lane;
                uint8_t *dst = stages + ss * stage_bytes + dst_lane;
                // sample index of (raw row grp, this lane's piece) of this K-step
                const long long n_row0 = tile0 + (long long)(grp - halo_rows4) * P.G + 32 * kc + 4 * piece;
                for (int q = wi; q < NQ; q += W) {
                    const int rb = 4 * q + grp - halo_rows4;
                    const long long n = n_row0 + (long long)4 * q * P.G;
                    uint4 v;
                    if (rb < -(J - 1)) {
                        v = make_uint4(0, 0, 0, 0);  // padding rows in front of the halo: never read by an MMA
                    } else if (n < 0 || (long long)(tt * P.nrb) + rb >= X.rows_full) {
                        // carried history (already mixed) / ragged end: straight from global memory
                        if (MIX) {
                            v.x = tc_sample_mix(P, x, hist, n, ph0, fr);
                            v.y = tc_sample_mix(P, x, hist, n + 1, ph0, fr);
                            v.z = tc_sample_mix(P, x, hist, n + 2, ph0, fr);
                            v.w = tc_sample_mix(P, x, hist, n + 3, ph0, fr);
                        } else {
                            v.x = tc_sample(x, hist, P.H, P.n_in, n);
                            v.y = tc_sample(x, hist, P.H, P.n_in, n + 1);
                            v.z = tc_sample(x, hist, P.H, P.n_in, n + 2);
                            v.w = tc_sample(x, hist, P.H, P.n_in, n + 3);
                        }
                        tma_convert_store<false>(v, dst + q * 128, hi_off, tab_smem, 0, 0);
                        continue;
                    } else {
                        v = *reinterpret_cast<const uint4 *>(src + q * 512);
                    }
                    tma_convert_store<MIX>(v, dst + q * 128, hi_off, tab_smem, (unsigned)n & P.seq_mask, P.seq_mask);
                }
            }
            // the MMA reads shared memory through the async proxy: fence this warp's stores, then
            // one arrival per warp on both rings
            if (DBG & 16) {
                const long long f0 = clock64();
                if (!SRCDSP_EXP(P, 128)) fence_async_smem();
                const long long f1 = clock64();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar_full + 8 * ss);
                    mbar_arrive(bar_rempty + 8 * rs);
                }
                w_fence += f1 - f0;
                w_arrive += clock64() - f1;
            } else {
                if (!SRCDSP_EXP(P, 128)) fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar_full + 8 * ss);
                    mbar_arrive(bar_rempty + 8 * rs);
                }
            }
            rs += NG;
            if (rs >= NR) {
                rs -= NR;
                rpar ^= 1;
            }
            ss += NG;
            if (ss >= NS) {
                ss -= NS;
                spar ^= 1;
            }
            kc += NG;
            while (kc >= KS) {
                kc -= KS;
                tile += tile_step;
            }
            if (++halo_turn == W) halo_turn = 0;
        }
        if ((DBG & 16) && lane == 0) {
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(P.counters);
            atomicAdd(cnt + 0, (unsigned long long)(clock64() - w_t0));  // converter total
            atomicAdd(cnt + 1, (unsigned long long)w_wait_split);         // waiting for a free byte-plane stage
            atomicAdd(cnt + 2, (unsigned long long)w_wait_raw);           // waiting for a raw stage (TMA / HBM)
            atomicAdd(cnt + 8, (unsigned long long)w_fence);              // fence.proxy.async
            atomicAdd(cnt + 9, (unsigned long long)w_arrive);             // syncwarp + arrivals
        }
    } else if (warp == 5) {
        // =====================================================================================
        // TMA issuer: one box per K-step = 32 samples x (J-1 + 128) row-blocks of one channel
        // =====================================================================================
        if (lane == 0) {
            const int pad_bytes = (X.raw_rows - X.box_rows) * 128;
            const uint32_t box_bytes = (uint32_t)X.box_rows * 128;
            const uint32_t raw_u32 = smem_u32(raw);
            int rs = 0;
            uint32_t rpar = 1;
            for (long long tile = first_tile; tile < tile_end; tile += tile_step) {
                const unsigned tl = (unsigned)tile;
                const unsigned ch = tl / (unsigned)P.tiles_per_ch;
                const unsigned tt = tl - ch * (unsigned)P.tiles_per_ch;
                const int row0 = (int)tt * P.nrb - (J - 1);  // negative / beyond the end: zero-filled by the TMA unit
                for (int kc = 0; kc < KS; ++kc) {
                    mbar_wait(bar_rempty + 8 * rs, rpar, P.error_flag);
                    if (!(DBG & 2)) {
                        mbar_expect_tx(bar_rfull + 8 * rs, box_bytes);
                        tma_load_3d(raw_u32 + rs * raw_bytes + pad_bytes, &in_map, 32 * kc, row0, (int)ch, bar_rfull + 8 * rs);
                    } else {  // timing experiment: no global loads
                        mbar_arrive(bar_rfull + 8 * rs);
                    }
                    if (++rs == NR) {
                        rs = 0;
                        rpar ^= 1;
                    }
                }
            }
        }
    } else if (warp == 4) {
        tc_mma_role<DBG>(P, role);
    } else {
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
