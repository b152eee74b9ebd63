// kernels_dec_band.cuh -- the tcgen05 int8 Toeplitz decimator with the operand roles SWAPPED and the lags merged:
// "band form".  Same arithmetic, same converters, same rings as dec_tma_kernel; what changes is which side of the MMA
// the Toeplitz structure lives on, and with it how much of every MMA is padding.
//
// dec_tma_kernel puts the taps on the M side (128 rows = 32 outputs x 4 weight slots) and 128 row-blocks of samples on
// the N side: every (K-step, lag) is a full 128 x 256 x 32 MMA although a K-step of 32 samples only meets ~18 of the 32
// outputs of a row-block (16 samples per output against a 255-tap filter), and a second, mostly empty MMA per plane for the
// outputs of the NEXT row-block that reach back into this one (lag 1).  48 MMAs per 4096 outputs where 8 dense ones would
// do; under the boxes' 1 kW power cap those MMAs are the largest removable term.
//
// Here the SAMPLES are the M side -- 128 rows = 64 row-blocks x (re, im), one byte plane per MMA -- and the taps the N side,
// where an MMA may be as narrow as the band is wide: N = 4 * (outputs a K-step meets), e.g. 80 columns instead of 256.
// The outputs of the next row-block that a row-block's samples reach ("extended" outputs b = 32 .. 32 + E - 1, E =
// floor((ntaps - 2) / M) + 1 <= 32) are accumulated in extra accumulator columns of the SAME row and added to the next
// row's outputs in the epilogue, so there is no lag MMA and no row shifting at all:
//
//   D[rho, 4 b + w] += sum_t  plane[rho, t] * d_w'[M (b - a) - t - r]        rho = (row-block, re / im), b < 32 + E
//   out[n, b] = limitScale16( own(n)[b] + ext(n - 1)[b],  shift ),  own / ext = sum_w 256^w D[., 4 b + w]   (mod 2^32)
//
// A tile is 64 rows of row-blocks: row 0 is the previous tile's last row-block (only its extended outputs are used: a
// tile therefore yields 63 row-blocks = 2016 outputs), which also covers the carried history (row-block -1 of a call).
// A shared-memory stage holds TWO K-steps (one TMA box of 64 samples x 64 rows = 16 KB keeps the HBM rate of 16 KB boxes);
// per stage the MMA warp issues 4 MMAs of 128 x N x 32.  The first MMA of a tile spans all accumulator columns with
// accumulate = 0 (rows of the tap master outside the band are zero), so no column is ever accumulated into before it
// is written.  The master is ~7 KB (one copy per residue (32 kc) mod M), which leaves the shared memory to the rings.
//
// Applicability: M even, ntaps <= 32 M + 1 (the filter reaches one row-block back), taps of at most 3 signed byte
// digits, 16-byte aligned rows; everything else runs dec_tma_kernel / dec_tc_kernel / dec_fir_kernel.
#pragma once

#include "kernels_dec_tma.cuh"

namespace srcdsp {

constexpr int BAND_ROWS = 64;          // row-blocks per tile (M side: 128 MMA rows)
constexpr int BAND_NEW = BAND_ROWS - 1;  // row-blocks a tile produces
constexpr int BAND_RBP = 2 * BAND_ROWS + 1;  // padded rows per (plane, kc) chunk: odd -> conflict-free byte-plane stores
constexpr int BAND_HALF_BYTES = 4 * BAND_RBP * 16;  // byte planes of one K-step: 2 planes x 2 kc chunks
constexpr int BAND_STAGE_BYTES = 2 * BAND_HALF_BYTES;
constexpr int BAND_RAW_BYTES = BAND_ROWS * 256;
constexpr int BAND_UOFF = 4;           // master rows in front of u = 0 (hi-plane shift, clamped windows)
constexpr int BAND_EPI_BAR = 2;        // named barrier of the 4 epilogue warps

struct BandExtra {
    int n_raw, n_conv, shared_raw;
    long long rows_full;   // row-blocks that are part of the TMA tensor: floor(n_in / G)
    int E;                 // extended outputs per row-block
    int a_rows;            // rows per kc chunk of a master copy (-> LBO of the tap operand)
    int ks2;               // stages per tile: M / 2
    int plan_off;          // byte offset of the per-K-step plan inside the master image: uint4 {B lo >> 4, B hi >> 4, N, D column}
};

template <bool MIX, int W>
__global__ void __launch_bounds__(32 * (TMA_CONV_WARP0 + TMA_MAX_CONV_MIX), 1)  // at most 3 converter groups of 4 warps: 113 registers
    dec_band_kernel(const __grid_constant__ TcParams P, const __grid_constant__ BandExtra X, const __grid_constant__ CUtensorMap in_map)
{
    extern __shared__ __align__(128) uint8_t tc_smem_raw[];
    uint8_t *smem = tc_smem_raw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = P.n_stages, NR = X.n_raw, NCW = X.n_conv;
    uint8_t *b_smem = smem;  // tap master (+ plan)
    uint32_t *tab_smem = reinterpret_cast<uint32_t *>(smem + ((P.master_bytes + 127) & ~127));
    uint8_t *raw = reinterpret_cast<uint8_t *>(tab_smem) + P.table_bytes;
    uint8_t *stages = raw + NR * BAND_RAW_BYTES;
    int *ext_x = reinterpret_cast<int *>(stages + NS * BAND_STAGE_BYTES);  // [2 tile parities][4 warps][2 lanes][32] boundary hand-over
    uint64_t *bars = reinterpret_cast<uint64_t *>(ext_x + 2 * 4 * 2 * 32);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * TC_MAX_STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * TC_MAX_STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t bar_rfull = bar_tempty + 16, bar_rempty = bar_rfull + 8 * TMA_MAX_RAW;
    __shared__ uint32_t tmem_base_s;

    // ---- setup ------------------------------------------------------------------------------
    for (int i = tid; i < P.master_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(b_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(P.master) + i);
    for (int i = tid; i < (NR * BAND_RAW_BYTES + NS * BAND_STAGE_BYTES) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(raw)[i] = make_uint4(0, 0, 0, 0);  // the padding row of every chunk must be defined
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + 8 * s, W);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < NR; ++s) {
            mbar_init(bar_rfull + 8 * s, 1);
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

    const long long first_tile = P.total_tiles * blockIdx.x / gridDim.x, tile_end = P.total_tiles * (blockIdx.x + 1) / gridDim.x;
    const int KS2 = X.ks2;

    if (warp >= TMA_CONV_WARP0) {
        // =====================================================================================
        // converters: raw stage (64 rows of 64 interleaved samples = two K-steps) -> two byte-plane half stages
        // =====================================================================================
        const int cw = warp - TMA_CONV_WARP0;
        const int NG = NCW / W;
        const int g = cw / W, wi = cw - g * W;
        const int piece = lane & 7, grp = lane >> 3;
        constexpr int chunk = BAND_RBP * 16;
        constexpr int hi_off = 2 * chunk;
        const int src_lane = grp * 256 + piece * 16;
        const int dst_lane = (piece >> 2) * chunk + grp * 32 + (piece & 3) * 4;
        // row groups of this warp: wi + W * i, i < 16 / W (both halves of the stage)
        const uint32_t src_main = smem_u32(raw) + wi * 1024 + src_lane;
        const uint32_t dst_main = smem_u32(stages) + wi * 128 + tma_stm_lane(lane, chunk);  // stmatrix row address of this lane
        const uint32_t tab_u32 = smem_u32(tab_smem);
        const unsigned mask4 = P.seq_mask << 2;
        int rs = g % NR, ss = g % NS;
        uint32_t rpar = 0, spar = 1;
        long long tile = first_tile, cur_tile = -1, tile0 = 0;
        int kk = g;  // stage index within the tile
        while (kk >= KS2 && tile < tile_end) {
            kk -= KS2;
            ++tile;
        }
        const uint32_t *x = nullptr, *hist = nullptr;
        unsigned tt = 0, ph0 = 0, fr = 0, didx4 = 0, n_lane = 0;
        int cur_ch = -1;
        bool edge = false;
        while (tile < tile_end) {
            if (tile != cur_tile) {
                const unsigned tl = (unsigned)tile;
                const unsigned ch = tl / (unsigned)P.tiles_per_ch;
                tt = tl - ch * (unsigned)P.tiles_per_ch;
                x = P.in + (size_t)ch * P.in_stride;
                hist = P.hist_in + (size_t)ch * P.H;
                tile0 = ((long long)tt * BAND_NEW - 1) * (long long)P.G;  // sample of (row 0, K-step 0); negative for the first tile
                edge = tt == 0 || (long long)tt * BAND_NEW - 1 + BAND_ROWS > X.rows_full;
                if (MIX) {
                    if ((int)ch != cur_ch) {
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
                    didx4 = ((unsigned)(4 * W * P.G) & P.seq_mask) << 2;
                    // sample index mod the sequence length of (row 4 * wi + grp, K-step 0, this lane's piece); tile0 may be negative
                    n_lane = (unsigned)((tile0 + (long long)(4 * wi + grp) * P.G + 4 * piece) & (long long)P.seq_mask);
                }
                cur_tile = tile;
            }
            if (X.shared_raw) mbar_wait(bar_rempty + 8 * rs, rpar ^ 1, P.error_flag);
            mbar_wait(bar_rfull + 8 * rs, rpar, P.error_flag);
            mbar_wait(bar_empty + 8 * ss, spar, P.error_flag);
            if (!edge) {
                const uint32_t src = src_main + rs * BAND_RAW_BYTES;
                const uint32_t dst = dst_main + ss * BAND_STAGE_BYTES;
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // the two K-steps of the stage: bytes [0, 128) and [128, 256) of every raw row
                    const unsigned idx4 = MIX ? ((n_lane + 64 * kk + 32 * h) & P.seq_mask) << 2 : 0u;
                    if (MIX && didx4 == 0) {
                        const MixPiece m = tma_mix_piece(tab_u32, idx4, mask4 + 4);
                        tma_convert4_same<W, 1024>(src + h * 128, dst + h * BAND_HALF_BYTES, m);
                    } else {
                        tma_convert4<MIX, W, 1024>(src + h * 128, dst + h * BAND_HALF_BYTES, tab_u32, idx4, didx4, mask4);
                    }
                }
            } else {
                // tiles with rows the TMA tensor does not hold (the carried history in front of the call, the ragged
                // end): those rows come straight from global memory, sample by sample
                const uint8_t *src = raw + rs * BAND_RAW_BYTES + src_lane;
                uint8_t *dst = stages + ss * BAND_STAGE_BYTES + dst_lane;
                for (int q = wi; q < BAND_ROWS / 4; q += W) {
                    const int row = 4 * q + grp;
                    const long long rb = (long long)tt * BAND_NEW - 1 + row;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const long long n = tile0 + (long long)row * P.G + 64 * kk + 32 * h + 4 * piece;
                        uint8_t *d = dst + q * 128 + h * BAND_HALF_BYTES;
                        uint4 v;
                        if (rb < 0 || rb >= X.rows_full) {
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
                            tma_convert_store<false>(v, d, hi_off, tab_smem, 0, 0);
                        } else {
                            v = *reinterpret_cast<const uint4 *>(src + q * 1024 + h * 128);
                            tma_convert_store<MIX>(v, d, hi_off, tab_smem, (unsigned)(n & (long long)P.seq_mask), P.seq_mask);
                        }
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_full + 8 * ss);
                mbar_arrive(bar_rempty + 8 * rs);
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
            kk += NG;
            while (kk >= KS2) {
                kk -= KS2;
                ++tile;
            }
        }
    } else if (warp == 5) {
        // =====================================================================================
        // TMA issuer: one box per stage = 64 samples x 64 row-blocks of one channel
        // =====================================================================================
        if (lane == 0) {
            const uint32_t raw_u32 = smem_u32(raw);
            int rs = 0;
            uint32_t rpar = 1;
            for (long long tile = first_tile; tile < tile_end; ++tile) {
                const unsigned tl = (unsigned)tile;
                const unsigned ch = tl / (unsigned)P.tiles_per_ch;
                const unsigned tt = tl - ch * (unsigned)P.tiles_per_ch;
                const int row0 = (int)tt * BAND_NEW - 1;  // row -1 / rows beyond the end: zero-filled by the TMA unit
                for (int kk = 0; kk < KS2; ++kk) {
                    mbar_wait(bar_rempty + 8 * rs, rpar, P.error_flag);
                    mbar_expect_tx(bar_rfull + 8 * rs, BAND_RAW_BYTES);
                    tma_load_3d(raw_u32 + rs * BAND_RAW_BYTES, &in_map, 64 * kk, row0, (int)ch, bar_rfull + 8 * rs);
                    if (++rs == NR) {
                        rs = 0;
                        rpar ^= 1;
                    }
                }
            }
        }
    } else if (warp == 4) {
        // =====================================================================================
        // MMA issuer: per stage 2 K-steps x (lo plane, hi plane) = 4 MMAs of 128 x N x 32, N = the band's width
        // =====================================================================================
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_const = (smem_u32(stages) >> 4) + (((uint32_t)BAND_RBP) << 16);
        const uint32_t b_const = (smem_u32(b_smem) >> 4) + ((((uint32_t)X.a_rows * 16) >> 4) << 16);
        const uint32_t idesc_lo = umma_idesc_i8(0, 1, 128, 0), idesc_hi = umma_idesc_i8(1, 1, 128, 0);  // samples: lo u8 / hi s8; taps s8
        const uint4 *plan = reinterpret_cast<const uint4 *>(b_smem + X.plan_off);
        auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
        int stage = 0;
        uint32_t sb16 = 0, phase = 0, acc_phases = 0;
        int acc = 0;
        for (long long tile = first_tile; tile < tile_end; ++tile) {
            mbar_wait(bar_tempty + 8 * acc, ((acc_phases >> acc) & 1) ^ 1, P.error_flag);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kk = 0; kk < KS2; ++kk) {
                const uint4 e0 = plan[2 * kk], e1 = plan[2 * kk + 1];
                if (P.mma_sleep_ns)
                    mbar_wait_backoff(bar_full + 8 * stage, phase, P.error_flag, P.mma_sleep_ns);
                else
                    mbar_wait(bar_full + 8 * stage, phase, P.error_flag);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t as = a_const + sb16;
                    const uint32_t n0 = (e0.z >> 3) << 17, n1 = (e1.z >> 3) << 17;
                    umma_i8(d_tmem + e0.w, desc(as), desc(b_const + e0.x), idesc_lo | n0, kk != 0);
                    umma_i8(d_tmem + e0.w, desc(as + (2 * BAND_RBP)), desc(b_const + e0.y), idesc_hi | n0, 1);
                    umma_i8(d_tmem + e1.w, desc(as + (BAND_HALF_BYTES >> 4)), desc(b_const + e1.x), idesc_lo | n1, 1);
                    umma_i8(d_tmem + e1.w, desc(as + (BAND_HALF_BYTES >> 4) + (2 * BAND_RBP)), desc(b_const + e1.y), idesc_hi | n1, 1);
                    tc_commit(bar_empty + 8 * stage);
                    if (kk == KS2 - 1) tc_commit(bar_tfull + 8 * acc);
                }
                __syncwarp();
                sb16 += BAND_STAGE_BYTES >> 4;
                if (++stage == NS) {
                    stage = 0;
                    sb16 = 0;
                    phase ^= 1;
                }
            }
            acc_phases ^= 1u << acc;
            acc ^= 1;
        }
    } else {
        // =====================================================================================
        // epilogue warps 0..3: TMEM lane 32 * warp + lane = tile row 16 * warp + lane / 2, component lane & 1
        // =====================================================================================
        const int comp = lane & 1;
        const int row = 16 * warp + (lane >> 1);
        const int E = X.E;
        const uint32_t swap_sel = comp ? 0x1032u : 0x3210u;  // the im lane of a pair holds (other, mine) = (re, im) swapped
        uint32_t acc_phases = 0;
        int acc = 0;
        unsigned par = 0;
        for (long long tile = first_tile; tile < tile_end; ++tile, par ^= 1) {
            const unsigned ch = (unsigned)tile / (unsigned)P.tiles_per_ch;
            const long long tt = (long long)((unsigned)tile - ch * (unsigned)P.tiles_per_ch);
            uint32_t *o = P.out + (size_t)ch * P.out_stride;
            if (P.epi_sleep_ns)
                mbar_wait_backoff(bar_tfull + 8 * acc, (acc_phases >> acc) & 1, P.error_flag, P.epi_sleep_ns);
            else
                mbar_wait(bar_tfull + 8 * acc, (acc_phases >> acc) & 1, P.error_flag);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(32 * warp) << 16) + acc * 256;
            auto load8 = [&](uint32_t col, uint32_t (&val)[8]) {  // 32 columns = 8 outputs x 4 weight slots -> 8 sums (mod 2^32)
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
                    : "r"(t_addr + col));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = v[4 * k] + (v[4 * k + 1] << 8) + (v[4 * k + 2] << 16) + (v[4 * k + 3] << 24);
            };
            // phase 1: the extended outputs of this warp's LAST row (lanes 30, 31) belong to the next warp's first row.
            // (Every extended chunk is whole: the tile's first MMA writes the accumulator up to a multiple of 8 outputs.)
            int *xo = ext_x + ((par * 4 + warp) * 2 + comp) * 32;
            for (int j = 0; 8 * j < E; ++j) {
                uint32_t e8[8];
                load8(128 + 32 * j, e8);
                if (lane >= 30) {
                    *reinterpret_cast<uint4 *>(xo + 8 * j) = make_uint4(e8[0], e8[1], e8[2], e8[3]);
                    *reinterpret_cast<uint4 *>(xo + 8 * j + 4) = make_uint4(e8[4], e8[5], e8[6], e8[7]);
                }
            }
            asm volatile("bar.sync %0, 128;" ::"n"(BAND_EPI_BAR) : "memory");
            // phase 2: own outputs + the previous row's extended outputs, scale, pack (re, im), store
            const int *xi = ext_x + ((par * 4 + (warp > 0 ? warp - 1 : 0)) * 2 + comp) * 32;
            const long long n_rb = tt * BAND_NEW + row - 1;  // output row-block of this row (row 0: none)
            const long long idx0 = n_rb * 32 + 4 * comp;
            uint32_t *op = o + idx0;
            // whole row inside the block and 16-byte aligned rows: no per-store checks
            const bool fast = row > 0 && P.vec_out && (n_rb + 1) * 32 <= P.n_out;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t v8[8];
                load8(32 * j, v8);
                if (8 * j < E) {
                    uint32_t e8[8];
                    load8(128 + 32 * j, e8);
                    uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
                    if (lane < 2) {
                        x0 = *reinterpret_cast<const uint4 *>(xi + 8 * j);
                        x1 = *reinterpret_cast<const uint4 *>(xi + 8 * j + 4);
                    }
                    const uint32_t xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, e8[k], 2);
                        v8[k] += lane < 2 ? xs[k] : up;
                    }
                }
                // the pair (re lane, im lane) shares the 8 outputs: the re lane packs and stores outputs 0..3 of the chunk,
                // the im lane outputs 4..7
                uint32_t w4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t mine = comp ? v8[4 + k] : v8[k];
                    const uint32_t send = comp ? v8[k] : v8[4 + k];
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, send, 1);
                    // limitScale16 (dsp_complex.cpp:63-73) of both components at once, then (re, im) into place
                    w4[k] = prmt(scale_pack_sym_sat((int)mine, (int)other, P.shift), 0u, swap_sel);
                }
                if (fast) {
                    *reinterpret_cast<uint4 *>(op + 8 * j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                } else if (row > 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (idx0 + 8 * j + k < P.n_out) op[8 * j + k] = w4[k];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            acc_phases ^= 1u << acc;
            acc ^= 1;
        }
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
