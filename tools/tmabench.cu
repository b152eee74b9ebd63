// tmabench.cu -- can a TMA-fed raw ring + converter warps stream HBM faster than LDG producers?
// One persistent CTA per SM.  Thread 0 of warp 0 issues 2D TMA box loads (ROWS rows of BOXW bytes,
// row pitch `pitch` bytes -- the tensor-core decimator's K-step access pattern) into an NS-stage
// shared-memory ring; CW consumer warps read each stage with LDS.128 and either xor-reduce it
// (mode 0) or byte-plane split it and store it to a second ring with STS.32 (mode 1), exactly the
// per-sample work of dec_tc_kernel's producers.  Prints achieved HBM read bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmabench tools/tmabench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > 200000000u) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// grid-stride over "steps"; a step = one box.  Boxes tile the 2D array [rows_total][pitch]:
// step s -> (tile = s / ksteps, kc = s % ksteps): rows tile*ROWS.., byte column kc*BOXW
template <int MODE>
__global__ void __launch_bounds__(1024, 1) tmastream(const __grid_constant__ CUtensorMap map, long long n_steps, int ksteps,
                                                     int rows, int boxw, int ns, int cw, uint32_t *sink, int sub)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_bytes = rows * boxw;
    uint8_t *raw = smem;
    uint8_t *split = raw + ns * stage_bytes;                // MODE 1: 2 split stages
    uint64_t *bars = reinterpret_cast<uint64_t *>(split + (MODE ? 2 * stage_bytes : 0));
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * 16;
    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, cw);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t parity = 1;
            for (long long s = blockIdx.x; s < n_steps; s += gridDim.x) {
                mbar_wait(bar_empty + 8 * stage, parity);
                mbar_expect_tx(bar_full + 8 * stage, stage_bytes);
                const long long tile = s / ksteps;
                const int kc = (int)(s - tile * ksteps);
                // sub > 1: the stage is filled by `sub` boxes of boxw / sub bytes per row (the map's box is that narrow),
                // each into its own dense region -- the in-place byte-plane layout needs 32-byte rows
                for (int q = 0; q < sub; ++q)
                    tma_load_2d(smem_u32(raw + stage * stage_bytes + q * (stage_bytes / sub)), &map, kc * (boxw / 4) + q * (boxw / sub / 4),
                                (int)(tile * rows), bar_full + 8 * stage);
                if (++stage == ns) {
                    stage = 0;
                    parity ^= 1;
                }
            }
        }
    } else if (warp <= cw) {
        const int c = warp - 1;
        int stage = 0;
        uint32_t parity = 0, acc = 0;
        const int n16 = stage_bytes / 16;
        int sp = 0;
        for (long long s = blockIdx.x; s < n_steps; s += gridDim.x) {
            mbar_wait(bar_full + 8 * stage, parity);
            const uint4 *src = reinterpret_cast<const uint4 *>(raw + stage * stage_bytes);
            // warp c takes 16-byte pieces c*32+lane, +cw*32, ...
            if (MODE == 0) {
#pragma unroll 4
                for (int i = c * 32 + lane; i < n16; i += cw * 32) {
                    const uint4 q = src[i];
                    acc ^= q.x ^ q.y ^ q.z ^ q.w;
                }
            } else {
                uint8_t *dst = split + sp * stage_bytes;
                const int plane = stage_bytes / 4;
#pragma unroll 4
                for (int i = c * 32 + lane; i < n16; i += cw * 32) {
                    const uint4 q = src[i];
                    const uint32_t a = prmt(q.x, q.y, 0x5140), b = prmt(q.z, q.w, 0x5140);
                    const uint32_t cc = prmt(q.x, q.y, 0x7362), d = prmt(q.z, q.w, 0x7362);
                    // piece i = (row i/8, piece i%8): word address keeps 32 lanes on 32 banks
                    uint32_t *o = reinterpret_cast<uint32_t *>(dst) + i;
                    o[0] = prmt(a, b, 0x5410);
                    o[plane / 4] = prmt(a, b, 0x7632);
                    o[2 * plane / 4] = prmt(cc, d, 0x5410);
                    o[3 * plane / 4] = prmt(cc, d, 0x7632);
                }
                sp ^= 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
            if (++stage == ns) {
                stage = 0;
                parity ^= 1;
            }
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const size_t bytes = (size_t)8 << 30;
    uint8_t *d;
    uint32_t *sink;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 1, bytes));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn encode = (EncodeFn)fn;
    CK(cudaFuncSetAttribute(tmastream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(tmastream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    struct Cfg { int mode, pitch, boxw, rows, ns, cw, l2promo, sub, swz; };
    const Cfg cfgs[] = {
        {0, 2048, 128, 128, 8, 8, 2},  {0, 2048, 128, 128, 12, 8, 2}, {0, 2048, 128, 128, 8, 16, 2}, {0, 2048, 256, 128, 6, 8, 2},
        {0, 2048, 128, 128, 8, 8, 0},  {0, 2048, 128, 128, 8, 8, 3},  {0, 2048, 128, 64, 16, 8, 2},  {0, 2048, 1024, 16, 8, 8, 2},
        {0, 512, 128, 128, 8, 8, 2},   {0, 8192, 128, 128, 8, 8, 2},
        {1, 2048, 128, 128, 8, 8, 2},  {1, 2048, 128, 128, 8, 12, 2}, {1, 2048, 128, 128, 8, 16, 2}, {1, 2048, 128, 128, 6, 24, 2},
        {1, 2048, 256, 128, 4, 16, 2}, {1, 512, 128, 128, 8, 16, 2},
        // a K-step as 2 boxes of 64-byte rows / 4 boxes of 32-byte rows (plain and with the 32-byte swizzle)
        {0, 2048, 128, 128, 8, 8, 2, 2, 0}, {0, 2048, 128, 128, 8, 8, 2, 4, 0}, {0, 2048, 128, 128, 8, 8, 2, 4, 1},
        {0, 2048, 128, 128, 12, 8, 2, 4, 1}, {0, 2048, 128, 128, 8, 8, 1, 4, 1}, {0, 2048, 128, 128, 8, 8, 3, 4, 1},
    };
    for (const Cfg &c : cfgs) {
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)c.pitch / 4, (cuuint64_t)(bytes / c.pitch)};
        const cuuint64_t gstr[1] = {(cuuint64_t)c.pitch};
        const int sub = c.sub ? c.sub : 1;
        const cuuint32_t box[2] = {(cuuint32_t)c.boxw / 4 / sub, (cuuint32_t)c.rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            c.swz ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)c.l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("encode failed %d\n", (int)r);
            continue;
        }
        const int ksteps = c.pitch / c.boxw;
        const long long n_steps = (long long)(bytes / ((size_t)c.rows * c.pitch)) * ksteps;
        const int smem = (c.ns + (c.mode ? 2 : 0)) * c.rows * c.boxw + 1024;
        const int threads = 32 * (1 + c.cw);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        auto launch = [&]() {
            if (c.mode == 0)
                tmastream<0><<<148, threads, smem>>>(map, n_steps, ksteps, c.rows, c.boxw, c.ns, c.cw, sink, sub);
            else
                tmastream<1><<<148, threads, smem>>>(map, n_steps, ksteps, c.rows, c.boxw, c.ns, c.cw, sink, sub);
        };
        launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 4; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("mode=%d pitch=%5d boxw=%4d rows=%3d stages=%2d (%3d KB in flight) consumers=%2d l2promo=%d sub=%d swz=%d : %.0f GB/s\n", c.mode, c.pitch,
               c.boxw, c.rows, c.ns, c.ns * c.rows * c.boxw / 1024, c.cw, c.l2promo, sub, c.swz, (double)bytes * 4 / (ms * 1e-3) / 1e9);
    }
    return 0;
}
