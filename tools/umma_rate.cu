// umma_rate.cu -- how many cycles does one tcgen05.mma take?  One CTA per SM, one thread issues
// REPS back-to-back MMAs (SWIZZLE_NONE K-major operands in shared memory, accumulator in TMEM),
// commits, waits; prints cycles per MMA for a few kinds / shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// kind: 0 = i8 (s8 x s8 -> s32), 1 = f16 (bf16 x bf16 -> f32)
template <int KIND>
__global__ void __launch_bounds__(128) rate(int m, int n, int reps, int distinct, int layout, int alt_acc, long long *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24))
                                         : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24));
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        t0 = clock64();
        // descriptors hoisted: the loop is nothing but MMA issues (layout 0: SWIZZLE_NONE, rows linear at 16 B,
        // what the Toeplitz kernel uses; 2: SWIZZLE_128B, rows of 128 B, SBO 1024)
        const uint64_t lay = (uint64_t)layout << 61;
        const uint64_t da = layout ? (make_desc(a0, 16, 1024) | lay) : make_desc(a0, 512 * 16, 128);
        const uint64_t db = layout ? (make_desc(b0, 16, 1024) | lay) : make_desc(b0, 512 * 16, 128);
        const uint32_t tb2 = tb + (alt_acc ? 256u : 0u);
        (void)distinct;
#pragma unroll 1
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t tbr = (u & 1) ? tb2 : tb;
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tbr), "l"(da), "l"(db), "r"(idesc), "r"(1));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tbr), "l"(da), "l"(db), "r"(idesc), "r"(1));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main()
{
    long long *d;
    CK(cudaMalloc(&d, 148 * 8));
    CK(cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct C { int kind, m, n, distinct, grid, layout, alt; } cs[] = {
        {0, 128, 256, 1, 148, 0, 0}, {0, 128, 256, 1, 148, 0, 1}, {0, 128, 256, 1, 148, 2, 0}, {0, 128, 256, 4, 148, 2, 1}, {0, 128, 64, 1, 148, 2, 0},
        {0, 128, 128, 1, 148, 2, 0}, {1, 128, 256, 1, 148, 2, 0}, {1, 128, 64, 1, 148, 2, 0},  {0, 128, 256, 1, 148, 6, 0}, {0, 128, 64, 1, 148, 6, 0},
        // narrow N with the SWIZZLE_NONE layout (samples on the M side, a band of the taps on the N side)
        {0, 128, 16, 1, 148, 0, 0},  {0, 128, 24, 1, 148, 0, 0},  {0, 128, 32, 1, 148, 0, 0},  {0, 128, 48, 1, 148, 0, 0},  {0, 128, 64, 1, 148, 0, 0},
        {0, 128, 96, 1, 148, 0, 0},  {0, 128, 128, 1, 148, 0, 0}, {0, 128, 48, 1, 148, 0, 1},  {0, 64, 48, 1, 148, 0, 0},   {0, 64, 256, 1, 148, 0, 0}};
    const int reps = 4000;
    for (auto c : cs) {
        for (int it = 0; it < 2; ++it) {
            if (c.kind == 0) rate<0><<<c.grid, 128, 200 * 1024>>>(c.m, c.n, reps, c.distinct, c.layout, c.alt, d);
            else rate<1><<<c.grid, 128, 200 * 1024>>>(c.m, c.n, reps, c.distinct, c.layout, c.alt, d);
            CK(cudaDeviceSynchronize());
        }
        long long h[148];
        CK(cudaMemcpy(h, d, c.grid * 8, cudaMemcpyDeviceToHost));
        double s = 0;
        for (int i = 0; i < c.grid; ++i) s += h[i];
        printf("kind=%s M=%3d N=%3d K=32B distinct=%2d grid=%3d layout=%d alt_acc=%d : %.1f cycles per MMA\n", c.kind ? "f16(bf16)" : "i8", c.m, c.n, c.distinct, c.grid, c.layout, c.alt,
               s / c.grid / reps);
    }
    return 0;
}
