// umma_probe.cu -- hardware-assumption probe for the tcgen05 int8 Toeplitz kernel.
// Checks, against a CPU int8 GEMM, on one CTA:
//   (1) kind::i8 MMA with SWIZZLE_NONE K-major operands laid out "linearly along M/N"
//       (row r of a 16-byte K chunk at base + 16*r, SBO = 128 B, LBO = rows*16 B);
//   (2) that moving the descriptor start address by 16*s bytes shifts the operand by s rows
//       (the trick that turns one resident tile into the shifted Toeplitz operands);
//   (3) mixed signedness per instruction (A u8 / s8 with B s8);
//   (4) the register <-> (lane, column) mapping of tcgen05.ld 32x32b and 16x256b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

constexpr int M = 128, N = 256, KB = 32;       // one MMA: K = 32 bytes
constexpr int A_ROWS = M + 16, B_ROWS = N + 16; // extra rows for the shift test

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ uint32_t make_idesc(int a_signed, int b_signed, int m, int n)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

__global__ void __launch_bounds__(128) probe(const int8_t *A, const int8_t *B, int32_t *D32, int32_t *D16, int a_shift,
                                             int b_shift, int a_signed, int second)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *As = smem;                       // [2 kc][A_ROWS][16]
    uint8_t *Bs = smem + 2 * A_ROWS * 16;     // [2 kc][B_ROWS][16]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    // global A: [A_ROWS][32] row-major bytes; B: [B_ROWS][32]
    for (int i = tid; i < A_ROWS * 2; i += 128) {
        int r = i >> 1, kc = i & 1;
        *reinterpret_cast<uint4 *>(As + (kc * A_ROWS + r) * 16) = *reinterpret_cast<const uint4 *>(A + r * 32 + kc * 16);
    }
    for (int i = tid; i < B_ROWS * 2; i += 128) {
        int r = i >> 1, kc = i & 1;
        *reinterpret_cast<uint4 *>(Bs + (kc * B_ROWS + r) * 16) = *reinterpret_cast<const uint4 *>(B + r * 32 + kc * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;

    if (tid == 0) {
        uint64_t da = make_desc(smem_u32(As) + 16 * a_shift, A_ROWS * 16, 128);
        uint64_t db = make_desc(smem_u32(Bs) + 16 * b_shift, B_ROWS * 16, 128);
        mma_i8(tb, da, db, make_idesc(a_signed, 1, M, N), 0);
        if (second) {  // accumulate a second product with unshifted operands and the other A signedness
            uint64_t da2 = make_desc(smem_u32(As), A_ROWS * 16, 128);
            uint64_t db2 = make_desc(smem_u32(Bs), B_ROWS * 16, 128);
            mma_i8(tb, da2, db2, make_idesc(1 - a_signed, 1, M, N), 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // (4a) 32x32b: thread = lane, 8 consecutive columns per load
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tb + ((uint32_t)(32 * warp) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 8; ++k) D32[(32 * warp + (tid & 31)) * N + c0 + k] = (int32_t)v[k];
    }
    // (4b) 16x256b.x1: hypothesis: regs {0,1} = row lane/4, cols 2*(lane%4)+{0,1}; regs {2,3} = row lane/4 + 8
    for (int half = 0; half < 2; ++half) {
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[4];
            const uint32_t taddr = tb + ((uint32_t)(32 * warp + 16 * half) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int lane = tid & 31;
            const int r0 = 32 * warp + 16 * half + lane / 4, cc = c0 + 2 * (lane % 4);
            D16[r0 * N + cc] = (int32_t)v[0];
            D16[r0 * N + cc + 1] = (int32_t)v[1];
            D16[(r0 + 8) * N + cc] = (int32_t)v[2];
            D16[(r0 + 8) * N + cc + 1] = (int32_t)v[3];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tb));
}

int main()
{
    std::vector<int8_t> A(A_ROWS * 32), B(B_ROWS * 32);
    srand(1);
    for (auto &v : A) v = (int8_t)(rand() & 0xFF);
    for (auto &v : B) v = (int8_t)(rand() & 0xFF);
    int8_t *dA, *dB;
    int32_t *d32, *d16;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&d32, M * N * 4));
    CK(cudaMalloc(&d16, M * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    const size_t smem = 2 * (A_ROWS + B_ROWS) * 16;
    int fails = 0;
    struct Case { int a_shift, b_shift, a_signed, second; } cases[] = {
        {0, 0, 1, 0}, {0, 0, 0, 0}, {8, 0, 1, 0}, {0, 2, 1, 0}, {5, 3, 0, 0}, {8, 2, 1, 1}, {3, 7, 0, 1}};
    for (auto c : cases) {
        CK(cudaMemset(d32, 0xFF, M * N * 4));
        CK(cudaMemset(d16, 0xFF, M * N * 4));
        probe<<<1, 128, smem>>>(dA, dB, d32, d16, c.a_shift, c.b_shift, c.a_signed, c.second);
        CK(cudaDeviceSynchronize());
        std::vector<int32_t> h32(M * N), h16(M * N);
        CK(cudaMemcpy(h32.data(), d32, M * N * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h16.data(), d16, M * N * 4, cudaMemcpyDeviceToHost));
        long bad32 = 0, bad16 = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                long acc = 0;
                for (int k = 0; k < KB; ++k) {
                    int a = c.a_signed ? (int)A[(m + c.a_shift) * 32 + k] : (int)(uint8_t)A[(m + c.a_shift) * 32 + k];
                    acc += (long)a * (int)B[(n + c.b_shift) * 32 + k];
                    if (c.second) {
                        int a2 = c.a_signed ? (int)(uint8_t)A[m * 32 + k] : (int)A[m * 32 + k];
                        acc += (long)a2 * (int)B[n * 32 + k];
                    }
                }
                if (h32[m * N + n] != (int32_t)acc) ++bad32;
                if (h16[m * N + n] != (int32_t)acc) ++bad16;
            }
        printf("case a_shift=%d b_shift=%d a_signed=%d second=%d : mismatches 32x32b=%ld 16x256b=%ld  (D[0][0]=%d D[5][9]=%d)\n",
               c.a_shift, c.b_shift, c.a_signed, c.second, bad32, bad16, h32[0], h32[5 * N + 9]);
        fails += (bad32 != 0) + (bad16 != 0);
    }
    printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
    return fails ? 1 : 0;
}
