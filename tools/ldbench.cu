// ldbench.cu -- how many bytes in flight does one persistent CTA per SM need to stream HBM?
// Each warp streams "steps" of ROWS lines of 128 B (row stride STRIDE bytes) with a register ring
// of (AHEAD+1) batches x BATCH 16-byte loads per lane, exactly like the tensor-core kernel's
// producers, and xor-reduces the data.  Prints achieved read bandwidth per configuration.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

template <int BATCH, int AHEAD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) stream(const uint4 *src, size_t n16, size_t row_stride16, uint32_t *sink)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t gw = (size_t)blockIdx.x * nw + warp, tw = (size_t)gridDim.x * nw;
    // a "batch" = BATCH iterations x 4 lines (8 lanes per 128-byte line); consecutive iterations are
    // 4 rows apart; rows are row_stride16 apart; consecutive batches of a warp continue down the rows
    const size_t lane_off = (size_t)(lane >> 3) * row_stride16 + (lane & 7);
    const size_t batch_span = (size_t)BATCH * 4 * row_stride16;
    const size_t n_batches = n16 / batch_span;
    uint4 ring[AHEAD + 1][BATCH];
    uint32_t acc = 0;
    size_t b = gw;
#pragma unroll
    for (int k = 0; k < AHEAD; ++k) {
        if (b + k * tw < n_batches)
#pragma unroll
            for (int it = 0; it < BATCH; ++it) {
                const uint4 *p = src + (b + k * tw) * batch_span + lane_off + (size_t)it * 4 * row_stride16;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ring[k][it].x), "=r"(ring[k][it].y), "=r"(ring[k][it].z), "=r"(ring[k][it].w) : "l"(p));
            }
    }
    for (; b < n_batches; b += (AHEAD + 1) * tw) {
#pragma unroll
        for (int r = 0; r <= AHEAD; ++r) {
            const size_t cur = b + r * tw, nxt = cur + AHEAD * tw;
            if (cur < n_batches) {
                if (nxt < n_batches)
#pragma unroll
                    for (int it = 0; it < BATCH; ++it) {
                        const uint4 *p = src + nxt * batch_span + lane_off + (size_t)it * 4 * row_stride16;
                        uint4 &q = ring[(r + AHEAD) % (AHEAD + 1)][it];
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p));
                    }
#pragma unroll
                for (int it = 0; it < BATCH; ++it) acc ^= ring[r][it].x ^ ring[r][it].y ^ ring[r][it].z ^ ring[r][it].w;
            }
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int BATCH, int AHEAD, int WARPS>
void run(const uint4 *d, size_t bytes, size_t row_stride_bytes, uint32_t *sink)
{
    const int warps = WARPS;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const size_t n16 = bytes / 16;
    for (int i = 0; i < 2; ++i) stream<BATCH, AHEAD, WARPS><<<148, warps * 32>>>(d, n16, row_stride_bytes / 16, sink);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) stream<BATCH, AHEAD, WARPS><<<148, warps * 32>>>(d, n16, row_stride_bytes / 16, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // each batch reads BATCH*512 B out of a span of BATCH*4*row_stride bytes
    const double read = (double)(n16 / ((size_t)BATCH * 4 * row_stride_bytes / 16)) * BATCH * 512.0;
    printf("warps=%2d batch=%2d ahead=%d row_stride=%5zu : %.0f GB/s  (%.1f KB in flight per SM)\n", warps, BATCH, AHEAD,
           row_stride_bytes, read * 5 / (ms * 1e-3) / 1e9, warps * 32.0 * 16 * BATCH * AHEAD / 1024);
}

int main()
{
    const size_t bytes = (size_t)8 << 30;
    uint4 *d;
    uint32_t *sink;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 1, bytes));
    for (size_t rs : {(size_t)128, (size_t)2048}) {
        run<8, 1, 7>(d, bytes, rs, sink);
        run<8, 2, 7>(d, bytes, rs, sink);
        run<8, 3, 7>(d, bytes, rs, sink);
        run<4, 3, 7>(d, bytes, rs, sink);
        run<8, 1, 8>(d, bytes, rs, sink);
        run<8, 1, 12>(d, bytes, rs, sink);
        run<8, 2, 12>(d, bytes, rs, sink);
        run<8, 1, 16>(d, bytes, rs, sink);
        run<4, 1, 16>(d, bytes, rs, sink);
        run<4, 2, 16>(d, bytes, rs, sink);
        run<8, 1, 24>(d, bytes, rs, sink);
        run<4, 1, 32>(d, bytes, rs, sink);
    }
    return 0;
}
