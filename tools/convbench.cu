// convbench.cu -- rate of dec_tma_kernel's converter loop in isolation (no TMA, no MMAs, no barriers): how many
// samples per clock per SM do N warps turn from a raw stage into a byte-plane stage, with and without the fused
// NCO mix, as a function of the number of warps and of the registers the compiler may use?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/convbench tools/convbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../srcdsp_b200/csrc/kernels_dec_tma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

namespace srcdsp {
std::string &last_error_ref() { static std::string s; return s; }
int fail(int code, const char *, ...) { return code; }
}
using namespace srcdsp;

// one "K-step" per iteration per group of 4 warps: 32 main row groups of 4 rows (8 per warp), raw stage 132 rows x 128 B
template <bool MIX, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) conv_kernel(int iters, unsigned seq_mask, unsigned long long *clk)
{
    extern __shared__ __align__(128) uint8_t sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nw = blockDim.x >> 5;
    uint32_t *tab = reinterpret_cast<uint32_t *>(sm);                 // 8 * (seq_mask + 1) bytes
    uint8_t *raw = sm + 8 * (seq_mask + 1);                           // 2 raw stages
    const int raw_bytes = 132 * 128, rbp = 265, chunk = rbp * 16, stage_bytes = 4 * chunk;
    uint8_t *stages = raw + 2 * raw_bytes;                            // 2 plane stages
    for (int i = tid; i < (int)(8 * (seq_mask + 1) + 2 * raw_bytes) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(sm)[i] = hash32(7, blockIdx.x, i);
    __syncthreads();
    const int wi = warp & 3, g = warp >> 2;
    const int piece = lane & 7, grp = lane >> 3;
    const int src_lane = grp * 128 + piece * 16;
    const uint32_t src_main = smem_u32(raw) + (1 + wi) * 512 + src_lane + (g & 1) * raw_bytes;
    const uint32_t dst_main = smem_u32(stages) + (1 + wi) * 128 + tma_stm_lane(lane, chunk) + (g & 1) * stage_bytes;
    const uint32_t tab_u32 = smem_u32(tab);
    const unsigned mask4 = seq_mask << 2;
    unsigned n_lane = ((4 * wi + grp) * 512 + 4 * piece) & seq_mask;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const unsigned idx4 = ((n_lane + 32 * (it & 15)) & seq_mask) << 2;
        if (MIX) {
            const MixPiece m = tma_mix_piece(tab_u32, idx4, mask4 + 4);
            tma_convert4_same<4>(src_main, dst_main, m);
            tma_convert4_same<4>(src_main + 16 * 512, dst_main + 16 * 128, m);
        } else {
            tma_convert4<false, 4>(src_main, dst_main, tab_u32, 0, 0, 0);
            tma_convert4<false, 4>(src_main + 16 * 512, dst_main + 16 * 128, tab_u32, 0, 0, 0);
        }
        fence_async_smem();
        __syncwarp();
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
    (void)nw;
}

template <bool MIX, int MAXT, int MINB>
void run(int nw, unsigned seq_len, unsigned long long *d_clk)
{
    const int iters = 4000;
    const size_t smem = 8 * seq_len + 2 * 132 * 128 + 2 * 4 * 265 * 16;
    CK(cudaFuncSetAttribute(conv_kernel<MIX, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, conv_kernel<MIX, MAXT, MINB>));
    conv_kernel<MIX, MAXT, MINB><<<148, 32 * nw, smem>>>(10, seq_len - 1, d_clk);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    conv_kernel<MIX, MAXT, MINB><<<148, 32 * nw, smem>>>(iters, seq_len - 1, d_clk);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long h[148];
    CK(cudaMemcpy(h, d_clk, sizeof h, cudaMemcpyDeviceToHost));
    double cyc = 0;
    for (int i = 0; i < 148; ++i) cyc += (double)h[i] / 148;
    // per iteration a warp converts 8 pieces x 128 samples
    const double samples_per_sm = (double)iters * nw * 8 * 128;
    printf("%s warps %2d regs %3d seq %4u: %6.2f samples/clk/SM  (%.3f ms; a 4.29 G-sample batch would take %.2f ms at 1.965 GHz)\n",
           MIX ? "mix  " : "plain", nw, fa.numRegs, seq_len, samples_per_sm / cyc, ms,
           4.295e9 / 148 / (samples_per_sm / cyc) / 1.965e9 * 1e3);
}

int main()
{
    unsigned long long *d_clk;
    CK(cudaMalloc(&d_clk, 148 * 8));
    // register budgets: 576 threads -> up to 113 registers (dec_tma_kernel's MIX instantiation uses 96), 768 -> 85, 1024 -> 64
    for (int nw : {4, 8, 12, 16, 18}) run<true, 576, 1>(nw, 512, d_clk);
    for (int nw : {12, 16, 20, 24}) run<true, 768, 1>(nw, 512, d_clk);
    for (int nw : {12, 16, 20, 24, 28, 32}) run<true, 1024, 1>(nw, 512, d_clk);
    for (int nw : {12, 16}) run<true, 576, 1>(nw, 4096, d_clk);
    for (int nw : {4, 8, 12, 16}) run<false, 576, 1>(nw, 512, d_clk);
    for (int nw : {16, 24, 32}) run<false, 1024, 1>(nw, 512, d_clk);
    return 0;
}
