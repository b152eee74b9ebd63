// pipebench.cu -- issue rate of a few integer instructions on sm_100a (thread-ops per clock per SM),
// alone and paired, to see which pipe they share.  Used to choose the NCO mixer's instruction mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipebench tools/pipebench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

template <int OP>
__device__ __forceinline__ void step(int &a, int &b, int c)
{
    if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));                 // IMAD
    if (OP == 1) asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));            // IDP.2A
    if (OP == 2) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(a) : "r"(b));                             // IMAD.HI
    if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));                       // PRMT
    if (OP == 4) asm volatile("shr.s32 %0, %0, 3;" : "+r"(a));                                          // SHF
    if (OP == 5) asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(a) : "r"(b));                   // I2IP
    if (OP == 6) asm volatile("dp4a.s32.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));               // IDP.4A
    if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));             // LOP3
    if (OP == 8) asm volatile("{.reg .b32 t; max.s16x2 t, %0, %1; mov.b32 %0, t;}" : "+r"(a) : "r"(b)); // VIMNMX.S16x2
    if (OP == 9) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(*(long long *)&a) : "r"(b), "r"(c)); // placeholder (unused)
}

template <int OP1, int OP2>
__global__ void __launch_bounds__(1024) k(int *out, int iters, int seed)
{
    int a[8], b = seed + threadIdx.x, c = seed * 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * (i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            step<OP1>(a[i], b, c);
            if (OP2 >= 0) step<OP2 < 0 ? 0 : OP2>(a[i], b, c);
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345) out[0] = s;
}

template <int OP1, int OP2>
void run(const char *name, int *d)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<OP1, OP2><<<148, 1024>>>(d, 16, 1);
    CK(cudaEventRecord(e0));
    k<OP1, OP2><<<148, 1024>>>(d, iters, 1);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = 1024.0 * iters * 8 * (OP2 >= 0 ? 2 : 1);  // thread-ops per SM
    printf("%-28s %6.1f thread-ops / clk / SM (at 1.965 GHz)\n", name, ops / (ms * 1e-3 * 1.965e9));
}

int main()
{
    int *d;
    CK(cudaMalloc(&d, 4));
    run<0, -1>("IMAD", d);
    run<1, -1>("IDP.2A", d);
    run<6, -1>("IDP.4A", d);
    run<2, -1>("IMAD.HI", d);
    run<3, -1>("PRMT", d);
    run<4, -1>("SHF", d);
    run<5, -1>("I2IP.SAT", d);
    run<7, -1>("LOP3", d);
    run<8, -1>("VIMNMX.S16x2", d);
    run<0, 3>("IMAD + PRMT", d);
    run<1, 3>("IDP.2A + PRMT", d);
    run<1, 0>("IDP.2A + IMAD", d);
    run<2, 3>("IMAD.HI + PRMT", d);
    run<2, 0>("IMAD.HI + IMAD", d);
    run<5, 3>("I2IP + PRMT", d);
    run<5, 0>("I2IP + IMAD", d);
    run<8, 3>("VIMNMX + PRMT", d);
    run<4, 3>("SHF + PRMT", d);
    return 0;
}
