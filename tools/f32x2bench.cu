// f32x2bench.cu -- issue / pipe rate of the packed FP32 instructions of sm_100 (FFMA2, FADD2, FMUL2) against the scalar
// ones, per SM:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2bench tools/f32x2bench.cu && ./f32x2bench
// Every thread runs 8 independent chains; 1 CTA of W warps per SM; the result is lane-operations per clock and SM
// (128 = the FP32 pipe's nominal rate).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmas(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float adds(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float muls(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

__device__ __forceinline__ u64 bc(float v) { u64 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
struct Taps { float c[64]; unsigned nzb, nzm; };

// the decimator's operand kinds: sample = register pair, tap = uniform scalar (parameter space), -0.0 = scalar register
template <int MODE>
__global__ void k2(u64 *out, long long *clk, int iters, const __grid_constant__ Taps T)
{
    u64 a[8];
    const float nzf = __uint_as_float(T.nzb | (threadIdx.x & T.nzm));
    const u64 nz = bc(nzf);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk((float)(threadIdx.x + i), (float)i + 0.5f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u64 x = a[(i + 3) & 7];
            const float kf = T.c[(i + 8 * (it & 7))];
            if (MODE == 0) a[i] = fma2(x, bc(kf), nz);                                   // FFMA2 R, R.x2, UR.F32, R.F32
            if (MODE == 1) a[i] = add2(a[i], fma2(x, bc(kf), nz));                       // MAC A: FFMA2 + FADD2
            if (MODE == 2) { const float2 v = up(x); a[i] = add2(a[i], pk(muls(v.x, kf), muls(v.y, kf))); }  // MAC D: 2 FMUL + FADD2
            if (MODE == 3) { const float2 p = up(mul2(x, bc(kf))), v = up(a[i]); a[i] = pk(adds(v.x, p.x), adds(v.y, p.y)); }  // MAC C: FMUL2 + 2 FADD
            if (MODE == 4) { const float2 x2 = up(x), v = up(a[i]); a[i] = pk(adds(v.x, muls(x2.x, kf)), adds(v.y, muls(x2.y, kf))); }  // scalar: 2 FMUL + 2 FADD
        }
    }
    const long long t1 = clock64();
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
__global__ void k(u64 *out, long long *clk, int iters, u64 seed)
{
    u64 a[8], x = seed + threadIdx.x, kk = seed * 3 + 1, nz = seed ^ 0x8000000080000000ull;
    float s[16], xs = (float)threadIdx.x, ks = 1.0001f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i, s[2 * i] = (float)i, s[2 * i + 1] = (float)i + 0.5f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma2(a[i], kk, x);                       // FFMA2 only: 2 lane-ops (x2 flops)
            if (MODE == 1) a[i] = add2(a[i], x);                           // FADD2 only
            if (MODE == 2) a[i] = mul2(a[i], kk);                          // FMUL2 only
            if (MODE == 3) a[i] = add2(a[i], fma2(a[(i + 3) & 7], kk, nz)); // the decimator's MAC: FFMA2 + FADD2
            if (MODE == 4) s[2 * i] = fmas(s[2 * i], ks, xs), s[2 * i + 1] = fmas(s[2 * i + 1], ks, xs);  // scalar FFMA
            if (MODE == 5) s[2 * i] = adds(s[2 * i], muls(s[(2 * i + 5) & 15], ks)), s[2 * i + 1] = adds(s[2 * i + 1], muls(s[(2 * i + 6) & 15], ks));  // FMUL + FADD
            if (MODE == 6) a[i] = add2(a[i], mul2(x, a[(i + 1) & 7]));     // FMUL2 + FADD2 (contracted by ptxas?)
        }
    }
    const long long t1 = clock64();
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ (u64)__float_as_uint(s[2 * i]) ^ (u64)__float_as_uint(s[2 * i + 1]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int lane_ops_per_iter_thread)
{
    u64 *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 4096;
        k<MODE><<<148, warps * 32>>>(out, clk, iters, 12345);
        k<MODE><<<148, warps * 32>>>(out, clk, iters, 12345);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-28s warps/SM=%2d  %.1f lane-ops/clk/SM  (%.2f instr/clk/SM)\n", name, warps,
               (double)lane_ops_per_iter_thread * iters * warps * 32 / avg, (double)lane_ops_per_iter_thread * iters * warps / avg / (MODE >= 4 && MODE <= 5 ? 1 : 2));
    }
    cudaFree(out); cudaFree(clk);
}
template <int MODE>
void run2(const char *name, int lane_ops_per_iter_thread)
{
    u64 *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
    Taps T;
    for (int i = 0; i < 64; ++i) T.c[i] = 1.0f + 1e-6f * i;
    T.nzb = 0x80000000u, T.nzm = 0;
    for (int warps : {8, 16, 24}) {
        const int iters = 4096;
        k2<MODE><<<148, warps * 32>>>(out, clk, iters, T);
        k2<MODE><<<148, warps * 32>>>(out, clk, iters, T);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-44s warps/SM=%2d  %.1f lane-ops/clk/SM\n", name, warps, (double)lane_ops_per_iter_thread * iters * warps * 32 / avg);
    }
    cudaFree(out); cudaFree(clk);
}
int main()
{
    run2<0>("FFMA2 (pair, uniform scalar, scalar)", 16); run2<1>("MAC A: FFMA2 + FADD2", 32); run2<2>("MAC D: 2 FMUL + FADD2", 32);
    run2<3>("MAC C: FMUL2 + 2 FADD", 32); run2<4>("MAC scalar: 2 FMUL + 2 FADD", 32);
    run<0>("FFMA2", 16); run<1>("FADD2", 16); run<2>("FMUL2", 16); run<3>("FFMA2+FADD2 (MAC)", 32);
    run<4>("FFMA scalar", 16); run<5>("FMUL+FADD scalar", 32); run<6>("mul2+add2 (as compiled)", 32);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
