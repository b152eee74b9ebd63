// common.cuh -- shared device helpers and host-side error plumbing for libsrcdsp_b200.
//
// Arithmetic contract (bit-exact with the reference):
//   limitScale16  (dsp_complex.cpp:63-73): arithmetic >> then symmetric clamp to +-32767
//   limitScale<>  (dsp_complex.h:83-108) : arithmetic >> then clamp to [-32768, 32767]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/srcdsp_b200.h"

// Timing-experiment switches (kernel variants that skip work and therefore give WRONG results) exist only in
// builds with -DSRCDSP_TIMING_EXPERIMENTS; in the shipped library the tests below are compile-time false.
#ifdef SRCDSP_TIMING_EXPERIMENTS
#define SRCDSP_EXP(params, bits) (((params).debug & (bits)) != 0)
#else
#define SRCDSP_EXP(params, bits) false
#endif

namespace srcdsp {

// ---------------------------------------------------------------------------------------------
// host: thread-local error message + status helpers
// ---------------------------------------------------------------------------------------------
std::string &last_error_ref();
int fail(int code, const char *fmt, ...);

#define SRCDSP_CUDA(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::srcdsp::fail(SRCDSP_E_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

#define SRCDSP_TRY(expr)                 \
    do {                                 \
        int _s = (expr);                 \
        if (_s != SRCDSP_OK) return _s;  \
    } while (0)

// No C++ exception crosses the C ABI ("never throw across the ABI"): every extern "C" entry point that
// does more than return a field is a function-try-block closed by SRCDSP_ABI_CATCH.  std::bad_alloc (a std::vector
// growing under memory pressure) becomes SRCDSP_E_NOMEM, anything else SRCDSP_E_INVALID, with the message kept.
int abi_exception() noexcept;
#define SRCDSP_ABI_CATCH \
    catch (...) { return ::srcdsp::abi_exception(); }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// packed sample word: low half = I (re), high half = Q (im), little endian == cs16 in memory
__device__ __forceinline__ int sx_lo(uint32_t w)
{
    int r;
    // PRMT with sign replication: bytes {b0, b1, sign(b1), sign(b1)}  (ALU pipe, 1 instruction)
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w));
    return r;
}
__device__ __forceinline__ int sx_hi(uint32_t w) { return ((int)w) >> 16; }

__device__ __forceinline__ uint32_t pack_iq(int re, int im)
{
    uint32_t r;
    // bytes {re.b0, re.b1, im.b0, im.b1}
    asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(re), "r"(im));
    return r;
}

// limitScale16 per component: symmetric clamp
__device__ __forceinline__ int clamp_sym(int v) { return min(max(v, -32767), 32767); }
// limitScale<cs16> per component: asymmetric clamp
__device__ __forceinline__ int clamp_asym(int v) { return min(max(v, -32768), 32767); }

template <bool SYM>
__device__ __forceinline__ uint32_t scale_pack(int re, int im, unsigned shift)
{
    re >>= shift;
    im >>= shift;
    if (SYM) {
        re = clamp_sym(re);
        im = clamp_sym(im);
    } else {
        re = clamp_asym(re);
        im = clamp_asym(im);
    }
    return pack_iq(re, im);
}

// limitScale16 of a complex value with the packed saturate: cvt.pack.sat clamps each half to
// [-32768, 32767], the per-half signed max with -32767 restores the symmetric clamp (dsp_complex.cpp:63-73)
__device__ __forceinline__ uint32_t scale_pack_sym_sat(int re, int im, unsigned shift)
{
    re >>= shift;
    im >>= shift;
    uint32_t p;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(p) : "r"(im), "r"(re));  // {hi = sat(im), lo = sat(re)}
    return __vmaxs2(p, 0x80018001u);
}

// limitScale<cs16> of a complex value (dsp_complex.h:83-108): arithmetic >> then clamp to [-32768, 32767] per
// component -- exactly what the saturating pack does
__device__ __forceinline__ uint32_t scale_pack_asym_sat(int re, int im, unsigned shift)
{
    re >>= shift;
    im >>= shift;
    uint32_t p;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(p) : "r"(im), "r"(re));  // {hi = sat(im), lo = sat(re)}
    return p;
}

// One NCO mix: mixers.h:175-176.  cs = packed (cos, sin) = (T[(phi + N/4) % N], T[phi]).
__device__ __forceinline__ uint32_t mix_sample(uint32_t x, uint32_t cs)
{
    const int xr = sx_lo(x), xi = sx_hi(x);
    const int c = sx_lo(cs), s = sx_hi(cs);
    const int r = xr * c - xi * s;   // dsp_complex.cpp:31-37
    const int i = xi * c + s * xr;
    return scale_pack<true>(r, i, 14);
}

// Same arithmetic with the packed saturate: cvt.pack.sat gives [-32768, 32767] per half, the
// per-half signed max with -32767 restores limitScale16's symmetric clamp (dsp_complex.cpp:63-73).
__device__ __forceinline__ uint32_t mix_sample_packed(uint32_t x, uint32_t cs)
{
    const int xr = sx_lo(x), xi = sx_hi(x);
    const int c = sx_lo(cs), s = sx_hi(cs);
    const int r = (xr * c - xi * s) >> 14;
    const int i = (xi * c + s * xr) >> 14;
    uint32_t p;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(p) : "r"(i), "r"(r));  // {hi = sat(i), lo = sat(r)}
    return __vmaxs2(p, 0x80018001u);
}

// NCO mix (mixers.h:172-177) on the integer dot-product pipe.  The sample word is already the packed pair
// (re, im) that dp2a takes; the oscillator value is kept as signed byte digits, cos = 256 * c1 + c0 etc.:
//   Bre = bytes {c0, n0, c1, n1}  (n = -sin):  xr*cos - xi*sin = dp2a.lo(x, Bre) + 256 * dp2a.hi(x, Bre)
//   Bim = bytes {s0, c0, s1, c1}            :  xr*sin + xi*cos = dp2a.lo(x, Bim) + 256 * dp2a.hi(x, Bim)
// exactly the int32 products of dsp_complex.cpp:31-37, with no unpacking of x or of the table entry: the
// multiply pipe (IMAD / IDP, 64 lanes/clk) takes 6 instructions per sample and the ALU pipe (PRMT / SHF / I2IP /
// VIMNMX, also 64 lanes/clk, the binding one -- tools/pipebench.cu) only the 2 shifts, the saturating pack and
// the symmetric clamp (+ the 2 PRMT of the byte-plane split).
__device__ __forceinline__ uint32_t mix_sample_dp2a(uint32_t x, uint32_t bre, uint32_t bim)
{
    const int r = __dp2a_lo((int)x, (int)bre, __dp2a_hi((int)x, (int)bre, 0) * 256) >> 14;
    const int i = __dp2a_lo((int)x, (int)bim, __dp2a_hi((int)x, (int)bim, 0) * 256) >> 14;
    uint32_t p;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(p) : "r"(i), "r"(r));  // {hi = sat(i), lo = sat(r)}
    return __vmaxs2(p, 0x80018001u);  // limitScale16's symmetric clamp (dsp_complex.cpp:63-73)
}
// The same mix for the 4 samples of a 16-byte piece, delivered as the 4 byte-plane words of the tcgen05 sample
// operand (byte k of a word = sample k): the saturating pack is given the SAME component of two samples, so the
// re / im de-interleave costs nothing and each plane is one PRMT of two packed pairs -- 4 PRMT per piece
// instead of the 8 that splitting 4 packed (re, im) words takes.
__device__ __forceinline__ int mix_component_dp2a(uint32_t x, uint32_t digits)
{
    return __dp2a_lo((int)x, (int)digits, __dp2a_hi((int)x, (int)digits, 0) * 256) >> 14;
}
__device__ __forceinline__ uint32_t pack2_sym_sat(int lo, int hi)
{
    uint32_t p;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(p) : "r"(hi), "r"(lo));
    return __vmaxs2(p, 0x80018001u);  // limitScale16's symmetric clamp (dsp_complex.cpp:63-73)
}
__device__ __forceinline__ void mix4_planes_dp2a(const uint4 q, const uint4 bre, const uint4 bim, uint32_t &re_lo, uint32_t &re_hi,
                                                 uint32_t &im_lo, uint32_t &im_hi)
{
    const uint32_t r01 = pack2_sym_sat(mix_component_dp2a(q.x, bre.x), mix_component_dp2a(q.y, bre.y));
    const uint32_t r23 = pack2_sym_sat(mix_component_dp2a(q.z, bre.z), mix_component_dp2a(q.w, bre.w));
    const uint32_t i01 = pack2_sym_sat(mix_component_dp2a(q.x, bim.x), mix_component_dp2a(q.y, bim.y));
    const uint32_t i23 = pack2_sym_sat(mix_component_dp2a(q.z, bim.z), mix_component_dp2a(q.w, bim.w));
    asm("prmt.b32 %0, %1, %2, 0x6420;" : "=r"(re_lo) : "r"(r01), "r"(r23));
    asm("prmt.b32 %0, %1, %2, 0x7531;" : "=r"(re_hi) : "r"(r01), "r"(r23));
    asm("prmt.b32 %0, %1, %2, 0x6420;" : "=r"(im_lo) : "r"(i01), "r"(i23));
    asm("prmt.b32 %0, %1, %2, 0x7531;" : "=r"(im_hi) : "r"(i01), "r"(i23));
}
// digits of one table entry cs = packed (cos, sin)
__device__ __forceinline__ void mix_digits(uint32_t cs, uint32_t &bre, uint32_t &bim)
{
    const int c = sx_lo(cs), s = sx_hi(cs), n = -s;
    const int c0 = ((c + 128) & 255) - 128, c1 = (c - c0) >> 8;
    const int s0 = ((s + 128) & 255) - 128, s1 = (s - s0) >> 8;
    const int n0 = ((n + 128) & 255) - 128, n1 = (n - n0) >> 8;
    bre = (uint32_t)(c0 & 255) | ((uint32_t)(n0 & 255) << 8) | ((uint32_t)(c1 & 255) << 16) | ((uint32_t)(n1 & 255) << 24);
    bim = (uint32_t)(s0 & 255) | ((uint32_t)(c0 & 255) << 8) | ((uint32_t)(s1 & 255) << 16) | ((uint32_t)(c1 & 255) << 24);
}

// phase of sample n of a block that started at phase phi0: (phi0 + n * freq) mod N
struct PhaseMod {
    unsigned n_table;
    unsigned mask;  // n_table - 1 when n_table is a power of two, else 0
    __device__ __forceinline__ unsigned operator()(unsigned v) const
    {
        return mask ? (v & mask) : (v % n_table);
    }
};

// counter-based synthetic sample (host twin: oracle/srcdsp_oracle.c:orc_hash32)
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t seed, uint32_t channel, uint64_t n)
{
    uint32_t x = seed ^ (channel * 0x9E3779B1u) ^ ((uint32_t)n * 0x85EBCA6Bu) ^
                 ((uint32_t)(n >> 32) * 0xC2B2AE35u);
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

}  // namespace srcdsp
