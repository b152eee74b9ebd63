// user_float.cpp -- a user program written against the REFERENCE's class API with FLOAT samples:
//   dsptl::FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M>
// (the one float instantiation of the hot path the reference headers can build; BASELINE configs[0] shape:
// decimate-by-8, 63 taps, streaming blocks).  It compiles unchanged against the reference headers
// (-I/root/reference + dsp_complex.cpp) and against the drop-in headers (-Iinclude/srcdsp + libsrcdsp_b200.so);
// both builds must print the same lines -- checksums over the raw float bits, so "the same" means bit-exact.
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dnsampling_filters.h"   // obsolete twin: accepts 63 taps
#include "filters.h"              // FilterFir (global namespace)

typedef std::complex<float> cf32;

static uint32_t hash32(uint32_t seed, uint32_t ch, uint64_t n)
{
    uint32_t x = seed ^ (ch * 0x9E3779B1u) ^ ((uint32_t)n * 0x85EBCA6Bu) ^ ((uint32_t)(n >> 32) * 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

static std::vector<float> lowpass(int ntaps, int ratio, double gain)
{
    std::vector<float> c(ntaps);
    const double pi = 3.14159265358979323846;
    double sum = 0;
    std::vector<double> h(ntaps);
    for (int k = 0; k < ntaps; ++k) {
        const double n = k - (ntaps - 1) / 2.0, t = n / ratio;
        const double s = (std::fabs(t) < 1e-12) ? 1.0 : std::sin(pi * t) / (pi * t);
        h[k] = s * (0.54 - 0.46 * std::cos(2 * pi * k / (ntaps - 1)));
        sum += h[k];
    }
    for (int k = 0; k < ntaps; ++k) c[k] = (float)(gain * h[k] / sum);
    return c;
}

static uint64_t checksum(const std::vector<cf32> &v)
{
    uint64_t a = 1469598103934665603ull;
    for (size_t i = 0; i < v.size(); ++i) {
        uint32_t b[2];
        std::memcpy(b, &v[i], 8);
        a = (a ^ b[0]) * 1099511628211ull;
        a = (a ^ b[1]) * 1099511628211ull;
    }
    return a;
}

template <unsigned M>
static void run(const char *name, int ntaps, double gain, int left_shift, float scale)
{
    dsptl::FilterDnsamplingFir<cf32, cf32, cf32, float, M> f(lowpass(ntaps, M, gain));
    if (left_shift) f.setLeftShiftBy2(left_shift);
    const size_t blocks[3] = {8192 * M / 8, 4096 * M / 8, 12288 * M / 8};
    uint64_t n0 = 0;
    for (int b = 0; b < 3; ++b) {
        std::vector<cf32> in(blocks[b]), out(blocks[b] / M);
        for (size_t i = 0; i < in.size(); ++i) {
            const uint32_t h = hash32(0x5EED0F10u, M, n0 + i);
            in[i] = cf32(scale * (float)(int16_t)(h & 0xFFFF), scale * (float)(int16_t)(h >> 16));
        }
        n0 += in.size();
        f.step(in, out);
        std::printf("%s block %d: %zu -> %zu  out[0]=(%.1f,%.1f) out[last]=(%.1f,%.1f) checksum %016llx\n", name, b, in.size(),
                    out.size(), out[0].real(), out[0].imag(), out.back().real(), out.back().imag(),
                    (unsigned long long)checksum(out));
    }
    f.reset();
    std::vector<cf32> in(2048, cf32(1000.25f, -500.5f)), out(2048 / M);  // >= ntaps - 1 samples (the reference needs that)
    f.step(in, out);
    std::printf("%s after reset, DC: out[last]=(%.1f,%.1f)\n", name, out.back().real(), out.back().imag());
}

// FilterFir<complex<float>, ...> (filters.h): the non-decimating FIR with float types, streaming blocks + copy
static void run_fir()
{
    FilterFir<cf32, cf32, cf32, float> f(lowpass(33, 4, 311.7));
    uint64_t n0 = 0;
    const size_t blocks[3] = {1000, 3096, 17};
    for (int b = 0; b < 3; ++b) {
        std::vector<cf32> in(blocks[b]), out(blocks[b]);
        for (size_t i = 0; i < in.size(); ++i) {
            const uint32_t h = hash32(0x5EED0F11u, 1, n0 + i);
            in[i] = cf32(0.25f * (float)(int16_t)(h & 0xFFFF), 0.25f * (float)(int16_t)(h >> 16));
        }
        n0 += in.size();
        f.step(in, out);
        std::printf("FilterFir float block %d: %zu  out[0]=(%.1f,%.1f) out[last]=(%.1f,%.1f) checksum %016llx\n", b, out.size(),
                    out[0].real(), out[0].imag(), out.back().real(), out.back().imag(), (unsigned long long)checksum(out));
    }
    FilterFir<cf32, cf32, cf32, float> g(f);  // copies taps and history
    std::vector<cf32> in(64, cf32(100.5f, -7.25f)), o1(64), o2(64);
    f.step(in, o1);
    g.step(in, o2);
    std::printf("FilterFir float copy: %016llx %016llx\n", (unsigned long long)checksum(o1), (unsigned long long)checksum(o2));
}

int main()
{
    run_fir();
    run<8>("unity /8 63 taps", 63, 1.0, 0, 0.37f);        // all |c| < 1: the reference applies no shift
    run<16>("gain-40000 /16 255 taps", 255, 40000.0, 0, 0.5f);  // integer parts: coeffScaling = 15
    run<4>("gain-97.3 /4 1023 taps ls1", 1023, 97.3, 1, 1.0f);
    return 0;
}
