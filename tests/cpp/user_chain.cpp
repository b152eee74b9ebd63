// user_chain.cpp -- a user program written against the REFERENCE's class API:
//   mixer.step(in, mixed); dec1.step(mixed, d1); dec2.step(d1, d2);  up.step(d2, back [, flush]);
// It compiles unchanged against the reference headers (-I/root/reference + dsp_complex.cpp) and
// against the drop-in headers (-Iinclude/srcdsp + libsrcdsp_b200.so); both builds must print the
// same lines.  The input is the counter-based synthetic baseband (same hash as the oracle).
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "mixers.h"
#include "dnsampling_filters.h"   // obsolete twin: accepts 63 taps (SURVEY.md section 0 trap (i))
#include "upsampling_filters.h"
#include "filters.h"

typedef std::complex<int16_t> cs16;
typedef std::complex<int32_t> cs32;

static uint32_t hash32(uint32_t seed, uint32_t ch, uint64_t n)
{
    uint32_t x = seed ^ (ch * 0x9E3779B1u) ^ ((uint32_t)n * 0x85EBCA6Bu) ^ ((uint32_t)(n >> 32) * 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

static std::vector<int32_t> lowpass(int ntaps, int ratio, double gain)
{
    std::vector<int32_t> c(ntaps);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < ntaps; ++k) {
        const double n = k - (ntaps - 1) / 2.0, t = n / ratio;
        const double s = (std::fabs(t) < 1e-12) ? 1.0 : std::sin(pi * t) / (pi * t);
        const double w = 0.54 - 0.46 * std::cos(2 * pi * k / (ntaps - 1));
        c[k] = (int32_t)std::lround(gain * s * w);
    }
    return c;
}

static uint64_t checksum(const std::vector<cs16> &v)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < v.size(); ++i) {
        h = (h ^ (uint16_t)v[i].real()) * 1099511628211ull;
        h = (h ^ (uint16_t)v[i].imag()) * 1099511628211ull;
    }
    return h;
}

int main()
{
    const size_t N = 32 * 2048;  // per block
    dsptl::Mixer<cs16, cs16, int16_t, 4096> mixer;
    mixer.setFrequency(-0.3217f);
    dsptl::FilterDnsamplingFir<cs16, cs16, cs32, int32_t, 8> dec1(lowpass(63, 8, 6000.0));
    dsptl::FilterDnsamplingFir<cs16, cs16, cs32, int32_t, 4> dec2(lowpass(63, 4, 11000.0));
    dec2.setLeftShiftBy2(1);
    std::vector<int32_t> ut = lowpass(64, 8, 4096.0);
    ut[63] = 0;
    dsptl::FilterUpsamplingFir<cs16, cs16, cs32, int32_t, 8> up(ut);
    std::printf("up length %d impLength %d ratio %d\n", up.getLength(), up.getImpLength(), up.getUpsamplingRatio());

    for (int blk = 0; blk < 3; ++blk) {
        std::vector<cs16> in(N), mixed(N), d1(N / 8), d2(N / 32);
        for (size_t n = 0; n < N; ++n) {
            const uint32_t h = hash32(0x5EED0001u, 0, blk * N + n);
            in[n] = cs16((int16_t)(h & 0xFFFF), (int16_t)(h >> 16));
        }
        if (blk == 1) mixer.adjustFrequency(0.0123f);
        mixer.step(in, mixed);
        dec1.step(mixed, d1);
        dec2.step(d1, d2);
        const bool flush = (blk == 2);
        std::vector<cs16> back(8 * (d2.size() + (flush ? up.getLength() / 8 : 0)));
        up.step(d2, back, flush);
        std::printf("block %d: mixed %016llx d1 %016llx d2 %016llx up %016llx (%zu)\n", blk,
                    (unsigned long long)checksum(mixed), (unsigned long long)checksum(d1),
                    (unsigned long long)checksum(d2), (unsigned long long)checksum(back), back.size());
    }
    // iterator overload (output shift 0) into the middle of a caller buffer, then reset()
    std::vector<cs16> small(256, cs16(3, -2)), dst(8 * 256 + 16, cs16(7, 7));
    up.reset();
    up.step(small, dst.begin() + 8);
    std::printf("iterator overload %016llx guard %d %d\n", (unsigned long long)checksum(dst), (int)dst[0].real(),
                (int)dst[dst.size() - 1].imag());
    // non-decimating FIR (filters.h), two blocks with carried history, then new taps (history cleared)
    FilterFir<cs16, cs16, cs32, int32_t> fir(lowpass(33, 4, 9000.0));
    std::vector<cs16> f1(1000), f2(1000);
    for (size_t n = 0; n < 2000; ++n) {
        const uint32_t h = hash32(0x5EED0007u, 0, n);
        (n < 1000 ? f1[n] : f2[n - 1000]) = cs16((int16_t)(h & 0xFFFF), (int16_t)(h >> 16));
    }
    std::vector<cs16> g1(1000), g2(1000);
    fir.step(f1, g1);
    fir.step(f2, g2);
    fir.setCoeffs(lowpass(33, 2, 20000.0));  // same length: the reference keeps `top`, a shorter filter would index out of bounds
    fir.step(f1, f1);  // in place
    std::printf("fir %016llx %016llx %016llx\n", (unsigned long long)checksum(g1), (unsigned long long)checksum(g2),
                (unsigned long long)checksum(f1));
    dec1.reset();
    mixer.reset(0.25f);
    std::vector<cs16> in2(64, cs16(1000, -500)), m2(64), o2(8);
    mixer.step(in2, m2);
    dec1.step(m2, o2);
    std::printf("after reset %016llx\n", (unsigned long long)checksum(o2));
    return 0;
}
