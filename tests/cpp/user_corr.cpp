// user_corr.cpp -- a user program written against the REFERENCE's correlators.h API:
// a burst with a known pattern (every S-th sample) buried in noise is searched block by block.  Compiles
// unchanged against the reference headers (+ dsp_complex.cpp) and against the drop-in headers; both builds
// must print the same lines.
#include <array>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <vector>

#include "dsp_complex.h"
#include "correlators.h"

typedef std::complex<int16_t> cs16;
typedef std::complex<int32_t> cs32;

static uint32_t hash32(uint32_t seed, uint64_t n)
{
    uint32_t x = seed ^ ((uint32_t)n * 0x85EBCA6Bu) ^ ((uint32_t)(n >> 32) * 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

int main()
{
    const size_t N = 32, S = 4;
    std::array<cs32, N> pattern;
    for (size_t k = 0; k < N; ++k) {
        const uint32_t h = hash32(77, k);
        pattern[k] = cs32((h & 1) ? 1500 : -1500, (h & 2) ? 1500 : -1500);
    }
    dsptl::FixedPatternCorrelator<int16_t, int32_t, N, S> corr;
    corr.setPattern(pattern, 0.8);
    std::vector<cs16> x(20000);
    for (size_t n = 0; n < x.size(); ++n) {
        const uint32_t h = hash32(0x5EEDC0, n);
        x[n] = cs16((int16_t)(h % 601) - 300, (int16_t)((h >> 16) % 601) - 300);
    }
    const size_t bursts[3] = {3000, 9111, 15020};
    for (int b = 0; b < 3; ++b)
        for (size_t k = 0; k < N; ++k)
            x[bursts[b] + k * S] += cs16((int16_t)(pattern[k].real() * 3 / 2), (int16_t)(pattern[k].imag() * 3 / 2));
    size_t pos = 0;
    const size_t block = 2500;
    while (pos < x.size()) {
        std::vector<cs16> in(x.begin() + pos, x.begin() + std::min(pos + block, x.size()));
        int idx = -1;
        const bool found = corr.step(in, idx);
        auto st = corr.getStatus();
        std::printf("block @%zu found %d idx %d corr %u %u %u energy %u %u %u\n", pos, (int)found, found ? idx : -1, st.corrValue[0],
                    st.corrValue[1], st.corrValue[2], st.energyValue[0], st.energyValue[1], st.energyValue[2]);
        if (found) {
            std::vector<cs16> bits = corr.getRefBitSamples();
            uint32_t sum = 0;
            for (size_t k = 0; k < bits.size(); ++k) sum = sum * 31u + (uint16_t)bits[k].real() * 7u + (uint16_t)bits[k].imag();
            std::printf("  peak at absolute sample %zu, bit samples %zu sum %08x\n", pos + idx, bits.size(), sum);
            pos += idx + 2;  // resume right behind the sample the correlator stopped at
        } else {
            pos += in.size();
        }
    }
    auto st = corr.getStatus();
    std::printf("coeffsEnergy %u coeffScaling %d\n", st.coeffsEnergy, st.coeffScaling);
    return 0;
}
