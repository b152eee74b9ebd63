// user_stream.cpp -- a user program written against the REFERENCE's buffers.h / dsptl_files.h API
// (SURVEY.md 8(f) #2, #3): a producer writes time-stamped blocks into a FifoWithTimeTrack, a consumer
// reads ranges back by time point, the capture is saved and re-read as a raw I/Q file.  It compiles
// unchanged against the reference headers and against the drop-in headers; both builds must print the
// same lines.
#include <complex>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <vector>

#include "buffers.h"
#include "dsptl_files.h"

typedef std::complex<int16_t> cs16;

static uint32_t hash32(uint32_t seed, uint64_t n)
{
    uint32_t x = seed ^ ((uint32_t)n * 0x85EBCA6Bu) ^ ((uint32_t)(n >> 32) * 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
static uint32_t checksum(const std::vector<cs16> &v)
{
    uint32_t s = 0;
    for (size_t k = 0; k < v.size(); ++k) s = s * 31u + (uint16_t)v[k].real() * 7u + (uint16_t)v[k].imag();
    return s;
}

int main(int argc, char **argv)
{
    const char *path = argc > 1 ? argv[1] : "/tmp/srcdsp_user_stream.iq";
    dsptl::FifoWithTimeTrack<cs16, 1024> fifo(48000.0);
    std::printf("count0 %zu\n", fifo.count());
    uint64_t n = 0;
    std::vector<cs16> capture;
    for (int blk = 0; blk < 9; ++blk) {
        std::vector<cs16> in(100 + 97 * blk);
        for (size_t k = 0; k < in.size(); ++k, ++n) {
            const uint32_t h = hash32(0x5EED0F1F, n);
            in[k] = cs16((int16_t)(h & 0xFFFF), (int16_t)(h >> 16));
        }
        fifo.write(in, 1000 + blk, 0.125 * blk);
        capture.insert(capture.end(), in.begin(), in.end());
        std::printf("blk %d count %zu\n", blk, fifo.count());
        std::vector<cs16> out(64 + 50 * blk);
        uint64_t start = n > 700 ? n - 700 : 1;
        const bool err = fifo.read(out, start);
        std::printf("  read %zu @%llu -> err %d sum %08x\n", out.size(), (unsigned long long)start, (int)err, err ? 0u : checksum(out));
        std::pair<unsigned, double> t = fifo.getAbsoluteTime(start + 10, 0.25);
        std::printf("  time %u %.9f\n", t.first, t.second);
        uint64_t late = n + 5;
        std::vector<cs16> none(8);
        std::printf("  beyond -> err %d\n", (int)fifo.read(none, late));
    }
    fifo.reset();
    std::printf("after reset count %zu\n", fifo.count());
    {
        std::ofstream os(path, std::ios::binary);
        dsptl::saveBinarySamples(capture, os);
    }
    std::vector<cs16> back(2, cs16(1, 2));
    {
        std::ifstream is(path, std::ios::binary);
        dsptl::readBinarySamples(is, back);
    }
    std::printf("saved %zu read %zu first (%d,%d) sum %08x\n", capture.size(), back.size(), back[0].real(), back[0].imag(), checksum(back));
    return 0;
}
