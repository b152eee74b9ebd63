/*
 * srcdsp/buffers.h -- drop-in for the reference's buffers.h: dsptl::FifoWithTimeTrack<T, N>
 * (buffers.h:58-459), the single-writer / single-reader ring with sample time stamps that sits
 * in front of the DSP chain.  Same class name, template parameters and member signatures; the
 * ring lives in PINNED host memory (libsrcdsp_b200.so, srcdsp_fifo_*), so that the GPU filter
 * banks can take a block straight out of it:
 *
 *     dsptl::FifoWithTimeTrack<std::complex<int16_t>, 1 << 24> fifo(fs);
 *     fifo.write(block, seconds, frac);                       // producer thread
 *     fifo.readSegments(n, start, seg);                       // consumer: zero-copy view
 *     dec.step(seg[0].ptr, seg[0].n, out) ...                 // pinned -> device DMA
 *
 * T must be trivially copyable (the reference std::copy's elements; every use in the reference
 * is an arithmetic or std::complex type).
 */
#ifndef SRCDSP_DROPIN_BUFFERS_H
#define SRCDSP_DROPIN_BUFFERS_H

#include <cassert>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <utility>

#include "detail.h"

namespace dsptl {

template <class T, size_t N>
class FifoWithTimeTrack
{
    static_assert(std::is_trivially_copyable<T>::value, "FifoWithTimeTrack<T>: T must be trivially copyable");

public:
    /* buffers.h:62-65 */
    FifoWithTimeTrack(double samplingFrequencyArg = 0)
    {
        srcdsp_dropin::check(srcdsp_fifo_create(&h_, sizeof(T), N, samplingFrequencyArg), "FifoWithTimeTrack()");
    }
    ~FifoWithTimeTrack() { srcdsp_fifo_destroy(h_); }
    FifoWithTimeTrack(const FifoWithTimeTrack &) = delete;
    FifoWithTimeTrack &operator=(const FifoWithTimeTrack &) = delete;

    /* buffers.h:139-217 */
    void write(std::vector<T> &in, unsigned int seconds = 0, double fracSeconds = 0)
    {
        srcdsp_dropin::check(srcdsp_fifo_write(h_, in.data(), in.size(), seconds, fracSeconds), "FifoWithTimeTrack::write");
    }
    /* buffers.h:284-352: true = the requested range is not (yet) available */
    bool read(std::vector<T> &out, uint64_t &start)
    {
        int err = 0;
        srcdsp_dropin::check(srcdsp_fifo_read(h_, out.data(), out.size(), &start, &err), "FifoWithTimeTrack::read");
        return err != 0;
    }
    /* buffers.h:361-377 */
    size_t count()
    {
        size_t c = 0;
        srcdsp_dropin::check(srcdsp_fifo_count(h_, &c), "FifoWithTimeTrack::count");
        return c;
    }
    /* buffers.h:262-276 */
    void reset() { srcdsp_dropin::check(srcdsp_fifo_reset(h_), "FifoWithTimeTrack::reset"); }
    /* buffers.h:227-251 */
    void dumpInfo(bool dumpData = false)
    {
        size_t wp = 0;
        uint64_t ts = 0, te = 0;
        int ro = 0;
        srcdsp_dropin::check(srcdsp_fifo_get_state(h_, &wp, &ts, &te, &ro), "FifoWithTimeTrack::dumpInfo");
        std::cout << "writePtr: " << wp << '\n';
        std::cout << "timeStart : " << ts << '\n';
        std::cout << "timeEnd : " << te << '\n';
        std::cout << "rolloverFlag : " << (ro != 0) << '\n';
        if (dumpData) {
            const T *s = static_cast<const T *>(srcdsp_fifo_storage(h_));
            for (size_t index = 0; index < N; ++index) std::cout << "index: " << index << " Value: " << s[index] << '\n';
        }
    }
    /* buffers.h:396-459 */
    std::pair<unsigned int, double> getAbsoluteTime(uint64_t timePoint, double fracTimePoint)
    {
        unsigned s = 0;
        double f = 0;
        srcdsp_dropin::check(srcdsp_fifo_get_absolute_time(h_, timePoint, fracTimePoint, &s, &f), "FifoWithTimeTrack::getAbsoluteTime");
        return std::make_pair(s, f);
    }

    /* ---- additions for the GPU path (not in the reference) ---- */
    struct Segment {
        const T *ptr;
        size_t n;
    };
    /* read() without the copy: the (at most two) contiguous pieces of the pinned ring that hold
     * [start, start + n).  Same return value and start adjustment as read(). */
    bool readSegments(size_t n, uint64_t &start, Segment seg[2])
    {
        int err = 0;
        const void *p0 = nullptr, *p1 = nullptr;
        size_t n0 = 0, n1 = 0;
        srcdsp_dropin::check(srcdsp_fifo_segments(h_, n, &start, &p0, &n0, &p1, &n1, &err), "FifoWithTimeTrack::readSegments");
        seg[0].ptr = static_cast<const T *>(p0), seg[0].n = n0;
        seg[1].ptr = static_cast<const T *>(p1), seg[1].n = n1;
        return err != 0;
    }
    /* Absolute time of output sample j of a decimate-by-M stage fed with the block read at `start`:
     * output j is aligned with input time point start + j * M. */
    std::pair<unsigned int, double> getAbsoluteTimeOfOutput(uint64_t start, uint64_t j, unsigned M)
    {
        return getAbsoluteTime(start + j * M, 0.0);
    }
    bool pinned() const { return srcdsp_fifo_is_pinned(h_) != 0; }
    srcdsp_fifo_t handle() const { return h_; }

private:
    srcdsp_fifo_t h_ = nullptr;
};

}  // namespace dsptl

#endif
