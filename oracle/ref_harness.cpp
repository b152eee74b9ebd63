/*
 * ref_harness.cpp -- the UNMODIFIED reference compiled as a checker.  TEST INFRASTRUCTURE ONLY.
 *
 * This file contains no reference code: it #includes the reference's own headers from the
 * read-only checkout (the Makefile passes -I$(REF), default /root/reference, and compiles
 * $(REF)/dsp_complex.cpp next to it) and wraps the three hot-path classes behind a small
 * extern "C" surface so that tests and bench.py can drive them through ctypes:
 *
 *   dsptl::Mixer<cs16,cs16,int16_t,N>                         mixers.h:130-188
 *   dsptl::FilterDnsamplingFir<cs16,cs16,cs32,int32_t,M>      dsptl_dnsampling_filters.h:43-220
 *   (obsolete twin, no taps%M assert)                         dnsampling_filters.h:43-172
 *   dsptl::FilterUpsamplingFir<cs16,cs16,cs32,int32_t,L>      upsampling_filters.h:35-323
 *   FilterFir<cs16,cs16,cs32,int32_t>                         filters.h:42-169
 *   dsptl::FifoWithTimeTrack<cs16,N>                          buffers.h:58-459
 *   dsptl::saveBinarySamples / readBinarySamples              dsptl_files.h:57-109, 250-262
 *   dsptl::FixedPatternCorrelator<int16_t,int32_t,N,S>        correlators.h:54-303
 *
 * The two decimator headers share an include guard and a class name, so the obsolete one is
 * included inside namespace ref_obsolete (its std/dsp_complex includes are already satisfied).
 * Output goes to oracle/_ref/ (git-ignored, but it travels to the GPU box).
 */
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "dsp_complex.h"
#include "constants.h"
#include "generators.h"

namespace ref_obsolete {
#include "dnsampling_filters.h"
}
#undef _DNSAMPLING_FILTER_FIR_H
#include "dsptl_dnsampling_filters.h"
#include "mixers.h"
#include "upsampling_filters.h"
#include "filters.h"
#include "buffers.h"
#include "dsptl_files.h"
#include "correlators.h"

typedef std::complex<int16_t> cs16;
typedef std::complex<int32_t> cs32;

static_assert(sizeof(cs16) == 4, "interleaved I/Q layout");

namespace {

// ---------------------------------------------------------------------------------------------
struct MixerBase {
    virtual ~MixerBase() {}
    virtual void setFrequency(float) = 0;
    virtual void reset(float) = 0;
    virtual void adjustFrequency(float) = 0;
    virtual void step(std::vector<cs16> &in, std::vector<cs16> &out) = 0;
    virtual int phi() = 0;
    virtual int freq() = 0;
};
template <unsigned N>
struct MixerT : MixerBase, dsptl::Mixer<cs16, cs16, int16_t, N> {
    typedef dsptl::Mixer<cs16, cs16, int16_t, N> B;
    void setFrequency(float f) { B::setFrequency(f); }
    void reset(float f) { B::reset(f); }
    void adjustFrequency(float f) { B::adjustFrequency(f); }
    void step(std::vector<cs16> &in, std::vector<cs16> &out) { B::step(in, out); }
    int phi() { return B::phi; }     // protected members, visible to the derived wrapper
    int freq() { return B::freq; }
};

struct DecBase {
    virtual ~DecBase() {}
    virtual void step(const std::vector<cs16> &in, std::vector<cs16> &out) = 0;
    virtual void reset() = 0;
    virtual void setLeftShiftBy2(int) = 0;
    virtual int setCoeffs(const std::vector<int32_t> &) { return -1; }  // the obsolete header has no setCoeffs()
};
template <class F>
struct DecT : DecBase {
    F f;
    explicit DecT(const std::vector<int32_t> &taps) : f(taps) {}
    void step(const std::vector<cs16> &in, std::vector<cs16> &out) { f.step(in, out); }
    void reset() { f.reset(); }
    void setLeftShiftBy2(int s) { f.setLeftShiftBy2(s); }
};
// dsptl_dnsampling_filters.h only: setCoeffs() on a live object (history.resize, :114-134)
template <class F>
struct DecTNew : DecT<F> {
    explicit DecTNew(const std::vector<int32_t> &taps) : DecT<F>(taps) {}
    int setCoeffs(const std::vector<int32_t> &t)
    {
        this->f.setCoeffs(t);
        return 0;
    }
};

struct UpBase {
    virtual ~UpBase() {}
    virtual void setCoefficients(const std::vector<int32_t> &) = 0;
    virtual void step(const std::vector<cs16> &in, std::vector<cs16> &out, bool flush) = 0;
    virtual void stepIt(const std::vector<cs16> &in, std::vector<cs16>::iterator out, bool flush) = 0;
    virtual void reset() = 0;
    virtual int getLength() = 0;
    virtual int getImpLength() = 0;
    virtual int getUpsamplingRatio() = 0;
};
template <unsigned L>
struct UpT : UpBase {
    dsptl::FilterUpsamplingFir<cs16, cs16, cs32, int32_t, L> f;
    explicit UpT(const std::vector<int32_t> &taps) : f(taps) {}
    void setCoefficients(const std::vector<int32_t> &t) { f.setCoefficients(t); }
    void step(const std::vector<cs16> &in, std::vector<cs16> &out, bool flush) { f.step(in, out, flush); }
    void stepIt(const std::vector<cs16> &in, std::vector<cs16>::iterator out, bool flush) { f.step(in, out, flush); }
    void reset() { f.reset(); }
    int getLength() { return f.getLength(); }
    int getImpLength() { return f.getImpLength(); }
    int getUpsamplingRatio() { return f.getUpsamplingRatio(); }
};

#define RATIOS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(10) X(12) X(16) X(20) X(24) X(32) X(64)

DecBase *make_dec(int variant, int M, const std::vector<int32_t> &taps)
{
    switch (M) {
#define X(m)                                                                                     \
    case m:                                                                                      \
        if (variant == 0)                                                                        \
            return new DecT<ref_obsolete::dsptl::FilterDnsamplingFir<cs16, cs16, cs32, int32_t, m> >(taps); \
        return new DecTNew<dsptl::FilterDnsamplingFir<cs16, cs16, cs32, int32_t, m> >(taps);
        RATIOS(X)
#undef X
    }
    return nullptr;
}

UpBase *make_up(int L, const std::vector<int32_t> &taps)
{
    switch (L) {
#define X(l) \
    case l:  \
        return new UpT<l>(taps);
        RATIOS(X)
#undef X
    }
    return nullptr;
}

MixerBase *make_mixer(unsigned N)
{
    switch (N) {
    case 256: return new MixerT<256>();
    case 1024: return new MixerT<1024>();
    case 4096: return new MixerT<4096>();
    case 8192: return new MixerT<8192>();
    }
    return nullptr;
}

inline std::vector<cs16> to_vec(const int16_t *iq, size_t n)
{
    std::vector<cs16> v(n);
    if (n) std::memcpy(v.data(), iq, n * sizeof(cs16));
    return v;
}

}  // namespace

// float decimator: the reference's own template with float types
typedef std::complex<float> cf32;
struct DecfBase {
    virtual ~DecfBase() {}
    virtual void step(const std::vector<cf32> &in, std::vector<cf32> &out) = 0;
    virtual void reset() = 0;
    virtual void setLeftShiftBy2(int) = 0;
};
template <class F>
struct DecfT : DecfBase {
    F f;
    explicit DecfT(const std::vector<float> &taps) : f(taps) {}
    void step(const std::vector<cf32> &in, std::vector<cf32> &out) { f.step(in, out); }
    void reset() { f.reset(); }
    void setLeftShiftBy2(int s) { f.setLeftShiftBy2(s); }
};
static DecfBase *make_decf(int variant, int M, const std::vector<float> &taps)
{
    switch (M) {
#define X(m)                                                                                                   \
    case m:                                                                                                    \
        if (variant == 0) return new DecfT<ref_obsolete::dsptl::FilterDnsamplingFir<cf32, cf32, cf32, float, m> >(taps); \
        return new DecfT<dsptl::FilterDnsamplingFir<cf32, cf32, cf32, float, m> >(taps);
        RATIOS(X)
#undef X
    }
    return nullptr;
}

extern "C" {

// --- mixer -----------------------------------------------------------------------------------
void *ref_mixer_create(unsigned n_table) { return make_mixer(n_table); }
void ref_mixer_destroy(void *h) { delete static_cast<MixerBase *>(h); }
int ref_mixer_set_frequency(void *h, float f)
{
    if (!(f <= 1 && f >= -1)) return -1;  // mixers.h:54 would assert
    static_cast<MixerBase *>(h)->setFrequency(f);
    return 0;
}
void ref_mixer_reset(void *h, float f) { static_cast<MixerBase *>(h)->reset(f); }
void ref_mixer_adjust_frequency(void *h, float f) { static_cast<MixerBase *>(h)->adjustFrequency(f); }
int ref_mixer_phi(void *h) { return static_cast<MixerBase *>(h)->phi(); }
int ref_mixer_freq(void *h) { return static_cast<MixerBase *>(h)->freq(); }
void ref_mixer_step(void *h, const int16_t *in_iq, int16_t *out_iq, size_t n)
{
    std::vector<cs16> in = to_vec(in_iq, n), out(n);
    static_cast<MixerBase *>(h)->step(in, out);
    if (n) std::memcpy(out_iq, out.data(), n * sizeof(cs16));
}

// --- decimator (variant 0 = dnsampling_filters.h, 1 = dsptl_dnsampling_filters.h) ---------------
void *ref_dec_create(int variant, int M, const int32_t *taps, int ntaps)
{
    if (ntaps < 1) return nullptr;
    if (variant != 0 && ntaps % M != 0) return nullptr;  // dsptl_dnsampling_filters.h:122 would assert
    return make_dec(variant, M, std::vector<int32_t>(taps, taps + ntaps));
}
void ref_dec_destroy(void *h) { delete static_cast<DecBase *>(h); }
void ref_dec_reset(void *h) { static_cast<DecBase *>(h)->reset(); }
void ref_dec_set_left_shift(void *h, int s) { static_cast<DecBase *>(h)->setLeftShiftBy2(s); }
// setCoeffs() on a live object; -1: the object is the obsolete header's (no such member) or the call would assert
int ref_dec_set_coeffs(void *h, int M, const int32_t *taps, int ntaps)
{
    if (ntaps < 1 || ntaps % M != 0) return -1;  // dsptl_dnsampling_filters.h:122 would assert
    return static_cast<DecBase *>(h)->setCoeffs(std::vector<int32_t>(taps, taps + ntaps));
}
void ref_dec_step(void *h, const int16_t *in_iq, size_t n_in, int M, int16_t *out_iq)
{
    std::vector<cs16> in = to_vec(in_iq, n_in), out(n_in / M);
    static_cast<DecBase *>(h)->step(in, out);
    if (!out.empty()) std::memcpy(out_iq, out.data(), out.size() * sizeof(cs16));
}

// --- float decimator: dsptl::FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M> ------
void *ref_decf_create(int variant, int M, const float *taps, int ntaps)
{
    if (ntaps < 1) return nullptr;
    if (variant != 0 && ntaps % M != 0) return nullptr;  // dsptl_dnsampling_filters.h:122 would assert
    return make_decf(variant, M, std::vector<float>(taps, taps + ntaps));
}
void ref_decf_destroy(void *h) { delete static_cast<DecfBase *>(h); }
void ref_decf_reset(void *h) { static_cast<DecfBase *>(h)->reset(); }
void ref_decf_set_left_shift(void *h, int s) { static_cast<DecfBase *>(h)->setLeftShiftBy2(s); }
void ref_decf_step(void *h, const float *in_iq, size_t n_in, int M, float *out_iq)
{
    std::vector<cf32> in(n_in), out(n_in / M);
    if (n_in) std::memcpy(in.data(), in_iq, n_in * sizeof(cf32));
    static_cast<DecfBase *>(h)->step(in, out);
    if (!out.empty()) std::memcpy(out_iq, out.data(), out.size() * sizeof(cf32));
}

// --- float plain FIR: FilterFir<complex<float>, complex<float>, complex<float>, float> (filters.h) ------------------
void *ref_firf_create(const float *taps, int ntaps)
{
    return new FilterFir<cf32, cf32, cf32, float>(std::vector<float>(taps, taps + ntaps));
}
void ref_firf_destroy(void *h) { delete static_cast<FilterFir<cf32, cf32, cf32, float> *>(h); }
void ref_firf_reset(void *h) { static_cast<FilterFir<cf32, cf32, cf32, float> *>(h)->reset(); }
void ref_firf_step(void *h, const float *in_iq, size_t n, float *out_iq)
{
    std::vector<cf32> in(n), out(n);
    if (n) std::memcpy(in.data(), in_iq, n * sizeof(cf32));
    static_cast<FilterFir<cf32, cf32, cf32, float> *>(h)->step(in, out);
    if (n) std::memcpy(out_iq, out.data(), n * sizeof(cf32));
}

// --- plain FIR (filters.h) ----------------------------------------------------------------------
void *ref_fir_create(const int32_t *taps, int ntaps)
{
    return new FilterFir<cs16, cs16, cs32, int32_t>(std::vector<int32_t>(taps, taps + ntaps));
}
void ref_fir_destroy(void *h) { delete static_cast<FilterFir<cs16, cs16, cs32, int32_t> *>(h); }
void ref_fir_reset(void *h) { static_cast<FilterFir<cs16, cs16, cs32, int32_t> *>(h)->reset(); }
void ref_fir_step(void *h, const int16_t *in_iq, size_t n, int16_t *out_iq)
{
    std::vector<cs16> in = to_vec(in_iq, n), out(n);
    static_cast<FilterFir<cs16, cs16, cs32, int32_t> *>(h)->step(in, out);
    if (n) std::memcpy(out_iq, out.data(), n * sizeof(cs16));
}

// --- upsampler ----------------------------------------------------------------------------------
void *ref_up_create(int L, const int32_t *taps, int ntaps)
{
    if (ntaps < 1 || ntaps % L != 0) return nullptr;  // upsampling_filters.h:110,113 would assert
    return make_up(L, std::vector<int32_t>(taps, taps + ntaps));
}
void ref_up_destroy(void *h) { delete static_cast<UpBase *>(h); }
void ref_up_reset(void *h) { static_cast<UpBase *>(h)->reset(); }
int ref_up_set_coefficients(void *h, const int32_t *taps, int ntaps)
{
    UpBase *u = static_cast<UpBase *>(h);
    if (ntaps < 1 || ntaps % u->getUpsamplingRatio() != 0) return -1;
    u->setCoefficients(std::vector<int32_t>(taps, taps + ntaps));
    return 0;
}
int ref_up_get_length(void *h) { return static_cast<UpBase *>(h)->getLength(); }
int ref_up_get_imp_length(void *h) { return static_cast<UpBase *>(h)->getImpLength(); }
int ref_up_get_ratio(void *h) { return static_cast<UpBase *>(h)->getUpsamplingRatio(); }
// shift_mode 0: vector overload (shift 15 - round(log2 L)); 1: iterator overload (shift 0).
// out_iq must hold L*(n_in + (flush ? length/L : 0)) samples.
void ref_up_step(void *h, const int16_t *in_iq, size_t n_in, int16_t *out_iq, int flush, int shift_mode)
{
    UpBase *u = static_cast<UpBase *>(h);
    const size_t L = (size_t)u->getUpsamplingRatio();
    const size_t n_out = L * (n_in + (flush ? (size_t)u->getLength() / L : 0));
    std::vector<cs16> in = to_vec(in_iq, n_in), out(n_out);
    if (shift_mode == 0)
        u->step(in, out, flush != 0);
    else
        u->stepIt(in, out.begin(), flush != 0);
    if (n_out) std::memcpy(out_iq, out.data(), n_out * sizeof(cs16));
}

// --- CPU baseline: time the reference's own step() calls on a bank of independent channels -----
// One Mixer / decimator object set per channel (that is how the reference models a channel),
// channels dealt round-robin over n_threads std::threads, streaming blocks of block_len samples.
// kind: 0 = dec only, 1 = mix + dec, 2 = mix + dec + dec2, 3 = upsampler (M1 is L).
// in_iq: [C][n_per_ch] interleaved I/Q; out_iq may be NULL (timing only) or [C][n_out_per_ch].
// Returns wall seconds spent inside step() calls (vector marshalling excluded), < 0 on error.
double ref_bench_bank(int kind, int C, size_t n_per_ch, size_t block_len, int n_threads,
                      const int16_t *in_iq, int16_t *out_iq,
                      int M1, const int32_t *taps1, int ntaps1,
                      int M2, const int32_t *taps2, int ntaps2,
                      const float *lo_freq /* [C] or NULL */)
{
    if (n_threads < 1) n_threads = 1;
    if (block_len == 0 || n_per_ch % block_len != 0) return -1.0;
    const size_t ratio_num = (kind == 3) ? (size_t)M1 : 1;
    const size_t ratio_den = (kind == 3) ? 1 : (size_t)M1 * (kind == 2 ? (size_t)M2 : 1);
    if (kind != 3 && block_len % ratio_den != 0) return -1.0;
    const size_t n_out_per_ch = n_per_ch * ratio_num / ratio_den;

    struct Chan {
        MixerBase *mix = nullptr;
        DecBase *d1 = nullptr, *d2 = nullptr;
        UpBase *up = nullptr;
    };
    std::vector<Chan> ch((size_t)C);
    for (int c = 0; c < C; ++c) {
        if (kind == 1 || kind == 2) {
            ch[c].mix = make_mixer(4096);
            ch[c].mix->setFrequency(lo_freq ? lo_freq[c] : 0.f);
        }
        if (kind <= 2) {
            ch[c].d1 = make_dec(0, M1, std::vector<int32_t>(taps1, taps1 + ntaps1));
            if (!ch[c].d1) return -2.0;
        }
        if (kind == 2) {
            ch[c].d2 = make_dec(0, M2, std::vector<int32_t>(taps2, taps2 + ntaps2));
            if (!ch[c].d2) return -2.0;
        }
        if (kind == 3) {
            if (ntaps1 % M1) return -2.0;
            ch[c].up = make_up(M1, std::vector<int32_t>(taps1, taps1 + ntaps1));
            if (!ch[c].up) return -2.0;
        }
    }

    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) {
        th.emplace_back([&, t]() {
            std::vector<cs16> in(block_len), mixed(block_len), o1, o2;
            for (int c = t; c < C; c += n_threads) {
                for (size_t b0 = 0; b0 < n_per_ch; b0 += block_len) {
                    std::memcpy(in.data(), in_iq + 2 * ((size_t)c * n_per_ch + b0), block_len * sizeof(cs16));
                    const std::vector<cs16> *res = nullptr;
                    if (kind == 3) {
                        o1.resize(block_len * (size_t)M1);
                        ch[c].up->step(in, o1, false);
                        res = &o1;
                    } else {
                        std::vector<cs16> *src = &in;
                        if (ch[c].mix) {
                            ch[c].mix->step(in, mixed);
                            src = &mixed;
                        }
                        o1.resize(block_len / (size_t)M1);
                        ch[c].d1->step(*src, o1);
                        res = &o1;
                        if (ch[c].d2) {
                            o2.resize(o1.size() / (size_t)M2);
                            ch[c].d2->step(o1, o2);
                            res = &o2;
                        }
                    }
                    if (out_iq) {
                        const size_t o0 = b0 * ratio_num / ratio_den;
                        std::memcpy(out_iq + 2 * ((size_t)c * n_out_per_ch + o0), res->data(), res->size() * sizeof(cs16));
                    }
                }
            }
        });
    }
    for (auto &x : th) x.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    for (auto &c : ch) {
        delete c.mix;
        delete c.d1;
        delete c.d2;
        delete c.up;
    }
    return secs;
}

// ---------------------------------------------------------------------------------------------
// FifoWithTimeTrack<cs16, N> for the capacities the tests use (N is a template parameter)
extern "C++" {
// The reference keeps its time counters private and has no way to start them near 2^64, so its rollover branches
// (buffers.h:179-207) could not be exercised through the public API.  The reference source stays untouched: an explicit
// template instantiation may name private members ([temp.explicit]), and the friend function it defines hands the
// member pointer to this file.
template <class Tag, typename Tag::type Member>
struct FifoMemberOf {
    friend typename Tag::type fifo_member(Tag) { return Member; }
};
template <size_t N>
struct FifoPeek;
#define REF_FIFO_MEMBER(N, member, mtype)                                         \
    struct FifoTag_##member##_##N {                                               \
        typedef mtype dsptl::FifoWithTimeTrack<cs16, N>::*type;                   \
        friend type fifo_member(FifoTag_##member##_##N);                          \
    };                                                                            \
    template struct FifoMemberOf<FifoTag_##member##_##N, &dsptl::FifoWithTimeTrack<cs16, N>::member>;
#define REF_FIFO_PEEK(N)                                                                          \
    REF_FIFO_MEMBER(N, writePtr, size_t)                                                          \
    REF_FIFO_MEMBER(N, timeStart, uint64_t)                                                       \
    REF_FIFO_MEMBER(N, timeEnd, uint64_t)                                                         \
    REF_FIFO_MEMBER(N, rolloverFlag, bool)                                                        \
    template <>                                                                                   \
    struct FifoPeek<N> {                                                                          \
        typedef dsptl::FifoWithTimeTrack<cs16, N> F;                                              \
        static size_t &write_ptr(F &f) { return f.*fifo_member(FifoTag_writePtr_##N()); }         \
        static uint64_t &time_start(F &f) { return f.*fifo_member(FifoTag_timeStart_##N()); }     \
        static uint64_t &time_end(F &f) { return f.*fifo_member(FifoTag_timeEnd_##N()); }         \
        static bool &rollover(F &f) { return f.*fifo_member(FifoTag_rolloverFlag_##N()); }        \
    };
REF_FIFO_PEEK(16)
REF_FIFO_PEEK(100)
REF_FIFO_PEEK(1024)
REF_FIFO_PEEK(65536)

struct FifoBase {
    virtual ~FifoBase() {}
    virtual void write(std::vector<cs16> &in, unsigned s, double f) = 0;
    virtual bool read(std::vector<cs16> &out, uint64_t &start) = 0;
    virtual size_t count() = 0;
    virtual void reset() = 0;
    virtual std::pair<unsigned, double> abs_time(uint64_t tp, double f) = 0;
    virtual void set_time(uint64_t start, uint64_t end) = 0;
    virtual void state(size_t *wp, uint64_t *start, uint64_t *end, int *rollover) = 0;
};
template <size_t N>
struct FifoImpl : FifoBase {
    dsptl::FifoWithTimeTrack<cs16, N> f;
    explicit FifoImpl(double fs) : f(fs) {}
    void write(std::vector<cs16> &in, unsigned s, double fr) override { f.write(in, s, fr); }
    bool read(std::vector<cs16> &out, uint64_t &start) override { return f.read(out, start); }
    size_t count() override { return f.count(); }
    void reset() override { f.reset(); }
    std::pair<unsigned, double> abs_time(uint64_t tp, double fr) override { return f.getAbsoluteTime(tp, fr); }
    void set_time(uint64_t start, uint64_t end) override
    {
        FifoPeek<N>::time_start(f) = start;
        FifoPeek<N>::time_end(f) = end;
    }
    void state(size_t *wp, uint64_t *start, uint64_t *end, int *rollover) override
    {
        *wp = FifoPeek<N>::write_ptr(f);
        *start = FifoPeek<N>::time_start(f);
        *end = FifoPeek<N>::time_end(f);
        *rollover = FifoPeek<N>::rollover(f) ? 1 : 0;
    }
};
}  // extern "C++"

void *ref_fifo_create(size_t capacity, double fs)
{
    switch (capacity) {
    case 16: return new FifoImpl<16>(fs);
    case 100: return new FifoImpl<100>(fs);
    case 1024: return new FifoImpl<1024>(fs);
    case 65536: return new FifoImpl<65536>(fs);
    default: return nullptr;
    }
}
void ref_fifo_destroy(void *h) { delete static_cast<FifoBase *>(h); }
void ref_fifo_write(void *h, const int16_t *iq, size_t n, unsigned s, double f)
{
    std::vector<cs16> v(n);
    if (n) std::memcpy(v.data(), iq, n * sizeof(cs16));
    static_cast<FifoBase *>(h)->write(v, s, f);
}
int ref_fifo_read(void *h, int16_t *iq, size_t n, uint64_t *start)
{
    std::vector<cs16> v(n);
    const bool err = static_cast<FifoBase *>(h)->read(v, *start);
    if (!err) std::memcpy(iq, v.data(), n * sizeof(cs16));
    return err ? 1 : 0;
}
size_t ref_fifo_count(void *h) { return static_cast<FifoBase *>(h)->count(); }
void ref_fifo_reset(void *h) { static_cast<FifoBase *>(h)->reset(); }
void ref_fifo_set_time(void *h, uint64_t start, uint64_t end) { static_cast<FifoBase *>(h)->set_time(start, end); }
void ref_fifo_state(void *h, size_t *wp, uint64_t *start, uint64_t *end, int *rollover)
{
    static_cast<FifoBase *>(h)->state(wp, start, end, rollover);
}
void ref_fifo_abs_time(void *h, uint64_t tp, double f, unsigned *s, double *fr)
{
    auto r = static_cast<FifoBase *>(h)->abs_time(tp, f);
    *s = r.first;
    *fr = r.second;
}

// ---------------------------------------------------------------------------------------------
// binary sample files (dsptl_files.h)
void ref_save_binary(const char *path, const int16_t *iq, size_t n)
{
    std::vector<cs16> v(n);
    if (n) std::memcpy(v.data(), iq, n * sizeof(cs16));
    std::ofstream os(path, std::ios::binary);
    dsptl::saveBinarySamples(v, os);
}
// returns the number of elements readBinarySamples produced when appending to `prefill` existing elements
size_t ref_read_binary(const char *path, int16_t *iq, size_t cap, size_t prefill)
{
    std::vector<cs16> v(prefill, cs16(7, -7));
    std::ifstream is(path, std::ios::binary);
    dsptl::readBinarySamples(is, v);
    const size_t n = v.size() < cap ? v.size() : cap;
    if (n) std::memcpy(iq, v.data(), n * sizeof(cs16));
    return v.size();
}

// ---------------------------------------------------------------------------------------------
// FixedPatternCorrelator<int16_t, int32_t, N, S>
extern "C++" {
struct CorrBase {
    virtual ~CorrBase() {}
    virtual void set_pattern(const int32_t *iq, double thr) = 0;
    virtual bool step(const std::vector<cs16> &in, int &idx) = 0;
    virtual void reset() = 0;
    virtual std::vector<cs16> bits() = 0;
    virtual void status(uint32_t *energy3, uint32_t *corr3, uint32_t *coeffs_energy, int *coeff_scaling) = 0;
};
template <size_t N, size_t S>
struct CorrImpl : CorrBase {
    dsptl::FixedPatternCorrelator<int16_t, int32_t, N, S> c;
    void set_pattern(const int32_t *iq, double thr) override
    {
        std::array<cs32, N> p;
        for (size_t k = 0; k < N; ++k) p[k] = cs32(iq[2 * k], iq[2 * k + 1]);
        c.setPattern(p, thr);
    }
    bool step(const std::vector<cs16> &in, int &idx) override { return c.step(in, idx); }
    void reset() override { c.reset(); }
    std::vector<cs16> bits() override { return c.getRefBitSamples(); }
    void status(uint32_t *e3, uint32_t *c3, uint32_t *ce, int *cs) override
    {
        auto st = c.getStatus();
        for (int k = 0; k < 3; ++k) e3[k] = st.energyValue[k], c3[k] = st.corrValue[k];
        *ce = st.coeffsEnergy;
        *cs = st.coeffScaling;
    }
};
}  // extern "C++"
void *ref_corr_create(int N, int S)
{
    if (N == 32 && S == 4) return new CorrImpl<32, 4>();
    if (N == 16 && S == 2) return new CorrImpl<16, 2>();
    if (N == 8 && S == 1) return new CorrImpl<8, 1>();
    if (N == 64 && S == 8) return new CorrImpl<64, 8>();
    return nullptr;
}
void ref_corr_destroy(void *h) { delete static_cast<CorrBase *>(h); }
void ref_corr_set_pattern(void *h, const int32_t *iq, double thr) { static_cast<CorrBase *>(h)->set_pattern(iq, thr); }
void ref_corr_reset(void *h) { static_cast<CorrBase *>(h)->reset(); }
int ref_corr_step(void *h, const int16_t *iq, size_t n, int *corr_index)
{
    std::vector<cs16> v(n);
    if (n) std::memcpy(v.data(), iq, n * sizeof(cs16));
    return static_cast<CorrBase *>(h)->step(v, *corr_index) ? 1 : 0;
}
void ref_corr_bits(void *h, int16_t *iq, size_t n)
{
    auto b = static_cast<CorrBase *>(h)->bits();
    std::memcpy(iq, b.data(), (b.size() < n ? b.size() : n) * sizeof(cs16));
}
void ref_corr_status(void *h, uint32_t *e3, uint32_t *c3, uint32_t *ce, int *cs) { static_cast<CorrBase *>(h)->status(e3, c3, ce, cs); }

const char *ref_build_info()
{
    return "SrcDsp reference headers, unmodified, g++ " __VERSION__
#ifdef __OPTIMIZE__
           " optimised"
#else
           " -O0 (makefile flags)"
#endif
        ;
}

}  // extern "C"
