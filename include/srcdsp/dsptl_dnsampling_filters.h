/*
 * srcdsp/dsptl_dnsampling_filters.h -- drop-in for the reference's dsptl_dnsampling_filters.h
 * (and, with SRCDSP_DNSAMPLING_OBSOLETE defined by dnsampling_filters.h next to this file, for
 * its obsolete twin dnsampling_filters.h, which has no `taps % M == 0` precondition).
 *
 * Same class name, template parameter list and members as dsptl::FilterDnsamplingFir (reference
 * dsptl_dnsampling_filters.h:43-70) for the instantiation the reference supports (:40-41):
 * FilterDnsamplingFir<complex<int16_t>, complex<int16_t>, complex<int32_t>, int32_t, M>
 * (step() runs on the GPU through srcdsp_dec_step, ../srcdsp_b200.h), and for the float instantiation
 * FilterDnsamplingFir<complex<float>, complex<float>, complex<float>, float, M> (srcdsp_decf_step).
 */
#ifndef SRCDSP_DROPIN_DNSAMPLING_FILTERS_H
#define SRCDSP_DROPIN_DNSAMPLING_FILTERS_H

#include "detail.h"

namespace dsptl {

template <class InType, class OutType, class InternalType, class CoefType, unsigned M>
class FilterDnsamplingFir;

template <unsigned M>
class FilterDnsamplingFir<std::complex<int16_t>, std::complex<int16_t>, std::complex<int32_t>, int32_t, M> {
    typedef std::complex<int16_t> Sample;

public:
#ifndef SRCDSP_DNSAMPLING_OBSOLETE
    /* dsptl_dnsampling_filters.h:81-83: no coefficients yet */
    FilterDnsamplingFir() : h_(nullptr), leftShift_(0) { create(); }
#endif
    /* dsptl_dnsampling_filters.h:93-101 / dnsampling_filters.h:83-97 */
    FilterDnsamplingFir(const std::vector<int32_t> &firCoeff) : h_(nullptr), leftShift_(0)
    {
        create();
        load(firCoeff);
    }
    FilterDnsamplingFir(const FilterDnsamplingFir &o) : h_(nullptr), leftShift_(0)
    {
        create();
        copy_from(o);
    }
    FilterDnsamplingFir &operator=(const FilterDnsamplingFir &o)
    {
        if (this != &o) copy_from(o);
        return *this;
    }
    ~FilterDnsamplingFir() { srcdsp_dec_destroy(h_); }

    /* dsptl_dnsampling_filters.h:172-220: filteredSignal.size() * M == input.size() */
    void step(const std::vector<Sample> &input, std::vector<Sample> &filteredSignal)
    {
        srcdsp_dropin::check(filteredSignal.size() * M == input.size() ? SRCDSP_OK : SRCDSP_E_SIZE,
                             "FilterDnsamplingFir::step (filteredSignal.size() * M != input.size())");
        if (input.empty()) return;
        srcdsp_dropin::check(srcdsp_dec_step(h_, srcdsp_dropin::iq(input), input.size(), input.size(),
                                             srcdsp_dropin::iq(filteredSignal), filteredSignal.size()),
                             "FilterDnsamplingFir::step");
    }
    /* dsptl_dnsampling_filters.h:55-59 */
    void reset() { srcdsp_dropin::check(srcdsp_dec_reset(h_), "FilterDnsamplingFir::reset"); }
#ifndef SRCDSP_DNSAMPLING_OBSOLETE
    /* dsptl_dnsampling_filters.h:114-134 */
    void setCoeffs(const std::vector<int32_t> &firCoeff) { load(firCoeff); }
#endif
    /* dsptl_dnsampling_filters.h:63 */
    void setLeftShiftBy2(int leftShiftBy2)
    {
        leftShift_ = leftShiftBy2;
        srcdsp_dropin::check(srcdsp_dec_set_left_shift(h_, leftShiftBy2), "FilterDnsamplingFir::setLeftShiftBy2");
    }

    /* extension: the C-ABI handle, e.g. to build a fused srcdsp_ddc chain */
    srcdsp_dec_t handle() const { return h_; }

private:
    void create()
    {
        srcdsp_dropin::check(srcdsp_dec_create(&h_, srcdsp_dropin::default_device(), 1, M), "FilterDnsamplingFir()");
    }
    void load(const std::vector<int32_t> &c)
    {
#ifdef SRCDSP_DNSAMPLING_OBSOLETE
        const int strict = 0;
#else
        const int strict = 1;
#endif
        srcdsp_dropin::check(srcdsp_dec_set_coeffs(h_, c.data(), static_cast<int>(c.size()), strict),
                             "FilterDnsamplingFir::setCoeffs");
        coeff_ = c;
        leftShift_ = 0;
    }
    void copy_from(const FilterDnsamplingFir &o)
    {
        if (o.coeff_.empty()) return;
        srcdsp_dropin::check(srcdsp_dec_set_coeffs(h_, o.coeff_.data(), static_cast<int>(o.coeff_.size()), 0), "copy");
        coeff_ = o.coeff_;
        setLeftShiftBy2(o.leftShift_);
        std::vector<Sample> hist(coeff_.size() - 1);
        size_t n = hist.size();
        if (n) {
            srcdsp_dropin::check(srcdsp_dec_get_state(o.h_, 0, srcdsp_dropin::iq(hist), &n), "copy");
            srcdsp_dropin::check(srcdsp_dec_set_state(h_, 0, srcdsp_dropin::iq(hist), n), "copy");
        }
    }
    srcdsp_dec_t h_;
    std::vector<int32_t> coeff_;
    int leftShift_;
};

/* The float instantiation (the only other one the reference header can build): same members, float samples and
 * taps, step() on the GPU through srcdsp_decf_step -- bit-exact with the reference's tap-order float sum, its int32
 * truncation + limitScale16 (:214) and its coeffScaling (:128-132). */
template <unsigned M>
class FilterDnsamplingFir<std::complex<float>, std::complex<float>, std::complex<float>, float, M> {
    typedef std::complex<float> Sample;
    static_assert(sizeof(Sample) == 2 * sizeof(float), "std::complex<float> must be interleaved I/Q");

public:
#ifndef SRCDSP_DNSAMPLING_OBSOLETE
    FilterDnsamplingFir() : h_(nullptr), leftShift_(0) { create(); }
#endif
    FilterDnsamplingFir(const std::vector<float> &firCoeff) : h_(nullptr), leftShift_(0)
    {
        create();
        load(firCoeff);
    }
    FilterDnsamplingFir(const FilterDnsamplingFir &o) : h_(nullptr), leftShift_(0)
    {
        create();
        copy_from(o);
    }
    FilterDnsamplingFir &operator=(const FilterDnsamplingFir &o)
    {
        if (this != &o) copy_from(o);
        return *this;
    }
    ~FilterDnsamplingFir() { srcdsp_decf_destroy(h_); }

    /* dsptl_dnsampling_filters.h:172-220 */
    void step(const std::vector<Sample> &input, std::vector<Sample> &filteredSignal)
    {
        srcdsp_dropin::check(filteredSignal.size() * M == input.size() ? SRCDSP_OK : SRCDSP_E_SIZE,
                             "FilterDnsamplingFir<float>::step (filteredSignal.size() * M != input.size())");
        if (input.empty()) return;
        srcdsp_dropin::check(srcdsp_decf_step(h_, reinterpret_cast<const float *>(input.data()), input.size(), input.size(),
                                              reinterpret_cast<float *>(filteredSignal.data()), filteredSignal.size()),
                             "FilterDnsamplingFir<float>::step");
    }
    void reset() { srcdsp_dropin::check(srcdsp_decf_reset(h_), "FilterDnsamplingFir<float>::reset"); }
#ifndef SRCDSP_DNSAMPLING_OBSOLETE
    void setCoeffs(const std::vector<float> &firCoeff) { load(firCoeff); }
#endif
    void setLeftShiftBy2(int leftShiftBy2)
    {
        leftShift_ = leftShiftBy2;
        srcdsp_dropin::check(srcdsp_decf_set_left_shift(h_, leftShiftBy2), "FilterDnsamplingFir<float>::setLeftShiftBy2");
    }
    srcdsp_decf_t handle() const { return h_; }

private:
    void create()
    {
        srcdsp_dropin::check(srcdsp_decf_create(&h_, srcdsp_dropin::default_device(), 1, M), "FilterDnsamplingFir<float>()");
    }
    void load(const std::vector<float> &c)
    {
#ifdef SRCDSP_DNSAMPLING_OBSOLETE
        const int strict = 0;
#else
        const int strict = 1;
#endif
        srcdsp_dropin::check(srcdsp_decf_set_coeffs(h_, c.data(), static_cast<int>(c.size()), strict),
                             "FilterDnsamplingFir<float>::setCoeffs");
        coeff_ = c;
        leftShift_ = 0;
    }
    void copy_from(const FilterDnsamplingFir &o)
    {
        if (o.coeff_.empty()) return;
        srcdsp_dropin::check(srcdsp_decf_set_coeffs(h_, o.coeff_.data(), static_cast<int>(o.coeff_.size()), 0), "copy");
        coeff_ = o.coeff_;
        setLeftShiftBy2(o.leftShift_);
        std::vector<Sample> hist(coeff_.size() - 1);
        size_t n = hist.size();
        if (n) {
            srcdsp_dropin::check(srcdsp_decf_get_state(o.h_, 0, reinterpret_cast<float *>(hist.data()), &n), "copy");
            srcdsp_dropin::check(srcdsp_decf_set_state(h_, 0, reinterpret_cast<const float *>(hist.data()), n), "copy");
        }
    }
    srcdsp_decf_t h_;
    std::vector<float> coeff_;
    int leftShift_;
};

}  // namespace dsptl

#endif
