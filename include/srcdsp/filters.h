/*
 * srcdsp/filters.h -- drop-in for the reference's filters.h (non-decimating FIR).
 *
 * Same name (global namespace, like the reference), template parameter list and members as
 * FilterFir (reference filters.h:42-60) for the 16-bit instantiation
 * FilterFir<complex<int16_t>, complex<int16_t>, complex<int32_t>, int32_t>.  In age order the
 * reference's circular buffer is the M = 1 case of the decimator, so step() runs
 * srcdsp_dec_step on a bank created with M = 1 (filters.h:130-169: out[j] =
 * limitScale16(sum_n c[n] * x[j-n], coeffScaling)).
 */
#ifndef SRCDSP_DROPIN_FILTERS_H
#define SRCDSP_DROPIN_FILTERS_H

#include "detail.h"

template <class InType, class OutType, class InternalType, class CoefType>
class FilterFir;

template <>
class FilterFir<std::complex<int16_t>, std::complex<int16_t>, std::complex<int32_t>, int32_t> {
    typedef std::complex<int16_t> Sample;

public:
    /* filters.h:48 */
    FilterFir() : h_(nullptr) { create(); }
    /* filters.h:71-77 */
    FilterFir(const std::vector<int32_t> &firCoeff) : h_(nullptr)
    {
        create();
        setCoeffs(firCoeff);
    }
    FilterFir(const FilterFir &o) : h_(nullptr)
    {
        create();
        copy_from(o);
    }
    FilterFir &operator=(const FilterFir &o)
    {
        if (this != &o) copy_from(o);
        return *this;
    }
    ~FilterFir() { srcdsp_dec_destroy(h_); }

    /* filters.h:130-169: signal.size() == filteredSignal.size(); filtering in place is allowed */
    void step(const std::vector<Sample> &signal, std::vector<Sample> &filteredSignal)
    {
        srcdsp_dropin::check(signal.size() == filteredSignal.size() ? SRCDSP_OK : SRCDSP_E_SIZE,
                             "FilterFir::step (signal.size() != filteredSignal.size())");
        if (signal.empty()) return;
        srcdsp_dropin::check(srcdsp_dec_step(h_, srcdsp_dropin::iq(signal), signal.size(), signal.size(),
                                             srcdsp_dropin::iq(filteredSignal), filteredSignal.size()),
                             "FilterFir::step");
    }
    /* filters.h:104-111 */
    void reset() { srcdsp_dropin::check(srcdsp_dec_reset(h_), "FilterFir::reset"); }
    /* filters.h:85-97: replaces the taps and clears the history */
    void setCoeffs(const std::vector<int32_t> &firCoeff)
    {
        srcdsp_dropin::check(srcdsp_dec_set_coeffs(h_, firCoeff.data(), static_cast<int>(firCoeff.size()), 0),
                             "FilterFir::setCoeffs");
        srcdsp_dropin::check(srcdsp_dec_reset(h_), "FilterFir::setCoeffs");
        coeff_ = firCoeff;
    }

    srcdsp_dec_t handle() const { return h_; }

private:
    void create() { srcdsp_dropin::check(srcdsp_dec_create(&h_, srcdsp_dropin::default_device(), 1, 1), "FilterFir()"); }
    void copy_from(const FilterFir &o)
    {
        if (o.coeff_.empty()) return;
        setCoeffs(o.coeff_);
        std::vector<Sample> hist(coeff_.size() - 1);
        size_t n = hist.size();
        if (n) {
            srcdsp_dropin::check(srcdsp_dec_get_state(o.h_, 0, srcdsp_dropin::iq(hist), &n), "copy");
            srcdsp_dropin::check(srcdsp_dec_set_state(h_, 0, srcdsp_dropin::iq(hist), n), "copy");
        }
    }
    srcdsp_dec_t h_;
    std::vector<int32_t> coeff_;
};

/* The float instantiation FilterFir<complex<float>, complex<float>, complex<float>, float>: in age order the M = 1
 * case of the float decimator (srcdsp_decf_*): the same tap-order float sum (filters.h:146-160), the same int32
 * truncation + limitScale16 (:161) and the same integer abs() in coeffScaling (:93-95) -- bit-exact with the
 * compiled reference (tests/cpp/user_float.cpp). */
template <>
class FilterFir<std::complex<float>, std::complex<float>, std::complex<float>, float> {
    typedef std::complex<float> Sample;

public:
    FilterFir() : h_(nullptr) { create(); }
    FilterFir(const std::vector<float> &firCoeff) : h_(nullptr)
    {
        create();
        setCoeffs(firCoeff);
    }
    FilterFir(const FilterFir &o) : h_(nullptr)
    {
        create();
        copy_from(o);
    }
    FilterFir &operator=(const FilterFir &o)
    {
        if (this != &o) copy_from(o);
        return *this;
    }
    ~FilterFir() { srcdsp_decf_destroy(h_); }

    void step(const std::vector<Sample> &signal, std::vector<Sample> &filteredSignal)
    {
        srcdsp_dropin::check(signal.size() == filteredSignal.size() ? SRCDSP_OK : SRCDSP_E_SIZE,
                             "FilterFir<float>::step (signal.size() != filteredSignal.size())");
        if (signal.empty()) return;
        srcdsp_dropin::check(srcdsp_decf_step(h_, reinterpret_cast<const float *>(signal.data()), signal.size(), signal.size(),
                                              reinterpret_cast<float *>(filteredSignal.data()), filteredSignal.size()),
                             "FilterFir<float>::step");
    }
    void reset() { srcdsp_dropin::check(srcdsp_decf_reset(h_), "FilterFir<float>::reset"); }
    void setCoeffs(const std::vector<float> &firCoeff)
    {
        srcdsp_dropin::check(srcdsp_decf_set_coeffs(h_, firCoeff.data(), static_cast<int>(firCoeff.size()), 0),
                             "FilterFir<float>::setCoeffs");
        srcdsp_dropin::check(srcdsp_decf_reset(h_), "FilterFir<float>::setCoeffs");
        coeff_ = firCoeff;
    }
    srcdsp_decf_t handle() const { return h_; }

private:
    void create() { srcdsp_dropin::check(srcdsp_decf_create(&h_, srcdsp_dropin::default_device(), 1, 1), "FilterFir<float>()"); }
    void copy_from(const FilterFir &o)
    {
        if (o.coeff_.empty()) return;
        setCoeffs(o.coeff_);
        std::vector<Sample> hist(coeff_.size() - 1);
        size_t n = hist.size();
        if (n) {
            srcdsp_dropin::check(srcdsp_decf_get_state(o.h_, 0, reinterpret_cast<float *>(hist.data()), &n), "copy");
            srcdsp_dropin::check(srcdsp_decf_set_state(h_, 0, reinterpret_cast<const float *>(hist.data()), n), "copy");
        }
    }
    srcdsp_decf_t h_;
    std::vector<float> coeff_;
};

#endif
