/*
 * srcdsp/upsampling_filters.h -- drop-in for the reference's upsampling_filters.h.
 *
 * Same class name, template parameter list and members as dsptl::FilterUpsamplingFir (reference
 * upsampling_filters.h:35-77) for the integer instantiation the reference supports
 * (dsp_complex.h:87-89): FilterUpsamplingFir<complex<int16_t>, complex<int16_t>,
 * complex<int32_t>, int32_t, L>.  step() runs on the GPU through srcdsp_up_step.
 */
#ifndef SRCDSP_DROPIN_UPSAMPLING_FILTERS_H
#define SRCDSP_DROPIN_UPSAMPLING_FILTERS_H

#include "detail.h"

namespace dsptl {

template <class InType, class OutType, class InternalType, class CoefType, unsigned L>
class FilterUpsamplingFir;

template <unsigned L>
class FilterUpsamplingFir<std::complex<int16_t>, std::complex<int16_t>, std::complex<int32_t>, int32_t, L> {
    typedef std::complex<int16_t> Sample;

public:
    /* upsampling_filters.h:89-96 */
    FilterUpsamplingFir(const std::vector<int32_t> &firCoeff = std::vector<int32_t>()) : h_(nullptr)
    {
        create();
        if (!firCoeff.empty()) setCoefficients(firCoeff);
    }
    FilterUpsamplingFir(const FilterUpsamplingFir &o) : h_(nullptr)
    {
        create();
        copy_from(o);
    }
    FilterUpsamplingFir &operator=(const FilterUpsamplingFir &o)
    {
        if (this != &o) copy_from(o);
        return *this;
    }
    ~FilterUpsamplingFir() { srcdsp_up_destroy(h_); }

    /* upsampling_filters.h:107-126 */
    void setCoefficients(const std::vector<int32_t> &firCoeff)
    {
        srcdsp_dropin::check(srcdsp_up_set_coefficients(h_, firCoeff.data(), static_cast<int>(firCoeff.size())),
                             "FilterUpsamplingFir::setCoefficients");
        coeff_ = firCoeff;
    }
    /* upsampling_filters.h:149-233 (output shift 15 - round(log2 L)) */
    void step(const std::vector<Sample> &signal, std::vector<Sample> &filteredSignal, bool flush = false)
    {
        const size_t need = L * (signal.size() + (flush ? static_cast<size_t>(getLength()) / L : 0));
        const bool ok = flush ? filteredSignal.size() >= need : filteredSignal.size() == need;
        srcdsp_dropin::check(ok ? SRCDSP_OK : SRCDSP_E_SIZE, "FilterUpsamplingFir::step (signal.size() * L != filteredSignal.size())");
        if (need == 0) return;
        srcdsp_dropin::check(srcdsp_up_step(h_, srcdsp_dropin::iq(signal), signal.size(), signal.size(),
                                            srcdsp_dropin::iq(filteredSignal), filteredSignal.size(), flush ? 1 : 0, 0),
                             "FilterUpsamplingFir::step");
    }
    /* upsampling_filters.h:240-323 (iterator destination; output shift 0).  The caller
     * guarantees room for L * (signal.size() + flush zeros) samples, as in the reference. */
    void step(const std::vector<Sample> &signal, typename std::vector<Sample>::iterator filteredSignal, bool flush = false)
    {
        const size_t need = L * (signal.size() + (flush ? static_cast<size_t>(getLength()) / L : 0));
        if (need == 0) return;
        srcdsp_dropin::check(srcdsp_up_step(h_, srcdsp_dropin::iq(signal), signal.size(), signal.size(),
                                            reinterpret_cast<int16_t *>(&*filteredSignal), need, flush ? 1 : 0, 1),
                             "FilterUpsamplingFir::step(iterator)");
    }
    /* upsampling_filters.h:50-55 */
    void reset() { srcdsp_dropin::check(srcdsp_up_reset(h_), "FilterUpsamplingFir::reset"); }
    /* upsampling_filters.h:57-67 */
    int getLength() const { return geti(srcdsp_up_get_length); }
    int getImpLength() const { return geti(srcdsp_up_get_imp_length); }
    int getUpsamplingRatio() const { return L; }

    srcdsp_up_t handle() const { return h_; }

private:
    void create() { srcdsp_dropin::check(srcdsp_up_create(&h_, srcdsp_dropin::default_device(), 1, L), "FilterUpsamplingFir()"); }
    int geti(int (*fn)(srcdsp_up_t, int *)) const
    {
        int v = 0;
        srcdsp_dropin::check(fn(h_, &v), "FilterUpsamplingFir getter");
        return v;
    }
    void copy_from(const FilterUpsamplingFir &o)
    {
        if (o.coeff_.empty()) return;
        setCoefficients(o.coeff_);
        std::vector<Sample> hist(coeff_.size() / L > 0 ? coeff_.size() / L - 1 : 0);
        size_t n = hist.size();
        if (n) {
            srcdsp_dropin::check(srcdsp_up_get_state(o.h_, 0, srcdsp_dropin::iq(hist), &n), "copy");
            srcdsp_dropin::check(srcdsp_up_set_state(h_, 0, srcdsp_dropin::iq(hist), n), "copy");
        }
    }
    srcdsp_up_t h_;
    std::vector<int32_t> coeff_;
};

}  // namespace dsptl

#endif
