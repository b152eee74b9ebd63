/*
 * srcdsp/correlators.h -- drop-in for the reference's correlators.h:
 * dsptl::FixedPatternCorrelator<InType, CompType, N, S> (correlators.h:54-303) for the instantiation the
 * reference uses, <int16_t, int32_t, N, S>.  Same template parameters and member signatures; the
 * sliding correlation, the energy and the peak search run on the GPU (libsrcdsp_b200.so,
 * srcdsp_corr_*, csrc/kernels_corr.cuh).
 */
#ifndef SRCDSP_DROPIN_CORRELATORS_H
#define SRCDSP_DROPIN_CORRELATORS_H

#include <array>
#include <sstream>
#include <string>

#include "detail.h"

namespace dsptl {

template <class InType = int16_t, class CompType = int32_t, size_t N = 32, size_t S = 4>
class FixedPatternCorrelator
{
    static_assert(sizeof(InType) == 2 && sizeof(CompType) == 4,
                  "only FixedPatternCorrelator<int16_t, int32_t, N, S> is provided (the reference's instantiation)");

public:
    /* correlators.h:58-82 */
    struct CorrState {
        static const int Nelements = 3;
        float InputEnergy;
        uint32_t coeffsEnergy;
        int coeffScaling;
        uint32_t energyValue[Nelements];
        uint32_t corrValue[Nelements];
        double thresholdFactor;
        std::string prettyString()
        {
            std::ostringstream os;
            os << "Input Energy: " << InputEnergy << '\n';
            os << "Coeffs Energy: " << coeffsEnergy << '\n';
            os << "Coeff Scaling: " << coeffScaling << '\n';
            os << "Threshold Factor: " << thresholdFactor << '\n';
            for (int index = 0; index < Nelements; ++index) os << "Energy Value " << index << ": " << energyValue[index] << '\n';
            for (int index = 0; index < Nelements; ++index) os << "CorrValue " << index << ": " << corrValue[index] << '\n';
            return os.str();
        }
    };

    /* correlators.h:124-133 */
    FixedPatternCorrelator()
    {
        srcdsp_dropin::check(srcdsp_corr_create(&h_, srcdsp_dropin::default_device(), 1, (int)N, (int)S), "FixedPatternCorrelator()");
    }
    ~FixedPatternCorrelator() { srcdsp_corr_destroy(h_); }
    FixedPatternCorrelator(const FixedPatternCorrelator &) = delete;
    FixedPatternCorrelator &operator=(const FixedPatternCorrelator &) = delete;

    /* correlators.h:209-303 */
    bool step(const std::vector<std::complex<InType>> &in, int &corrIndex)
    {
        int found = 0, idx = 0;
        srcdsp_dropin::check(srcdsp_corr_step(h_, reinterpret_cast<const int16_t *>(in.data()), in.size(), in.size(), &found, &idx),
                             "FixedPatternCorrelator::step");
        if (found) corrIndex = idx;
        return found != 0;
    }
    /* correlators.h:167-194 */
    void setPattern(const std::array<std::complex<CompType>, N> &in, double thresholdCoeff = 0.8)
    {
        static_assert(sizeof(std::complex<CompType>) == 2 * sizeof(CompType), "");
        srcdsp_dropin::check(srcdsp_corr_set_pattern(h_, reinterpret_cast<const int32_t *>(in.data()), thresholdCoeff),
                             "FixedPatternCorrelator::setPattern");
    }
    /* correlators.h:143-155 */
    void reset() { srcdsp_dropin::check(srcdsp_corr_reset(h_), "FixedPatternCorrelator::reset"); }
    /* correlators.h:311-316 */
    std::vector<std::complex<InType>> getRefBitSamples()
    {
        std::vector<std::complex<InType>> b(N);
        srcdsp_dropin::check(srcdsp_corr_get_ref_bit_samples(h_, 0, reinterpret_cast<int16_t *>(b.data())),
                             "FixedPatternCorrelator::getRefBitSamples");
        return b;
    }
    /* correlators.h:101 */
    CorrState getStatus()
    {
        CorrState st = CorrState();
        srcdsp_dropin::check(srcdsp_corr_get_status(h_, 0, st.energyValue, st.corrValue, &st.coeffsEnergy, &st.coeffScaling,
                                                    &st.thresholdFactor),
                             "FixedPatternCorrelator::getStatus");
        return st;
    }

private:
    srcdsp_corr_t h_ = nullptr;
};

}  // namespace dsptl

#endif
