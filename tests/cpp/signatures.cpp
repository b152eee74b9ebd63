// signatures.cpp -- compile-time statement of SURVEY.md 8(b) "signatures to keep": every public member of the three
// hot-path class templates (canonical 16-bit instantiations), with its default arguments and return type, written as
// unevaluated calls.  The file must compile (-fsyntax-only) against the reference headers AND against include/srcdsp/
// (tests/test_dropin_cpp.py::test_public_signatures_compile_against_both_trees); nothing here runs.
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <array>
#include <type_traits>
#include <utility>
#include <vector>

#include "mixers.h"
#ifdef SRCDSP_SIGNATURES_OBSOLETE_HEADER
#include "dnsampling_filters.h"
#else
#include "dsptl_dnsampling_filters.h"
#endif
#include "upsampling_filters.h"
#include "filters.h"
#include "buffers.h"
#include "correlators.h"

typedef std::complex<int16_t> cs16;
typedef std::complex<int32_t> cs32;
typedef std::vector<cs16> vec;
typedef std::vector<int32_t> taps_t;

typedef dsptl::Mixer<cs16, cs16, int16_t, 4096> MixerT;
typedef dsptl::FilterDnsamplingFir<cs16, cs16, cs32, int32_t, 8> DecT;
typedef dsptl::FilterUpsamplingFir<cs16, cs16, cs32, int32_t, 8> UpT;

#define RETURNS(type, expr) static_assert(std::is_same<decltype(expr), type>::value, #expr " must return " #type)

// Mixer (mixers.h:31-34, 134-135): default-constructible; setFrequency(float); reset(float = 0); adjustFrequency(float = 0);
// step(vector&, vector&)
static_assert(std::is_default_constructible<MixerT>::value, "Mixer()");
RETURNS(void, std::declval<MixerT &>().setFrequency(0.25f));
RETURNS(void, std::declval<MixerT &>().reset());
RETURNS(void, std::declval<MixerT &>().reset(0.5f));
RETURNS(void, std::declval<MixerT &>().adjustFrequency());
RETURNS(void, std::declval<MixerT &>().adjustFrequency(0.01f));
RETURNS(void, std::declval<MixerT &>().step(std::declval<vec &>(), std::declval<vec &>()));

// FilterDnsamplingFir (dsptl_dnsampling_filters.h:50-63; obsolete twin dnsampling_filters.h): (), (taps), setCoeffs,
// step(const vector&, vector&), reset, setLeftShiftBy2(int)
static_assert(std::is_constructible<DecT, const taps_t &>::value, "FilterDnsamplingFir(const std::vector<Coef>&)");
#ifdef SRCDSP_SIGNATURES_OBSOLETE_HEADER
// the obsolete header has neither a default constructor nor setCoeffs (dnsampling_filters.h:50-60): the taps come with the object
static_assert(!std::is_default_constructible<DecT>::value, "the obsolete FilterDnsamplingFir has no default constructor");
#else
static_assert(std::is_default_constructible<DecT>::value, "FilterDnsamplingFir()");
RETURNS(void, std::declval<DecT &>().setCoeffs(std::declval<const taps_t &>()));
#endif
RETURNS(void, std::declval<DecT &>().step(std::declval<const vec &>(), std::declval<vec &>()));
RETURNS(void, std::declval<DecT &>().reset());
RETURNS(void, std::declval<DecT &>().setLeftShiftBy2(1));

// FilterUpsamplingFir (upsampling_filters.h:42-67): (taps = {}), setCoefficients, the two step overloads with
// flush = false, reset, the three getters
static_assert(std::is_constructible<UpT, const taps_t &>::value, "FilterUpsamplingFir(const std::vector<Coef>&)");
static_assert(std::is_default_constructible<UpT>::value, "FilterUpsamplingFir(taps = {})");
RETURNS(void, std::declval<UpT &>().setCoefficients(std::declval<const taps_t &>()));
RETURNS(void, std::declval<UpT &>().step(std::declval<const vec &>(), std::declval<vec &>()));
RETURNS(void, std::declval<UpT &>().step(std::declval<const vec &>(), std::declval<vec &>(), true));
RETURNS(void, std::declval<UpT &>().step(std::declval<const vec &>(), std::declval<vec::iterator>()));
RETURNS(void, std::declval<UpT &>().step(std::declval<const vec &>(), std::declval<vec::iterator>(), true));
RETURNS(void, std::declval<UpT &>().reset());
RETURNS(int, std::declval<const UpT &>().getLength());  // const members returning int (upsampling_filters.h:57-67)
RETURNS(int, std::declval<const UpT &>().getImpLength());
RETURNS(int, std::declval<const UpT &>().getUpsamplingRatio());

// SURVEY.md 8(f) rows.  FilterFir (global namespace, filters.h:42-60): (), (taps), step, reset, setCoeffs
typedef FilterFir<cs16, cs16, cs32, int32_t> FirT;
static_assert(std::is_default_constructible<FirT>::value, "FilterFir()");
static_assert(std::is_constructible<FirT, const taps_t &>::value, "FilterFir(const std::vector<Coef>&)");
RETURNS(void, std::declval<FirT &>().step(std::declval<const vec &>(), std::declval<vec &>()));
RETURNS(void, std::declval<FirT &>().reset());
RETURNS(void, std::declval<FirT &>().setCoeffs(std::declval<const taps_t &>()));

// FifoWithTimeTrack<T, N> (buffers.h:58-83): (fs = 0), write(vector&, seconds = 0, frac = 0), read(vector&, uint64_t&) -> bool,
// count() -> size_t, reset(), dumpInfo(bool = false), getAbsoluteTime(uint64_t, double) -> pair<unsigned, double>
typedef dsptl::FifoWithTimeTrack<cs16, 1024> FifoT;
static_assert(std::is_default_constructible<FifoT>::value && std::is_constructible<FifoT, double>::value, "FifoWithTimeTrack(fs = 0)");
RETURNS(void, std::declval<FifoT &>().write(std::declval<vec &>()));
RETURNS(void, std::declval<FifoT &>().write(std::declval<vec &>(), 5u, 0.25));
RETURNS(bool, std::declval<FifoT &>().read(std::declval<vec &>(), std::declval<uint64_t &>()));
RETURNS(size_t, std::declval<FifoT &>().count());
RETURNS(void, std::declval<FifoT &>().reset());
RETURNS(void, std::declval<FifoT &>().dumpInfo());
RETURNS(void, std::declval<FifoT &>().dumpInfo(true));
typedef std::pair<unsigned int, double> abs_time_t;
RETURNS(abs_time_t, std::declval<FifoT &>().getAbsoluteTime(uint64_t(1), 0.5));

// FixedPatternCorrelator<int16_t, int32_t, N, S> (correlators.h:95-101): (), step(const vector&, int&) -> bool,
// setPattern(array, threshold = 0.8), reset(), getRefBitSamples() -> vector<complex<int16_t>>, getStatus()
typedef dsptl::FixedPatternCorrelator<int16_t, int32_t, 32, 4> CorrT;
typedef std::array<cs32, 32> pattern_t;
static_assert(std::is_default_constructible<CorrT>::value, "FixedPatternCorrelator()");
RETURNS(bool, std::declval<CorrT &>().step(std::declval<const vec &>(), std::declval<int &>()));
RETURNS(void, std::declval<CorrT &>().setPattern(std::declval<const pattern_t &>()));
RETURNS(void, std::declval<CorrT &>().setPattern(std::declval<const pattern_t &>(), 0.7));
RETURNS(void, std::declval<CorrT &>().reset());
RETURNS(vec, std::declval<CorrT &>().getRefBitSamples());
static_assert(std::is_same<std::decay<decltype(std::declval<CorrT &>().getStatus().corrValue[0])>::type, uint32_t>::value, "getStatus().corrValue");
static_assert(std::is_same<decltype(std::declval<CorrT &>().getStatus().thresholdFactor), double>::value, "getStatus().thresholdFactor");

int main() { return 0; }
