/*
 * srcdsp/dsp_complex.h -- drop-in for the reference's dsp_complex.h / dsp_complex.cpp: the sample
 * types' scalar helpers.  User code that includes "dsp_complex.h" for
 * operator*, scale32, limitScale16 or limitScale<> keeps compiling against the drop-in headers;
 * here they are inline host functions (the reference links dsp_complex.cpp), with the same
 * arithmetic: full int32 complex products, arithmetic >> before the clamp, +-32767 for
 * limitScale16 and [lowest, max] for limitScale<>.  The GPU kernels restate the same rules in
 * csrc/common.cuh.
 */
#ifndef SRCDSP_DROPIN_DSP_COMPLEX_H
#define SRCDSP_DROPIN_DSP_COMPLEX_H

#include <complex>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <type_traits>
#include <vector>

/* dsp_complex.cpp:23-29 */
inline std::complex<int32_t> operator*(std::complex<int32_t> a, std::complex<int16_t> b)
{
    return std::complex<int32_t>(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
/* dsp_complex.cpp:31-37 */
inline std::complex<int32_t> operator*(std::complex<int16_t> a, std::complex<int32_t> b)
{
    return std::complex<int32_t>(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
/* dsp_complex.cpp:43-46, :52-55 */
inline std::complex<int32_t> scale32(std::complex<int32_t> z, unsigned shift)
{
    return std::complex<int32_t>(z.real() >> shift, z.imag() >> shift);
}
inline std::complex<uint32_t> scale32(std::complex<uint32_t> z, unsigned shift)
{
    return std::complex<uint32_t>(z.real() >> shift, z.imag() >> shift);
}
/* dsp_complex.cpp:63-73 */
inline std::complex<int16_t> limitScale16(std::complex<int32_t> z, unsigned shift)
{
    int32_t a = z.real() >> shift, b = z.imag() >> shift;
    if (std::abs(a) > INT16_MAX) a = a > 0 ? INT16_MAX : -INT16_MAX;
    if (std::abs(b) > INT16_MAX) b = b > 0 ? INT16_MAX : -INT16_MAX;
    return std::complex<int16_t>((int16_t)a, (int16_t)b);
}
/* dsp_complex.h:45-63 (scalar) */
template <class T, class U, typename std::enable_if<std::numeric_limits<T>::is_integer, int>::type * = nullptr>
T limitScale(U z, unsigned shift)
{
    static_assert(std::numeric_limits<U>::is_integer && std::numeric_limits<T>::is_integer, "");
    static_assert(std::numeric_limits<U>::digits >= std::numeric_limits<T>::digits, "");
    U a = z >> shift;
    if (a > std::numeric_limits<T>::max())
        a = std::numeric_limits<T>::max();
    else if (a < std::numeric_limits<T>::lowest())
        a = std::numeric_limits<T>::lowest();
    return a;
}
/* dsp_complex.h:83-108 (complex) */
template <class T, class U, typename std::enable_if<!std::numeric_limits<T>::is_integer, int>::type * = nullptr>
T limitScale(U z, unsigned shift)
{
    typedef typename U::value_type UV;
    typedef typename T::value_type TV;
    static_assert(std::numeric_limits<UV>::is_integer && std::numeric_limits<TV>::is_integer, "");
    static_assert(std::numeric_limits<UV>::digits >= std::numeric_limits<TV>::digits, "");
    UV a = z.real() >> shift, b = z.imag() >> shift;
    if (a > std::numeric_limits<TV>::max())
        a = std::numeric_limits<TV>::max();
    else if (a < std::numeric_limits<TV>::lowest())
        a = std::numeric_limits<TV>::lowest();
    if (b > std::numeric_limits<TV>::max())
        b = std::numeric_limits<TV>::max();
    else if (b < std::numeric_limits<TV>::lowest())
        b = std::numeric_limits<TV>::lowest();
    return T(a, b);
}

#endif
