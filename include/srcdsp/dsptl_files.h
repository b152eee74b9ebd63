/*
 * srcdsp/dsptl_files.h -- drop-in for the binary sample file helpers of the reference's
 * dsptl_files.h (saveBinarySamples :57-109, readBinarySamples :250-262): the de-facto on-disk /
 * wire format of the library, raw interleaved I/Q with no header.  Host
 * code; it exists here so that recorded captures can be replayed through the GPU path and the
 * outputs diffed against files the reference produced.
 *
 * readBinarySamples reproduces the reference's observable behaviour, which a caller may depend on:
 *   - `out.empty()` at :254 is a no-op, so the samples are APPENDED to `out`;
 *   - the `while (is)` loop (:255-261) pushes one more element after the last complete sample,
 *     built from variables the failed read left untouched -- with the compilers the reference
 *     builds with that is a copy of the last sample (or of a partial trailing one).
 * readBinarySamplesExact is the same reader without the trailing element.
 */
#ifndef SRCDSP_DROPIN_DSPTL_FILES_H
#define SRCDSP_DROPIN_DSPTL_FILES_H

#include <complex>
#include <fstream>
#include <vector>

namespace dsptl {

/* dsptl_files.h:57-63 */
template <class Type>
void saveBinarySamples(std::vector<Type> &in, std::ofstream &os)
{
    os.write(reinterpret_cast<char *>(in.data()), in.size() * sizeof(Type));
    os.flush();
}

/* dsptl_files.h:76-90: any container of complex values, element by element */
template <class Type, class Allocator, template <class, class> class Container>
void saveBinarySamples(Container<std::complex<Type>, Allocator> &in, std::ofstream &os)
{
    static_assert(sizeof(std::complex<Type>) == 2 * sizeof(Type), "");
    for (auto &e : in) os.write(reinterpret_cast<char *>(&e), sizeof(std::complex<Type>));
    os.flush();
}

/* dsptl_files.h:101-109 */
template <class Type>
void saveBinarySamples(std::vector<std::complex<Type>> &in, std::ofstream &os)
{
    static_assert(sizeof(std::complex<Type>) == 2 * sizeof(Type), "");
    os.write(reinterpret_cast<char *>(in.data()), in.size() * sizeof(std::complex<Type>));
    os.flush();
}

/* complete samples only, appended to out */
template <class Type>
void readBinarySamplesExact(std::ifstream &is, std::vector<std::complex<Type>> &out)
{
    Type iq[2];
    while (is.read(reinterpret_cast<char *>(iq), sizeof iq)) out.push_back(std::complex<Type>(iq[0], iq[1]));
}

/* dsptl_files.h:250-262 */
template <class Type>
void readBinarySamples(std::ifstream &is, std::vector<std::complex<Type>> &out)
{
    Type i = Type(), q = Type();
    while (is) {
        is.read(reinterpret_cast<char *>(&i), sizeof(Type));
        is.read(reinterpret_cast<char *>(&q), sizeof(Type));
        out.push_back(std::complex<Type>(i, q));
    }
}

}  // namespace dsptl

#endif
