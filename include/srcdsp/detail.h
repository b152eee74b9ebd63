/*
 * srcdsp/detail.h -- glue shared by the drop-in class headers in this directory.
 *
 * The drop-in headers keep the reference's class names, template parameter lists and member
 * signatures (namespace dsptl) for the canonical 16-bit instantiations and forward every call to
 * the C ABI in ../srcdsp_b200.h (libsrcdsp_b200.so, hand-written sm_100a kernels).
 *
 * Error behaviour mirrors the reference: the reference assert()s its preconditions (abort in a
 * debug build); here a non-zero status prints srcdsp_last_error() and aborts.
 */
#ifndef SRCDSP_DROPIN_DETAIL_H
#define SRCDSP_DROPIN_DETAIL_H

#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../srcdsp_b200.h"

namespace srcdsp_dropin {

typedef std::complex<int16_t> cs16;
static_assert(sizeof(cs16) == 2 * sizeof(int16_t), "std::complex<int16_t> must be interleaved I/Q");

/* GPU used by objects constructed afterwards (default: $SRCDSP_DEVICE or 0). */
inline int &default_device()
{
    static int dev = [] {
        const char *e = std::getenv("SRCDSP_DEVICE");
        return e ? std::atoi(e) : 0;
    }();
    return dev;
}

inline void check(int status, const char *what)
{
    if (status != SRCDSP_OK) {
        std::fprintf(stderr, "srcdsp: %s failed (%d): %s\n", what, status, srcdsp_last_error());
        std::abort();
    }
}

inline const int16_t *iq(const std::vector<cs16> &v) { return reinterpret_cast<const int16_t *>(v.data()); }
inline int16_t *iq(std::vector<cs16> &v) { return reinterpret_cast<int16_t *>(v.data()); }

}  // namespace srcdsp_dropin

#endif
