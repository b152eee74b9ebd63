/*
 * srcdsp/dnsampling_filters.h -- drop-in for the reference's OBSOLETE dnsampling_filters.h
 * (reference dnsampling_filters.h:43-172): same class, but no default constructor, no setCoeffs
 * and no `taps % M == 0` precondition ("no constraints on the number of coefficients", :78).
 * Like the reference pair, only one of the two decimator headers can be included per
 * translation unit.
 */
#ifndef SRCDSP_DROPIN_DNSAMPLING_FILTERS_H
#define SRCDSP_DNSAMPLING_OBSOLETE 1
#include "dsptl_dnsampling_filters.h"
#else
#ifndef SRCDSP_DNSAMPLING_OBSOLETE
#error "dsptl_dnsampling_filters.h and dnsampling_filters.h cannot be included together (as in the reference)"
#endif
#endif
