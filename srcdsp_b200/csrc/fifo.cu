// fifo.cu -- FifoWithTimeTrack (reference buffers.h:58-459) as a ring in PINNED host memory: the
// streaming ingest stage in front of the GPU path.  Host code only.
//
// The observable behaviour is the reference's member by member (time points, wrap at 2^64 - 1, error returns);
// what changes is where the samples live: page-locked memory, so that the decimator / DDC banks
// can DMA a block straight out of the ring (srcdsp_fifo_segments + srcdsp_*_step on the segment
// pointers) instead of going through a pageable std::vector first.  When no CUDA device is
// present the ring falls back to ordinary memory: the FIFO is bookkeeping, not compute.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

struct srcdsp_fifo_s {
    size_t elem = 0, N = 0;      // bytes per element, ring capacity in elements
    uint8_t *storage = nullptr;
    bool pinned = false;
    size_t head = 0;             // ring slot the next write starts at                (reference: writePtr)
    // Time points number the samples from 1 and live in [1, 2^64 - 1]; 0 is "nothing written yet".  [oldest, newest]
    // is the window the ring still holds                                              (reference: timeStart / timeEnd)
    uint64_t oldest = 0, newest = 0;
    double fs = 0;               // samples per second
    struct {                     // absolute time of the first sample of the last block written (buffers.h:174-176)
        uint64_t point = 0;
        unsigned seconds = 0;
        double frac = 0;
    } anchor;
    std::mutex mx;               // guards head, oldest, newest and anchor; the sample copies run outside it
};

// A time counter that would pass 2^64 - 1 continues from 1 (the reference: buffers.h:179-186 for the newest point,
// :196-205 for the oldest one).
static inline uint64_t advance_point(uint64_t t, size_t n, bool *passed_top)
{
    const uint64_t room = UINT64_MAX - t;
    if (room >= n) return t + n;
    if (passed_top) *passed_top = true;
    return n - room;
}

// the (at most two) contiguous pieces of the ring that hold `n` elements from slot `slot` on
static inline size_t ring_first_piece(const srcdsp_fifo_s *f, size_t slot, size_t n)
{
    const size_t to_top = f->N - slot;
    return n < to_top ? n : to_top;
}

using namespace srcdsp;

extern "C" {

int srcdsp_fifo_create(srcdsp_fifo_t *h, size_t elem_bytes, size_t capacity, double sampling_frequency)
try {
    if (!h || elem_bytes == 0 || capacity == 0) return fail(SRCDSP_E_INVALID, "fifo: element size and capacity must be > 0");
    srcdsp_fifo_s *f = new (std::nothrow) srcdsp_fifo_s;
    if (!f) return fail(SRCDSP_E_CUDA, "out of memory");
    f->elem = elem_bytes;
    f->N = capacity;
    f->fs = sampling_frequency;
    void *p = nullptr;
    if (cudaHostAlloc(&p, elem_bytes * capacity, cudaHostAllocPortable) == cudaSuccess) {
        f->pinned = true;
    } else {
        cudaGetLastError();
        p = aligned_alloc(64, (elem_bytes * capacity + 63) / 64 * 64);
        if (!p) {
            delete f;
            return fail(SRCDSP_E_CUDA, "fifo: cannot allocate %zu bytes", elem_bytes * capacity);
        }
    }
    memset(p, 0, elem_bytes * capacity);  // std::vector<T> storage(N) value-initialises (buffers.h:64)
    f->storage = static_cast<uint8_t *>(p);
    *h = f;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_fifo_destroy(srcdsp_fifo_t f)
try {
    if (!f) return SRCDSP_OK;
    if (f->pinned)
        cudaFreeHost(f->storage);
    else
        free(f->storage);
    delete f;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_fifo_is_pinned(srcdsp_fifo_t f) { return f && f->pinned; }

/* write: buffers.h:139-217.  n must be < capacity (the reference asserts, :144). */
int srcdsp_fifo_write(srcdsp_fifo_t f, const void *in, size_t n, unsigned seconds, double frac_seconds)
try {
    if (!f || (!in && n)) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    if (n >= f->N) return fail(SRCDSP_E_SIZE, "fifo write of %zu elements into a fifo of %zu: must be smaller [buffers.h:144]", n, f->N);
    // samples first, outside the critical section: the single writer and the single reader work on different parts of
    // the ring (buffers.h:146-158)
    const uint8_t *src = static_cast<const uint8_t *>(in);
    const size_t a = ring_first_piece(f, f->head, n);
    memcpy(f->storage + f->head * f->elem, src, a * f->elem);
    if (a < n) memcpy(f->storage, src + a * f->elem, (n - a) * f->elem);

    std::lock_guard<std::mutex> lk(f->mx);
    f->head = (f->head + n) % f->N;
    f->anchor.point = f->newest + 1;  // this block's first sample carries the caller's absolute time
    f->anchor.seconds = seconds;
    f->anchor.frac = frac_seconds;
    bool passed_top = false;
    f->newest = advance_point(f->newest, n, &passed_top);
    if (passed_top) {
        // the one call in which the newest point wraps: the reference slides the window's start by the block length,
        // whatever the fill (:193-205; its rollover flag is raised and lowered again inside that call, :186 / :207)
        f->oldest = advance_point(f->oldest, n, nullptr);
    } else {
        // a window longer than the ring loses its oldest samples; a shorter one starts at point 1 (:189-192, unsigned
        // arithmetic as there)
        const uint64_t held = f->newest - f->oldest + 1;
        f->oldest = held > f->N ? f->newest - f->N + 1 : 1;
    }
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* Where in the ring [*start, *start + n) lives (the bookkeeping half of read, buffers.h:284-320).  A start before the
 * window is moved to its first point (the reference prints a warning and carries on, :296-301); a range that ends
 * behind the newest point is not available (the reference returns true). */
struct FifoRange {
    bool available = false, moved = false;
    size_t slot = 0;
};
static FifoRange fifo_locate(srcdsp_fifo_t f, size_t n, uint64_t *start)
{
    FifoRange r;
    std::lock_guard<std::mutex> lk(f->mx);
    if (*start < f->oldest) {
        *start = f->oldest;
        r.moved = true;
    }
    if (*start + n - 1 > f->newest) return r;
    r.available = true;
    const uint64_t behind_head = f->newest - *start + 1;  // the newest point sits one slot below the head
    r.slot = (f->head + f->N - behind_head) % f->N;
    return r;
}

/* read: buffers.h:284-352.  *error = the reference's return value (true = range not available). */
int srcdsp_fifo_read(srcdsp_fifo_t f, void *out, size_t n, uint64_t *start, int *error)
try {
    if (!f || !out || !start || !error) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    if (n == 0) return fail(SRCDSP_E_SIZE, "fifo read of 0 elements [buffers.h:288]");
    const FifoRange r = fifo_locate(f, n, start);
    if (r.moved) fputs("******* REQUESTED START BEFORE FIRST AVAILABLE SAMPLE *****", stderr);
    *error = !r.available;
    if (!r.available) return SRCDSP_OK;
    const size_t a = ring_first_piece(f, r.slot, n);
    memcpy(out, f->storage + r.slot * f->elem, a * f->elem);
    if (a < n) memcpy(static_cast<uint8_t *>(out) + a * f->elem, f->storage, (n - a) * f->elem);
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* zero-copy read: the (at most two) contiguous pieces of the ring that hold [start, start + n) */
int srcdsp_fifo_segments(srcdsp_fifo_t f, size_t n, uint64_t *start, const void **p0, size_t *n0, const void **p1,
                         size_t *n1, int *error)
try {
    if (!f || !start || !p0 || !n0 || !p1 || !n1 || !error) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    if (n == 0) return fail(SRCDSP_E_SIZE, "fifo read of 0 elements [buffers.h:288]");
    const FifoRange r = fifo_locate(f, n, start);
    *p0 = *p1 = nullptr;
    *n0 = *n1 = 0;
    *error = !r.available;
    if (!r.available) return SRCDSP_OK;
    const size_t a = ring_first_piece(f, r.slot, n);
    *p0 = f->storage + r.slot * f->elem;
    *n0 = a;
    if (a < n) {
        *p1 = f->storage;
        *n1 = n - a;
    }
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* count: buffers.h:361-377.  newest - oldest + 1: an empty fifo (both 0) reports 1, as the reference does.  The
 * reference's second formula applies while its rollover flag is up, which no caller can observe (see write). */
int srcdsp_fifo_count(srcdsp_fifo_t f, size_t *count)
try {
    if (!f || !count) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    std::lock_guard<std::mutex> lk(f->mx);
    *count = (size_t)(f->newest - f->oldest + 1);
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* reset: buffers.h:262-276 (the stored values and the time anchor are kept) */
int srcdsp_fifo_reset(srcdsp_fifo_t f)
try {
    if (!f) return fail(SRCDSP_E_INVALID, "fifo: null handle");
    std::lock_guard<std::mutex> lk(f->mx);
    f->head = 0;
    f->oldest = f->newest = 0;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* getAbsoluteTime: buffers.h:396-459.  Time of (time_point + frac_time_point) relative to the anchor of the last write:
 * whole seconds by floor (points before the anchor give negative differences), the fraction carried into the seconds
 * by truncation.  The order of the floating-point operations is the reference's: the doubles come out bit-identical. */
int srcdsp_fifo_get_absolute_time(srcdsp_fifo_t f, uint64_t time_point, double frac_time_point, unsigned *seconds,
                                  double *frac_seconds)
try {
    if (!f || !seconds || !frac_seconds) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    std::lock_guard<std::mutex> lk(f->mx);
    const double elapsed = (int64_t)(time_point - f->anchor.point) / f->fs;
    const double whole = floor(elapsed);
    const double frac = f->anchor.frac + (elapsed - whole) + (frac_time_point / f->fs);
    const int32_t carry = static_cast<int32_t>(frac);
    *seconds = f->anchor.seconds + (uint32_t) static_cast<int32_t>(whole) + (uint32_t)carry;
    *frac_seconds = frac - carry;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* head slot, oldest / newest time point, rollover flag (what dumpInfo prints, buffers.h:227-251; the flag is never up
 * between calls) */
int srcdsp_fifo_get_state(srcdsp_fifo_t f, size_t *write_ptr, uint64_t *time_start, uint64_t *time_end, int *rollover)
try {
    if (!f) return fail(SRCDSP_E_INVALID, "fifo: null handle");
    std::lock_guard<std::mutex> lk(f->mx);
    if (write_ptr) *write_ptr = f->head;
    if (time_start) *time_start = f->oldest;
    if (time_end) *time_end = f->newest;
    if (rollover) *rollover = 0;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* test hook: place the time counters near the 64-bit rollover (the reference has no such entry point; its
 * rollover branches are otherwise unreachable in a test) */
int srcdsp_fifo_set_time(srcdsp_fifo_t f, uint64_t time_start, uint64_t time_end)
try {
    if (!f) return fail(SRCDSP_E_INVALID, "fifo: null handle");
    std::lock_guard<std::mutex> lk(f->mx);
    f->oldest = time_start;
    f->newest = time_end;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

const void *srcdsp_fifo_storage(srcdsp_fifo_t f) { return f ? f->storage : nullptr; }

}  // extern "C"
