// fifo.cu -- FifoWithTimeTrack (reference buffers.h:58-459) as a ring in PINNED host memory: the
// streaming ingest stage in front of the GPU path.  Host code only.
//
// Semantics follow the reference member by member (time points, rollover flag, error returns);
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
    size_t elem = 0, N = 0;
    uint8_t *storage = nullptr;
    bool pinned = false;
    // buffers.h:97-113
    size_t writePtr = 0;
    uint64_t timeStart = 0, timeEnd = 0;
    bool rolloverFlag = false;
    double samplingFrequency = 0;
    struct {
        uint64_t timePoint = 0;
        unsigned seconds = 0;
        double frac = 0;
    } ref;
    std::mutex mx;
};

using namespace srcdsp;

extern "C" {

int srcdsp_fifo_create(srcdsp_fifo_t *h, size_t elem_bytes, size_t capacity, double sampling_frequency)
try {
    if (!h || elem_bytes == 0 || capacity == 0) return fail(SRCDSP_E_INVALID, "fifo: element size and capacity must be > 0");
    srcdsp_fifo_s *f = new (std::nothrow) srcdsp_fifo_s;
    if (!f) return fail(SRCDSP_E_CUDA, "out of memory");
    f->elem = elem_bytes;
    f->N = capacity;
    f->samplingFrequency = sampling_frequency;
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
    const size_t N = f->N;
    if (n >= N) return fail(SRCDSP_E_SIZE, "fifo write of %zu elements into a fifo of %zu: must be smaller [buffers.h:144]", n, N);
    const size_t upToTop = N - f->writePtr;
    const uint8_t *src = static_cast<const uint8_t *>(in);
    // the copy is outside the critical section: reader and writer touch different parts (buffers.h:146-158)
    if (n <= upToTop) {
        memcpy(f->storage + f->writePtr * f->elem, src, n * f->elem);
    } else {
        memcpy(f->storage + f->writePtr * f->elem, src, upToTop * f->elem);
        memcpy(f->storage, src + upToTop * f->elem, (n - upToTop) * f->elem);
    }
    std::lock_guard<std::mutex> lk(f->mx);
    f->writePtr = (f->writePtr + n) % N;
    const uint64_t diff = UINT64_MAX - f->timeEnd;
    f->ref.timePoint = f->timeEnd + 1;  // :174-176
    f->ref.seconds = seconds;
    f->ref.frac = frac_seconds;
    if (diff >= n) {
        f->timeEnd += n;
    } else {
        f->timeEnd = n - diff;
        f->rolloverFlag = true;
    }
    if (!f->rolloverFlag) {
        if ((f->timeEnd - f->timeStart + 1) > N)
            f->timeStart = f->timeEnd - N + 1;
        else
            f->timeStart = 1;
    } else {
        const uint64_t d2 = UINT64_MAX - f->timeStart;
        if (d2 >= n)
            f->timeStart += n;
        else
            f->timeStart = n - d2;
        f->rolloverFlag = false;  // :207 (outside the inner else in the reference)
    }
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* the bookkeeping half of read (buffers.h:284-320): returns 1 in *error when the range is not available */
static void fifo_locate(srcdsp_fifo_t f, size_t n, uint64_t *start, size_t *startPtr, int *error, int *adjusted)
{
    std::lock_guard<std::mutex> lk(f->mx);
    *adjusted = 0;
    if (*start < f->timeStart) {
        *start = f->timeStart;  // the reference prints a warning and carries on (:296-301)
        *adjusted = 1;
    }
    if ((*start + n - 1) > f->timeEnd) {
        *error = 1;
        return;
    }
    *error = 0;
    *startPtr = (f->writePtr + f->N - (f->timeEnd - *start) - 1) % f->N;
}

/* read: buffers.h:284-352.  *error = the reference's return value (true = range not available). */
int srcdsp_fifo_read(srcdsp_fifo_t f, void *out, size_t n, uint64_t *start, int *error)
try {
    if (!f || !out || !start || !error) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    if (n == 0) return fail(SRCDSP_E_SIZE, "fifo read of 0 elements [buffers.h:288]");
    size_t sp = 0;
    int adj = 0;
    fifo_locate(f, n, start, &sp, error, &adj);
    if (adj) fputs("******* REQUESTED START BEFORE FIRST AVAILABLE SAMPLE *****", stderr);
    if (*error) return SRCDSP_OK;
    const size_t first = (sp + n <= f->N) ? n : f->N - sp;
    memcpy(out, f->storage + sp * f->elem, first * f->elem);
    if (first < n) memcpy(static_cast<uint8_t *>(out) + first * f->elem, f->storage, (n - first) * f->elem);
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* zero-copy read: the (at most two) contiguous pieces of the ring that hold [start, start + n) */
int srcdsp_fifo_segments(srcdsp_fifo_t f, size_t n, uint64_t *start, const void **p0, size_t *n0, const void **p1,
                         size_t *n1, int *error)
try {
    if (!f || !start || !p0 || !n0 || !p1 || !n1 || !error) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    if (n == 0) return fail(SRCDSP_E_SIZE, "fifo read of 0 elements [buffers.h:288]");
    size_t sp = 0;
    int adj = 0;
    fifo_locate(f, n, start, &sp, error, &adj);
    *p0 = *p1 = nullptr;
    *n0 = *n1 = 0;
    if (*error) return SRCDSP_OK;
    const size_t first = (sp + n <= f->N) ? n : f->N - sp;
    *p0 = f->storage + sp * f->elem;
    *n0 = first;
    if (first < n) {
        *p1 = f->storage;
        *n1 = n - first;
    }
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* count: buffers.h:361-377 */
int srcdsp_fifo_count(srcdsp_fifo_t f, size_t *count)
try {
    if (!f || !count) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    std::lock_guard<std::mutex> lk(f->mx);
    if (!f->rolloverFlag)
        *count = (size_t)((f->timeEnd - f->timeStart) + 1);
    else
        *count = (size_t)((UINT64_MAX - f->timeStart) + f->timeEnd + 1);
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* reset: buffers.h:262-276 (the stored values are not cleared) */
int srcdsp_fifo_reset(srcdsp_fifo_t f)
try {
    if (!f) return fail(SRCDSP_E_INVALID, "fifo: null handle");
    std::lock_guard<std::mutex> lk(f->mx);
    f->writePtr = 0;
    f->timeStart = 0;
    f->timeEnd = 0;
    f->rolloverFlag = false;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* getAbsoluteTime: buffers.h:396-459 */
int srcdsp_fifo_get_absolute_time(srcdsp_fifo_t f, uint64_t time_point, double frac_time_point, unsigned *seconds,
                                  double *frac_seconds)
try {
    if (!f || !seconds || !frac_seconds) return fail(SRCDSP_E_INVALID, "fifo: null argument");
    std::lock_guard<std::mutex> lk(f->mx);
    const int64_t sampleDiff = (int64_t)(time_point - f->ref.timePoint);
    const double timeDiff = sampleDiff / f->samplingFrequency;
    const int32_t timeDiffInt = static_cast<int32_t>(floor(timeDiff));
    const double timeDiffFrac = timeDiff - floor(timeDiff);
    uint32_t s = f->ref.seconds + timeDiffInt;
    double fr = f->ref.frac + timeDiffFrac + (frac_time_point / f->samplingFrequency);
    const int32_t tmp = static_cast<int32_t>(fr);
    fr -= tmp;
    s += tmp;
    *seconds = s;
    *frac_seconds = fr;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* writePtr, timeStart, timeEnd, rolloverFlag (what dumpInfo prints, buffers.h:227-251) */
int srcdsp_fifo_get_state(srcdsp_fifo_t f, size_t *write_ptr, uint64_t *time_start, uint64_t *time_end, int *rollover)
try {
    if (!f) return fail(SRCDSP_E_INVALID, "fifo: null handle");
    std::lock_guard<std::mutex> lk(f->mx);
    if (write_ptr) *write_ptr = f->writePtr;
    if (time_start) *time_start = f->timeStart;
    if (time_end) *time_end = f->timeEnd;
    if (rollover) *rollover = f->rolloverFlag;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

/* test hook: place the time counters near the 64-bit rollover (the reference has no such entry point; its
 * rollover branches are otherwise unreachable in a test) */
int srcdsp_fifo_set_time(srcdsp_fifo_t f, uint64_t time_start, uint64_t time_end)
try {
    if (!f) return fail(SRCDSP_E_INVALID, "fifo: null handle");
    std::lock_guard<std::mutex> lk(f->mx);
    f->timeStart = time_start;
    f->timeEnd = time_end;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

const void *srcdsp_fifo_storage(srcdsp_fifo_t f) { return f ? f->storage : nullptr; }

}  // extern "C"
