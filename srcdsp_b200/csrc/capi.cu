// capi.cu -- host side of libsrcdsp_b200.so: the extern "C" layer declared in
// include/srcdsp_b200.h.  Owns device state (taps, carried history, NCO phase), picks kernel
// instantiations, and stages host buffers through a chunked H2D -> kernel -> D2H pipeline.
// There is no CPU compute path in this file: every step launches sm_100a kernels or fails.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_dec.cuh"
#include "kernels_dec_tc.cuh"
#include "kernels_dec_tma.cuh"
#include "kernels_dec_band.cuh"
#include "kernels_up.cuh"
#include "kernels_up_tc.cuh"
#include "kernels_corr.cuh"

namespace srcdsp {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
std::string &last_error_ref()
{
    static thread_local std::string s;
    return s;
}

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

int abi_exception() noexcept
{
    try {
        throw;  // re-throw the exception the entry point's handler caught
    } catch (const std::bad_alloc &) {
        try {
            return fail(SRCDSP_E_NOMEM, "out of host memory");
        } catch (...) {
            return SRCDSP_E_NOMEM;
        }
    } catch (const std::exception &e) {
        try {
            return fail(SRCDSP_E_INVALID, "unexpected C++ exception: %s", e.what());
        } catch (...) {
            return SRCDSP_E_INVALID;
        }
    } catch (...) {
        return SRCDSP_E_INVALID;
    }
}

static std::atomic<uint64_t> g_launches{0};
static inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// tuning knobs
// ---------------------------------------------------------------------------------------------
// Development sweeps override the measured defaults through SRCDSP_* environment variables.  A bank reads
// them ONCE, when it is created (Bank::init) -- never on the step path -- and none of them changes a result:
// the timing-experiment kernel instantiations that do (SRCDSP_TC_DEBUG, SRCDSP_UPT_DEBUG) exist only in
// builds with -DSRCDSP_TIMING_EXPERIMENTS; the shipped library ignores both variables.
struct Tuning {
    int tma_stages = 0;    // SRCDSP_TMA_STAGES   byte-plane stages of dec_tma_kernel (0: groups + 1)
    int tma_p2 = -1;       // SRCDSP_TMA_P2       force the 2-digit tile geometry on (1) / off (0)
    int tma_grouped = -1;  // SRCDSP_TMA_GROUPED  force the grouped (1) / interleaved (0) master layout
    int tma_w = 0;         // SRCDSP_TMA_W        converter warps per group (4 / 8)
    int tma_groups = 0;    // SRCDSP_TMA_GROUPS   converter groups
    int tma_raw = 0;       // SRCDSP_TMA_RAW      cap on the raw ring
    int no_tma = 0;        // SRCDSP_NO_TMA       register-staged dec_tc_kernel instead of dec_tma_kernel
    int epi_sleep = -1;    // SRCDSP_EPI_SLEEP    ns between epilogue polls
    int conv_sleep = -1;   // SRCDSP_CONV_SLEEP   ns between converter polls of the raw ring
    int mma_sleep = -1;    // SRCDSP_MMA_SLEEP    ns between MMA-warp polls of the byte-plane ring
    int pf_dist = -1;      // SRCDSP_PF_DIST      dec_tc_kernel's L2 prefetch distance
    int verbose = 0;       // SRCDSP_TMA_VERBOSE  print the chosen shared-memory plan
    int seq_period = -1;   // SRCDSP_SEQ_PERIOD   0: always the full-table oscillator sequence
    int band = -1;         // SRCDSP_BAND         0: never the band form (dec_band_kernel), 1: whenever applicable
    int up_tc = -1;        // SRCDSP_UP_TC        tcgen05 interpolator on (1) / off (0)
    int up_tc_form = 0;    // SRCDSP_UP_TC_FORM   2: operand-swapped form
    int up_generic = 0;    // SRCDSP_UP_GENERIC
    int corr_generic = 0;  // SRCDSP_CORR_GENERIC
    int decf_pairs = 0;    // SRCDSP_DECF_PAIRS   1 / 2
    int decf_tile = 0;     // SRCDSP_DECF_TILE
    int decf_ct = -1;      // SRCDSP_DECF_CT
    int decf_blocks = -1;  // SRCDSP_DECF_BLOCKS
    int decf_prefetch = -1;  // SRCDSP_DECF_PREFETCH
    int decf_quad = -1;    // SRCDSP_DECF_QUAD    0: pairs of outputs per thread only
    int tc_debug = 0;      // SRCDSP_TC_DEBUG / SRCDSP_UPT_DEBUG: timing experiments, -DSRCDSP_TIMING_EXPERIMENTS builds only
    int upt_debug = 0;
};
static Tuning tuning_from_env()
{
    Tuning v;
    auto get = [](const char *name, int &dst) {
        if (const char *e = getenv(name)) dst = atoi(e);
    };
    get("SRCDSP_TMA_STAGES", v.tma_stages);
    get("SRCDSP_TMA_P2", v.tma_p2);
    get("SRCDSP_TMA_GROUPED", v.tma_grouped);
    get("SRCDSP_TMA_W", v.tma_w);
    get("SRCDSP_TMA_GROUPS", v.tma_groups);
    get("SRCDSP_TMA_RAW", v.tma_raw);
    v.no_tma = getenv("SRCDSP_NO_TMA") != nullptr;
    get("SRCDSP_EPI_SLEEP", v.epi_sleep);
    get("SRCDSP_CONV_SLEEP", v.conv_sleep);
    get("SRCDSP_MMA_SLEEP", v.mma_sleep);
    get("SRCDSP_PF_DIST", v.pf_dist);
    v.verbose = getenv("SRCDSP_TMA_VERBOSE") != nullptr;
    get("SRCDSP_SEQ_PERIOD", v.seq_period);
    get("SRCDSP_BAND", v.band);
    get("SRCDSP_UP_TC", v.up_tc);
    get("SRCDSP_UP_TC_FORM", v.up_tc_form);
    v.up_generic = getenv("SRCDSP_UP_GENERIC") != nullptr;
    v.corr_generic = getenv("SRCDSP_CORR_GENERIC") != nullptr;
    get("SRCDSP_DECF_PAIRS", v.decf_pairs);
    get("SRCDSP_DECF_TILE", v.decf_tile);
    get("SRCDSP_DECF_CT", v.decf_ct);
    get("SRCDSP_DECF_BLOCKS", v.decf_blocks);
    get("SRCDSP_DECF_PREFETCH", v.decf_prefetch);
    get("SRCDSP_DECF_QUAD", v.decf_quad);
#ifdef SRCDSP_TIMING_EXPERIMENTS
    get("SRCDSP_TC_DEBUG", v.tc_debug);
    get("SRCDSP_UPT_DEBUG", v.upt_debug);
#endif
    return v;
}

#define SRCDSP_LAUNCH_CHECK()                                                                     \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess)                                                                    \
            return fail(SRCDSP_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            changed = true;
        }
    }
    ~DeviceGuard()
    {
        if (changed) cudaSetDevice(prev);
    }
};

// device allocation that is freed unless its owner takes it over (error paths of the set-up code)
struct DeviceBuffer {
    void *p = nullptr;
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer()
    {
        if (p) cudaFree(p);
    }
    void *release()
    {
        void *q = p;
        p = nullptr;
        return q;
    }
};

static int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SRCDSP_E_NOGPU, "no CUDA device available (%s): libsrcdsp_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(SRCDSP_E_INVALID, "device %d out of range [0, %d)", device, n);
    cudaDeviceProp prop;
    SRCDSP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SRCDSP_E_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    return SRCDSP_OK;
}

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------------------------------------
// bank base: device, stream, host-staging pipeline
// ---------------------------------------------------------------------------------------------
constexpr int NBUF = 3;
constexpr size_t STAGE_TARGET_BYTES = 48u << 20;  // per chunk, larger of in/out

struct Bank {
    int device = 0;
    int C = 0;
    Tuning tune;  // environment overrides, captured at creation
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // A DDC chain runs all of its members on the leader's (first decimator's) stream, so that phase / history
    // ping-pong stays ordered.  The borrow is tracked: the leader's set_stream() reaches its followers, and a
    // leader that goes away -- or a chain that is destroyed -- hands every follower a private stream again.
    Bank *leader = nullptr;
    std::vector<Bank *> followers;
    // host staging
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[NBUF] = {}, ev_comp[NBUF] = {}, ev_d2h[NBUF] = {};
    uint32_t *d_in[NBUF] = {}, *d_out[NBUF] = {};
    size_t cap_in = 0, cap_out = 0;  // words per buffer

    int init(int dev, int channels)
    {
        SRCDSP_TRY(check_device(dev));
        if (channels < 1) return fail(SRCDSP_E_INVALID, "channels must be >= 1 (got %d)", channels);
        device = dev;
        C = channels;
        tune = tuning_from_env();
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
        return SRCDSP_OK;
    }

    int set_stream(void *s)
    {
        DeviceGuard g(device);
        // only an owned stream is drained and destroyed here; a borrowed one belongs to the caller or the leader
        if (own_stream) {
            SRCDSP_CUDA(cudaStreamSynchronize(stream));
            cudaStreamDestroy(stream);
        }
        own_stream = false;
        if (s == nullptr) {
            SRCDSP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            own_stream = true;
        } else {
            stream = static_cast<cudaStream_t>(s);
        }
        for (Bank *f : followers) {  // chain members keep running on the leader's stream, whatever it is now
            if (f->own_stream) {
                cudaStreamSynchronize(f->stream);
                cudaStreamDestroy(f->stream);
                f->own_stream = false;
            }
            f->stream = stream;
        }
        return SRCDSP_OK;
    }

    // chain membership (srcdsp_ddc_create / _destroy)
    void follow(Bank *l)
    {
        unfollow();
        if (own_stream) {
            DeviceGuard g(device);
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
            own_stream = false;
        }
        leader = l;
        stream = l->stream;
        l->followers.push_back(this);
    }
    void unfollow()
    {
        if (!leader) return;
        auto &v = leader->followers;
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == this) {
                v.erase(v.begin() + i);
                break;
            }
        leader = nullptr;
        DeviceGuard g(device);
        cudaStreamSynchronize(stream);  // the leader's stream is still alive here
        own_stream = false;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess) own_stream = true;
    }

    int sync()
    {
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaStreamSynchronize(stream));
        return SRCDSP_OK;
    }

    int ensure_staging(size_t words_in, size_t words_out)
    {
        if (!s_h2d) {
            SRCDSP_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
            SRCDSP_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
            for (int k = 0; k < NBUF; ++k) {
                SRCDSP_CUDA(cudaEventCreateWithFlags(&ev_h2d[k], cudaEventDisableTiming));
                SRCDSP_CUDA(cudaEventCreateWithFlags(&ev_comp[k], cudaEventDisableTiming));
                SRCDSP_CUDA(cudaEventCreateWithFlags(&ev_d2h[k], cudaEventDisableTiming));
            }
        }
        if (words_in > cap_in) {
            for (int k = 0; k < NBUF; ++k) {
                if (d_in[k]) cudaFree(d_in[k]);
                d_in[k] = nullptr;
                SRCDSP_CUDA(cudaMalloc(&d_in[k], words_in * 4));
            }
            cap_in = words_in;
        }
        if (words_out > cap_out) {
            for (int k = 0; k < NBUF; ++k) {
                if (d_out[k]) cudaFree(d_out[k]);
                d_out[k] = nullptr;
                SRCDSP_CUDA(cudaMalloc(&d_out[k], words_out * 4));
            }
            cap_out = words_out;
        }
        return SRCDSP_OK;
    }

    void release()
    {
        DeviceGuard g(device);
        // leave the chains this bank takes part in BEFORE its stream goes away
        while (!followers.empty()) followers.back()->unfollow();
        if (leader) {
            auto &v = leader->followers;
            for (size_t i = 0; i < v.size(); ++i)
                if (v[i] == this) {
                    v.erase(v.begin() + i);
                    break;
                }
            leader = nullptr;
            stream = nullptr;  // borrowed: nothing to destroy
            own_stream = false;
        }
        // a borrowed stream may already have been destroyed by its owner; cudaFree below
        // synchronises the device before releasing memory, so no work can still touch it
        if (own_stream && stream) cudaStreamSynchronize(stream);
        for (int k = 0; k < NBUF; ++k) {
            if (d_in[k]) cudaFree(d_in[k]);
            if (d_out[k]) cudaFree(d_out[k]);
            if (ev_h2d[k]) cudaEventDestroy(ev_h2d[k]);
            if (ev_comp[k]) cudaEventDestroy(ev_comp[k]);
            if (ev_d2h[k]) cudaEventDestroy(ev_d2h[k]);
        }
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        if (own_stream && stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

// Runs `fn(d_in, in_stride, n_chunk, d_out, out_stride, last)` over consecutive chunks of a host
// buffer.  Outputs per chunk: n_chunk * num / den (+ tail_out on the last chunk).
template <class Fn>
static int staged_run(Bank &b, const int16_t *in, size_t in_stride, size_t n_in, int16_t *out,
                      size_t out_stride, size_t quantum, size_t num, size_t den, size_t tail_out, Fn fn)
{
    DeviceGuard g(b.device);
    const size_t C = (size_t)b.C;
    // chunk length in input samples per channel
    const size_t per_in = 4 * C, per_out = 4 * C * num / den + 1;
    size_t chunk = STAGE_TARGET_BYTES / (per_in > per_out ? per_in : per_out);
    if (chunk >= n_in) chunk = n_in;
    chunk -= chunk % quantum;
    if (chunk == 0) chunk = (n_in < quantum) ? n_in : quantum;
    const size_t n_chunks = n_in == 0 ? 1 : (n_in + chunk - 1) / chunk;
    const size_t in_pitch = (chunk + 3) & ~(size_t)3;                           // words, 16-byte rows
    const size_t out_pitch = ((chunk * num / den + tail_out) + 3) & ~(size_t)3;
    SRCDSP_TRY(b.ensure_staging(in_pitch * C + 4, out_pitch * C + 4));
    // order the copy streams behind whatever is already queued on the compute stream
    SRCDSP_CUDA(cudaEventRecord(b.ev_comp[0], b.stream));
    SRCDSP_CUDA(cudaStreamWaitEvent(b.s_h2d, b.ev_comp[0], 0));

    const uint32_t *hin = reinterpret_cast<const uint32_t *>(in);
    uint32_t *hout = reinterpret_cast<uint32_t *>(out);
    for (size_t i = 0; i < n_chunks; ++i) {
        const int k = (int)(i % NBUF);
        const size_t off = i * chunk;
        const size_t len = (off + chunk <= n_in) ? chunk : n_in - off;
        const bool last = (i + 1 == n_chunks);
        const size_t ooff = off * num / den;
        const size_t olen = len * num / den + (last ? tail_out : 0);
        if (i >= NBUF) {
            SRCDSP_CUDA(cudaStreamWaitEvent(b.s_h2d, b.ev_comp[k], 0));   // d_in[k] consumed
            SRCDSP_CUDA(cudaStreamWaitEvent(b.stream, b.ev_d2h[k], 0));   // d_out[k] drained
        }
        if (len)
            SRCDSP_CUDA(cudaMemcpy2DAsync(b.d_in[k], in_pitch * 4, hin + off, in_stride * 4, len * 4, C,
                                          cudaMemcpyHostToDevice, b.s_h2d));
        SRCDSP_CUDA(cudaEventRecord(b.ev_h2d[k], b.s_h2d));
        SRCDSP_CUDA(cudaStreamWaitEvent(b.stream, b.ev_h2d[k], 0));
        SRCDSP_TRY(fn(b.d_in[k], in_pitch, len, b.d_out[k], out_pitch, last));
        SRCDSP_CUDA(cudaEventRecord(b.ev_comp[k], b.stream));
        SRCDSP_CUDA(cudaStreamWaitEvent(b.s_d2h, b.ev_comp[k], 0));
        if (olen)
            SRCDSP_CUDA(cudaMemcpy2DAsync(hout + ooff, out_stride * 4, b.d_out[k], out_pitch * 4, olen * 4, C,
                                          cudaMemcpyDeviceToHost, b.s_d2h));
        SRCDSP_CUDA(cudaEventRecord(b.ev_d2h[k], b.s_d2h));
    }
    SRCDSP_CUDA(cudaStreamSynchronize(b.s_d2h));
    SRCDSP_CUDA(cudaStreamSynchronize(b.stream));
    return SRCDSP_OK;
}

static int classify(const void *in, const void *out, bool *on_device)
{
    const bool di = is_device_ptr(in), dout = is_device_ptr(out);
    if (di != dout)
        return fail(SRCDSP_E_INVALID, "input and output must both be device or both be host pointers");
    *on_device = di;
    return SRCDSP_OK;
}

// Device buffers of a filter step must not overlap: a CTA reads halo samples that a neighbouring CTA may already have
// overwritten, and the history kernel reads the input after the main kernel has written the output.  (The reference
// documents FilterFir::step as usable in place; through this ABI that holds for HOST buffers, which are staged, and
// for the mixer, which is element-wise.)  Ranges are [p, p + ((C - 1) * stride + n) * bytes).
static bool device_ranges_overlap(const void *a, size_t a_stride, size_t a_n, const void *b, size_t b_stride, size_t b_n, int C,
                                  size_t bytes_per_sample)
{
    const uintptr_t a0 = (uintptr_t)a, a1 = a0 + ((size_t)(C - 1) * a_stride + a_n) * bytes_per_sample;
    const uintptr_t b0 = (uintptr_t)b, b1 = b0 + ((size_t)(C - 1) * b_stride + b_n) * bytes_per_sample;
    return a0 < b1 && b0 < a1;
}

static inline bool aligned16(const void *p, size_t stride_words)
{
    return ((uintptr_t)p % 16 == 0) && (stride_words % 4 == 0);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library links the
// static CUDA runtime only, not libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder()
{
    static TensorMapEncodeFn fn = []() -> TensorMapEncodeFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (TensorMapEncodeFn)p;
    }();
    return fn;
}

// ---------------------------------------------------------------------------------------------
// mixer bank
// ---------------------------------------------------------------------------------------------
struct MixerBank : Bank {
    unsigned n_table = 4096;
    uint32_t *d_cs = nullptr;  // packed (cos, sin)
    int *d_phi[2] = {nullptr, nullptr};
    int *d_freq = nullptr;
    int cur = 0;
    bool dirty = true;
    std::vector<int> h_phi, h_freq;
    std::vector<float> h_nominal;
    unsigned seq_period = 0;  // common period of the channels' oscillator sequences (power-of-two tables), see upload_if_dirty

    PhaseMod pm() const
    {
        PhaseMod m;
        m.n_table = n_table;
        m.mask = (n_table & (n_table - 1)) == 0 ? n_table - 1 : 0;
        return m;
    }

    // _Mixer::setFrequency, mixers.h:51-67 (float arithmetic, round half away from zero)
    static int quantise(float lo, unsigned N)
    {
        int16_t f;
        if (lo >= 0)
            f = static_cast<int16_t>(roundf(lo * N / 2));
        else {
            f = static_cast<int16_t>(roundf(N - roundf(-lo * N / 2)));
            if (f == static_cast<int16_t>(N)) f = 0;
        }
        return f;
    }

    int create(int dev, int channels, unsigned N)
    {
        if (N < 4 || N > 16384) return fail(SRCDSP_E_INVALID, "n_table must be in [4, 16384] (int16 phase), got %u", N);
        SRCDSP_TRY(init(dev, channels));
        n_table = N;
        DeviceGuard g(device);
        // Mixer ctor, mixers.h:149-160: T[k] = (int16) (16383 * sin(2 pi k / N)), truncating cast
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944592307816406286;
        std::vector<int16_t> T(N);
        const int16_t mx = (INT16_MAX >> 1);
        for (unsigned k = 0; k < N; ++k) T[k] = static_cast<int16_t>(mx * sin(2 * pi * static_cast<double>(k) / N));
        std::vector<uint32_t> cs(N);
        for (unsigned k = 0; k < N; ++k)
            cs[k] = (uint32_t)(uint16_t)T[(k + N / 4) % N] | ((uint32_t)(uint16_t)T[k] << 16);
        SRCDSP_CUDA(cudaMalloc(&d_cs, N * 4));
        SRCDSP_CUDA(cudaMemcpy(d_cs, cs.data(), N * 4, cudaMemcpyHostToDevice));
        SRCDSP_CUDA(cudaMalloc(&d_phi[0], C * sizeof(int)));
        SRCDSP_CUDA(cudaMalloc(&d_phi[1], C * sizeof(int)));
        SRCDSP_CUDA(cudaMalloc(&d_freq, C * sizeof(int)));
        h_phi.assign(C, 0);
        h_freq.assign(C, 0);
        h_nominal.assign(C, 0.f);
        dirty = true;
        return SRCDSP_OK;
    }

    int upload_if_dirty()
    {
        if (!dirty) return SRCDSP_OK;
        // The oscillator of a channel, T[(phi0 + n * freq) mod N] in time order, has period N / gcd(freq, N).  For a
        // power-of-two table every such period is a power of two, so the largest one is a period of all channels: the
        // fused mixer keeps that many sequence entries in shared memory instead of N (e.g. 512 for frequencies on a
        // grid of 8 table steps) and gives the rest to the sample rings.
        seq_period = n_table;
        if ((n_table & (n_table - 1)) == 0) {
            unsigned g = n_table;  // gcd of all frequencies and N
            for (int c = 0; c < C; ++c) {
                const unsigned f = (unsigned)h_freq[c] & (n_table - 1);
                if (f) g = std::min(g, f & (~f + 1u));  // lowest set bit = gcd(f, 2^k)
            }
            seq_period = std::max(16u, n_table / g);  // 128 bytes at least: the sample rings behind it stay 128-byte aligned
        }
        // pageable sources: the driver stages them before returning, so the vectors may change
        SRCDSP_CUDA(cudaMemcpyAsync(d_phi[cur], h_phi.data(), C * sizeof(int), cudaMemcpyHostToDevice, stream));
        SRCDSP_CUDA(cudaMemcpyAsync(d_freq, h_freq.data(), C * sizeof(int), cudaMemcpyHostToDevice, stream));
        dirty = false;
        return SRCDSP_OK;
    }

    // host mirror of "phi advanced by n samples" (mixers.h:177 applied n times)
    void advance(size_t n)
    {
        const unsigned long long nm = n % n_table;
        for (int c = 0; c < C; ++c)
            h_phi[c] = (int)(((unsigned long long)h_phi[c] + nm * (unsigned long long)h_freq[c]) % n_table);
        cur ^= 1;
    }

    int step_device(const uint32_t *in, size_t in_stride, uint32_t *out, size_t out_stride, size_t n)
    {
        DeviceGuard g(device);
        SRCDSP_TRY(upload_if_dirty());
        const bool vec = aligned16(in, in_stride) && aligned16(out, out_stride);
        long long work = vec ? (long long)((n + 3) / 4) : (long long)n;
        int bx = (int)std::min<long long>((work + 255) / 256, (148ll * 16 + C - 1) / C);
        if (bx < 1) bx = 1;
        dim3 grid(bx, C);
        const PhaseMod m = pm();
        if (vec && m.mask && n_table <= 8192 && n >= 4 * (size_t)n_table) {
            // power-of-two table and a block worth building the per-channel oscillator sequence for
            const int per_ch = (int)std::max<long long>(1, std::min<long long>((148ll * 6 + C - 1) / C, (long long)(n / (4 * (size_t)n_table))));
            const size_t smem = (size_t)n_table * 8;
            if (smem > 48 * 1024) SRCDSP_CUDA(cudaFuncSetAttribute(mixer_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mixer_seq_kernel<<<dim3(per_ch, C), MIXSEQ_THREADS, smem, stream>>>(in, in_stride, out, out_stride, (long long)n, d_cs,
                                                                               d_phi[cur], d_phi[cur ^ 1], d_freq, m.mask);
        } else if (vec)
            mixer_kernel<true><<<grid, 256, 0, stream>>>(in, in_stride, out, out_stride, (long long)n, d_cs,
                                                         d_phi[cur], d_phi[cur ^ 1], d_freq, pm());
        else
            mixer_kernel<false><<<grid, 256, 0, stream>>>(in, in_stride, out, out_stride, (long long)n, d_cs,
                                                          d_phi[cur], d_phi[cur ^ 1], d_freq, pm());
        SRCDSP_LAUNCH_CHECK();
        count_launch();
        advance(n);
        return SRCDSP_OK;
    }

    void destroy()
    {
        DeviceGuard g(device);
        release();
        if (d_cs) cudaFree(d_cs);
        if (d_phi[0]) cudaFree(d_phi[0]);
        if (d_phi[1]) cudaFree(d_phi[1]);
        if (d_freq) cudaFree(d_freq);
    }
};

// ---------------------------------------------------------------------------------------------
// decimator bank
// ---------------------------------------------------------------------------------------------
struct DecBank : Bank {
    int M = 1;
    int ntaps = 0;
    int H = 0;  // ntaps - 1
    int Qp = 0, JP = 0;
    int coeff_scaling = 0;
    int left_shift = 0;
    int kernel_kind = 0;
    int last_kernel = 0;
    std::vector<int32_t> taps;
    int32_t *d_taps_poly = nullptr;
    uint32_t *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    size_t smem_bytes = 0;
    int nt_threads = DEC_NT;  // threads per CTA of the FIR kernel; tile = 8 * nt_threads outputs
    // tcgen05 int8 Toeplitz path (kernels_dec_tc.cuh)
    bool tc_ok = false;
    const char *tc_why = "no coefficients";
    // register-staged variant / TMA variant / TMA variant with the fused mixer (each with its own master layout)
    TcParams tc{}, tc_tma{}, tc_tma_mix{};
    bool tma_ok = false, tma_mix_ok = false;
    std::vector<std::vector<int8_t>> tc_dig;  // signed base-256 digits of the taps
    int tc_P = 1;                             // digits per tap
    bool tc_want_p2 = false;
    int tc_mix_table = 32768;                 // shared-memory bytes of the oscillator sequence the mixer variant was planned for
    int build_variant(int variant, bool p2, int mix_table);
    // band form (kernels_dec_band.cuh): samples on the M side, a band of the taps on the N side, lags merged
    TcParams tc_band{};
    BandExtra band{};
    bool band_ok = false;
    uint8_t *d_master_band = nullptr;
    size_t band_fixed = 0;
    int build_band();
    int band_layout(int table_bytes, int groups, int *n_raw, int *n_stages, size_t *smem) const
    {
        const size_t avail = (size_t)226 * 1024;
        if (band_fixed + (size_t)table_bytes + 2 * (size_t)BAND_STAGE_BYTES + 2 * (size_t)BAND_RAW_BYTES > avail) return SRCDSP_E_SIZE;
        int ns = groups + 1;
        if (tune.tma_stages > 0) ns = std::max(2, std::min(tune.tma_stages, TC_MAX_STAGES));
        while (ns > 2 && band_fixed + (size_t)table_bytes + (size_t)ns * BAND_STAGE_BYTES + 4 * (size_t)BAND_RAW_BYTES > avail) --ns;
        int nr = (int)((avail - band_fixed - (size_t)table_bytes - (size_t)ns * BAND_STAGE_BYTES) / BAND_RAW_BYTES);
        if (nr > 8) nr = 8;
        if (nr < 2) return SRCDSP_E_SIZE;
        if (n_raw) *n_raw = nr;
        if (n_stages) *n_stages = ns;
        if (smem) *smem = band_fixed + (size_t)table_bytes + (size_t)ns * BAND_STAGE_BYTES + (size_t)nr * BAND_RAW_BYTES;
        return SRCDSP_OK;
    }
    uint8_t *d_master = nullptr, *d_master_tma = nullptr, *d_master_tma_mix = nullptr;
    size_t tc_fixed_tma = 0, tc_fixed_tma_mix = 0;
    int *d_error = nullptr;  // wait-cycle counters of the timing variants (device)
    int *h_diag = nullptr;   // mapped host memory: record of a timed-out barrier wait (survives the trap)
    int *d_diag = nullptr;
    size_t tc_fixed = 0;  // master + barriers
    int sm_count = 148;

    int prepare_tc();
    // shared-memory plan of the tensor-core kernel for a given NCO table size (0: no fused mixer)
    int tc_layout(int table_bytes, int *n_stages, size_t *smem) const
    {
        const size_t avail = (size_t)226 * 1024;
        if (tc_fixed + (size_t)table_bytes > avail) return SRCDSP_E_SIZE;
        int ns = (int)((avail - tc_fixed - (size_t)table_bytes) / ((size_t)64 * tc.rbp));
        if (ns < TC_OWNERS + 1) return SRCDSP_E_SIZE;
        if (ns > TC_MAX_STAGES) ns = TC_MAX_STAGES;
        if (n_stages) *n_stages = ns;
        if (smem) *smem = tc_fixed + (size_t)table_bytes + (size_t)ns * 64 * tc.rbp;
        return SRCDSP_OK;
    }

    // shared-memory plan of the TMA-fed variant: byte-plane stages only decouple the converters from
    // the MMAs (3 are enough), everything else goes to the raw ring = the bytes in flight from HBM
    int tma_layout(const TcParams &T, size_t fixed, int table_bytes, int groups, int *n_raw, int *n_stages, size_t *smem) const
    {
        const size_t avail = (size_t)226 * 1024;
        const size_t split = (size_t)64 * T.rbp;
        const size_t rawb = (size_t)(4 * ((T.J - 1 + 3) / 4) + T.nrb) * 128;
        if (fixed + (size_t)table_bytes + 2 * split + 2 * rawb > avail) return SRCDSP_E_SIZE;
        // byte-plane stages: one per converter group being filled + 1 for the MMAs; the rest of the
        // shared memory is the raw ring = the bytes in flight from HBM (at least 4 stages wanted)
        int ns = groups + 1;
        if (tune.tma_stages > 0) ns = std::max(2, std::min(tune.tma_stages, TC_MAX_STAGES));
        while (ns > 2 && fixed + (size_t)table_bytes + ns * split + 4 * rawb > avail) --ns;
        int nr = (int)((avail - fixed - (size_t)table_bytes - ns * split) / rawb);
        if (nr > 8) nr = 8;  // more bytes in flight than ~128 KB per SM lowers the HBM rate (tools/tmabench.cu)
        if (nr < 2) return SRCDSP_E_SIZE;
        if (n_raw) *n_raw = nr;
        if (n_stages) *n_stages = ns;
        if (smem) *smem = fixed + (size_t)table_bytes + ns * split + nr * rawb;
        return SRCDSP_OK;
    }

    int set_coeffs(const int32_t *t, int n, int require_multiple)
    {
        if (!t || n < 1) return fail(SRCDSP_E_SIZE, "ntaps must be >= 1");
        if (require_multiple && n % M != 0)
            return fail(SRCDSP_E_SIZE, "ntaps (%d) must be a multiple of M (%d) [dsptl_dnsampling_filters.h:122]", n, M);
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaStreamSynchronize(stream));
        // coeffScaling: dsptl_dnsampling_filters.h:126-132
        double sum = 0;
        for (int k = 0; k < n; ++k) sum += std::abs(t[k]);
        if (!(sum >= 1)) return fail(SRCDSP_E_SIZE, "sum |taps| must be >= 1 (coeffScaling = floor(log2(sum)))");
        const int cs = static_cast<int>(floor(log2(sum)));

        const int newH = n - 1;
        const int q = (n + M - 1) / M;
        const int newQp = (q + DEC_QC - 1) / DEC_QC * DEC_QC;
        // tile size: the largest CTA (128/64/32 threads x 8 outputs) whose polyphase-transposed
        // tile lets three CTAs share an SM; large M falls back to whatever still fits 227 KB
        int nthr = 0, newJP = 0;
        size_t newSmem = 0;
        for (int pass = 0; pass < 2 && !nthr; ++pass) {
            for (int cand = DEC_NT; cand >= 32; cand >>= 1) {
                int jp = (cand * DEC_R + newQp + 7) / 8 * 8;
                while (jp % 32 != 8) jp += 8;
                const size_t sm = ((size_t)M * newQp + (size_t)M * jp) * 4;
                if (sm <= (pass == 0 ? (size_t)72 * 1024 : (size_t)227 * 1024)) {
                    nthr = cand, newJP = jp, newSmem = sm;
                    break;
                }
            }
        }
        if (!nthr)
            return fail(SRCDSP_E_SIZE, "M=%d with %d taps does not fit the 227 KB shared-memory tile", M, n);
        std::vector<int32_t> poly((size_t)M * newQp, 0);
        for (int k = 0; k < n; ++k) poly[(size_t)(k % M) * newQp + k / M] = t[k];
        // the new buffers are owned locally until everything that can fail has succeeded: a failed call leaves the
        // bank as it was and leaks nothing
        DeviceBuffer nd, nh[2];
        SRCDSP_CUDA(cudaMalloc(&nd.p, poly.size() * 4));
        SRCDSP_CUDA(cudaMemcpy(nd.p, poly.data(), poly.size() * 4, cudaMemcpyHostToDevice));
        // history.resize(n-1) (dsptl_dnsampling_filters.h:127): keeps the first min(old,new)
        // entries, value-initialises the rest
        const size_t hw = (size_t)C * (newH > 0 ? newH : 1);
        for (int k = 0; k < 2; ++k) {
            SRCDSP_CUDA(cudaMalloc(&nh[k].p, hw * 4));
            SRCDSP_CUDA(cudaMemset(nh[k].p, 0, hw * 4));
        }
        if (d_hist[cur] && H > 0 && newH > 0) {
            const int keep = H < newH ? H : newH;
            SRCDSP_CUDA(cudaMemcpy2D(nh[0].p, (size_t)newH * 4, d_hist[cur], (size_t)H * 4, (size_t)keep * 4, C,
                                     cudaMemcpyDeviceToDevice));
        }
        if (d_taps_poly) cudaFree(d_taps_poly);
        if (d_hist[0]) cudaFree(d_hist[0]);
        if (d_hist[1]) cudaFree(d_hist[1]);
        d_taps_poly = static_cast<int32_t *>(nd.release());
        d_hist[0] = static_cast<uint32_t *>(nh[0].release());
        d_hist[1] = static_cast<uint32_t *>(nh[1].release());
        cur = 0;
        taps.assign(t, t + n);
        ntaps = n;
        H = newH;
        Qp = newQp;
        JP = newJP;
        nt_threads = nthr;
        smem_bytes = newSmem;
        coeff_scaling = cs;
        left_shift = 0;
        return prepare_tc();
    }

    int reset()
    {
        if (!d_hist[0]) return SRCDSP_OK;
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaMemsetAsync(d_hist[cur], 0, (size_t)C * (H > 0 ? H : 1) * 4, stream));
        return SRCDSP_OK;
    }

    int shift_now(unsigned *s) const
    {
        const int sh = coeff_scaling - left_shift;
        if (sh < 0 || sh > 31)
            return fail(SRCDSP_E_STATE, "coeffScaling - leftShift = %d is outside [0, 31] (undefined in the reference)", sh);
        *s = (unsigned)sh;
        return SRCDSP_OK;
    }

    // mixer == nullptr: plain decimator.  Device pointers only.
    int step_device(const uint32_t *in, size_t in_stride, size_t n_in, uint32_t *out, size_t out_stride,
                    MixerBank *mixer);

    void destroy()
    {
        DeviceGuard g(device);
        release();
        if (d_taps_poly) cudaFree(d_taps_poly);
        if (d_hist[0]) cudaFree(d_hist[0]);
        if (d_hist[1]) cudaFree(d_hist[1]);
        if (d_master) cudaFree(d_master);
        if (d_master_tma) cudaFree(d_master_tma);
        if (d_master_tma_mix) cudaFree(d_master_tma_mix);
        if (d_master_band) cudaFree(d_master_band);
        if (d_error) cudaFree(d_error);
        if (h_diag) cudaFreeHost(h_diag);
    }
};

// One layout variant of the tensor-core decimator (see prepare_tc): 0 = dec_tc_kernel, 1 = dec_tma_kernel,
// 2 = dec_tma_kernel with the fused mixer, whose shared-memory plan depends on the oscillator sequence's size.
int DecBank::build_variant(int variant, bool p2, int mix_table)
{
    const std::vector<std::vector<int8_t>> &dig = tc_dig;
    const int P = tc_P;
    uint8_t **pimg = variant == 2 ? &d_master_tma_mix : variant ? &d_master_tma : &d_master;
    if (*pimg) {
        DeviceGuard g0(device);
        cudaFree(*pimg);
        *pimg = nullptr;
    }
    if (variant == 2) tc_mix_table = mix_table;
    TcParams &T = variant == 2 ? tc_tma_mix : variant ? tc_tma : tc;
    uint8_t *&d_img = variant == 2 ? d_master_tma_mix : variant ? d_master_tma : d_master;
    size_t &fixed = variant == 2 ? tc_fixed_tma_mix : variant ? tc_fixed_tma : tc_fixed;
    bool &ok = variant == 2 ? tma_mix_ok : variant ? tma_ok : tc_ok;
    ok = false;
    const int nrb = p2 ? 64 : TC_NRB, bout = p2 ? 64 : TC_BOUT, S = p2 ? 2 : 4;
    const int G = bout * M, ksteps = G / 32;
    const int J = 1 + (ntaps - 1 + G - 1) / G;
    const int OFF = bout;  // a = (32 * kc) / M < bout
    const int front_pad = 2 * (4 * ((J - 1 + 3) / 4) - (J - 1));
    const int rbp = (front_pad + 2 * (nrb + J - 1)) | 1;  // odd: conflict-free byte-plane stores
    T = TcParams{};
    T.rbp = rbp, T.J = J, T.nrb = nrb;  // the layout planners need these
    std::vector<int> copy_of_kc(ksteps);
    std::vector<std::pair<int, int>> copies;  // (s, r)
    int grouped = 1, a_rows = 0;
    size_t master_bytes = 0, image_bytes = 0;
    for (; grouped >= (p2 ? 1 : 0); --grouped) {
        copies.clear();
        for (int kc = 0; kc < ksteps; ++kc) {
            const int a = (32 * kc) / M, r = (32 * kc) % M;
            const std::pair<int, int> key(grouped ? a % 8 : 0, r);
            int idx = -1;
            for (size_t i = 0; i < copies.size(); ++i)
                if (copies[i] == key) idx = (int)i;
            if (idx < 0) {
                idx = (int)copies.size();
                copies.push_back(key);
            }
            copy_of_kc[kc] = idx;
        }
        // rows: guard + S * OFF (shift range) + 128 per lag + slack
        a_rows = grouped ? 128 * J + 8 * S + S * OFF + 64 : 128 * J + 136;
        image_bytes = copies.size() * (size_t)a_rows * 32;
        // the MMA plan behind the image: one 8-byte header per K-step, then one 16-byte entry per (K-step, lag) + 1 pad
        master_bytes = image_bytes + (((size_t)ksteps * 8 + 15) & ~(size_t)15) + ((size_t)ksteps * J + 1) * 16;
        fixed = ((master_bytes + 127) & ~(size_t)127) + 512;  // + mbarriers (at most 416 B)
        if (variant) {
            // grouped (shuffle-free epilogue, but several master copies) when enough raw stages remain: 4 for the
            // plain decimator; 5 next to the fused mixer's oscillator sequence (32 KB for N = 4096), whose
            // converters are slower and need the deeper ring more than the cheaper epilogue (sweeps on ddc16 / ddc8:
            // /16 is faster interleaved with 6 raw stages, /8 grouped with 5).  Otherwise the interleaved layout,
            // when it fits at all.
            int nr = 0;
            const int table = variant == 2 ? mix_table : 0, want = variant == 2 ? 5 : 4;
            if (tune.tma_grouped >= 0) {  // tuning override
                if (((tune.tma_grouped != 0) || p2) == (grouped != 0) && tma_layout(T, fixed, table, 1, nullptr, nullptr, nullptr) == SRCDSP_OK) break;
                continue;
            }
            if (tma_layout(T, fixed, table, variant == 2 ? 3 : 2, &nr, nullptr, nullptr) == SRCDSP_OK && (nr >= want || p2)) break;
            if (!grouped && tma_layout(T, fixed, table, 1, nullptr, nullptr, nullptr) == SRCDSP_OK) break;
        } else {
            tc.rbp = rbp;  // tc_layout reads it
            if (tc_layout(16384, nullptr, nullptr) == SRCDSP_OK) break;  // leave room for a 4096-entry sine table
            if (!grouped && tc_layout(0, nullptr, nullptr) == SRCDSP_OK) break;
        }
    }
    if (grouped < (p2 ? 1 : 0)) {
        if (!variant) tc_why = "Toeplitz master + stages exceed the shared memory of one SM";
        return SRCDSP_OK;
    }
    std::vector<uint8_t> img(master_bytes, 0);
    const int guard = grouped ? 8 * S : 4;
    for (size_t ci = 0; ci < copies.size(); ++ci) {
        uint8_t *base = img.data() + ci * (size_t)a_rows * 32;
        const int s8 = copies[ci].first, r = copies[ci].second;
        for (int row = guard; row < a_rows; ++row) {
            int v, w;
            if (grouped) {
                const int q = row - guard;
                v = 8 * (q / (8 * S)) + (q % 8);
                w = (q % (8 * S)) / 8;
            } else {
                v = (row - guard) / 4;
                w = (row - guard) % 4;
            }
            if (w >= P) continue;
            const int u = v - s8 - OFF;
            for (int t = 0; t < 32; ++t) {
                const long long k = (long long)M * u - t - r;
                if (k < 0 || k >= ntaps) continue;
                // SWIZZLE_NONE K-major image: [kc = t / 16][row][t % 16]
                base[(size_t)(t / 16) * a_rows * 16 + (size_t)row * 16 + (t % 16)] = (uint8_t)dig[w][k];
            }
        }
    }
    // per K-step: master row of (b = 0, slot 0, lag 0), v0 = OFF - a + s (a multiple of 8 when grouped), and the lags
    // whose taps reach this K-step's samples
    auto a_row_of = [&](int a) { return grouped ? S * (OFF - 8 * (a / 8)) + guard : 4 * (OFF - a) + 4; };
    auto lag_active = [&](int a, int r, int j) {
        const long long kmax = (long long)M * (bout - 1 + bout * j - a) - r;
        const long long kmin = (long long)M * (bout * j - a) - r - 31;
        return kmax >= 0 && kmin <= ntaps - 1;
    };
    // MMA plan (tc_mma_role): operand addresses >> 4, A relative to the master, B relative to the stage
    const int plan_hdr_off = (int)image_bytes;
    const int plan_ent_off = plan_hdr_off + (int)((((size_t)ksteps * 8 + 15) & ~(size_t)15));
    {
        uint32_t *hdr = reinterpret_cast<uint32_t *>(img.data() + plan_hdr_off);
        uint32_t *ent = reinterpret_cast<uint32_t *>(img.data() + plan_ent_off);
        // hi byte plane: one weight slot up = 8 rows (grouped) / 1 row (interleaved) back; p2: same rows, own columns
        const int hi_shift = p2 ? 0 : grouped ? 128 : 16;
        uint32_t n_ent = 0;
        for (int kc = 0; kc < ksteps; ++kc) {
            const int a = (32 * kc) / M, r = (32 * kc) % M;
            const int res_off = copy_of_kc[kc] * a_rows * 32;
            hdr[2 * kc] = n_ent;
            uint32_t cnt = 0;
            for (int j = 0; j < J; ++j) {
                if (!lag_active(a, r, j)) continue;
                const int a_addr = res_off + (a_row_of(a) + 128 * j) * 16;
                const int b_addr = (front_pad + 2 * (J - 1 - j)) * 16;
                ent[4 * n_ent + 0] = (uint32_t)a_addr >> 4;
                ent[4 * n_ent + 1] = (uint32_t)(a_addr - hi_shift) >> 4;
                ent[4 * n_ent + 2] = (uint32_t)b_addr >> 4;
                ent[4 * n_ent + 3] = (uint32_t)(b_addr + 2 * rbp * 16) >> 4;
                ++n_ent;
                ++cnt;
            }
            hdr[2 * kc + 1] = cnt;
        }
    }
    DeviceGuard g(device);
    SRCDSP_CUDA(cudaMalloc(&d_img, master_bytes));
    SRCDSP_CUDA(cudaMemcpy(d_img, img.data(), master_bytes, cudaMemcpyHostToDevice));
    T.M = M;
    T.G = G;
    T.p2 = p2;
    T.ksteps = ksteps;
    T.master = d_img;
    T.master_bytes = (int)master_bytes;
    T.plan_hdr_off = plan_hdr_off;
    T.plan_ent_off = plan_ent_off;
    T.a_rows = a_rows;
    T.front_pad = front_pad;
    T.grouped = grouped;
    T.error_flag = d_diag;
    T.counters = d_error;
    for (int kc = 0; kc < ksteps && kc < TC_MAX_KSTEPS; ++kc) {
        const int a = (32 * kc) / M, r = (32 * kc) % M;
        TcKstep &ks = T.ks[kc];
        ks.a_row = a_row_of(a);
        ks.res_off = copy_of_kc[kc] * a_rows * 32;
        ks.jmask = 0;
        for (int j = 0; j < J; ++j)
            if (lag_active(a, r, j)) ks.jmask |= 1u << j;
    }
    ok = true;
    return SRCDSP_OK;
}

// The band form's tap operand and MMA plan (kernels_dec_band.cuh).  One master copy per residue r = (32 kc) mod M:
// row 4 (u + UOFF) + w holds, for the 32 samples t of a K-step, the digit d_w[M u - t - r] of the tap that output
// (a + u) of the row-block applies to sample t (a = (32 kc) / M); rows outside the filter's support are zero.  The plan
// gives, per K-step, the window of outputs the K-step meets: the first master row (lo plane; the hi plane reads one row
// earlier = one weight slot up), the MMA's N and its first accumulator column.
int DecBank::build_band()
{
    band_ok = false;
    if (d_master_band) {
        DeviceGuard g0(device);
        cudaFree(d_master_band);
        d_master_band = nullptr;
    }
    const int P = tc_P;
    if (M < 2 || (M & 1) || M > TC_MAX_KSTEPS || P > 3) return SRCDSP_OK;
    const int E = ntaps >= 2 ? (ntaps - 2) / M + 1 : 0;
    if (E > 32) return SRCDSP_OK;  // the filter reaches more than one row-block back
    const int G = 32 * M;
    std::vector<int> residues, copy_of_kc(M);
    for (int kc = 0; kc < M; ++kc) {
        const int r = (32 * kc) % M;
        int idx = -1;
        for (size_t i = 0; i < residues.size(); ++i)
            if (residues[i] == r) idx = (int)i;
        if (idx < 0) {
            idx = (int)residues.size();
            residues.push_back(r);
        }
        copy_of_kc[kc] = idx;
    }
    const int a_rows = 4 * (64 + 2 * BAND_UOFF);
    const size_t copy_bytes = (size_t)a_rows * 32;
    const size_t image_bytes = residues.size() * copy_bytes;
    const size_t master_bytes = image_bytes + (size_t)M * 16;
    band_fixed = ((master_bytes + 127) & ~(size_t)127) + 2048 + 512;  // + boundary hand-over + mbarriers
    if (band_fixed + 2 * (size_t)BAND_STAGE_BYTES + 2 * (size_t)BAND_RAW_BYTES > (size_t)226 * 1024) return SRCDSP_OK;
    std::vector<uint8_t> img(master_bytes, 0);
    for (size_t ci = 0; ci < residues.size(); ++ci) {
        uint8_t *base = img.data() + ci * copy_bytes;
        const int r = residues[ci];
        for (int row = 0; row < a_rows; ++row) {
            const int u = row / 4 - BAND_UOFF, w = row % 4;
            if (w >= P) continue;
            for (int t = 0; t < 32; ++t) {
                const long long k = (long long)M * u - t - r;
                if (k < 0 || k >= ntaps) continue;
                base[(size_t)(t / 16) * a_rows * 16 + (size_t)row * 16 + (t % 16)] = (uint8_t)tc_dig[w][k];
            }
        }
    }
    uint32_t *plan = reinterpret_cast<uint32_t *>(img.data() + image_bytes);
    for (int kc = 0; kc < M; ++kc) {
        const int a = (32 * kc) / M, r = (32 * kc) % M;
        int b0, nb;
        if (kc == 0) {  // first MMA of a tile: all columns, accumulate = 0
            b0 = 0;
            nb = (32 + E + 7) & ~7;  // whole chunks of 8 outputs: the epilogue adds extended chunks without per-output checks
        } else {
            const int b_lo = a + (r > 0 ? 1 : 0);
            const int b_hi = std::min(32 + E - 1, (32 * kc + 31 + ntaps - 1) / M);
            b0 = b_lo & ~1;
            nb = std::max(4, (b_hi - b0 + 1 + 3) & ~3);
        }
        if (b0 + nb > 64) b0 = 64 - nb;
        const int u0 = b0 - a;
        if (u0 < -BAND_UOFF + 1 || u0 + nb > 64 + BAND_UOFF || 4 * nb > 256 || b0 < 0) return SRCDSP_OK;  // (cannot happen for E <= 32)
        const uint32_t addr = (uint32_t)(copy_of_kc[kc] * copy_bytes) + (uint32_t)(4 * (u0 + BAND_UOFF)) * 16;
        plan[4 * kc + 0] = addr >> 4;
        plan[4 * kc + 1] = (addr - 16) >> 4;
        plan[4 * kc + 2] = (uint32_t)(4 * nb);
        plan[4 * kc + 3] = (uint32_t)(4 * b0);
    }
    DeviceGuard g(device);
    SRCDSP_CUDA(cudaMalloc(&d_master_band, master_bytes));
    SRCDSP_CUDA(cudaMemcpy(d_master_band, img.data(), master_bytes, cudaMemcpyHostToDevice));
    TcParams &T = tc_band;
    T = TcParams{};
    T.M = M, T.G = G, T.ksteps = M, T.nrb = BAND_NEW, T.J = 2;
    T.master = d_master_band, T.master_bytes = (int)master_bytes;
    T.a_rows = a_rows, T.rbp = BAND_RBP;
    T.error_flag = d_diag, T.counters = d_error;
    band = BandExtra{};
    band.E = E, band.a_rows = a_rows, band.ks2 = M / 2, band.plan_off = (int)image_bytes;
    band_ok = true;
    return SRCDSP_OK;
}

// Builds the resident Toeplitz "master" operand and the per-K-step table of the tensor-core
// kernel (see kernels_dec_tc.cuh for the layout).  Sets tc_ok = false (with a reason) when the
// taps / ratio do not fit it; the IMAD kernel then handles every case.
int DecBank::prepare_tc()
{
    tc_ok = tma_ok = tma_mix_ok = false;
    uint8_t **imgs[3] = {&d_master, &d_master_tma, &d_master_tma_mix};
    for (auto pp : imgs)
        if (*pp) {
            cudaFree(*pp);
            *pp = nullptr;
        }
    if (M > TC_MAX_KSTEPS) { tc_why = "M > 64"; return SRCDSP_OK; }
    // signed base-256 digits of every tap: c = sum_pl 256^pl * d_pl, d_pl in [-128, 127]
    int P = 1;
    std::vector<std::vector<int8_t>> &dig = tc_dig;
    dig.assign(4, std::vector<int8_t>(ntaps, 0));
    for (int k = 0; k < ntaps; ++k) {
        long long c = taps[k];
        for (int pl = 0; pl < 4; ++pl) {
            long long d = ((c + 128) & 255) - 128;
            dig[pl][k] = (int8_t)d;
            c = (c - d) / 256;
            if (d != 0 && pl + 1 > P) P = pl + 1;
        }
        if (c != 0) P = 5;
    }
    if (P > 3) { tc_why = "taps need more than 3 signed byte digits (|c| >= 2^23)"; return SRCDSP_OK; }
    if (1 + (ntaps - 1 + 32 * M - 1) / (32 * M) > TC_MAX_J) { tc_why = "filter spans more than 16 row-blocks"; return SRCDSP_OK; }
    if (!d_error) {
        DeviceGuard g0(device);
        SRCDSP_CUDA(cudaMalloc(&d_error, 128));
        SRCDSP_CUDA(cudaMemset(d_error, 0, 128));
        SRCDSP_CUDA(cudaHostAlloc(&h_diag, 64, cudaHostAllocMapped));
        memset(h_diag, 0, 64);
        SRCDSP_CUDA(cudaHostGetDevicePointer(&d_diag, h_diag, 0));
    }
    // "p2" tile geometry for the TMA variants: when every tap fits 2 signed digits, half of the 4 weight-slot rows
    // of each MMA are zero.  With 64 outputs x 2 digit slots per row-block (row-blocks of 64*M samples, 64 of them
    // per tile) and the two byte planes accumulating into separate column halves, all 128 rows work and the filter
    // spans half as many row-blocks: ~1.8x fewer MMA cycles per output for long filters (the tensor pipe binds
    // there: /4 with 1023 taps).  Short filters lose more to the doubled K-step count than they gain, so it is
    // used from 4 lags on (SRCDSP_TMA_P2 = 0 / 1 overrides).
    bool want_p2 = P <= 2 && 2 * M <= TC_MAX_KSTEPS && 1 + (ntaps - 1 + 32 * M - 1) / (32 * M) >= 4;
    if (tune.tma_p2 >= 0) want_p2 = tune.tma_p2 != 0 && P <= 2 && 2 * M <= TC_MAX_KSTEPS;
    tc_P = P;
    tc_want_p2 = want_p2;

    // One resident "master" Toeplitz image per distinct tap alignment.  S = weight slots per output (4, or the 2
    // digit slots in p2 mode).  Preferred layout ("grouped"): row = 8*S*(v>>3) + 8*w + (v&7) (+ 8*S guard rows), the
    // slots of an output 8 rows apart, so that the epilogue needs no shuffles; a K-step whose output shift is a
    // needs the copy with s = a mod 8.  Fallback ("interleaved", 4 slots only): row = 4*v + w (+4 guard rows), one copy
    // per residue r.
    // The kernels have different shared-memory budgets (the TMA variant needs few byte-plane stages, the
    // register-staged one at least TC_OWNERS + 1), so each gets the best master layout that fits ITS budget:
    // variant 0 = dec_tc_kernel (tc), 1 = dec_tma_kernel (tc_tma), 2 = dec_tma_kernel with the fused mixer (tc_tma_mix).
    SRCDSP_TRY(build_variant(0, false, 0));
    for (int variant = 1; variant <= 2; ++variant) {
        // p2 first where it pays; the 4-slot geometry when its master does not fit next to the rings.  The mixer
        // variant is planned for a full 4096-entry oscillator sequence here and re-planned by step_device() when the
        // mixer's frequencies allow a shorter one.
        if (want_p2) SRCDSP_TRY(build_variant(variant, true, 32768));
        if (!(variant == 2 ? tma_mix_ok : tma_ok)) SRCDSP_TRY(build_variant(variant, false, 32768));
    }
    if (!tc_ok) return SRCDSP_OK;
    cudaDeviceProp prop;
    SRCDSP_CUDA(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    const int max_smem = 226 * 1024;  // 227 KB minus the kernel's static shared memory
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
#ifdef SRCDSP_TIMING_EXPERIMENTS
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tc_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tc_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
#endif
#define SRCDSP_TMA_ATTR(D, MX)                                                                                             \
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tma_kernel<D, MX, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));    \
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_tma_kernel<D, MX, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))
    SRCDSP_TMA_ATTR(0, false);
    SRCDSP_TMA_ATTR(0, true);
#ifdef SRCDSP_TIMING_EXPERIMENTS
    SRCDSP_TMA_ATTR(2, false);
    SRCDSP_TMA_ATTR(16, false);
    SRCDSP_TMA_ATTR(16, true);
#endif
#undef SRCDSP_TMA_ATTR
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_band_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SRCDSP_CUDA(cudaFuncSetAttribute(dec_band_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SRCDSP_TRY(build_band());
    tc_why = "";
    return SRCDSP_OK;
}

template <int MT, bool MIX>
static int launch_dec(const DecParams &P, int grid, int threads, size_t smem, cudaStream_t stream)
{
    static thread_local int configured_dev = -1;
    static thread_local size_t configured_smem = 0;
    int dev;
    cudaGetDevice(&dev);
    if (configured_dev != dev || smem > configured_smem) {
        SRCDSP_CUDA(cudaFuncSetAttribute(dec_fir_kernel<MT, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024));
        configured_dev = dev;
        configured_smem = 227 * 1024;
    }
    dec_fir_kernel<MT, MIX><<<grid, threads, smem, stream>>>(P);
    SRCDSP_LAUNCH_CHECK();
    count_launch();
    return SRCDSP_OK;
}

template <bool MIX>
static int launch_dec_m(const DecParams &P, int grid, int threads, size_t smem, cudaStream_t stream)
{
    switch (P.M) {
    case 1: return launch_dec<1, MIX>(P, grid, threads, smem, stream);
    case 2: return launch_dec<2, MIX>(P, grid, threads, smem, stream);
    case 3: return launch_dec<3, MIX>(P, grid, threads, smem, stream);
    case 4: return launch_dec<4, MIX>(P, grid, threads, smem, stream);
    case 5: return launch_dec<5, MIX>(P, grid, threads, smem, stream);
    case 8: return launch_dec<8, MIX>(P, grid, threads, smem, stream);
    case 10: return launch_dec<10, MIX>(P, grid, threads, smem, stream);
    case 16: return launch_dec<16, MIX>(P, grid, threads, smem, stream);
    case 32: return launch_dec<32, MIX>(P, grid, threads, smem, stream);
    default: return launch_dec<0, MIX>(P, grid, threads, smem, stream);
    }
}

int DecBank::step_device(const uint32_t *in, size_t in_stride, size_t n_in, uint32_t *out, size_t out_stride,
                         MixerBank *mixer)
{
    if (ntaps == 0) return fail(SRCDSP_E_STATE, "decimator has no coefficients (call srcdsp_dec_set_coeffs)");
    if (h_diag && h_diag[0])
        return fail(SRCDSP_E_CUDA, "a tensor-core kernel timed out waiting on mbarrier 0x%x (parity %d, CTA %d, thread %d) and trapped",
                    h_diag[1], h_diag[2], h_diag[3], h_diag[4]);
    if (n_in % (size_t)M != 0)
        return fail(SRCDSP_E_SIZE, "n_in (%zu) must be a multiple of M (%d) [dsptl_dnsampling_filters.h:181]", n_in, M);
    if (n_in == 0) return SRCDSP_OK;
    DeviceGuard g(device);
    DecParams P{};
    SRCDSP_TRY(shift_now(&P.shift));
    P.in = in;
    P.out = out;
    P.in_stride = in_stride;
    P.out_stride = out_stride;
    P.n_in = (long long)n_in;
    P.n_out = (long long)(n_in / (size_t)M);
    P.M = M;
    P.ntaps = ntaps;
    P.Qp = Qp;
    P.JP = JP;
    P.taps_poly = d_taps_poly;
    P.hist_in = d_hist[cur];
    P.H = H;
    const int TB = nt_threads * DEC_R;
    P.tiles_per_ch = (int)((P.n_out + TB - 1) / TB);
    P.vec_in = aligned16(in, in_stride);
    P.vec_out = aligned16(out, out_stride);
    const long long grid = (long long)P.tiles_per_ch * C;
    if (grid > 0x7fffffffll) return fail(SRCDSP_E_SIZE, "step too large: %lld tiles", grid);
    dim3 hgrid((unsigned)std::max(1, std::min((H + 255) / 256, 64)), (unsigned)C);
    // kernel choice: the tcgen05 Toeplitz kernel when the taps fit it and there are enough
    // 4096-output tiles to occupy the machine; the IMAD kernel otherwise
    const long long tc_tiles = (long long)C * ((P.n_out + TC_NRB * TC_BOUT - 1) / (TC_NRB * TC_BOUT));
    int tc_stages = 0;
    size_t tc_smem = 0;
    bool use_tc = tc_ok && tc_tiles < 0x7fffffffll && (kernel_kind >= 2 || (kernel_kind == 0 && tc_tiles >= sm_count / 2));  // (4 / 5: forms of 2)
    const char *why = tc_why;
    if (use_tc && mixer) {
        const PhaseMod pm = mixer->pm();
        if (!pm.mask) use_tc = false, why = "fused mixer needs a power-of-two sine table";
    }
    if (mixer) {
        if (mixer->C != C || mixer->device != device)
            return fail(SRCDSP_E_INVALID, "mixer and decimator banks must have the same channels and device");
        // the mixer's pending uploads must be ordered on *this* stream
        cudaStream_t ms = mixer->stream;
        mixer->stream = stream;
        int st = mixer->upload_if_dirty();
        mixer->stream = ms;
        SRCDSP_TRY(st);
        P.cs_table = mixer->d_cs;
        P.phi = mixer->d_phi[mixer->cur];
        P.freq = mixer->d_freq;
        P.pm = mixer->pm();
    }
    // TMA-fed variant: needs 16-byte aligned rows and at least one whole row-block in the tensor (row-blocks of
    // G samples: 32 * M, or 64 * M in p2 mode -- the variants have their own tile geometry)
    const int tbl_bytes = mixer ? (int)mixer->n_table * 4 : 0;      // LDG variant: copy of the packed (cos, sin) table
    // TMA variant: oscillator sequence in time order, two digit words per sample, one common period of all channels
    // (SRCDSP_SEQ_PERIOD=0: always the whole table).  The mixer variant's shared-memory plan (master layout, ring
    // depths) depends on its size, so it is re-planned when that changes -- a frequency change, not a per-step event.
    const unsigned seq_len = mixer ? (tune.seq_period == 0 || !mixer->seq_period ? mixer->n_table : mixer->seq_period) : 0;
    const int tma_tbl_bytes = (int)seq_len * 8;
    if (use_tc && mixer && tma_tbl_bytes != tc_mix_table) {
        SRCDSP_CUDA(cudaStreamSynchronize(stream));
        tma_mix_ok = false;
        if (tc_want_p2) SRCDSP_TRY(build_variant(2, true, tma_tbl_bytes));
        if (!tma_mix_ok) SRCDSP_TRY(build_variant(2, false, tma_tbl_bytes));
    }
    const TcParams &Ttma = mixer ? tc_tma_mix : tc_tma;
    const bool tma_built = mixer ? tma_mix_ok : tma_ok;
    const long long rows_tma = tma_built ? (long long)(n_in / (size_t)Ttma.G) : 0;
    int tma_raw = 0, tma_stages = 0;
    size_t tma_smem = 0;
    const bool can_map = use_tc && P.vec_in && tensor_map_encoder() != nullptr;
    // converter warps: groups of W warps, one K-step per group at a time
    // defaults from sweeps on cfg2 / ddc16 (256 ch x 16 Mi, /16, 255 taps): W = 4; plain decimator 2 groups + 3 byte-plane
    // stages (the raw ring gets the rest), fused mixer 3 groups + 4 stages (more ALU work per K-step)
    int tma_w = 4, tma_groups = mixer ? 3 : 2;
    if (tune.tma_w) tma_w = tune.tma_w == 8 ? 8 : 4;
    if (tma_built && Ttma.p2) tma_w = 4;  // 16 main row groups per K-step: 4 per warp
    if (tune.tma_groups > 0) tma_groups = tune.tma_groups;
    tma_groups = std::min(tma_groups, (mixer ? TMA_MAX_CONV_MIX : TMA_MAX_CONV) / tma_w);
    bool use_tma = can_map && tma_built && rows_tma >= 1 && rows_tma < 0x7fffffffll && kernel_kind != 3 && !tune.no_tma &&
                   tma_layout(Ttma, mixer ? tc_fixed_tma_mix : tc_fixed_tma, tma_tbl_bytes, tma_groups, &tma_raw, &tma_stages,
                              &tma_smem) == SRCDSP_OK;
    const long long rows_full = use_tma ? rows_tma : (long long)(n_in / (size_t)(32 * M));
    const bool have_map = can_map && rows_full >= 1 && rows_full < 0x7fffffffll;
    if (use_tc && !use_tma && tc_layout(tbl_bytes, &tc_stages, &tc_smem) != SRCDSP_OK)
        use_tc = false, why = "sine table + Toeplitz master + stages exceed 227 KB of shared memory";
    if (kernel_kind >= 2 && !use_tc)
        return fail(SRCDSP_E_STATE, "tcgen05 kernel forced but not applicable: %s", tc_ok ? why : tc_why);

    // band form (kernels_dec_band.cuh): 2x fewer tensor-pipe cycles and ~4x fewer MACs per output than dec_tma_kernel, but
    // an epilogue of ~28 instead of ~12 instructions per output (32x32b loads, row hand-over, re / im exchange): it wins
    // where outputs are sparse (tools/bandbench.py, 256 ch x 8 Mi, with and without the mixer: /10 .. /64 0.65-0.91x the
    // time of the original form, /8 0.91x with 63 taps and 1.01x with 255, /6 1.02-1.05x, /4 1.07-1.5x, /2 1.6-1.8x --
    // profiles/r2_bandbench.txt), so it is chosen from /10 on and for /8 with short filters (at most 16 extended outputs).
    // kernel_kind 4 forces it, 5 forces dec_tma_kernel's form, SRCDSP_BAND = 0 / 1 overrides.
    const long long band_rbs = (P.n_out + 31) / 32;
    const long long band_tiles = (long long)C * ((band_rbs + BAND_NEW - 1) / BAND_NEW);
    const long long band_rows_full = (long long)(n_in / (size_t)(32 * M));
    int band_raw = 0, band_stages = 0;
    size_t band_smem = 0;
    const int band_groups = std::max(1, std::min(tune.tma_groups > 0 ? tune.tma_groups : (mixer ? 3 : 2),
                                                 TMA_MAX_CONV_MIX / 4));
    const bool use_band = band_ok && tc_ok && P.vec_in && tensor_map_encoder() != nullptr && tune.band != 0 &&
                          (kernel_kind == 0 || kernel_kind == 2 || kernel_kind == 4) && band_rows_full >= 1 &&
                          band_rows_full < 0x7fffffffll && band_tiles < 0x7fffffffll &&
                          (kernel_kind == 4 || tune.band == 1 || M >= 10 || (M >= 8 && band.E <= 16)) &&
                          (kernel_kind != 0 || tune.band == 1 || band_tiles >= sm_count / 2) && (!mixer || mixer->pm().mask) &&
                          band_layout(tma_tbl_bytes, band_groups, &band_raw, &band_stages, &band_smem) == SRCDSP_OK;
    if (kernel_kind == 4 && !use_band)
        return fail(SRCDSP_E_STATE, "band-form tcgen05 kernel forced but not applicable (needs even M, ntaps <= 32 M + 1, taps < 2^23, 16-byte aligned rows, at least 32 M samples per call)");
    if (use_band) {
        TcParams T = tc_band;
        T.in = in, T.out = out, T.in_stride = in_stride, T.out_stride = out_stride;
        T.n_in = P.n_in, T.n_out = P.n_out;
        T.tiles_per_ch = (int)(band_tiles / C);
        T.total_tiles = band_tiles;
        T.hist_in = d_hist[cur];
        T.H = H;
        T.shift = P.shift;
        T.vec_in = P.vec_in, T.vec_out = P.vec_out;
        T.epi_sleep_ns = tune.epi_sleep >= 0 ? (unsigned)tune.epi_sleep : 400;
        T.mma_sleep_ns = tune.mma_sleep >= 0 ? (unsigned)tune.mma_sleep : (mixer ? 100 : 0);
        T.table_bytes = tma_tbl_bytes;
        if (mixer) {
            T.cs_table = mixer->d_cs;
            T.phi = P.phi;
            T.freq = P.freq;
            T.mix_mask = P.pm.mask;
            T.seq_mask = seq_len - 1;
        }
        BandExtra X = band;
        X.n_raw = tune.tma_raw > 0 ? std::max(2, std::min(tune.tma_raw, band_raw)) : band_raw;
        const int groups = std::max(1, std::min(band_groups, std::min(std::min(X.n_raw, band_stages - 1), X.ks2)));
        X.shared_raw = (X.n_raw % groups) != 0;
        X.n_conv = 4 * groups;
        X.rows_full = band_rows_full;
        T.n_stages = band_stages;
        CUtensorMap map;
        memset(&map, 0, sizeof map);
        const cuuint64_t gdim[3] = {(cuuint64_t)T.G, (cuuint64_t)band_rows_full, (cuuint64_t)C};
        const cuuint64_t gstr[2] = {(cuuint64_t)T.G * 4, C > 1 ? (cuuint64_t)in_stride * 4 : (cuuint64_t)band_rows_full * T.G * 4};
        const cuuint32_t box[3] = {64, (cuuint32_t)BAND_ROWS, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult cr = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)in, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(SRCDSP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
        const int tgrid = (int)std::min<long long>(band_tiles, sm_count);
        const int threads = 32 * (TMA_CONV_WARP0 + X.n_conv);
        if (tune.verbose)
            fprintf(stderr, "dec_band_kernel: M=%d E=%d master=%d B table=%d B raw=%d split=%d groups=%d shared_raw=%d smem=%zu tiles=%lld\n",
                    T.M, X.E, T.master_bytes, T.table_bytes, X.n_raw, T.n_stages, groups, X.shared_raw, band_smem, band_tiles);
        if (mixer)
            dec_band_kernel<true, 4><<<tgrid, threads, band_smem, stream>>>(T, X, map);
        else
            dec_band_kernel<false, 4><<<tgrid, threads, band_smem, stream>>>(T, X, map);
        SRCDSP_LAUNCH_CHECK();
        count_launch();
        last_kernel = 4;
    } else if (use_tc) {
        TcParams T = use_tma ? (mixer ? tc_tma_mix : tc_tma) : tc;
        T.in = in;
        T.out = out;
        T.in_stride = in_stride;
        T.out_stride = out_stride;
        T.n_in = P.n_in;
        T.n_out = P.n_out;
        T.tiles_per_ch = (int)(tc_tiles / C);
        T.total_tiles = tc_tiles;
        T.hist_in = d_hist[cur];
        T.H = H;
        T.shift = P.shift;
        T.vec_in = P.vec_in;
        T.n_stages = tc_stages;
        T.rb_stride = T.G;
        T.kc_stride = 32;
        T.epi_sleep_ns = 400;
        if (tune.epi_sleep >= 0) T.epi_sleep_ns = (unsigned)tune.epi_sleep;
        if (tune.conv_sleep >= 0) T.conv_sleep_ns = (unsigned)tune.conv_sleep;
        T.mma_sleep_ns = mixer ? 100 : 0;
        if (tune.mma_sleep >= 0) T.mma_sleep_ns = (unsigned)tune.mma_sleep;
        if (mixer) {
            T.cs_table = mixer->d_cs;
            T.phi = P.phi;
            T.freq = P.freq;
            T.mix_mask = P.pm.mask;
            T.seq_mask = seq_len - 1;
            T.table_bytes = (int)mixer->n_table * 4;
        }
        const int tgrid = (int)std::min<long long>(tc_tiles, sm_count);
#ifdef SRCDSP_TIMING_EXPERIMENTS
        if (tune.tc_debug && (!mixer || use_tma)) {  // timing experiments only (wrong results)
            T.debug = tune.tc_debug;
            if ((T.debug & 16) && !use_tma) {  // contiguous 128-byte lines per K-step instead of a 4*G-byte stride
                T.rb_stride = 32;
                T.kc_stride = 32 * TC_NRB;
            }
        }
#endif
        // the input as a tensor [C][rows_full][G] of 32-bit words (one complex int16 sample each); a box is one
        // K-step of one tile: 32 samples x (J-1 halo + nrb) row-blocks x 1 channel
        CUtensorMap map;
        memset(&map, 0, sizeof map);
        if (have_map) {
            const cuuint64_t gdim[3] = {(cuuint64_t)T.G, (cuuint64_t)rows_full, (cuuint64_t)C};
            const cuuint64_t gstr[2] = {(cuuint64_t)T.G * 4,
                                        C > 1 ? (cuuint64_t)in_stride * 4 : (cuuint64_t)rows_full * T.G * 4};
            const cuuint32_t box[3] = {32, (cuuint32_t)(T.J - 1 + T.nrb), 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult cr = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)in, gdim, gstr, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return fail(SRCDSP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
            if (!use_tma) {  // register-staged producers + TMA prefetch into L2
                T.pf_dist = T.n_stages + 12;
                if (tune.pf_dist >= 0) T.pf_dist = tune.pf_dist;
            }
        }
        if (use_tma) {
            TmaExtra X{};
            X.n_raw = tma_raw;
            if (tune.tma_raw > 0) X.n_raw = std::max(2, std::min(tune.tma_raw, tma_raw));
            tma_groups = std::max(1, std::min(tma_groups, std::min(X.n_raw, tma_stages - 1)));
            tma_groups = std::min(tma_groups, T.ksteps);  // every group must see every tile (per-channel rebuild barrier of the fused mixer)
            // a raw stage that different groups convert in turn needs an extra wait (see the converters)
            X.shared_raw = (X.n_raw % tma_groups) != 0;
            X.n_conv = tma_w * tma_groups;
            X.raw_rows = 4 * ((T.J - 1 + 3) / 4) + T.nrb;
            X.box_rows = T.J - 1 + T.nrb;
            X.rows_full = rows_full;
            T.n_stages = tma_stages;
            T.table_bytes = tma_tbl_bytes;
            const int threads = 32 * (TMA_CONV_WARP0 + X.n_conv);
            if (tune.verbose)
                fprintf(stderr, "dec_tma_kernel: M=%d p2=%d J=%d grouped=%d master=%d B table=%d B raw=%d split=%d W=%d groups=%d shared_raw=%d smem=%zu\n",
                        T.M, T.p2, T.J, T.grouped, T.master_bytes, T.table_bytes, X.n_raw, T.n_stages, tma_w, tma_groups, X.shared_raw, tma_smem);
#define SRCDSP_TMA_LAUNCH(D, MX)                                                         \
    do {                                                                                 \
        if (tma_w == 8)                                                                  \
            dec_tma_kernel<D, MX, 8><<<tgrid, threads, tma_smem, stream>>>(T, X, map);   \
        else                                                                             \
            dec_tma_kernel<D, MX, 4><<<tgrid, threads, tma_smem, stream>>>(T, X, map);   \
    } while (0)
#ifdef SRCDSP_TIMING_EXPERIMENTS
            if (T.debug & 32) {  // wait-cycle accounting variant; counters printed at the next launch
                unsigned long long c[10];
                cudaMemcpy(c, d_error, sizeof c, cudaMemcpyDeviceToHost);
                fprintf(stderr, "tma counters (cycles summed over CTAs): conv total %llu wait_split_empty %llu wait_raw_full %llu fence %llu arrive %llu | mma total %llu wait_full %llu wait_tempty %llu | epi total %llu wait_tfull %llu\n", c[0], c[1], c[2], c[8], c[9], c[3], c[4], c[5], c[6], c[7]);
                cudaMemset(d_error, 0, sizeof c);
                if (mixer)
                    SRCDSP_TMA_LAUNCH(16, true);
                else
                    SRCDSP_TMA_LAUNCH(16, false);
            } else if (!mixer && (T.debug & 2)) {
                SRCDSP_TMA_LAUNCH(2, false);
            } else
#endif
            if (mixer) {
                SRCDSP_TMA_LAUNCH(0, true);
            } else {
                SRCDSP_TMA_LAUNCH(0, false);
            }
#undef SRCDSP_TMA_LAUNCH
        } else if (mixer) {
            dec_tc_kernel<0, true><<<tgrid, TC_THREADS, tc_smem, stream>>>(T, map);
#ifdef SRCDSP_TIMING_EXPERIMENTS
        } else if (T.debug & 32) {  // wait-cycle accounting variant; counters printed at the next launch
            unsigned long long c[8];
            cudaMemcpy(c, d_error, sizeof c, cudaMemcpyDeviceToHost);
            fprintf(stderr, "tc counters (cycles summed over CTAs): prod total %llu wait_empty %llu fence %llu | mma total %llu wait_full %llu wait_tempty %llu | epi total %llu wait_tfull %llu\n", c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
            cudaMemset(d_error, 0, sizeof c);
            dec_tc_kernel<16, false><<<tgrid, TC_THREADS, tc_smem, stream>>>(T, map);
        } else if (T.debug & 10) {  // 2 / 8: timing-experiment instantiations (wrong results)
            switch (T.debug & 10) {
            case 2: dec_tc_kernel<2, false><<<tgrid, TC_THREADS, tc_smem, stream>>>(T, map); break;
            case 8: dec_tc_kernel<8, false><<<tgrid, TC_THREADS, tc_smem, stream>>>(T, map); break;
            default: dec_tc_kernel<10, false><<<tgrid, TC_THREADS, tc_smem, stream>>>(T, map); break;
            }
#endif
        } else {
            dec_tc_kernel<0, false><<<tgrid, TC_THREADS, tc_smem, stream>>>(T, map);
        }
        SRCDSP_LAUNCH_CHECK();
        count_launch();
        last_kernel = use_tma ? 3 : 2;
    } else if (mixer) {
        SRCDSP_TRY(launch_dec_m<true>(P, (int)grid, nt_threads, smem_bytes, stream));
        last_kernel = 1;
    } else {
        SRCDSP_TRY(launch_dec_m<false>(P, (int)grid, nt_threads, smem_bytes, stream));
        last_kernel = 1;
    }
    if (mixer) {
        dec_history_kernel<true><<<hgrid, 256, 0, stream>>>(in, in_stride, (long long)n_in, d_hist[cur], d_hist[cur ^ 1],
                                                            H, mixer->d_cs, mixer->d_phi[mixer->cur],
                                                            mixer->d_phi[mixer->cur ^ 1], mixer->d_freq, mixer->pm());
        SRCDSP_LAUNCH_CHECK();
        count_launch();
        mixer->advance(n_in);
    } else {
        dec_history_kernel<false><<<hgrid, 256, 0, stream>>>(in, in_stride, (long long)n_in, d_hist[cur],
                                                             d_hist[cur ^ 1], H, nullptr, nullptr, nullptr, nullptr,
                                                             PhaseMod{1, 0});
        SRCDSP_LAUNCH_CHECK();
        count_launch();
    }
    cur ^= 1;
    return SRCDSP_OK;
}

// ---------------------------------------------------------------------------------------------
// fused DDC chain
// ---------------------------------------------------------------------------------------------
struct DdcChain {
    MixerBank *mixer = nullptr;
    DecBank *d1 = nullptr, *d2 = nullptr;
    uint32_t *d_mid = nullptr;  // [C][mid_pitch] stage-1 output when d2 is present
    size_t mid_cap = 0;

    int step_device(const uint32_t *in, size_t in_stride, size_t n_in, uint32_t *out, size_t out_stride)
    {
        const size_t Mt = (size_t)d1->M * (d2 ? (size_t)d2->M : 1);
        if (n_in % Mt != 0)
            return fail(SRCDSP_E_SIZE, "n_in (%zu) must be a multiple of the total decimation %zu", n_in, Mt);
        if (!d2) return d1->step_device(in, in_stride, n_in, out, out_stride, mixer);
        DeviceGuard g(d1->device);
        const size_t n_mid = n_in / (size_t)d1->M;
        const size_t pitch = (n_mid + 3) & ~(size_t)3;
        if (pitch * d1->C > mid_cap) {
            SRCDSP_CUDA(cudaStreamSynchronize(d1->stream));
            if (d_mid) cudaFree(d_mid);
            d_mid = nullptr;
            SRCDSP_CUDA(cudaMalloc(&d_mid, pitch * d1->C * 4 + 16));
            mid_cap = pitch * d1->C;
        }
        SRCDSP_TRY(d1->step_device(in, in_stride, n_in, d_mid, pitch, mixer));
        // stage 2 runs on stage 1's stream so that it is ordered behind it
        cudaStream_t s2 = d2->stream;
        d2->stream = d1->stream;
        int st = d2->step_device(d_mid, pitch, n_mid, out, out_stride, nullptr);
        d2->stream = s2;
        return st;
    }
};

// ---------------------------------------------------------------------------------------------
// upsampler bank
// ---------------------------------------------------------------------------------------------
struct UpBank : Bank {
    int L = 1;
    int ntaps = 0, H = 0, Hp = 0, G = 0;
    int length = 0, imp_length = 0, left_shift_factor = 0;
    unsigned long long top = 0;  // the reference's insertion index (only matters across setCoefficients)
    int32_t *d_taps_poly = nullptr;
    uint32_t *d_hist[2] = {nullptr, nullptr};  // [C][H] age order, [H-1] newest, [0] stale slot
    int cur = 0;
    size_t smem_bytes = 0;
    // tcgen05 form (kernels_up_tc.cuh): one MMA per 4096 outputs; applicable when 32 % L == 0, the group window
    // 32 / L + H - 1 fits 15 samples and the taps fit 3 signed byte digits
    uint8_t *d_a_image = nullptr;
    bool tc_ok = false;
    int tc_digits3 = 0, tc_KW = 0;
    int *h_diag = nullptr, *d_diag = nullptr;
    int sm_count = 148;
    int last_kernel = 0;  // 1 = up_fir_kernel, 2 = up_fir4_kernel, 3 = up_tc_kernel, 4 = up_tc2_kernel
    bool last_kernel_form2 = false;

    int prepare_tc(const int32_t *t, int n)
    {
        tc_ok = false;
        if (d_a_image) {
            cudaFree(d_a_image);
            d_a_image = nullptr;
        }
        const int Hh = n / L;
        if (L > 32 || 32 % L != 0) return SRCDSP_OK;
        const int JJ = 32 / L, KW = JJ + Hh - 1;
        if (KW > 15) return SRCDSP_OK;
        auto digits = [](long long v, int nd, int *d) {  // signed base-256 digits, returns the remainder
            for (int w = 0; w < nd; ++w) {
                const int dg = (int)(((v + 128) & 255) - 128);
                d[w] = dg;
                v = (v - dg) >> 8;
            }
            return v;
        };
        // digit tables [g][w][k]: low-plane columns k = m (and the bias column 15), high-plane columns k = 16 + m
        std::vector<int8_t> lo(32 * 4 * 16, 0), hi(32 * 4 * 16, 0);
        int d3 = 0;
        for (int g = 0; g < 32; ++g) {
            const int jj = g / L, p = g % L;
            long long sum = 0;
            for (int i = 0; i < Hh; ++i) sum += t[p + i * L];
            int bd[4];
            digits((long long)(int32_t)(uint32_t)(128ull * (unsigned long long)sum), 4, bd);  // mod 2^32
            if (bd[3] != 0) d3 = 1;  // the bias alone can reach the fourth slot
            for (int w = 0; w < 4; ++w) {
                for (int m = 0; m < KW; ++m) {
                    const int i = jj - m + Hh - 1;
                    if (i < 0 || i >= Hh) continue;
                    int d[3];
                    if (digits(t[p + i * L], 3, d) != 0) return SRCDSP_OK;  // |c| >= 2^23: CUDA-core kernels only
                    if (d[2] != 0) d3 = 1;
                    if (w < 3) lo[(g * 4 + w) * 16 + m] = (int8_t)d[w];
                    if (w >= 1) hi[(g * 4 + w) * 16 + m] = (int8_t)d[w - 1];
                }
                lo[(g * 4 + w) * 16 + 15] = (int8_t)bd[w];  // bias column (the sample operand holds a 1 there)
            }
        }
        // two images of the same digits: up_tc_kernel's M-side operand [2][128][16] (row 32 * (g / 8) + 8 * w + g % 8) and
        // up_tc2_kernel's N-side operand [2][32 * S][16] (row 32 * w + g; S = 3 slots unless the fourth is in use)
        std::vector<uint8_t> img(2 * UPT_A_BYTES, 0);
        const int S2 = d3 ? 4 : 3, N2 = 32 * S2;
        for (int g = 0; g < 32; ++g)
            for (int w = 0; w < 4; ++w) {
                const int row = 32 * (g >> 3) + 8 * w + (g & 7), row2 = 32 * w + g;
                memcpy(&img[(size_t)row * 16], &lo[(g * 4 + w) * 16], 16);
                memcpy(&img[(size_t)2048 + (size_t)row * 16], &hi[(g * 4 + w) * 16], 16);
                if (w < S2) {
                    memcpy(&img[(size_t)UPT_A_BYTES + (size_t)row2 * 16], &lo[(g * 4 + w) * 16], 16);
                    memcpy(&img[(size_t)UPT_A_BYTES + (size_t)(N2 + row2) * 16], &hi[(g * 4 + w) * 16], 16);
                }
            }
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaMalloc(&d_a_image, 2 * UPT_A_BYTES));
        SRCDSP_CUDA(cudaMemcpy(d_a_image, img.data(), 2 * UPT_A_BYTES, cudaMemcpyHostToDevice));
        if (!h_diag) {
            SRCDSP_CUDA(cudaHostAlloc(&h_diag, 64, cudaHostAllocMapped));
            memset(h_diag, 0, 64);
            SRCDSP_CUDA(cudaHostGetDevicePointer(&d_diag, h_diag, 0));
            cudaDeviceProp prop;
            SRCDSP_CUDA(cudaGetDeviceProperties(&prop, device));
            sm_count = prop.multiProcessorCount;
        }
        tc_digits3 = d3;
        tc_KW = KW;
        tc_ok = true;
        return SRCDSP_OK;
    }

    int set_coefficients(const int32_t *t, int n)
    {
        if (!t || n < 1) return fail(SRCDSP_E_SIZE, "taps must not be empty [upsampling_filters.h:110]");
        if (n % L != 0) return fail(SRCDSP_E_SIZE, "ntaps (%d) must be a multiple of L (%d) [upsampling_filters.h:113]", n, L);
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaStreamSynchronize(stream));
        const int newH = n / L;
        const int newHp = (newH + UP_HC - 1) / UP_HC * UP_HC;
        const int HP = newHp + 4;
        // every check that can fail comes before any state is replaced: a refused call leaves the bank as it was
        const size_t new_smem = ((size_t)L * HP + (size_t)(UP_NT / L) * UP_R * UP_NI + newHp + 8) * 4;
        if (new_smem > 227 * 1024)
            return fail(SRCDSP_E_SIZE, "L=%d with %d taps needs %zu bytes of shared memory per CTA", L, n, new_smem);
        std::vector<int32_t> poly((size_t)L * HP, 0);
        for (int k = 0; k < n; ++k) poly[(size_t)(k % L) * HP + k / L] = t[k];
        int len = n;
        while (len > 0 && t[len - 1] == 0) --len;  // upsampling_filters.h:122-123
        if (len == 0) return fail(SRCDSP_E_SIZE, "all taps are zero (the reference reads coeff[-1] here)");
        DeviceBuffer nd;  // freed on every early return below
        SRCDSP_CUDA(cudaMalloc(&nd.p, poly.size() * 4));
        SRCDSP_CUDA(cudaMemcpy(nd.p, poly.data(), poly.size() * 4, cudaMemcpyHostToDevice));
        if (newH != H) {
            // buffer.resize(newH) on the RAW circular buffer with `top` unchanged
            // (upsampling_filters.h:118): rebuild the raw buffer from age order, resize, re-age.
            std::vector<uint32_t> nh((size_t)C * newH, 0);
            unsigned long long new_top = top;
            if (H > 0 && d_hist[cur]) {
                std::vector<uint32_t> oh((size_t)C * H);
                SRCDSP_CUDA(cudaMemcpy(oh.data(), d_hist[cur], oh.size() * 4, cudaMemcpyDeviceToHost));
                const int tp = (int)(top % (unsigned long long)H);
                if (tp >= newH)
                    return fail(SRCDSP_E_STATE,
                                "setCoefficients shrinks the buffer below the insertion index (top=%d, new size %d): "
                                "the reference writes out of bounds here; call reset() first", tp, newH);
                std::vector<uint32_t> raw(std::max(H, newH));
                for (int c = 0; c < C; ++c) {
                    std::fill(raw.begin(), raw.end(), 0u);
                    for (int k = 0; k < H; ++k) raw[(tp + k) % H] = oh[(size_t)c * H + k];
                    for (int k = 0; k < newH; ++k) nh[(size_t)c * newH + k] = raw[(tp + k) % newH];
                }
                new_top = (unsigned long long)tp;
            }
            DeviceBuffer b0, b1;
            SRCDSP_CUDA(cudaMalloc(&b0.p, nh.size() * 4));
            SRCDSP_CUDA(cudaMalloc(&b1.p, nh.size() * 4));
            SRCDSP_CUDA(cudaMemcpy(b0.p, nh.data(), nh.size() * 4, cudaMemcpyHostToDevice));
            if (d_hist[0]) cudaFree(d_hist[0]);
            if (d_hist[1]) cudaFree(d_hist[1]);
            d_hist[0] = static_cast<uint32_t *>(b0.release());
            d_hist[1] = static_cast<uint32_t *>(b1.release());
            cur = 0;
            top = new_top;
        }
        if (d_taps_poly) cudaFree(d_taps_poly);
        d_taps_poly = static_cast<int32_t *>(nd.release());
        ntaps = n;
        H = newH;
        Hp = newHp;
        G = UP_NT / L;
        length = len;
        imp_length = n;
        left_shift_factor = static_cast<int>(round(log2((double)L)));  // upsampling_filters.h:120
        smem_bytes = new_smem;
        return prepare_tc(t, n);
    }

    int reset()
    {
        top = 0;
        if (!d_hist[0]) return SRCDSP_OK;
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaMemsetAsync(d_hist[cur], 0, (size_t)C * H * 4, stream));
        return SRCDSP_OK;
    }

    int step_device(const uint32_t *in, size_t in_stride, size_t n_in, size_t n_flush, uint32_t *out,
                    size_t out_stride, int shift_mode)
    {
        if (ntaps == 0) return fail(SRCDSP_E_STATE, "upsampler has no coefficients [upsampling_filters.h:155]");
        const size_t n_tot = n_in + n_flush;
        if (n_tot == 0) return SRCDSP_OK;
        DeviceGuard g(device);
        UpParams P{};
        const int sh = shift_mode == 0 ? 15 - left_shift_factor : 0;  // :189 vs :244
        if (sh < 0) return fail(SRCDSP_E_STATE, "15 - round(log2 L) < 0 (undefined in the reference)");
        P.shift = (unsigned)sh;
        P.in = in;
        P.out = out;
        P.in_stride = in_stride;
        P.out_stride = out_stride;
        P.n_in = (long long)n_in;
        P.n_tot = (long long)n_tot;
        P.L = L;
        P.H = H;
        P.Hp = Hp;
        P.G = G;
        P.taps_poly = d_taps_poly;
        P.hist_in = d_hist[cur];
        P.vec_in = aligned16(in, in_stride);
        // register-blocked kernel (4 phases x 4 inputs per thread, 16-byte stores) for the usual ratios; the
        // generic one for any other L or unaligned output rows
        if (h_diag && h_diag[0])
            return fail(SRCDSP_E_CUDA, "the tcgen05 upsampler timed out waiting on mbarrier 0x%x (parity %d, CTA %d, thread %d) and trapped",
                        h_diag[1], h_diag[2], h_diag[3], h_diag[4]);
        // tcgen05 kernel: when applicable, faster and the batch fills the machine (SRCDSP_UP_TC = 1 forces it, 0 disables it)
        {
            const int JJ = 32 / std::max(1, std::min(L, 32));
            const long long tiles = tc_ok ? (long long)C * (((long long)n_tot + UPT_GROUPS * JJ - 1) / (UPT_GROUPS * JJ)) : 0;
            const bool tc_able = tc_ok && aligned16(in, in_stride) && tiles > 0 && tiles < 0x7fffffffll;  // cp.async 16-byte chunks
            // Its rate does not depend on the filter length (one MMA per 4096 outputs; the 32 bytes of accumulator per
            // output that the epilogue reads from TMEM bound it at ~0.85-0.9 T out/s), the CUDA-core kernels pay 2 IMADs
            // per tap: they win up to 8 taps per phase (one tap block, ~1.0 T out/s) and lose beyond (0.5 T out/s).
            bool use_tc = tc_able && tiles >= sm_count / 2 && H > UP_HC;
            if (tune.up_tc >= 0) use_tc = tc_able && tune.up_tc != 0;
            if (use_tc) {
                UpTcParams T{};
                T.in = in, T.out = out, T.in_stride = in_stride, T.out_stride = out_stride;
                T.n_in = (long long)n_in, T.n_tot = (long long)n_tot;
                T.L = L, T.JJ = JJ, T.H = H, T.KW = tc_KW;
                T.halo = (H - 1 + 3) & ~3;
                T.raw_words = upt_raw_words(JJ, H);
                T.a_image = d_a_image;
                T.hist_in = d_hist[cur];
                T.shift = P.shift;
                T.tiles_per_ch = (int)(tiles / C);
                T.total_tiles = tiles;
                T.error_flag = d_diag;
                T.debug = tune.upt_debug;  // 0 unless built with -DSRCDSP_TIMING_EXPERIMENTS
                const int tgrid = (int)std::min<long long>(tiles, sm_count);
                const size_t tsmem = upt_smem_bytes(JJ, H);
#ifdef SRCDSP_TIMING_EXPERIMENTS
                static unsigned long long *d_cnt = nullptr;
                if (T.debug & 8) {
                    if (!d_cnt) cudaMalloc(&d_cnt, 64);
                    unsigned long long c[8];
                    cudaMemcpy(c, d_cnt, 64, cudaMemcpyDeviceToHost);
                    fprintf(stderr, "up_tc counters (cycles summed over warps): conv total %llu wait_empty %llu wait_data %llu | mma total %llu wait_full %llu wait_tempty %llu | epi total %llu wait_tfull %llu\n", c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
                    cudaMemset(d_cnt, 0, 64);
                    T.counters = d_cnt;
                }
#endif
                // form 2 (operand roles swapped: S accumulator slots per output instead of 4, no cross-lane reduction,
                // 16-byte stores) is bit-identical but measured 3-12 % SLOWER than form 1 on B200 (x8/96 taps 2.93 vs
                // 2.61 ms: its 32x32b TMEM loads deliver ~60 B/clk/SM, form 1's 16x256b loads ~90), so it only runs
                // on request: SRCDSP_UP_TC_FORM = 2
                bool form2 = false;
                if (tune.up_tc_form == 2) form2 = aligned16(out, out_stride);
                T.n_image = d_a_image + UPT_A_BYTES;
                if (form2 && tc_digits3) {
                    SRCDSP_CUDA(cudaFuncSetAttribute(up_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    up_tc2_kernel<4><<<tgrid, UPT_THREADS, tsmem, stream>>>(T);
                } else if (form2) {
                    SRCDSP_CUDA(cudaFuncSetAttribute(up_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    up_tc2_kernel<3><<<tgrid, UPT_THREADS, tsmem, stream>>>(T);
                } else if (tc_digits3) {
                    SRCDSP_CUDA(cudaFuncSetAttribute(up_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    up_tc_kernel<true><<<tgrid, UPT_THREADS, tsmem, stream>>>(T);
                } else {
                    SRCDSP_CUDA(cudaFuncSetAttribute(up_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    up_tc_kernel<false><<<tgrid, UPT_THREADS, tsmem, stream>>>(T);
                }
                last_kernel_form2 = form2;
                SRCDSP_LAUNCH_CHECK();
                last_kernel = last_kernel_form2 ? 4 : 3;
                dim3 hgrid((unsigned)std::max(1, std::min((H + 255) / 256, 64)), (unsigned)C);
                up_history_kernel<<<hgrid, 256, 0, stream>>>(in, in_stride, (long long)n_in, (long long)n_tot, d_hist[cur],
                                                             d_hist[cur ^ 1], H);
                SRCDSP_LAUNCH_CHECK();
                count_launch(2);
                cur ^= 1;
                top += n_tot;
                return SRCDSP_OK;
            }
        }
        const bool blocked = (L == 4 || L == 8 || L == 16) && aligned16(out, out_stride) && !tune.up_generic;
        last_kernel = blocked ? 2 : 1;
        const int span = blocked ? up4_span(L) : G * UP_R * UP_NI;
        const size_t smem = blocked ? ((size_t)L * (Hp + 4) + (size_t)span + Hp + 8) * 4 : smem_bytes;
        P.tiles_per_ch = (int)((n_tot + span - 1) / span);
        const long long grid = (long long)P.tiles_per_ch * C;
        if (grid > 0x7fffffffll) return fail(SRCDSP_E_SIZE, "step too large: %lld tiles", grid);
        if (smem > 227 * 1024) return fail(SRCDSP_E_SIZE, "L=%d with %d taps needs %zu bytes of shared memory per CTA", L, ntaps, smem);
#define SRCDSP_UP4(LL, SG)                                                                                                   \
    do {                                                                                                                     \
        if (smem > 48 * 1024)                                                                                                \
            SRCDSP_CUDA(cudaFuncSetAttribute(up_fir4_kernel<LL, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        up_fir4_kernel<LL, SG><<<(int)grid, UP_NT, smem, stream>>>(P);                                                        \
    } while (0)
        if (blocked) {
            const bool single = Hp == UP_HC;
            if (L == 4) {
                if (single) SRCDSP_UP4(4, true); else SRCDSP_UP4(4, false);
            } else if (L == 8) {
                if (single) SRCDSP_UP4(8, true); else SRCDSP_UP4(8, false);
            } else {
                if (single) SRCDSP_UP4(16, true); else SRCDSP_UP4(16, false);
            }
        } else {
            if (smem > 48 * 1024)
                SRCDSP_CUDA(cudaFuncSetAttribute(up_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            up_fir_kernel<<<(int)grid, UP_NT, smem, stream>>>(P);
        }
#undef SRCDSP_UP4
        SRCDSP_LAUNCH_CHECK();
        dim3 hgrid((unsigned)std::max(1, std::min((H + 255) / 256, 64)), (unsigned)C);
        up_history_kernel<<<hgrid, 256, 0, stream>>>(in, in_stride, (long long)n_in, (long long)n_tot, d_hist[cur],
                                                     d_hist[cur ^ 1], H);
        SRCDSP_LAUNCH_CHECK();
        count_launch(2);
        cur ^= 1;
        top += n_tot;
        return SRCDSP_OK;
    }

    void destroy()
    {
        DeviceGuard g(device);
        release();
        if (d_taps_poly) cudaFree(d_taps_poly);
        if (d_hist[0]) cudaFree(d_hist[0]);
        if (d_hist[1]) cudaFree(d_hist[1]);
        if (d_a_image) cudaFree(d_a_image);
        if (h_diag) cudaFreeHost(h_diag);
    }
};

}  // namespace srcdsp

// =============================================================================================
// extern "C"
// =============================================================================================
using namespace srcdsp;

// ---------------------------------------------------------------------------------------------
// correlator bank (correlators.h)
// ---------------------------------------------------------------------------------------------
struct CorrBank : Bank {
    int N = 0, S = 0, H = 0;
    bool has_pattern = false;
    uint32_t coeffs_energy = 0;
    int coeff_scaling = 0;
    double threshold_factor = 0;
    int *d_coef = nullptr;
    uint32_t *d_hist[2] = {nullptr, nullptr};
    uint32_t *d_reg[2] = {nullptr, nullptr};
    int cur = 0;
    int *d_found = nullptr;
    uint32_t *d_bits = nullptr;
    unsigned long long *d_cnt = nullptr;
    uint32_t *d_scratch = nullptr;  // device copy of a host input block
    size_t scratch_words = 0;
    std::vector<int> h_found;

    int create(int dev, int channels, int n, int s)
    {
        if (n < 1 || n > CORR_MAX_N || s < 1 || (long long)n * s > 8192)
            return fail(SRCDSP_E_SIZE, "correlator: need 1 <= N <= %d, S >= 1, N * S <= 8192 (got N=%d S=%d)", CORR_MAX_N, n, s);
        SRCDSP_TRY(init(dev, channels));
        N = n, S = s, H = n * s - 1;
        DeviceGuard g(device);
        const size_t hw = (size_t)C * (H > 0 ? H : 1);
        SRCDSP_CUDA(cudaMalloc(&d_coef, (size_t)2 * N * sizeof(int)));
        for (int k = 0; k < 2; ++k) {
            SRCDSP_CUDA(cudaMalloc(&d_hist[k], hw * 4));
            SRCDSP_CUDA(cudaMalloc(&d_reg[k], (size_t)C * 6 * 4));
        }
        SRCDSP_CUDA(cudaMalloc(&d_found, (size_t)C * sizeof(int)));
        SRCDSP_CUDA(cudaMalloc(&d_bits, (size_t)C * N * 4));
        SRCDSP_CUDA(cudaMalloc(&d_cnt, (size_t)C * 8));
        h_found.resize(C);
        SRCDSP_CUDA(cudaFuncSetAttribute(corr_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        SRCDSP_CUDA(cudaFuncSetAttribute(corr_scan_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        return reset();
    }

    // correlators.h:143-155
    int reset()
    {
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaMemsetAsync(d_hist[cur], 0, (size_t)C * (H > 0 ? H : 1) * 4, stream));
        SRCDSP_CUDA(cudaMemsetAsync(d_reg[cur], 0, (size_t)C * 6 * 4, stream));
        SRCDSP_CUDA(cudaMemsetAsync(d_bits, 0, (size_t)C * N * 4, stream));
        SRCDSP_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)C * 8, stream));
        return SRCDSP_OK;
    }

    // correlators.h:167-194: conjugate, energy (asserted <= 1073217600), threshold factor, coeffScaling
    int set_pattern(const int32_t *iq, double threshold_coeff)
    {
        if (!iq) return fail(SRCDSP_E_INVALID, "pattern is null");
        std::vector<int> c(2 * N);
        double tmp = 0;
        for (int k = 0; k < N; ++k) {
            c[2 * k] = iq[2 * k];
            c[2 * k + 1] = -iq[2 * k + 1];
            tmp += (double)((int32_t)(c[2 * k] * c[2 * k] + c[2 * k + 1] * c[2 * k + 1]));  // int32 arithmetic like the reference
        }
        if (!(tmp <= 1073217600.0))
            return fail(SRCDSP_E_SIZE, "pattern energy %.0f exceeds 1073217600: each value must be below 13 bits [correlators.h:183]", tmp);
        if (!(tmp >= 1)) return fail(SRCDSP_E_SIZE, "pattern energy must be >= 1 (coeffScaling = floor(log2(sqrt(energy))))");
        DeviceGuard g(device);
        SRCDSP_CUDA(cudaStreamSynchronize(stream));
        SRCDSP_CUDA(cudaMemcpy(d_coef, c.data(), c.size() * sizeof(int), cudaMemcpyHostToDevice));
        coeffs_energy = static_cast<uint32_t>(tmp);
        threshold_factor = threshold_coeff * sqrt((double)coeffs_energy);
        coeff_scaling = static_cast<int>(floor(log2(sqrt((double)coeffs_energy))));
        has_pattern = true;
        return SRCDSP_OK;
    }

    // step: correlators.h:209-303 for every channel; found[ch] / corr_index[ch] = the reference's return value / corrIndex
    int step(const int16_t *in, size_t in_stride, size_t n, int *found, int *corr_index)
    {
        if (!has_pattern) return fail(SRCDSP_E_STATE, "correlator has no pattern (call srcdsp_corr_set_pattern)");
        if (!found || !corr_index) return fail(SRCDSP_E_INVALID, "found / corr_index are null");
        if (n >= 0x7f7f7f7fu)  // the "not found" sentinel written by cudaMemsetAsync(0x7f) must stay above every index
            return fail(SRCDSP_E_SIZE, "correlator block too long (int indices, correlators.h:211; at most 0x7f7f7f7e samples per call)");
        for (int c = 0; c < C; ++c) found[c] = 0;
        if (n == 0) return SRCDSP_OK;
        DeviceGuard g(device);
        const uint32_t *d_in = reinterpret_cast<const uint32_t *>(in);
        size_t stride = in_stride;
        if (!is_device_ptr(in)) {
            const size_t pitch = (n + 3) & ~(size_t)3;
            if (pitch * C > scratch_words) {
                SRCDSP_CUDA(cudaStreamSynchronize(stream));
                if (d_scratch) cudaFree(d_scratch);
                d_scratch = nullptr;
                SRCDSP_CUDA(cudaMalloc(&d_scratch, pitch * C * 4));
                scratch_words = pitch * C;
            }
            SRCDSP_CUDA(cudaMemcpy2DAsync(d_scratch, pitch * 4, in, in_stride * 4, n * 4, C, cudaMemcpyHostToDevice, stream));
            d_in = d_scratch;
            stride = pitch;
        }
        CorrParams P{};
        P.in = d_in;
        P.in_stride = stride;
        P.n = (int)n;
        P.N = N, P.S = S, P.H = H;
        P.coef = d_coef;
        P.hist_in = d_hist[cur];
        P.hist_out = d_hist[cur ^ 1];
        P.reg_in = d_reg[cur];
        P.reg_out = d_reg[cur ^ 1];
        P.found = d_found;
        P.bits = d_bits;
        P.cnt = d_cnt;
        P.coeff_scaling = coeff_scaling;
        SRCDSP_CUDA(cudaMemsetAsync(d_found, 0x7f, (size_t)C * sizeof(int), stream));  // 0x7f7f7f7f > any index
        const int halo = (N - 1) * S + 2;
        const size_t smem = (3 * ((size_t)halo + CORR_THREADS) + 2 * N + 2 * (CORR_THREADS + 2)) * 4;
        // register-blocked scan (8 outputs S apart per thread) for power-of-two strides; the one-output-per-thread
        // kernel for any other stride
        const bool blocked = (S & (S - 1)) == 0 && S <= CORR_THREADS && !tune.corr_generic &&
                             corr_blocked_smem(N, S) <= 200 * 1024;
        if (blocked) {
            const int TT = CORR_THREADS * CORR_R;
            dim3 grid((unsigned)((n + TT - 1) / TT), (unsigned)C);
            corr_scan_blocked_kernel<<<grid, CORR_THREADS, corr_blocked_smem(N, S), stream>>>(P, aligned16(d_in, stride) ? 1 : 0);
        } else {
            dim3 grid((unsigned)((n + CORR_THREADS - 1) / CORR_THREADS), (unsigned)C);
            corr_scan_kernel<<<grid, CORR_THREADS, smem, stream>>>(P);
        }
        SRCDSP_LAUNCH_CHECK();
        corr_finish_kernel<<<C, CORR_THREADS, 0, stream>>>(P);
        SRCDSP_LAUNCH_CHECK();
        count_launch(2);
        SRCDSP_CUDA(cudaMemcpyAsync(h_found.data(), d_found, (size_t)C * sizeof(int), cudaMemcpyDeviceToHost, stream));
        SRCDSP_CUDA(cudaStreamSynchronize(stream));
        cur ^= 1;
        for (int c = 0; c < C; ++c) {
            found[c] = h_found[c] < (int)n;
            if (found[c]) corr_index[c] = h_found[c] - 1;  // "-1 to refer to the previous sample", correlators.h:269
        }
        return SRCDSP_OK;
    }

    void destroy()
    {
        DeviceGuard g(device);
        release();
        if (d_coef) cudaFree(d_coef);
        for (int k = 0; k < 2; ++k) {
            if (d_hist[k]) cudaFree(d_hist[k]);
            if (d_reg[k]) cudaFree(d_reg[k]);
        }
        if (d_found) cudaFree(d_found);
        if (d_bits) cudaFree(d_bits);
        if (d_cnt) cudaFree(d_cnt);
        if (d_scratch) cudaFree(d_scratch);
    }
};

struct srcdsp_mixer_s : MixerBank {};
struct srcdsp_dec_s : DecBank {};
struct srcdsp_up_s : UpBank {};
struct srcdsp_ddc_s : DdcChain {};
struct srcdsp_corr_s : CorrBank {};

#define CHECK_HANDLE(h)                                                   \
    do {                                                                  \
        if (!(h)) return fail(SRCDSP_E_INVALID, "%s: null handle", __func__); \
    } while (0)

extern "C" {

const char *srcdsp_last_error(void) { return last_error_ref().c_str(); }
int srcdsp_version(void) { return 100; }
uint64_t srcdsp_launch_count(void) { return g_launches.load(); }

int srcdsp_device_count(int *count)
try {
    if (!count) return fail(SRCDSP_E_INVALID, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(SRCDSP_E_NOGPU, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_host_alloc(void **ptr, size_t bytes)
try {
    if (!ptr) return fail(SRCDSP_E_INVALID, "ptr is null");
    SRCDSP_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_host_free(void *ptr)
try {
    SRCDSP_CUDA(cudaFreeHost(ptr));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_device_alloc(int device, void **ptr, size_t bytes)
try {
    if (!ptr) return fail(SRCDSP_E_INVALID, "ptr is null");
    SRCDSP_TRY(check_device(device));
    DeviceGuard g(device);
    SRCDSP_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_device_free(int device, void *ptr)
try {
    DeviceGuard g(device);
    SRCDSP_CUDA(cudaFree(ptr));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_memcpy(int device, void *dst, const void *src, size_t bytes)
try {
    DeviceGuard g(device);
    SRCDSP_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_synth_fill(int device, void *stream, int16_t *d_iq, size_t stride, int channels, size_t n_per_ch,
                      uint32_t seed, uint32_t ch0, uint64_t n0, int amp_shift)
try {
    SRCDSP_TRY(check_device(device));
    if (!d_iq || channels < 1 || amp_shift < 0 || amp_shift > 15) return fail(SRCDSP_E_INVALID, "bad synth arguments");
    if (!is_device_ptr(d_iq)) return fail(SRCDSP_E_INVALID, "srcdsp_synth_fill needs a device pointer");
    if (n_per_ch == 0) return SRCDSP_OK;
    DeviceGuard g(device);
    int bx = (int)std::min<size_t>((n_per_ch + 255) / 256, 1024);
    synth_kernel<<<dim3(bx, channels), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<uint32_t *>(d_iq), stride, (long long)n_per_ch, seed, ch0, n0, amp_shift);
    SRCDSP_LAUNCH_CHECK();
    count_launch();
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

// ---- mixer ------------------------------------------------------------------------------------
int srcdsp_mixer_create(srcdsp_mixer_t *h, int device, int channels, unsigned n_table)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "handle pointer is null");
    *h = nullptr;
    srcdsp_mixer_s *m = new (std::nothrow) srcdsp_mixer_s();
    if (!m) return fail(SRCDSP_E_NOMEM, "out of host memory");
    int st = m->create(device, channels, n_table);
    if (st != SRCDSP_OK) {
        m->destroy();
        delete m;
        return st;
    }
    *h = m;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_destroy(srcdsp_mixer_t h)
try {
    if (!h) return SRCDSP_OK;
    h->destroy();
    delete h;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
static int mixer_for_channels(srcdsp_mixer_t h, int ch, int *c0, int *c1)
{
    CHECK_HANDLE(h);
    if (ch < -1 || ch >= h->C) return fail(SRCDSP_E_INVALID, "channel %d out of range", ch);
    *c0 = ch < 0 ? 0 : ch;
    *c1 = ch < 0 ? h->C : ch + 1;
    return SRCDSP_OK;
}
int srcdsp_mixer_set_frequency(srcdsp_mixer_t h, int ch, float lo)
try {
    int c0, c1;
    SRCDSP_TRY(mixer_for_channels(h, ch, &c0, &c1));
    if (!(lo <= 1 && lo >= -1)) return fail(SRCDSP_E_SIZE, "loFreq %g outside [-1, 1] [mixers.h:54]", (double)lo);
    const int f = MixerBank::quantise(lo, h->n_table);
    for (int c = c0; c < c1; ++c) {
        h->h_nominal[c] = lo;
        h->h_freq[c] = f;
    }
    h->dirty = true;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_set_frequencies(srcdsp_mixer_t h, const float *lo)
try {
    CHECK_HANDLE(h);
    if (!lo) return fail(SRCDSP_E_INVALID, "lo_freq is null");
    for (int c = 0; c < h->C; ++c)
        if (!(lo[c] <= 1 && lo[c] >= -1)) return fail(SRCDSP_E_SIZE, "loFreq[%d] = %g outside [-1, 1]", c, (double)lo[c]);
    for (int c = 0; c < h->C; ++c) {
        h->h_nominal[c] = lo[c];
        h->h_freq[c] = MixerBank::quantise(lo[c], h->n_table);
    }
    h->dirty = true;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_reset(srcdsp_mixer_t h, int ch, float lo)
try {
    int c0, c1;
    SRCDSP_TRY(mixer_for_channels(h, ch, &c0, &c1));
    if (!(lo <= 1 && lo >= -1)) return fail(SRCDSP_E_SIZE, "loFreq %g outside [-1, 1] [mixers.h:54]", (double)lo);
    for (int c = c0; c < c1; ++c) h->h_phi[c] = 0;
    return srcdsp_mixer_set_frequency(h, ch, lo);
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_adjust_frequency(srcdsp_mixer_t h, int ch, float adjust)
try {
    int c0, c1;
    SRCDSP_TRY(mixer_for_channels(h, ch, &c0, &c1));
    for (int c = c0; c < c1; ++c) {
        float nf = h->h_nominal[c];  // mixers.h:91-98
        nf += adjust;
        if (nf > 1) nf -= 2;
        if (nf < -1) nf += 2;
        if (!(nf <= 1 && nf >= -1)) return fail(SRCDSP_E_SIZE, "adjusted frequency %g outside [-1, 1]", (double)nf);
        h->h_nominal[c] = nf;
        h->h_freq[c] = MixerBank::quantise(nf, h->n_table);
    }
    h->dirty = true;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_get_state(srcdsp_mixer_t h, int ch, int *phi, int *freq, float *nominal)
try {
    CHECK_HANDLE(h);
    if (ch < 0 || ch >= h->C) return fail(SRCDSP_E_INVALID, "channel %d out of range", ch);
    if (phi) *phi = h->h_phi[ch];
    if (freq) *freq = h->h_freq[ch];
    if (nominal) *nominal = h->h_nominal[ch];
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_set_state(srcdsp_mixer_t h, int ch, int phi, int freq, float nominal)
try {
    CHECK_HANDLE(h);
    if (ch < 0 || ch >= h->C) return fail(SRCDSP_E_INVALID, "channel %d out of range", ch);
    if (phi < 0 || phi >= (int)h->n_table || freq < 0 || freq >= (int)h->n_table)
        return fail(SRCDSP_E_INVALID, "phi/freq must be in [0, n_table)");
    h->h_phi[ch] = phi;
    h->h_freq[ch] = freq;
    h->h_nominal[ch] = nominal;
    h->dirty = true;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_step(srcdsp_mixer_t h, const int16_t *in, size_t in_stride, int16_t *out, size_t out_stride,
                      size_t n)
try {
    CHECK_HANDLE(h);
    if (n == 0) return SRCDSP_OK;
    if (!in || !out) return fail(SRCDSP_E_INVALID, "null buffer");
    if (h->C > 1 && (in_stride < n || out_stride < n)) return fail(SRCDSP_E_SIZE, "stride smaller than n_per_ch");
    bool dev;
    SRCDSP_TRY(classify(in, out, &dev));
    if (dev)
        return h->step_device(reinterpret_cast<const uint32_t *>(in), in_stride, reinterpret_cast<uint32_t *>(out),
                              out_stride, n);
    return staged_run(*h, in, in_stride, n, out, out_stride, 1, 1, 1, 0,
                      [&](const uint32_t *di, size_t dis, size_t len, uint32_t *dout, size_t dos, bool) {
                          return len ? h->step_device(di, dis, dout, dos, len) : SRCDSP_OK;
                      });
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_set_stream(srcdsp_mixer_t h, void *s)
try {
    CHECK_HANDLE(h);
    return h->set_stream(s);
}
SRCDSP_ABI_CATCH
int srcdsp_mixer_sync(srcdsp_mixer_t h)
try {
    CHECK_HANDLE(h);
    return h->sync();
}
SRCDSP_ABI_CATCH

// ---- decimator --------------------------------------------------------------------------------
int srcdsp_dec_create(srcdsp_dec_t *h, int device, int channels, int M)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "handle pointer is null");
    *h = nullptr;
    if (M < 1 || M > 1024) return fail(SRCDSP_E_INVALID, "M must be in [1, 1024] (got %d)", M);
    srcdsp_dec_s *d = new (std::nothrow) srcdsp_dec_s();
    if (!d) return fail(SRCDSP_E_NOMEM, "out of host memory");
    int st = d->init(device, channels);
    if (st != SRCDSP_OK) {
        delete d;
        return st;
    }
    d->M = M;
    *h = d;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_dec_destroy(srcdsp_dec_t h)
try {
    if (!h) return SRCDSP_OK;
    h->destroy();
    delete h;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_dec_set_coeffs(srcdsp_dec_t h, const int32_t *taps, int ntaps, int require_multiple_of_m)
try {
    CHECK_HANDLE(h);
    return h->set_coeffs(taps, ntaps, require_multiple_of_m);
}
SRCDSP_ABI_CATCH
int srcdsp_dec_set_left_shift(srcdsp_dec_t h, int s)
try {
    CHECK_HANDLE(h);
    h->left_shift = s;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_dec_reset(srcdsp_dec_t h)
try {
    CHECK_HANDLE(h);
    return h->reset();
}
SRCDSP_ABI_CATCH
int srcdsp_dec_get_coeff_scaling(srcdsp_dec_t h, int *cs)
try {
    CHECK_HANDLE(h);
    if (!cs) return fail(SRCDSP_E_INVALID, "null pointer");
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "no coefficients");
    *cs = h->coeff_scaling;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_dec_set_kernel(srcdsp_dec_t h, int kind)
try {
    CHECK_HANDLE(h);
    if (kind < 0 || kind > 5) return fail(SRCDSP_E_INVALID, "kernel kind must be in [0, 5]");
    h->kernel_kind = kind;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_dec_get_last_kernel(srcdsp_dec_t h, int *kind)
try {
    CHECK_HANDLE(h);
    if (!kind) return fail(SRCDSP_E_INVALID, "null pointer");
    *kind = h->last_kernel;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_dec_step(srcdsp_dec_t h, const int16_t *in, size_t in_stride, size_t n_in, int16_t *out,
                    size_t out_stride)
try {
    CHECK_HANDLE(h);
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "decimator has no coefficients (call srcdsp_dec_set_coeffs)");
    if (n_in % (size_t)h->M != 0)
        return fail(SRCDSP_E_SIZE, "n_in (%zu) must be a multiple of M (%d) [dsptl_dnsampling_filters.h:181]", n_in, h->M);
    if (n_in == 0) return SRCDSP_OK;
    if (!in || !out) return fail(SRCDSP_E_INVALID, "null buffer");
    if (h->C > 1 && (in_stride < n_in || out_stride < n_in / h->M)) return fail(SRCDSP_E_SIZE, "stride smaller than block");
    bool dev;
    SRCDSP_TRY(classify(in, out, &dev));
    if (dev) {
        if (device_ranges_overlap(in, in_stride, n_in, out, out_stride, n_in / h->M, h->C, 4))
            return fail(SRCDSP_E_INVALID, "device input and output buffers overlap (in-place filtering needs host buffers)");
        return h->step_device(reinterpret_cast<const uint32_t *>(in), in_stride, n_in,
                              reinterpret_cast<uint32_t *>(out), out_stride, nullptr);
    }
    return staged_run(*h, in, in_stride, n_in, out, out_stride, (size_t)h->M, 1, (size_t)h->M, 0,
                      [&](const uint32_t *di, size_t dis, size_t len, uint32_t *dout, size_t dos, bool) {
                          return h->step_device(di, dis, len, dout, dos, nullptr);
                      });
}
SRCDSP_ABI_CATCH
static int hist_io(Bank &b, uint32_t *d_hist, int H, int ch, int16_t *get, const int16_t *set, size_t *n)
{
    if (ch < 0 || ch >= b.C) return fail(SRCDSP_E_INVALID, "channel %d out of range", ch);
    DeviceGuard g(b.device);
    SRCDSP_CUDA(cudaStreamSynchronize(b.stream));
    if (get) {
        if (n) *n = (size_t)H;
        if (H > 0) SRCDSP_CUDA(cudaMemcpy(get, d_hist + (size_t)ch * H, (size_t)H * 4, cudaMemcpyDeviceToHost));
    } else {
        if (!n || *n != (size_t)H) return fail(SRCDSP_E_SIZE, "state must hold exactly %d samples", H);
        if (H > 0) SRCDSP_CUDA(cudaMemcpy(d_hist + (size_t)ch * H, set, (size_t)H * 4, cudaMemcpyHostToDevice));
    }
    return SRCDSP_OK;
}
int srcdsp_dec_get_state(srcdsp_dec_t h, int ch, int16_t *hist, size_t *n)
try {
    CHECK_HANDLE(h);
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "no coefficients");
    if (!hist) {
        if (n) *n = (size_t)h->H;
        return SRCDSP_OK;
    }
    return hist_io(*h, h->d_hist[h->cur], h->H, ch, hist, nullptr, n);
}
SRCDSP_ABI_CATCH
int srcdsp_dec_set_state(srcdsp_dec_t h, int ch, const int16_t *hist, size_t n)
try {
    CHECK_HANDLE(h);
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "no coefficients");
    if (!hist && h->H > 0) return fail(SRCDSP_E_INVALID, "null state");
    return hist_io(*h, h->d_hist[h->cur], h->H, ch, nullptr, hist, &n);
}
SRCDSP_ABI_CATCH
int srcdsp_dec_set_stream(srcdsp_dec_t h, void *s)
try {
    CHECK_HANDLE(h);
    return h->set_stream(s);
}
SRCDSP_ABI_CATCH
int srcdsp_dec_sync(srcdsp_dec_t h)
try {
    CHECK_HANDLE(h);
    return h->sync();
}
SRCDSP_ABI_CATCH

// ---- fused DDC chain ----------------------------------------------------------------------------
int srcdsp_ddc_create(srcdsp_ddc_t *h, srcdsp_mixer_t mixer, srcdsp_dec_t dec1, srcdsp_dec_t dec2)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "handle pointer is null");
    *h = nullptr;
    if (!dec1) return fail(SRCDSP_E_INVALID, "dec1 is required");
    if (mixer && (mixer->C != dec1->C || mixer->device != dec1->device))
        return fail(SRCDSP_E_INVALID, "mixer and dec1 must have the same channels and device");
    if (dec2 && (dec2->C != dec1->C || dec2->device != dec1->device))
        return fail(SRCDSP_E_INVALID, "dec1 and dec2 must have the same channels and device");
    srcdsp_ddc_s *d = new (std::nothrow) srcdsp_ddc_s();
    if (!d) return fail(SRCDSP_E_NOMEM, "out of host memory");
    d->mixer = mixer;
    d->d1 = dec1;
    d->d2 = dec2;
    // one stream for the whole chain so that phase / history ping-pong stays ordered: the members follow dec1
    // (tracked borrow, see Bank::follow) until the chain -- or dec1 -- is destroyed
    if (mixer) mixer->follow(dec1);
    if (dec2) dec2->follow(dec1);
    *h = d;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_ddc_destroy(srcdsp_ddc_t h)
try {
    if (!h) return SRCDSP_OK;
    if (h->d_mid) {
        DeviceGuard g(h->d1->device);
        cudaStreamSynchronize(h->d1->stream);
        cudaFree(h->d_mid);
    }
    // the members get private streams back (no-ops for members that were destroyed first or re-chained since)
    if (h->mixer && h->mixer->leader == h->d1) h->mixer->unfollow();
    if (h->d2 && h->d2->leader == h->d1) h->d2->unfollow();
    delete h;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_ddc_step(srcdsp_ddc_t h, const int16_t *in, size_t in_stride, size_t n_in, int16_t *out,
                    size_t out_stride)
try {
    CHECK_HANDLE(h);
    const size_t Mt = (size_t)h->d1->M * (h->d2 ? (size_t)h->d2->M : 1);
    if (h->d1->ntaps == 0 || (h->d2 && h->d2->ntaps == 0)) return fail(SRCDSP_E_STATE, "a decimator has no coefficients");
    if (n_in % Mt != 0) return fail(SRCDSP_E_SIZE, "n_in (%zu) must be a multiple of the total decimation %zu", n_in, Mt);
    if (n_in == 0) return SRCDSP_OK;
    if (!in || !out) return fail(SRCDSP_E_INVALID, "null buffer");
    if (h->d1->C > 1 && (in_stride < n_in || out_stride < n_in / Mt)) return fail(SRCDSP_E_SIZE, "stride smaller than block");
    bool dev;
    SRCDSP_TRY(classify(in, out, &dev));
    if (dev) {
        if (device_ranges_overlap(in, in_stride, n_in, out, out_stride, n_in / Mt, h->d1->C, 4))
            return fail(SRCDSP_E_INVALID, "device input and output buffers overlap (in-place filtering needs host buffers)");
        return h->step_device(reinterpret_cast<const uint32_t *>(in), in_stride, n_in,
                              reinterpret_cast<uint32_t *>(out), out_stride);
    }
    return staged_run(*h->d1, in, in_stride, n_in, out, out_stride, Mt, 1, Mt, 0,
                      [&](const uint32_t *di, size_t dis, size_t len, uint32_t *dout, size_t dos, bool) {
                          return h->step_device(di, dis, len, dout, dos);
                      });
}
SRCDSP_ABI_CATCH
int srcdsp_ddc_set_stream(srcdsp_ddc_t h, void *s)
try {
    CHECK_HANDLE(h);
    return h->d1->set_stream(s);  // reaches the members: they follow dec1 (Bank::follow)
}
SRCDSP_ABI_CATCH
int srcdsp_ddc_sync(srcdsp_ddc_t h)
try {
    CHECK_HANDLE(h);
    return h->d1->sync();
}
SRCDSP_ABI_CATCH

// ---- upsampler ----------------------------------------------------------------------------------
int srcdsp_up_create(srcdsp_up_t *h, int device, int channels, int L)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "handle pointer is null");
    *h = nullptr;
    if (L < 1 || L > UP_NT) return fail(SRCDSP_E_INVALID, "L must be in [1, %d] (got %d)", UP_NT, L);
    srcdsp_up_s *u = new (std::nothrow) srcdsp_up_s();
    if (!u) return fail(SRCDSP_E_NOMEM, "out of host memory");
    int st = u->init(device, channels);
    if (st != SRCDSP_OK) {
        delete u;
        return st;
    }
    u->L = L;
    *h = u;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_destroy(srcdsp_up_t h)
try {
    if (!h) return SRCDSP_OK;
    h->destroy();
    delete h;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_set_coefficients(srcdsp_up_t h, const int32_t *taps, int ntaps)
try {
    CHECK_HANDLE(h);
    return h->set_coefficients(taps, ntaps);
}
SRCDSP_ABI_CATCH
int srcdsp_up_reset(srcdsp_up_t h)
try {
    CHECK_HANDLE(h);
    return h->reset();
}
SRCDSP_ABI_CATCH
int srcdsp_up_step(srcdsp_up_t h, const int16_t *in, size_t in_stride, size_t n_in, int16_t *out,
                   size_t out_stride, int flush, int shift_mode)
try {
    CHECK_HANDLE(h);
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "upsampler has no coefficients [upsampling_filters.h:155]");
    if (shift_mode != 0 && shift_mode != 1) return fail(SRCDSP_E_INVALID, "shift_mode must be 0 or 1");
    const size_t L = (size_t)h->L;
    const size_t n_flush = flush ? (size_t)h->length / L : 0;  // upsampling_filters.h:199
    const size_t n_tot = n_in + n_flush;
    if (n_tot == 0) return SRCDSP_OK;
    if (!out || (!in && n_in)) return fail(SRCDSP_E_INVALID, "null buffer");
    if (h->C > 1 && (in_stride < n_in || out_stride < n_tot * L)) return fail(SRCDSP_E_SIZE, "stride smaller than block");
    bool dev;
    if (n_in) {
        SRCDSP_TRY(classify(in, out, &dev));
    } else {
        dev = is_device_ptr(out);
    }
    if (dev) {
        if (n_in && device_ranges_overlap(in, in_stride, n_in, out, out_stride, n_tot * L, h->C, 4))
            return fail(SRCDSP_E_INVALID, "device input and output buffers overlap");
        return h->step_device(reinterpret_cast<const uint32_t *>(in), in_stride, n_in, n_flush,
                              reinterpret_cast<uint32_t *>(out), out_stride, shift_mode);
    }
    return staged_run(*h, in, in_stride, n_in, out, out_stride, 1, L, 1, n_flush * L,
                      [&](const uint32_t *di, size_t dis, size_t len, uint32_t *dout, size_t dos, bool last) {
                          return h->step_device(di, dis, len, last ? n_flush : 0, dout, dos, shift_mode);
                      });
}
SRCDSP_ABI_CATCH
int srcdsp_up_get_length(srcdsp_up_t h, int *v)
try {
    CHECK_HANDLE(h);
    if (!v) return fail(SRCDSP_E_INVALID, "null pointer");
    *v = h->length;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_get_imp_length(srcdsp_up_t h, int *v)
try {
    CHECK_HANDLE(h);
    if (!v) return fail(SRCDSP_E_INVALID, "null pointer");
    *v = h->imp_length;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_get_ratio(srcdsp_up_t h, int *v)
try {
    CHECK_HANDLE(h);
    if (!v) return fail(SRCDSP_E_INVALID, "null pointer");
    *v = h->L;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_get_last_kernel(srcdsp_up_t h, int *v)
try {
    CHECK_HANDLE(h);
    if (!v) return fail(SRCDSP_E_INVALID, "null pointer");
    *v = h->last_kernel;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
// the public state is the H-1 samples the next outputs depend on, oldest first
int srcdsp_up_get_state(srcdsp_up_t h, int ch, int16_t *hist, size_t *n)
try {
    CHECK_HANDLE(h);
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "no coefficients");
    if (!hist) {
        if (n) *n = (size_t)(h->H - 1);
        return SRCDSP_OK;
    }
    if (ch < 0 || ch >= h->C) return fail(SRCDSP_E_INVALID, "channel %d out of range", ch);
    DeviceGuard g(h->device);
    SRCDSP_CUDA(cudaStreamSynchronize(h->stream));
    if (n) *n = (size_t)(h->H - 1);
    if (h->H > 1)
        SRCDSP_CUDA(cudaMemcpy(hist, h->d_hist[h->cur] + (size_t)ch * h->H + 1, (size_t)(h->H - 1) * 4,
                               cudaMemcpyDeviceToHost));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_set_state(srcdsp_up_t h, int ch, const int16_t *hist, size_t n)
try {
    CHECK_HANDLE(h);
    if (h->ntaps == 0) return fail(SRCDSP_E_STATE, "no coefficients");
    if (ch < 0 || ch >= h->C) return fail(SRCDSP_E_INVALID, "channel %d out of range", ch);
    if (n != (size_t)(h->H - 1)) return fail(SRCDSP_E_SIZE, "state must hold exactly %d samples", h->H - 1);
    DeviceGuard g(h->device);
    SRCDSP_CUDA(cudaStreamSynchronize(h->stream));
    if (h->H > 1)
        SRCDSP_CUDA(cudaMemcpy(h->d_hist[h->cur] + (size_t)ch * h->H + 1, hist, (size_t)(h->H - 1) * 4,
                               cudaMemcpyHostToDevice));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_up_set_stream(srcdsp_up_t h, void *s)
try {
    CHECK_HANDLE(h);
    return h->set_stream(s);
}
SRCDSP_ABI_CATCH
int srcdsp_up_sync(srcdsp_up_t h)
try {
    CHECK_HANDLE(h);
    return h->sync();
}
SRCDSP_ABI_CATCH


// ---- correlator ---------------------------------------------------------------------------------
int srcdsp_corr_create(srcdsp_corr_t *h, int device, int channels, int N, int S)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "handle pointer is null");
    srcdsp_corr_s *b = new (std::nothrow) srcdsp_corr_s;
    if (!b) return fail(SRCDSP_E_NOMEM, "out of memory");
    int st = b->create(device, channels, N, S);
    if (st != SRCDSP_OK) {
        delete b;
        return st;
    }
    *h = b;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_corr_destroy(srcdsp_corr_t h)
try {
    if (!h) return SRCDSP_OK;
    h->destroy();
    delete h;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_corr_set_pattern(srcdsp_corr_t h, const int32_t *pattern_iq, double threshold_coeff)
try {
    CHECK_HANDLE(h);
    return h->set_pattern(pattern_iq, threshold_coeff);
}
SRCDSP_ABI_CATCH
int srcdsp_corr_reset(srcdsp_corr_t h)
try {
    CHECK_HANDLE(h);
    return h->reset();
}
SRCDSP_ABI_CATCH
int srcdsp_corr_step(srcdsp_corr_t h, const int16_t *in_iq, size_t in_stride, size_t n_per_ch, int *found, int *corr_index)
try {
    CHECK_HANDLE(h);
    return h->step(in_iq, in_stride, n_per_ch, found, corr_index);
}
SRCDSP_ABI_CATCH
int srcdsp_corr_get_ref_bit_samples(srcdsp_corr_t h, int ch, int16_t *bits_iq)
try {
    CHECK_HANDLE(h);
    if (ch < 0 || ch >= h->C || !bits_iq) return fail(SRCDSP_E_INVALID, "bad channel or null pointer");
    DeviceGuard g(h->device);
    SRCDSP_CUDA(cudaStreamSynchronize(h->stream));
    SRCDSP_CUDA(cudaMemcpy(bits_iq, h->d_bits + (size_t)ch * h->N, (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_corr_get_status(srcdsp_corr_t h, int ch, uint32_t *energy_value3, uint32_t *corr_value3, uint32_t *coeffs_energy,
                           int *coeff_scaling, double *threshold_factor)
try {
    CHECK_HANDLE(h);
    if (ch < 0 || ch >= h->C) return fail(SRCDSP_E_INVALID, "bad channel");
    DeviceGuard g(h->device);
    uint32_t r[6];
    SRCDSP_CUDA(cudaStreamSynchronize(h->stream));
    SRCDSP_CUDA(cudaMemcpy(r, h->d_reg[h->cur] + (size_t)ch * 6, sizeof r, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) {
        if (corr_value3) corr_value3[k] = r[k];
        if (energy_value3) energy_value3[k] = r[3 + k];
    }
    if (coeffs_energy) *coeffs_energy = h->coeffs_energy;
    if (coeff_scaling) *coeff_scaling = h->coeff_scaling;
    if (threshold_factor) *threshold_factor = h->threshold_factor;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH
int srcdsp_corr_set_stream(srcdsp_corr_t h, void *cuda_stream)
try {
    CHECK_HANDLE(h);
    return h->set_stream(cuda_stream);
}
SRCDSP_ABI_CATCH

}  // extern "C"

#include "capi_decf.inc"
