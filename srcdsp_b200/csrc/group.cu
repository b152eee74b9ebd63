// group.cu -- one DDC chain (optional NCO mixer -> decimator [-> decimator]) spread over several devices of one box and
// driven from ONE process (north_star: "work partitions across the 8 GPUs of one box by independent
// channel batches (or time-sliced blocks with halo for a single very long stream) ... no cross-GPU reduction ... results
// are gathered to the host with async copies from pinned memory").
//
// There is no collective and no peer traffic on this path.  Each device gets its own banks (srcdsp_mixer / srcdsp_dec /
// srcdsp_ddc handles of the C ABI) and, for the duration of a step, its own host thread; every thread runs the usual
// staged host pipeline of its chain (pinned H2D -> kernels -> D2H on three streams of its device) on ITS part of the
// caller's buffers and writes its outputs straight to their final place in the caller's output buffer:
//
//   channel batches (SRCDSP_GROUP_CHANNELS): device g owns the contiguous channels [C*g/G, C*(g+1)/G) with their streaming
//       state resident on it; nothing is shared.
//   time slices (SRCDSP_GROUP_SLICES): every device holds all channels; a block of n samples is cut into G slices at
//       multiples of the total decimation.  Slice g > 0 is preceded by a warm-up run over the `warm` samples in front
//       of it (warm = the chain's delay-line reach (N1-1) + (N2-1)*M1 rounded up to the total decimation; outputs
//       discarded) that starts from zero history and from the NCO phase (phi0 + (start - warm) * freq) mod N in closed
//       form -- after it, history and phase are exactly what the reference's sequential run would carry at `start`
//       (dsptl_dnsampling_filters.h:198-205,218-219; mixers.h:177).  After the block the state of the last slice is
//       copied to device 0's banks, which always hold the stream's state between calls.
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct Member {
    int device = 0;
    int ch0 = 0, C = 0;  // channels [ch0, ch0 + C) of the group (all of them in slice mode)
    srcdsp_mixer_t mixer = nullptr;
    srcdsp_dec_t d1 = nullptr, d2 = nullptr;
    srcdsp_ddc_t chain = nullptr;
    std::vector<int16_t> scratch;  // warm-up outputs (discarded)
    int status = SRCDSP_OK;
    std::string error;
};

}  // namespace

struct srcdsp_group_s {
    int mode = 0, channels = 0, M1 = 1, M2 = 0;
    unsigned n_table = 0;
    int ntaps1 = 0, ntaps2 = 0;
    std::vector<Member> members;
    int last_used = 0;  // devices that took part in the last step
};

using srcdsp::fail;

#define GROUP_TRY_MEMBER(m, expr)                          \
    do {                                                   \
        int _s = (expr);                                   \
        if (_s != SRCDSP_OK) {                             \
            (m).status = _s;                               \
            (m).error = srcdsp_last_error();               \
            return;                                        \
        }                                                  \
    } while (0)

// Starting a host thread can fail (std::system_error) and a member's part can run out of host memory: nothing may be
// thrown across the C ABI or out of a std::thread (that would be std::terminate), so a part that cannot get a thread
// runs on the calling thread, and an exception inside a part becomes that member's status.
template <class F>
static void start_member(std::vector<std::thread> &threads, Member *m, F fn)
{
    auto guarded = [m, fn] {
        try {
            fn();
        } catch (...) {
            m->status = srcdsp::abi_exception();
            try {
                m->error = srcdsp_last_error();
            } catch (...) {
            }
        }
    };
    try {
        threads.emplace_back(guarded);
    } catch (...) {
        guarded();
    }
}

static int collect(srcdsp_group_s *g, int used)
{
    for (int i = 0; i < used; ++i)
        if (g->members[i].status != SRCDSP_OK)
            return fail(g->members[i].status, "device %d: %s", g->members[i].device, g->members[i].error.c_str());
    return SRCDSP_OK;
}

extern "C" {

int srcdsp_group_create(srcdsp_group_t *h, int mode, const int *devices, int n_devices, int channels, unsigned n_table, int M1, int M2)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "handle pointer is null");
    *h = nullptr;
    if (mode != SRCDSP_GROUP_CHANNELS && mode != SRCDSP_GROUP_SLICES) return fail(SRCDSP_E_INVALID, "mode must be SRCDSP_GROUP_CHANNELS or SRCDSP_GROUP_SLICES");
    if (!devices || n_devices < 1) return fail(SRCDSP_E_INVALID, "need at least one device");
    if (channels < 1) return fail(SRCDSP_E_INVALID, "channels must be >= 1");
    if (mode == SRCDSP_GROUP_CHANNELS && n_devices > channels) n_devices = channels;  // a device without channels does nothing
    std::unique_ptr<srcdsp_group_s> g(new (std::nothrow) srcdsp_group_s());
    if (!g) return fail(SRCDSP_E_NOMEM, "out of host memory");
    g->mode = mode, g->channels = channels, g->M1 = M1, g->M2 = M2, g->n_table = n_table;
    g->members.resize(n_devices);
    int st = SRCDSP_OK;
    for (int i = 0; i < n_devices && st == SRCDSP_OK; ++i) {
        Member &m = g->members[i];
        m.device = devices[i];
        if (mode == SRCDSP_GROUP_CHANNELS) {  // contiguous, balanced batches (the first channels % G devices take one more)
            const int base = channels / n_devices, extra = channels % n_devices;
            m.ch0 = i * base + (i < extra ? i : extra);
            m.C = base + (i < extra ? 1 : 0);
        } else {
            m.ch0 = 0, m.C = channels;
        }
        if (n_table) st = srcdsp_mixer_create(&m.mixer, m.device, m.C, n_table);
        if (st == SRCDSP_OK) st = srcdsp_dec_create(&m.d1, m.device, m.C, M1);
        if (st == SRCDSP_OK && M2 > 0) st = srcdsp_dec_create(&m.d2, m.device, m.C, M2);
        if (st == SRCDSP_OK) st = srcdsp_ddc_create(&m.chain, m.mixer, m.d1, m.d2);
    }
    if (st != SRCDSP_OK) {
        const std::string msg = srcdsp_last_error();
        srcdsp_group_destroy(g.release());
        return fail(st, "%s", msg.c_str());
    }
    *h = g.release();
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_destroy(srcdsp_group_t h)
try {
    if (!h) return SRCDSP_OK;
    for (Member &m : h->members) {
        if (m.chain) srcdsp_ddc_destroy(m.chain);  // before its members
        if (m.mixer) srcdsp_mixer_destroy(m.mixer);
        if (m.d2) srcdsp_dec_destroy(m.d2);
        if (m.d1) srcdsp_dec_destroy(m.d1);
    }
    delete h;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_set_coeffs(srcdsp_group_t h, int stage, const int32_t *taps, int ntaps, int require_multiple_of_m)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "null handle");
    if (stage != 1 && !(stage == 2 && h->M2 > 0)) return fail(SRCDSP_E_INVALID, "stage must be 1%s", h->M2 > 0 ? " or 2" : " (single-stage group)");
    for (Member &m : h->members) SRCDSP_TRY(srcdsp_dec_set_coeffs(stage == 1 ? m.d1 : m.d2, taps, ntaps, require_multiple_of_m));
    (stage == 1 ? h->ntaps1 : h->ntaps2) = ntaps;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_set_frequencies(srcdsp_group_t h, const float *lo_freq)
try {
    if (!h || !lo_freq) return fail(SRCDSP_E_INVALID, "null handle / pointer");
    if (!h->n_table) return fail(SRCDSP_E_STATE, "the group was created without a mixer (n_table == 0)");
    for (Member &m : h->members) SRCDSP_TRY(srcdsp_mixer_set_frequencies(m.mixer, lo_freq + m.ch0));
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_reset(srcdsp_group_t h)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "null handle");
    for (Member &m : h->members) {
        if (m.mixer)
            for (int c = 0; c < m.C; ++c) {  // _Mixer::reset keeps no frequency: phase 0 at the current frequency
                int phi, freq;
                float nominal;
                SRCDSP_TRY(srcdsp_mixer_get_state(m.mixer, c, &phi, &freq, &nominal));
                SRCDSP_TRY(srcdsp_mixer_set_state(m.mixer, c, 0, freq, nominal));
            }
        SRCDSP_TRY(srcdsp_dec_reset(m.d1));
        if (m.d2) SRCDSP_TRY(srcdsp_dec_reset(m.d2));
    }
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_get_layout(srcdsp_group_t h, int index, int *device, int *ch0, int *n_channels)
try {
    if (!h || index < 0 || index >= (int)h->members.size()) return fail(SRCDSP_E_INVALID, "bad handle / index");
    if (device) *device = h->members[index].device;
    if (ch0) *ch0 = h->members[index].ch0;
    if (n_channels) *n_channels = h->members[index].C;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_size(srcdsp_group_t h, int *n_members, int *used_in_last_step)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "null handle");
    if (n_members) *n_members = (int)h->members.size();
    if (used_in_last_step) *used_in_last_step = h->last_used;
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

int srcdsp_group_step(srcdsp_group_t h, const int16_t *in, size_t in_stride, size_t n_in, int16_t *out, size_t out_stride)
try {
    if (!h) return fail(SRCDSP_E_INVALID, "null handle");
    if (!in || !out) return fail(SRCDSP_E_INVALID, "null buffer");
    const size_t Mt = (size_t)h->M1 * (size_t)(h->M2 > 0 ? h->M2 : 1);
    if (n_in % Mt != 0) return fail(SRCDSP_E_SIZE, "n_in (%zu) must be a multiple of the total decimation %zu", n_in, Mt);
    if (h->ntaps1 == 0 || (h->M2 > 0 && h->ntaps2 == 0)) return fail(SRCDSP_E_STATE, "a decimator has no coefficients");
    if (n_in == 0) return SRCDSP_OK;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, in) == cudaSuccess && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged))
        return fail(SRCDSP_E_INVALID, "a group takes HOST buffers (pinned for speed: srcdsp_host_alloc); every device copies its own part");
    cudaGetLastError();
    const int G = (int)h->members.size();
    std::vector<std::thread> threads;
    threads.reserve(G);

    if (h->mode == SRCDSP_GROUP_CHANNELS) {
        for (int i = 0; i < G; ++i) {
            Member *m = &h->members[i];
            m->status = SRCDSP_OK;
            start_member(threads, m, [=] {
                GROUP_TRY_MEMBER(*m, srcdsp_ddc_step(m->chain, in + 2 * (size_t)m->ch0 * in_stride, in_stride, n_in,
                                                     out + 2 * (size_t)m->ch0 * out_stride, out_stride));
            });
        }
        for (auto &t : threads) t.join();
        h->last_used = G;
        return collect(h, G);
    }

    // ---- time slices -----------------------------------------------------------------------------------------------
    const size_t halo = (size_t)(h->ntaps1 - 1) + (h->M2 > 0 ? (size_t)(h->ntaps2 - 1) * h->M1 : 0);
    const size_t warm = (halo + Mt - 1) / Mt * Mt;
    const size_t n_out = n_in / Mt;
    // every slice but the first needs `warm` samples of THIS block in front of it
    int used = G;
    while (used > 1 && (n_out / used) * Mt < warm) --used;
    const int C = h->channels;
    Member &home = h->members[0];
    // state at the start of the block: device 0's banks
    std::vector<int> phi0(C, 0), freq(C, 0);
    std::vector<float> nominal(C, 0.f);
    if (home.mixer)
        for (int c = 0; c < C; ++c) SRCDSP_TRY(srcdsp_mixer_get_state(home.mixer, c, &phi0[c], &freq[c], &nominal[c]));
    for (int i = 0; i < used; ++i) {
        Member *m = &h->members[i];
        m->status = SRCDSP_OK;
        const size_t o0 = n_out * i / used, o1 = n_out * (i + 1) / used;
        const size_t start = o0 * Mt, len = (o1 - o0) * Mt;
        start_member(threads, m, [=, &phi0, &freq, &nominal] {
            if (i > 0) {
                for (int c = 0; c < C; ++c) {
                    if (m->mixer) {
                        const unsigned long long adv = (unsigned long long)((start - warm) % h->n_table) * (unsigned long long)freq[c];
                        GROUP_TRY_MEMBER(*m, srcdsp_mixer_set_state(m->mixer, c, (int)(((unsigned long long)phi0[c] + adv) % h->n_table), freq[c], nominal[c]));
                    }
                }
                GROUP_TRY_MEMBER(*m, srcdsp_dec_reset(m->d1));
                if (m->d2) GROUP_TRY_MEMBER(*m, srcdsp_dec_reset(m->d2));
                if (warm) {
                    m->scratch.resize((size_t)2 * C * (warm / Mt));
                    GROUP_TRY_MEMBER(*m, srcdsp_ddc_step(m->chain, in + 2 * (start - warm), in_stride, warm, m->scratch.data(), warm / Mt));
                }
            }
            GROUP_TRY_MEMBER(*m, srcdsp_ddc_step(m->chain, in + 2 * start, in_stride, len, out + 2 * o0, out_stride));
        });
    }
    for (auto &t : threads) t.join();
    h->last_used = used;
    SRCDSP_TRY(collect(h, used));
    // the stream's state lives on device 0 between calls: bring the last slice's state home
    if (used > 1) {
        Member &last = h->members[used - 1];
        std::vector<int16_t> hist;
        for (int c = 0; c < C; ++c) {
            if (home.mixer) {
                int p, f;
                float nm;
                SRCDSP_TRY(srcdsp_mixer_get_state(last.mixer, c, &p, &f, &nm));
                SRCDSP_TRY(srcdsp_mixer_set_state(home.mixer, c, p, f, nm));
            }
            for (int s = 0; s < (h->M2 > 0 ? 2 : 1); ++s) {
                size_t n = 0;
                SRCDSP_TRY(srcdsp_dec_get_state(s ? last.d2 : last.d1, c, nullptr, &n));
                hist.resize(2 * n + 2);
                SRCDSP_TRY(srcdsp_dec_get_state(s ? last.d2 : last.d1, c, hist.data(), &n));
                SRCDSP_TRY(srcdsp_dec_set_state(s ? home.d2 : home.d1, c, hist.data(), n));
            }
        }
    }
    return SRCDSP_OK;
}
SRCDSP_ABI_CATCH

}  // extern "C"
