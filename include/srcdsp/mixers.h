/*
 * srcdsp/mixers.h -- drop-in for the reference's mixers.h (table-lookup complex NCO mixer).
 *
 * Same name, template parameters and members as dsptl::Mixer (reference mixers.h:130-137) and its
 * base dsptl::_Mixer (mixers.h:26-41); only the specialisation the reference defines exists:
 * Mixer<std::complex<int16_t>, std::complex<int16_t>, int16_t, N>.  step() runs on the GPU
 * through srcdsp_mixer_step (../srcdsp_b200.h).
 */
#ifndef SRCDSP_DROPIN_MIXERS_H
#define SRCDSP_DROPIN_MIXERS_H

#include "detail.h"

namespace dsptl {

/* dsptl::_Mixer (mixers.h:26-41): the base that carries frequency and phase.  Same template parameters and public
 * members; the protected data members of the reference (phi, freq, nominalFreq, ptable) live in the GPU bank here,
 * so a class derived from _Mixer reaches them through state() / set_state() instead of by name.  PhaseType must be
 * able to hold N (reference :23-25); the table has N entries of amplitude 16383 (mixers.h:149-160). */
template <class InType, class OutType, class PhaseType, unsigned N = 4096>
class _Mixer {
public:
    _Mixer() : h_(nullptr) { srcdsp_dropin::check(srcdsp_mixer_create(&h_, srcdsp_dropin::default_device(), 1, N), "_Mixer()"); }
    _Mixer(const _Mixer &o) : h_(nullptr)
    {
        srcdsp_dropin::check(srcdsp_mixer_create(&h_, srcdsp_dropin::default_device(), 1, N), "_Mixer(copy)");
        copy_state(o);
    }
    _Mixer &operator=(const _Mixer &o)
    {
        if (this != &o) copy_state(o);
        return *this;
    }
    ~_Mixer() { srcdsp_mixer_destroy(h_); }

    /* mixers.h:51-67 */
    void setFrequency(float loFreq) { srcdsp_dropin::check(srcdsp_mixer_set_frequency(h_, 0, loFreq), "Mixer::setFrequency"); }
    /* mixers.h:76-81 */
    void reset(float loFreq = 0) { srcdsp_dropin::check(srcdsp_mixer_reset(h_, 0, loFreq), "Mixer::reset"); }
    /* mixers.h:91-98 */
    void adjustFrequency(float loFreq = 0)
    {
        srcdsp_dropin::check(srcdsp_mixer_adjust_frequency(h_, 0, loFreq), "Mixer::adjustFrequency");
    }

    /* extension: the C-ABI handle, e.g. to build a fused srcdsp_ddc chain */
    srcdsp_mixer_t handle() const { return h_; }

protected:
    /* the reference's protected phi / freq / nominalFreq */
    void state(PhaseType &phi, PhaseType &freq, float &nominalFreq) const
    {
        int p, f;
        srcdsp_dropin::check(srcdsp_mixer_get_state(h_, 0, &p, &f, &nominalFreq), "_Mixer::state");
        phi = static_cast<PhaseType>(p), freq = static_cast<PhaseType>(f);
    }
    void set_state(PhaseType phi, PhaseType freq, float nominalFreq)
    {
        srcdsp_dropin::check(srcdsp_mixer_set_state(h_, 0, (int)phi, (int)freq, nominalFreq), "_Mixer::set_state");
    }
    void copy_state(const _Mixer &o)
    {
        int phi, freq;
        float nominal;
        srcdsp_dropin::check(srcdsp_mixer_get_state(o.h_, 0, &phi, &freq, &nominal), "Mixer copy");
        srcdsp_dropin::check(srcdsp_mixer_set_state(h_, 0, phi, freq, nominal), "Mixer copy");
    }
    srcdsp_mixer_t h_;
};

template <class InType, class OutType, class PhaseType, unsigned N>
class Mixer;  // like the reference, the primary template is only declared (mixers.h:120-121)

/* mixers.h:130-137: the one specialisation the reference defines, derived from _Mixer like there */
template <unsigned N>
class Mixer<std::complex<int16_t>, std::complex<int16_t>, int16_t, N>
    : public dsptl::_Mixer<std::complex<int16_t>, std::complex<int16_t>, int16_t, N> {
    typedef dsptl::_Mixer<std::complex<int16_t>, std::complex<int16_t>, int16_t, N> Base;

public:
    /* mixers.h:149-160: builds the N-entry sine table; phase and frequency start at zero */
    Mixer() : Base() {}

    /* mixers.h:168-188: out must hold at least in.size() samples; out may be in */
    void step(std::vector<std::complex<int16_t>> &in, std::vector<std::complex<int16_t>> &out)
    {
        if (in.empty()) return;
        srcdsp_dropin::check(out.size() >= in.size() ? SRCDSP_OK : SRCDSP_E_SIZE, "Mixer::step (out too small)");
        srcdsp_dropin::check(srcdsp_mixer_step(this->h_, srcdsp_dropin::iq(in), in.size(), srcdsp_dropin::iq(out),
                                               out.size(), in.size()),
                             "Mixer::step");
    }
};

}  // namespace dsptl

#endif
