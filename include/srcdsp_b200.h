/*
 * srcdsp_b200.h -- C ABI of the B200-native SrcDsp DDC hot path (libsrcdsp_b200.so).
 *
 * The reference (dogjin/SrcDsp) has no FFI boundary: its "interface" is the public API of three
 * header-only class templates.  This ABI is what the drop-in classes in include/srcdsp/ (mixers.h, ...)
 * forward to; each entry point cites the reference member it replaces (file:line into the
 * reference checkout).  Template parameters of the reference (M, L, N) are runtime arguments.
 *
 * Conventions
 *   - every function returns an int status: 0 = SRCDSP_OK, negative = SRCDSP_E_*;
 *     srcdsp_last_error() returns a thread-local message for the last failure.  Nothing throws:
 *     every entry point catches C++ exceptions at the boundary (std::bad_alloc -> SRCDSP_E_NOMEM).
 *   - a handle is a BANK of `channels` independent streams that share taps (the reference
 *     models a channel as one object; a bank with channels == 1 is exactly one object).
 *   - samples are interleaved I/Q int16 (byte-identical to std::vector<std::complex<int16_t>>
 *     ::data()); a bank buffer is channel-major: channel c starts at ptr + 2*c*stride int16
 *     (stride in complex samples).  Pointers may be device, pinned-host or pageable-host memory
 *     (detected with cudaPointerGetAttributes); host buffers are staged through a chunked
 *     H2D -> kernel -> D2H pipeline.  Input and output must be the same kind.
 *   - one host thread per handle at a time (same contract as the reference objects); different
 *     handles are independent.  step() with host buffers returns when the output is complete;
 *     with device buffers it is asynchronous on the handle's stream (srcdsp_*_sync to wait).
 *   - there is NO CPU fallback: every step runs hand-written sm_100a kernels or fails.
 */
#ifndef SRCDSP_B200_H
#define SRCDSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCDSP_OK 0
#define SRCDSP_E_INVALID (-1) /* bad argument / handle                                         */
#define SRCDSP_E_SIZE (-2)    /* a size precondition the reference assert()s on is violated    */
#define SRCDSP_E_CUDA (-3)    /* CUDA runtime error (message in srcdsp_last_error)            */
#define SRCDSP_E_NOMEM (-4)
#define SRCDSP_E_STATE (-5)   /* e.g. step() before coefficients were set                      */
#define SRCDSP_E_NOGPU (-6)   /* no sm_100 device / driver: the library refuses to run        */

typedef struct srcdsp_mixer_s *srcdsp_mixer_t;
typedef struct srcdsp_dec_s *srcdsp_dec_t;
typedef struct srcdsp_up_s *srcdsp_up_t;
typedef struct srcdsp_ddc_s *srcdsp_ddc_t;
typedef struct srcdsp_fifo_s *srcdsp_fifo_t;
typedef struct srcdsp_corr_s *srcdsp_corr_t;
typedef struct srcdsp_decf_s *srcdsp_decf_t;
typedef struct srcdsp_group_s *srcdsp_group_t;

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                    */
const char *srcdsp_last_error(void);
int srcdsp_version(void);
int srcdsp_device_count(int *count);
/* kernel launch counter (all handles, this process): used by bench.py for "gpu_launches". */
uint64_t srcdsp_launch_count(void);

/* pinned host memory for zero-staging transfers, and plain device memory */
int srcdsp_host_alloc(void **ptr, size_t bytes);
int srcdsp_host_free(void *ptr);
int srcdsp_device_alloc(int device, void **ptr, size_t bytes);
int srcdsp_device_free(int device, void *ptr);
int srcdsp_memcpy(int device, void *dst, const void *src, size_t bytes); /* any <-> any, sync */

/* synthetic multi-channel complex baseband, counter-based: sample n of channel c depends only
 * on (seed, ch0 + c, n0 + n); oracle/srcdsp_oracle.c:orc_synth_fill is the host twin.
 * d_iq: device pointer, [channels][stride] complex int16.  stream may be NULL. */
int srcdsp_synth_fill(int device, void *stream, int16_t *d_iq, size_t stride, int channels,
                      size_t n_per_ch, uint32_t seed, uint32_t ch0, uint64_t n0, int amp_shift);

/* ------------------------------------------------------------------------------------------ */
/* Mixer bank -- dsptl::Mixer<complex<int16_t>, complex<int16_t>, int16_t, N>                 */
/* ctor: mixers.h:149-160 (sine table of n_table entries, amplitude 16383, phase = freq = 0)  */
int srcdsp_mixer_create(srcdsp_mixer_t *h, int device, int channels, unsigned n_table);
int srcdsp_mixer_destroy(srcdsp_mixer_t h);
/* _Mixer::setFrequency  mixers.h:51-67   (ch = -1: all channels; |lo_freq| <= 1 else E_SIZE) */
int srcdsp_mixer_set_frequency(srcdsp_mixer_t h, int ch, float lo_freq);
/* per-channel frequencies in one call: lo_freq[channels]                                      */
int srcdsp_mixer_set_frequencies(srcdsp_mixer_t h, const float *lo_freq);
/* _Mixer::reset         mixers.h:76-81                                                        */
int srcdsp_mixer_reset(srcdsp_mixer_t h, int ch, float lo_freq);
/* _Mixer::adjustFrequency mixers.h:91-98 (phase continuous)                                   */
int srcdsp_mixer_adjust_frequency(srcdsp_mixer_t h, int ch, float adjust);
/* Mixer::step           mixers.h:168-188.  out may alias in.                                  */
int srcdsp_mixer_step(srcdsp_mixer_t h, const int16_t *in_iq, size_t in_stride, int16_t *out_iq,
                      size_t out_stride, size_t n_per_ch);
/* streaming state (the reference's protected phi / freq / nominalFreq): checkpoint, migration */
int srcdsp_mixer_get_state(srcdsp_mixer_t h, int ch, int *phi, int *freq, float *nominal);
int srcdsp_mixer_set_state(srcdsp_mixer_t h, int ch, int phi, int freq, float nominal);
int srcdsp_mixer_set_stream(srcdsp_mixer_t h, void *cuda_stream);
int srcdsp_mixer_sync(srcdsp_mixer_t h);

/* ------------------------------------------------------------------------------------------ */
/* Decimator bank -- dsptl::FilterDnsamplingFir<cs16, cs16, cs32, int32_t, M>                 */
/* default ctor: dsptl_dnsampling_filters.h:81-83 (no taps yet; step() -> SRCDSP_E_STATE)      */
int srcdsp_dec_create(srcdsp_dec_t *h, int device, int channels, int M);
int srcdsp_dec_destroy(srcdsp_dec_t h);
/* setCoeffs: dsptl_dnsampling_filters.h:114-134 when require_multiple_of_m != 0 (ntaps % M
 * must be 0, else E_SIZE, the reference assert :122); ctor of the obsolete twin
 * dnsampling_filters.h:83-97 when 0 (any ntaps >= 1).  Zeroes the history (history.resize on a
 * fresh object) and leftShift, computes coeffScaling = floor(log2(sum |c|)).
 * Called again on a live object it behaves like the reference's history.resize(ntaps-1): the
 * first min(old, new) history entries are kept, the rest are zero. */
int srcdsp_dec_set_coeffs(srcdsp_dec_t h, const int32_t *taps, int ntaps, int require_multiple_of_m);
/* setLeftShiftBy2: dsptl_dnsampling_filters.h:63 */
int srcdsp_dec_set_left_shift(srcdsp_dec_t h, int left_shift);
/* reset: dsptl_dnsampling_filters.h:55-59 */
int srcdsp_dec_reset(srcdsp_dec_t h);
/* step: dsptl_dnsampling_filters.h:172-220 == dnsampling_filters.h:128-172.
 * n_in_per_ch % M must be 0 (E_SIZE; the reference assert :181).  Writes n_in_per_ch / M
 * samples per channel.  Blocks shorter than ntaps-1 are allowed (the reference reads out of
 * bounds there); the history then is the last ntaps-1 samples of history ++ input. */
int srcdsp_dec_step(srcdsp_dec_t h, const int16_t *in_iq, size_t in_stride, size_t n_in_per_ch,
                    int16_t *out_iq, size_t out_stride);
int srcdsp_dec_get_coeff_scaling(srcdsp_dec_t h, int *coeff_scaling);
/* history of one channel, oldest first, ntaps-1 complex samples (host pointers) */
int srcdsp_dec_get_state(srcdsp_dec_t h, int ch, int16_t *history_iq, size_t *n_samples);
int srcdsp_dec_set_state(srcdsp_dec_t h, int ch, const int16_t *history_iq, size_t n_samples);
int srcdsp_dec_set_stream(srcdsp_dec_t h, void *cuda_stream);
int srcdsp_dec_sync(srcdsp_dec_t h);
/* kernel selection: 0 = automatic, 1 = force the INT32-FMA (IMAD) kernel, 2 = force the
 * tcgen05 int8 Toeplitz kernel (E_STATE at step if the taps do not fit it; TMA-fed when the input
 * rows are 16-byte aligned -- in the band form where that applies --, register-staged loads otherwise),
 * 3 = force its register-staged variant, 4 = force the band form (dec_band_kernel: even M,
 * ntaps <= 32 M + 1), 5 = force the TMA-fed kernel in its original form (dec_tma_kernel). */
int srcdsp_dec_set_kernel(srcdsp_dec_t h, int kind);
/* which FIR kernel the last step launched: 0 none yet, 1 IMAD (dec_fir_kernel), 2 tcgen05 with
 * register-staged loads (dec_tc_kernel), 3 tcgen05 fed by TMA (dec_tma_kernel), 4 its band form (dec_band_kernel) */
int srcdsp_dec_get_last_kernel(srcdsp_dec_t h, int *kind);

/* ------------------------------------------------------------------------------------------ */
/* Float decimator bank -- dsptl::FilterDnsamplingFir<complex<float>, complex<float>,          */
/* complex<float>, float, M>, the only float instantiation of the hot path the reference can   */
/* build (Mixer has no float specialisation, the upsampler static_asserts integers).           */
/* Bit-exact with the compiled reference: the sum runs in tap order with one rounded multiply  */
/* and one rounded add per component (dsptl_dnsampling_filters.h:195-210), is truncated to      */
/* int32, shifted by coeffScaling - leftShift, clamped to +-32767 (limitScale16,               */
/* dsp_complex.cpp:63-73) and converted back to float (:214).  Samples: interleaved float I/Q  */
/* (std::vector<std::complex<float>>::data()), strides and sizes in complex samples; host or   */
/* device pointers.  A sum with |sum| >= 2^31 converts as the reference's x86-64 build does    */
/* (cvttss2si: 0x80000000 for either sign -- tested); samples must be FINITE: the kernel       */
/* multiplies zero padding taps, so an infinity or NaN reaches outputs the reference keeps clean. */
int srcdsp_decf_create(srcdsp_decf_t *h, int device, int channels, int M);
int srcdsp_decf_destroy(srcdsp_decf_t h);
/* setCoeffs (:114-134).  coeffScaling follows the reference literally: its unqualified abs()  */
/* on a float tap is ::abs(int), so coeffScaling = floor(log2(sum |int(c_k)|)); when that sum  */
/* is 0 (all |c_k| < 1) the reference's value is undefined and, as built by g++ for x86-64,    */
/* results in a shift of (0x80000000 - leftShift) & 31 -- reproduced here.                      */
int srcdsp_decf_set_coeffs(srcdsp_decf_t h, const float *taps, int ntaps, int require_multiple_of_m);
int srcdsp_decf_set_left_shift(srcdsp_decf_t h, int left_shift);
int srcdsp_decf_reset(srcdsp_decf_t h);
/* step (:172-220): n_in_per_ch % M must be 0 (E_SIZE); writes n_in_per_ch / M samples per channel */
int srcdsp_decf_step(srcdsp_decf_t h, const float *in_iq, size_t in_stride, size_t n_in_per_ch, float *out_iq,
                     size_t out_stride);
int srcdsp_decf_get_coeff_scaling(srcdsp_decf_t h, unsigned *coeff_scaling);
/* which kernel the current coefficients select: 1 = decf_fir_kernel (an output pair per thread), */
/* 2 = decf_quad_kernel (four outputs per thread: M in {4, 8, 16, 32} and ntaps > 2 M)          */
int srcdsp_decf_get_last_kernel(srcdsp_decf_t h, int *kind);
/* history of one channel, oldest first, ntaps-1 complex samples (host pointers) */
int srcdsp_decf_get_state(srcdsp_decf_t h, int ch, float *history_iq, size_t *n_samples);
int srcdsp_decf_set_state(srcdsp_decf_t h, int ch, const float *history_iq, size_t n_samples);
int srcdsp_decf_set_stream(srcdsp_decf_t h, void *cuda_stream);
int srcdsp_decf_sync(srcdsp_decf_t h);

/* ------------------------------------------------------------------------------------------ */
/* Fused DDC chain: mixer -> dec1 [-> dec2].  One call == the separate steps, bit for bit:     */
/*   mixer.step(in, t0); dec1.step(t0, t1); dec2.step(t1, out);                               */
/* (mixers.h:168-188 feeding dsptl_dnsampling_filters.h:172-220 twice).  The NCO mix is fused  */
/* into the first decimator's load stage.  mixer may be NULL (dec1 -> dec2 only); dec2 may be  */
/* NULL.  The handles stay owned by the caller and keep their streaming state.                */
/* Streams: while the chain exists its members run on dec1's stream (srcdsp_dec_set_stream on   */
/* dec1 and srcdsp_ddc_set_stream reach them); srcdsp_ddc_destroy -- or destroying dec1 --      */
/* hands the mixer and dec2 private streams again.  Destroy the chain BEFORE dec1.             */
/* Aliasing (all filter step functions): DEVICE input and output ranges must not overlap       */
/* (SRCDSP_E_INVALID); HOST buffers are staged and may alias where the reference allows it      */
/* (FilterFir::step in place, filters.h:130-169); srcdsp_mixer_step works in place either way. */
int srcdsp_ddc_create(srcdsp_ddc_t *h, srcdsp_mixer_t mixer, srcdsp_dec_t dec1, srcdsp_dec_t dec2);
int srcdsp_ddc_destroy(srcdsp_ddc_t h);
int srcdsp_ddc_step(srcdsp_ddc_t h, const int16_t *in_iq, size_t in_stride, size_t n_in_per_ch,
                    int16_t *out_iq, size_t out_stride);
int srcdsp_ddc_set_stream(srcdsp_ddc_t h, void *cuda_stream);
int srcdsp_ddc_sync(srcdsp_ddc_t h);

/* ------------------------------------------------------------------------------------------ */
/* Device group: ONE chain (mixer -> dec1 [-> dec2]) spread over several GPUs of one box and    */
/* driven from one process -- the reference has no such object; a user of it runs one           */
/* Mixer / FilterDnsamplingFir set per stream on one thread.  No collective   */
/* and no peer traffic: every device gets its own banks and, during a step, its own host       */
/* thread that stages ITS part of the caller's HOST buffers (pinned for speed:                 */
/* srcdsp_host_alloc) and writes its outputs to their final place in the caller's output.      */
/*   SRCDSP_GROUP_CHANNELS  independent channels: member g owns the contiguous channel batch   */
/*       [C*g/G, C*(g+1)/G) (balanced) together with its streaming state.                      */
/*   SRCDSP_GROUP_SLICES    one (or a few) very long stream(s): every member holds all         */
/*       channels; a block is cut into G time slices at multiples of the total decimation;     */
/*       slice g > 0 is preceded by a warm-up run over the (N1-1) + (N2-1)*M1 samples in front */
/*       of it (rounded up to the total decimation, outputs discarded) starting from zero      */
/*       history and the closed-form NCO phase, so that outputs, history and phase equal the   */
/*       reference's sequential run bit for bit (dsptl_dnsampling_filters.h:198-205,218-219;   */
/*       mixers.h:177).  Blocks too short for that use fewer members.  Between calls the        */
/*       stream's state lives in member 0's banks.                                             */
/* devices[] may name a device more than once (several members on one GPU).  n_table == 0:      */
/* no mixer; M2 == 0: a single decimator.  Calls mirror the chain's: set_coeffs(stage 1 | 2),  */
/* set_frequencies([channels]), reset, step (n_in_per_ch % (M1*M2) == 0, returns when every     */
/* output is in host memory).                                                                  */
#define SRCDSP_GROUP_CHANNELS 0
#define SRCDSP_GROUP_SLICES 1
int srcdsp_group_create(srcdsp_group_t *h, int mode, const int *devices, int n_devices, int channels, unsigned n_table,
                        int M1, int M2);
int srcdsp_group_destroy(srcdsp_group_t h);
int srcdsp_group_set_coeffs(srcdsp_group_t h, int stage, const int32_t *taps, int ntaps, int require_multiple_of_m);
int srcdsp_group_set_frequencies(srcdsp_group_t h, const float *lo_freq);
int srcdsp_group_reset(srcdsp_group_t h);
int srcdsp_group_step(srcdsp_group_t h, const int16_t *in_iq, size_t in_stride, size_t n_in_per_ch, int16_t *out_iq,
                      size_t out_stride);
/* member `index`: its device and channel batch; number of members and how many the last step used */
int srcdsp_group_get_layout(srcdsp_group_t h, int index, int *device, int *ch0, int *n_channels);
int srcdsp_group_size(srcdsp_group_t h, int *n_members, int *used_in_last_step);

/* ------------------------------------------------------------------------------------------ */
/* Upsampler bank -- dsptl::FilterUpsamplingFir<cs16, cs16, cs32, int32_t, L>                 */
/* ctor with empty taps: upsampling_filters.h:89-96 */
int srcdsp_up_create(srcdsp_up_t *h, int device, int channels, int L);
int srcdsp_up_destroy(srcdsp_up_t h);
/* setCoefficients: upsampling_filters.h:107-126 (ntaps >= 1 and ntaps % L == 0, else E_SIZE).
 * Like the reference it does NOT clear the history when the size is unchanged. */
int srcdsp_up_set_coefficients(srcdsp_up_t h, const int32_t *taps, int ntaps);
/* reset: upsampling_filters.h:50-55 */
int srcdsp_up_reset(srcdsp_up_t h);
/* step: shift_mode 0 = vector overload upsampling_filters.h:149-233 (shift 15 - round(log2 L)),
 *       shift_mode 1 = iterator overload :240-323 (shift 0).
 * Writes L * (n_in_per_ch + (flush ? getLength()/L : 0)) samples per channel. */
int srcdsp_up_step(srcdsp_up_t h, const int16_t *in_iq, size_t in_stride, size_t n_in_per_ch,
                   int16_t *out_iq, size_t out_stride, int flush, int shift_mode);
/* getLength / getImpLength / getUpsamplingRatio: upsampling_filters.h:57-67 */
int srcdsp_up_get_length(srcdsp_up_t h, int *length);
int srcdsp_up_get_imp_length(srcdsp_up_t h, int *imp_length);
int srcdsp_up_get_ratio(srcdsp_up_t h, int *ratio);
/* which FIR kernel the last step launched: 0 none yet, 1 up_fir_kernel (any L), 2 up_fir4_kernel (register
 * blocked, L in {4, 8, 16}), 3 up_tc_kernel (tcgen05 int8: one MMA per 4096 outputs; applicable when 32 % L == 0, the
 * filter has at most 16 - 32 / L taps per phase, |taps| < 2^23 and the channel rows are 16-byte aligned; chosen when
 * the filter has more than 8 taps per phase and the batch fills the machine) */
int srcdsp_up_get_last_kernel(srcdsp_up_t h, int *kind);
/* history of one channel in age order, oldest first, ntaps/L - 1 complex samples */
int srcdsp_up_get_state(srcdsp_up_t h, int ch, int16_t *history_iq, size_t *n_samples);
int srcdsp_up_set_state(srcdsp_up_t h, int ch, const int16_t *history_iq, size_t n_samples);
int srcdsp_up_set_stream(srcdsp_up_t h, void *cuda_stream);
int srcdsp_up_sync(srcdsp_up_t h);

/* ------------------------------------------------------------------------------------------ */
/* FifoWithTimeTrack<T, N> -- buffers.h:58-459: single-writer / single-reader ring with sample  */
/* time stamps, here in PINNED host memory so that a block can be DMA'd straight out of the    */
/* ring.  Elements are opaque (elem_bytes each).  Host code only; falls    */
/* back to ordinary memory when there is no CUDA device (bookkeeping, not compute).            */
/* ctor: buffers.h:62-65 */
int srcdsp_fifo_create(srcdsp_fifo_t *h, size_t elem_bytes, size_t capacity, double sampling_frequency);
int srcdsp_fifo_destroy(srcdsp_fifo_t h);
int srcdsp_fifo_is_pinned(srcdsp_fifo_t h);
/* write: buffers.h:139-217 (n < capacity, else E_SIZE -- the reference asserts) */
int srcdsp_fifo_write(srcdsp_fifo_t h, const void *in, size_t n, unsigned seconds, double frac_seconds);
/* read: buffers.h:284-352.  *start may be moved up to the first available time point (the reference
 * warns on stderr); *error = the reference's return value: 1 when [start, start+n) is not available. */
int srcdsp_fifo_read(srcdsp_fifo_t h, void *out, size_t n, uint64_t *start, int *error);
/* the same range without the copy: at most two contiguous pieces of the (pinned) ring */
int srcdsp_fifo_segments(srcdsp_fifo_t h, size_t n, uint64_t *start, const void **p0, size_t *n0, const void **p1,
                         size_t *n1, int *error);
/* count: buffers.h:361-377; reset: :262-276; getAbsoluteTime: :396-459 */
int srcdsp_fifo_count(srcdsp_fifo_t h, size_t *count);
int srcdsp_fifo_reset(srcdsp_fifo_t h);
int srcdsp_fifo_get_absolute_time(srcdsp_fifo_t h, uint64_t time_point, double frac_time_point, unsigned *seconds,
                                  double *frac_seconds);
/* writePtr, timeStart, timeEnd, rolloverFlag: what dumpInfo prints (buffers.h:227-251) */
int srcdsp_fifo_get_state(srcdsp_fifo_t h, size_t *write_ptr, uint64_t *time_start, uint64_t *time_end, int *rollover);
/* test hook: move the time counters (e.g. next to the 64-bit rollover) */
int srcdsp_fifo_set_time(srcdsp_fifo_t h, uint64_t time_start, uint64_t time_end);
const void *srcdsp_fifo_storage(srcdsp_fifo_t h);

/* ------------------------------------------------------------------------------------------ */
/* Correlator bank -- dsptl::FixedPatternCorrelator<int16_t, int32_t, N, S>  correlators.h:54-303 */
/* (the stage behind the DDC): N-point correlation with a fixed pattern over */
/* every S-th sample, 3-point peak test against the signal energy, stops at the first peak.      */
/* ctor: correlators.h:124-133 (N <= 256, N * S <= 8192) */
int srcdsp_corr_create(srcdsp_corr_t *h, int device, int channels, int N, int S);
int srcdsp_corr_destroy(srcdsp_corr_t h);
/* setPattern: correlators.h:167-194 -- pattern_iq = N complex int32 (not conjugated); E_SIZE when the
 * energy exceeds 1073217600 (the reference asserts) */
int srcdsp_corr_set_pattern(srcdsp_corr_t h, const int32_t *pattern_iq, double threshold_coeff);
/* reset: correlators.h:143-155 */
int srcdsp_corr_reset(srcdsp_corr_t h);
/* step: correlators.h:209-303 per channel.  found[ch] = the return value, corr_index[ch] = corrIndex (only
 * written when found).  Like the reference, the samples behind the detection are not consumed and the
 * detection sample's slot is reused by the next call.  Synchronous (found / corr_index are host arrays). */
int srcdsp_corr_step(srcdsp_corr_t h, const int16_t *in_iq, size_t in_stride, size_t n_per_ch, int *found, int *corr_index);
/* getRefBitSamples: correlators.h:311-316 -- N complex int16 (host) */
int srcdsp_corr_get_ref_bit_samples(srcdsp_corr_t h, int ch, int16_t *bits_iq);
/* getStatus: correlators.h:58-82 (energyValue[3], corrValue[3], coeffsEnergy, coeffScaling, thresholdFactor) */
int srcdsp_corr_get_status(srcdsp_corr_t h, int ch, uint32_t *energy_value3, uint32_t *corr_value3, uint32_t *coeffs_energy,
                           int *coeff_scaling, double *threshold_factor);
int srcdsp_corr_set_stream(srcdsp_corr_t h, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SRCDSP_B200_H */
