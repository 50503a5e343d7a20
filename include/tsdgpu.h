/* tsdgpu.h — C ABI of the B200-native filtering hot path (libtsdgpu.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each group of
 * entry points replaces one reference interface of libtsd (paths relative to the reference's
 * core/ directory); INTEGRATION.md shows the C++ adapters that subclass the reference's own
 * FiltreGen<T> / FFTPlan and forward to these calls.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is available from
 *     tsdgpu_last_error() (thread-local).  The reference reports errors by exceptions only
 *     (commun.hpp:152-163); the adapters turn a non-zero status into the same `échec`.
 *   - cf32 samples are interleaved (re, im) floats, exactly std::complex<float> / Veccf storage.
 *   - batches are channel-major: sample i of channel c is at base + c*stride + i (in samples).
 *     One channel == one reference filter object (the reference has no batching).
 *   - mem = TSDGPU_HOST: pointers are host memory, the call copies in and out and returns when
 *     y is valid.  mem = TSDGPU_DEVICE: pointers are device memory on the current device, work
 *     is enqueued on the library stream (tsdgpu_set_stream) and the call returns immediately.
 *   - there is no CPU fallback: every entry fails if no CUDA device is usable.
 */
#ifndef TSDGPU_H
#define TSDGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDGPU_HOST   0
#define TSDGPU_DEVICE 1

/* sample / tap kinds of filtre_rif<Tc,T> (filtre-rt.cc:816-818) */
#define TSDGPU_FIR_F32_F32   0   /* float data,  float taps  : filtre_rif<float,float>   */
#define TSDGPU_FIR_CF32_F32  1   /* cfloat data, float taps  : filtre_rif<float,cfloat>  */
#define TSDGPU_FIR_CF32_CF32 2   /* cfloat data, cfloat taps : filtre_rif<cfloat,cfloat> */

typedef struct tsdgpu_fir_s    *tsdgpu_fir_t;
typedef struct tsdgpu_fft_s    *tsdgpu_fft_t;
typedef struct tsdgpu_ola_s    *tsdgpu_ola_t;
typedef struct tsdgpu_resamp_s *tsdgpu_resamp_t;
typedef struct tsdgpu_poly_s   *tsdgpu_poly_t;

/* ---- runtime ------------------------------------------------------------------------------ */
/* Selects the CUDA device for the calling thread and creates that device's runtime (library stream, copy streams).
 * One runtime per device: a process may drive all the GPUs of a box; every object lives on the device that was current
 * when it was created and its entry points switch the calling thread to it.  Calls that target the same device are
 * serialised by a per-device lock (independent objects may be stepped from different threads, like the reference's —
 * SURVEY 8b); calls on different devices run concurrently. */
int tsdgpu_init(int device);
/* Initialises several devices at once (channel-sharded batches); the calling thread ends up on devices[0]. */
int tsdgpu_init_devices(const int *devices, int n);
/* Switches the calling thread to `device` (initialising it if needed) / returns its current device (-1: none yet). */
int tsdgpu_set_device(int device);
int tsdgpu_current_device(void);
/* Releases every runtime of the process (streams, events, staging buffers, tables) after synchronising the devices.
 * Objects must have been destroyed before.  The library can be initialised again afterwards. */
int tsdgpu_shutdown(void);
/* Final gather of per-device result shards (the only cross-GPU step of the hot path): shard i
 * (bytes[i] bytes at srcs[i] on src_devices[i]) is copied to dst + dst_offsets[i] on dst_device by peer copies over
 * NVLink, each on its source device's library stream — i.e. behind the kernels that produce the shard.  Returns when
 * all shards have arrived. */
int tsdgpu_gather(void *dst, int dst_device, const long long *dst_offsets, const void *const *srcs, const int *src_devices,
                  const long long *bytes, int n);
/* Use an existing cudaStream_t (e.g. torch's current stream) for all subsequent launches;
 * NULL restores the library's own stream. */
int tsdgpu_set_stream(void *cuda_stream);
/* Blocks until everything enqueued by the library on its stream has completed. */
int tsdgpu_synchronize(void);
const char *tsdgpu_last_error(void);
/* Number of kernels launched by this library since the last call with reset != 0. */
long long tsdgpu_launch_count(int reset);
/* Device-side timing of the dominant kernel of each path (fir_tc / fir_direct, fft64k, ola64k stages, resamp_tc / resamp_banded):
 * when enabled, every such launch is bracketed by a CUDA event pair on the launch stream;
 * tsdgpu_timing_read synchronises, returns the summed duration and the number of launches, and
 * clears the list.  Used by bench.py for the roofline figure. */
int tsdgpu_timing_enable(int on);
int tsdgpu_timing_read(double *total_ms, long long *launches);
/* Bookkeeping shared with the reference, computed on the host with the same expressions:
 * prochaine_puissance_de_2 (tsd.cc:287-291). */
int tsdgpu_p2(int i);
/* Cost model of the block filter (fourier.cc:708-735): FLOP per input sample for pattern length M and block
 * length Ne, and the block length 2^k - (M-1) that minimises it (what filtre_rif_fft-style callers use to pick
 * dim_blocs_temporel).  Host arithmetic, same expressions as the reference. */
int tsdgpu_ola_complexite(int M, int Ne, float *C, int *Nf, int *Nz);
int tsdgpu_ola_complexite_optimise(int M, float *C, int *Nf, int *Nz, int *Ne);

/* ---- direct-form FIR: replaces filtre_rif<Tc,T>(h) and FiltreRIF::step ---------------------- */
/* (filtrage.hpp:1367-1368, filtre-rt.cc:53-109,171-175).  State per channel = the last K-1
 * inputs, kept on the device across step() calls; `index` = (samples so far) mod K. */
int tsdgpu_fir_create(int kind, const float *taps, int K, int nchan, tsdgpu_fir_t *out);
/* y[c][i] = sum_k h[k] x[c][i-k] for i in [0,n); len(y) == len(x) (filtre-rt.cc:67-108).
 * x == y (same pointer and stride) is allowed, like the reference (filtre-rt.cc:76-80). */
int tsdgpu_fir_step(tsdgpu_fir_t f, const void *x, long long x_stride, int n,
                    void *y, long long y_stride, int mem);
/* Reference-layout state (filtre-rt.cc:56-58): fen[nchan][K] ring and the common ring index. */
int tsdgpu_fir_get_state(tsdgpu_fir_t f, void *fen_host, int *index);
int tsdgpu_fir_set_state(tsdgpu_fir_t f, const void *fen_host, int index);
/* Same state as a time-ordered history: hist[nchan][K-1] = the last K-1 inputs of every channel, oldest first, and the
 * number of samples the channel has consumed so far.  This is what a halo-split segment of a long stream starts from
 * (the halo is read from the source buffer, nothing is exchanged). */
int tsdgpu_fir_set_history(tsdgpu_fir_t f, const void *hist_host, long long samples_so_far);
int tsdgpu_fir_destroy(tsdgpu_fir_t f);

/* ---- FFT plan: replaces FFTPlan / tfrplan_création / fft() / ifft() ------------------------- */
/* (fourier.hpp:19-32,69,163-205; fourier.cc:360-481).  Always unitary: the reference ignores
 * `normalize` (fourier.cc:119-120,362).  Any n >= 1 like TFRPlanDefaut::configure (fourier.cc:372-405): powers of two
 * directly, even n through two transforms of n/2 (fourier.cc:438-462), odd n through the chirp-z plan of size
 * p2(2n-1) with the reference's float32 chirp (fourier.cc:237-255,392-398). */
int tsdgpu_fft_plan(int n, int batch, tsdgpu_fft_t *out);
/* y[b] = unitary DFT (forward != 0) or inverse DFT of x[b], b in [0,batch). x == y allowed. */
int tsdgpu_fft_exec(tsdgpu_fft_t p, const void *x, long long x_stride,
                    void *y, long long y_stride, int forward, int mem);
int tsdgpu_fft_destroy(tsdgpu_fft_t p);

/* ---- FFT-domain filter: replaces filtre_fft(FiltreFFTConfig) / OLA<cfloat> ------------------ */
/* (fourier.hpp:305-320,370; fourier.cc:737-882,935-940), plain mode, with the spectral
 * callback given as data: traitement_freq = "X *= H" (fourier.cc:956-959).
 *   Ne = dim_blocs_temporel (<= 0 -> 512), N = p2(Ne + nb_zeros_min), N_zeros = N - Ne.
 *   H        : N cfloat gains, or NULL for the identity callback.
 *   fir_len  : 0 -> arbitrary H, true overlap-add (the partial sums `svg` are carried).
 *              K > 0 -> caller guarantees H = fft(h2)*sqrt(N) with h2 = [0^(N-K), h] (the
 *              FiltreFFTRIF convention, fourier.cc:962-965) and K <= N_zeros + 1; the same
 *              samples are then produced in overlap-save form (single pass, plain stores).
 * Fails with the reference's own precondition when N_zeros > Ne (fourier.cc:870). */
int tsdgpu_ola_create(int dim_blocs_temporel, int nb_zeros_min, const float *H, int fir_len,
                      int nchan, tsdgpu_ola_t *out);
/* Same object with FiltreFFTConfig::avec_fenetrage = oui (fourier.cc:794-798,884-930): every block is processed
 * twice — the window [previous half, new half] and the block itself — each multiplied by `fenetre` (Ne floats; the
 * reference uses fenêtre("hn", Ne, non)), transformed, multiplied by H (NULL = identity), transformed back and
 * recombined with weights 1/2.  The first block of the stream emits nothing (cnt_ech < 0, :900-903); later blocks
 * emit Ne samples.  Reproduces the reference as it behaves, including the aliasing of its `svg` buffer with the
 * inverse-transform output from the second block on (:923).  Even Ne only. */
int tsdgpu_ola_create_fen(int dim_blocs_temporel, int nb_zeros_min, const float *H, const float *fenetre,
                          int nchan, tsdgpu_ola_t *out);
/* Generic spectral callback: FiltreFFTConfig::traitement_freq is an arbitrary host std::function<void(Veccf&)>
 * (fourier.hpp:319) that the reference calls once per transformed block (fourier.cc:863; twice per block in the windowed
 * mode, :895,915).  cb(user, chan, X, N) receives the N unitary-scaled bins of one block of channel `chan` in host memory
 * (interleaved cfloat) and may modify them in place; calls arrive in the reference's order for every channel.  The path
 * is FFT -> D2H -> callback -> H2D -> IFFT -> overlap-add, so it is bounded by the host link, not by the GPU; callers
 * whose callback is "X *= H" should hand H over as data (tsdgpu_ola_create).  fenetre: NULL = plain mode, else the Ne
 * window values of the windowed mode (see tsdgpu_ola_create_fen). */
typedef void (*tsdgpu_spectral_cb)(void *user, int chan, float *X, int N);
int tsdgpu_ola_create_cb(int dim_blocs_temporel, int nb_zeros_min, tsdgpu_spectral_cb cb, void *user, const float *fenetre,
                         int nchan, tsdgpu_ola_t *out);
int tsdgpu_ola_dims(tsdgpu_ola_t f, int *Ne, int *N, int *N_zeros, int *residual);
/* Number of samples the next step(n) will emit: Ne * ((residual + n) / Ne), one block less for the
 * first block of a windowed filter (TamponNv2 re-blocking, tsd.cc:332-370; fourier.cc:813-833). */
long long tsdgpu_ola_out_count(tsdgpu_ola_t f, int n);
/* Feeds n samples per channel; writes *n_out = tsdgpu_ola_out_count(f, n) samples per channel. */
int tsdgpu_ola_step(tsdgpu_ola_t f, const void *x, long long x_stride, int n,
                    void *y, long long y_stride, long long *n_out, int mem);
/* State hand-over (checkpoint / halo split of a long stream, SURVEY 5 + 8e): the re-blocking residual (TamponNv2 windex,
 * tsd.cc:310), the number of Ne-blocks emitted so far, carry[nchan][carry_len] = the last carry_len input samples of every
 * channel (oldest first; zeros before the stream start), and — only for the forms that have them — svg[nchan][Ne]
 * (overlap-add partial sums, fourier.cc:870-872) and last[nchan][Ne] (windowed mode, fourier.cc:905-922).  A segment
 * that starts at stream sample S is set up with residual = S mod Ne, blocks_done = S / Ne and the carry_len samples
 * before S; its step() then emits exactly the samples the one-shot call emits for those blocks. */
int tsdgpu_ola_state_dims(tsdgpu_ola_t f, int *carry_len, int *svg_len, int *last_len);
int tsdgpu_ola_get_state(tsdgpu_ola_t f, int *residual, long long *blocks_done, void *carry_host, void *svg_host, void *last_host);
int tsdgpu_ola_set_state(tsdgpu_ola_t f, int residual, long long blocks_done, const void *carry_host, const void *svg_host,
                         const void *last_host);
int tsdgpu_ola_destroy(tsdgpu_ola_t f);

/* periodogramme_tfd(x, N) (fourier.hpp:967, fourier.cc:1451-1481): short-time spectra of the frames of the windowed
 * filter_fft object (dim_blocs_temporel = N, nb_zeros_min = 0, avec_fenetrage, callback recording
 * 10 log10(|X|^2 + 1e-20) of the first N2/2 bins, N2 = p2(N)).  Every full block of N samples gives two frames: the
 * window [previous half, new half] and the block itself, both times `fenetre` (N floats; the reference uses
 * fenêtre("hn", N, non)).  out: per channel [2 * (n / N)][N2 / 2] floats (dB), frames in time order, channels
 * out_stride floats apart.  Even N only (see tsdgpu_ola_create_fen).  *n_frames, *n_bins receive the matrix shape. */
int tsdgpu_periodogramme_tfd(const void *x, long long x_stride, int n, int nchan, int N, const float *fenetre,
                             float *out, long long out_stride, int *n_frames, int *n_bins, int mem);

/* ---- rt_spectrum(SpectrumConfig): averaged power spectrum in dB ------------------------------------------------------ */
/* (fourier.hpp:909-952; Spectrum, src/fourier/fourier.cc:1162-1343).  A block of BS samples per channel is cut into nsubs
 * sub-blocks of Nf = BS / nsubs samples; each is multiplied by `fenetre` (Nf floats, ALREADY normalised to energy Nf as
 * fourier.cc:1203 does: sqrt(Nf / sum f^2) f), transformed by the unitary plan, |X|^2 is fft-shifted and accumulated — in
 * place, or at offset i * sweep_step and times the edge (masque_hf) / centre (masque_bf) mask when sweep_active.  The block
 * that completes nmeans blocks returns Ns values per channel, 10 log10(sum / (nmeans nsubs Nf) [/ coverage] + FLT_MIN), and
 * clears the sums (*n_out = Ns); the other blocks return nothing (*n_out = 0; the reference resizes y to 0).
 * Ns = Nf, or Nf + (nsubs - 1) sweep_step in sweep mode.  step() refuses n != BS like the reference's assertion. */
typedef struct tsdgpu_spectrum_s *tsdgpu_spectrum_t;
int tsdgpu_spectrum_create(int BS, int nmeans, int nsubs, int sweep_active, int sweep_step, int masque_bf, int masque_hf,
                           const float *fenetre, int nchan, tsdgpu_spectrum_t *out);
int tsdgpu_spectrum_dims(tsdgpu_spectrum_t s, int *Nf, int *Ns);
int tsdgpu_spectrum_step(tsdgpu_spectrum_t s, const void *x, long long x_stride, int n, float *y, long long y_stride, int *n_out, int mem);
int tsdgpu_spectrum_destroy(tsdgpu_spectrum_t s);

/* ---- normalised-correlation detector: the hot part of détecteur_création(config) / Detecteur::step -------------------- */
/* (fourier.hpp:577-679; src/fourier/detection.cc:120-260, MODE_OLA).  The reference correlates the stream with a fixed
 * motif through its block filter (traitement_freq = "X *= conj(fft(motif))", nb_zeros_min = M - 1, :146-170), tracks the
 * energy of the same M samples with a moving average and a delay line (:132,165,209-217) and normalises
 * (:231,246): score[i] = sqrt(N/M) |corr[i]| / sqrt(en[i] + 1e-20) in [0, 1], corr and score delayed by Ne samples.  Those
 * three steps run on the device (single-SM overlap-save correlator for motifs up to 8193 samples, N-point path beyond);
 * the peak logic that follows (:262-500: erosion over M samples, quadratic interpolation, gain / phase / SNR of every
 * detection) is sparse serial host work on `score` and `corr` and stays with the caller (libtsd_b200/detection.py).
 *   motif : M cfloat; Ne <= 0 -> ola_complexite_optimise(M) like the reference; n must be a multiple of Ne.
 *   corr  : the correlation signal (cfloat), with the reference's clean-up of |corr| <= 1e-6 applied in place. */
typedef struct tsdgpu_detect_s *tsdgpu_detect_t;
int tsdgpu_detect_create(const float *motif, int M, int Ne, int nchan, tsdgpu_detect_t *out);
int tsdgpu_detect_dims(tsdgpu_detect_t d, int *Ne, int *N, int *M, int *delais_corr, float *norme_motif);
int tsdgpu_detect_step(tsdgpu_detect_t d, const void *x, long long x_stride, int n, float *score, long long score_stride,
                       void *corr, long long corr_stride, int mem);
int tsdgpu_detect_destroy(tsdgpu_detect_t d);

/* ---- arbitrary-ratio resampler: replaces filtre_itrp<cfloat>(ratio, itrp) ------------------- */
/* (filtrage.hpp:2039; ra.cc:13-79; InterpolateurRIF::step filtrage.hpp:1873-1881; LUT of
 * InterpolateurSinc itrp.cc:16-54).  lut[p*K + i], p in [0,nphases], is the interpolator's
 * coefficient table handed over as data.  The float32 phase recurrence (ra.cc:64-73) is run
 * on the host, once for all channels. */
int tsdgpu_resamp_create(float ratio, const float *lut, int K, int nphases, int nchan,
                         tsdgpu_resamp_t *out);
/* Same with the sample type chosen: data_complex = 0 is filtre_itrp<float> / filtre_reechan<float> (ra.cc:190-195, the
 * instantiation the reference's own acceptance test drives, tests/test-ra.cc:148-165), 1 is cfloat. */
int tsdgpu_resamp_create_ex(float ratio, const float *lut, int K, int nphases, int data_complex, int nchan,
                            tsdgpu_resamp_t *out);
/* Interpolators that evaluate their coefficients at the EXACT fractional delay instead of reading a LUT column:
 * itrp_lineaire (itrp.cc:82-95, K = 2, {1 - tau, tau}) and itrp_lagrange(d) (itrp.cc:97-127, K = d + 1).  The host
 * schedule ships the float32 phase of every output; the device evaluates the coefficients with the reference's float
 * operations. */
#define TSDGPU_ITRP_LINEAIRE 1
#define TSDGPU_ITRP_LAGRANGE 2
int tsdgpu_resamp_create_exact(float ratio, int kind, int degree, int data_complex, int nchan, tsdgpu_resamp_t *out);
/* Output count of the next step(n) (identical for every channel) and the current phase. */
long long tsdgpu_resamp_out_count(tsdgpu_resamp_t f, int n);
float tsdgpu_resamp_phase(tsdgpu_resamp_t f);
int tsdgpu_resamp_step(tsdgpu_resamp_t f, const void *x, long long x_stride, int n,
                       void *y, long long y_stride, long long y_capacity, long long *n_out, int mem);
/* State hand-over: the float32 phase (ra.cc:16-22) and hist[nchan][K-1] = the last K-1 inputs, oldest first.  The phase
 * at any input index of a stream comes from tsdgpu_resamp_schedule (data-independent), which is how a halo-split segment
 * gets its start state. */
int tsdgpu_resamp_get_state(tsdgpu_resamp_t f, float *phase, void *hist_host);
int tsdgpu_resamp_set_state(tsdgpu_resamp_t f, float phase, const void *hist_host);
int tsdgpu_resamp_destroy(tsdgpu_resamp_t f);
/* The host-side schedule on its own (no device needed): runs the reference's float32 phase
 * recurrence (ra.cc:58-73) over n inputs starting from *phase, writes for every output j the index
 * of the newest input of its window (in_idx[j]) and its LUT column (lut_idx[j] = (int)(phase*nphases)),
 * updates *phase.  in_idx / lut_idx may be NULL to count only. */
int tsdgpu_resamp_schedule(float *phase, float ratio, int nphases, int n, int32_t *in_idx,
                           int32_t *lut_idx, long long capacity, long long *n_out);

/* ---- polyphase rate-change stages: replace filtre_rif_ups / filtre_rif_demi_bande / filtre_rif_decim ---- */
/* (filtrage.hpp polyphase factories; polyphase.cc:54-149 half-band decimator, :156-239 FIR + decimation by R,
 * :246-341 xR polyphase interpolator, factories :344-360) — the stages filtre_reechan puts in front of the LUT
 * interpolator when the ratio leaves [0.5, 2) (ra.cc:124-141).  Real coefficients; data float (data_complex = 0)
 * or cfloat.  State per channel: the reference's delay line (K samples, K'/R for the interpolator) kept on the
 * device, plus the decimation counter (`odd` / `cnt`), which decides the number of outputs of a call. */
#define TSDGPU_POLY_UPS        0   /* filtre_rif_ups<float,T>(c, R): n*R outputs, coefficients scaled by R, zero-padded */
#define TSDGPU_POLY_DEMI_BANDE 1   /* filtre_rif_demi_bande<float,T>(c): R = 2, even coefficients + literal 0.5 centre */
#define TSDGPU_POLY_DECIM      2   /* filtre_rif_decim<float,T>(c, R): (n + cnt) / R outputs */
int tsdgpu_poly_create(int kind, const float *coefs, int K, int R, int data_complex, int nchan,
                       tsdgpu_poly_t *out);
long long tsdgpu_poly_out_count(tsdgpu_poly_t f, int n);
/* ring index ((samples so far) mod delay-line length) and decimation counter of the reference object */
int tsdgpu_poly_state(tsdgpu_poly_t f, int *index, int *cnt);
int tsdgpu_poly_step(tsdgpu_poly_t f, const void *x, long long x_stride, int n,
                     void *y, long long y_stride, long long *n_out, int mem);
/* State hand-over: samples consumed so far (ring index = total mod L), the decimation counter, and hist[nchan][L-1] = the
 * last L-1 inputs, oldest first. */
int tsdgpu_poly_get_state(tsdgpu_poly_t f, long long *total, int *cnt, void *hist_host);
int tsdgpu_poly_set_state(tsdgpu_poly_t f, long long total, int cnt, const void *hist_host);
int tsdgpu_poly_destroy(tsdgpu_poly_t f);

#ifdef __cplusplus
}
#endif
#endif /* TSDGPU_H */
