/* impb.h — C ABI of libimpulcifer_b200.so (B200 / sm_100a hot path of 115dkk/Impulcifer-pip313).
 *
 * The reference is pure Python: its "plugin interface" for this path is the Python class API
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds with ctypes/cffi from those classes;
 * INTEGRATION.md shows the stub.  Each entry point names the reference code it replaces (file:line,
 * relative to the reference root).
 *
 * Conventions
 *   - every function returns int32 status: 0 = IMPB_OK, negative = error; text via impb_last_error()
 *   - all data pointers are DEVICE pointers owned by the caller unless the name ends in _host
 *   - every launch takes the cudaStream_t (as void*) it must be ordered on; nothing synchronises
 *   - arrays are row-major [batch, ld]; int64 index arrays live on the device
 *   - a context is bound to one device; calls on different streams may come from different host threads
 *   - no CPU fallback exists anywhere: if the device or the kernel image is missing, calls fail
 */
#ifndef IMPB_H
#define IMPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define IMPB_API __attribute__((visibility("default")))
#else
#define IMPB_API
#endif

#define IMPB_OK 0
#define IMPB_EINVAL (-1)    /* bad argument */
#define IMPB_ECUDA (-2)     /* CUDA runtime error (text in impb_last_error) */
#define IMPB_ENOMEM (-3)
#define IMPB_EUNSUPPORTED (-4) /* size that the planner cannot factor */

/* sample formats of impb_deconv_execute input */
#define IMPB_FMT_F64 0
#define IMPB_FMT_F32 1
#define IMPB_FMT_I32PCM 2   /* int32 PCM, scaled by 2^-31 on load (libsndfile PCM_32 -> float64) */

typedef struct impb_ctx impb_ctx;
typedef struct impb_plan impb_plan;
typedef struct impb_fft impb_fft;

/* Optional per-row "view" for the crop/align kernels: the array the reference would hold after crop_heads'
 * fade-in and a chain of ImpulseResponse.shift calls, as index arithmetic over the untouched samples
 *   v[j] = fade_in[j - shift] * x[offset + j - shift]   for valid_lo <= j < valid_hi,   0 elsewhere
 * (fade_in applies where j - shift < fade_in_len).  Every pointer may be NULL (no shift / all valid / no fade).
 * hint_max / hint_argmax (both or neither): max |x| over a range of the untouched samples that contains the whole
 * view, and one index (relative to `offset`) where it is attained, e.g. absmax_out / argmax_out of
 * impb_deconv_execute minus the head crop.  A row whose hinted sample survives inside the view outside the faded
 * head takes hint_max as its max |v| without a scan; every other row is scanned as usual, so results do not
 * depend on the hint being usable. */
typedef struct impb_view {
  const int64_t* shift;
  const int64_t* valid_lo;
  const int64_t* valid_hi;
  const double* fade_in;
  int64_t fade_in_len;
  const double* hint_max;
  const int64_t* hint_argmax;
  /* block maxima of the untouched samples (rows without shift / fade only): blk_max[r * nblk + p] = max |x| over
   * x[offset + blk_first + p * blk_len : ... + blk_len) (clipped to the row), as written by impb_deconv_execute's
   * blockmax_out.  impb_peak_index then starts its search at the first block that can contain a peak. */
  const double* blk_max;
  int64_t nblk, blk_len, blk_first;
} impb_view;

IMPB_API uint32_t impb_version(void);
IMPB_API int32_t impb_ctx_create(int32_t device, impb_ctx** out);
IMPB_API void impb_ctx_destroy(impb_ctx* ctx);
IMPB_API const char* impb_last_error(void); /* thread-local */

/* ---- K1: sweep deconvolution ---------------------------------------------------------------------
 * Replaces ImpulseResponseEstimator.estimate(), core/impulse_response_estimator.py:149-151
 * (scipy.signal.convolve(recording, inverse_filter, mode='same', method='auto')), and the serial
 * per-column loop that calls it in core/hrir.py:307-355.
 *
 * A plan caches the spectrum of the inverse filter for recordings of up to L samples.  precision is
 * 64 (FP64 arithmetic) or 32 (FP32 fast path, FP64-computed spectrum rounded to FP32).
 * execute(): rec is [batch, ld_in] in in_fmt, every row holds `len` valid samples (len <= L); with
 * in_offsets != NULL (device int64[batch]) row r instead starts at rec + in_offsets[r] elements, which is
 * how the track/column split of core/hrir.py:194-216 is consumed without a gather copy.  Row r of
 * out ([batch, ld_out] f64 for precision 64, f32 for precision 32 unless out_f64 != 0) receives
 *   convolve(rec[r, :len], inverse_filter, 'same')[out_offset : out_offset + out_len].
 * absmax_out (nullable, device double[batch]) receives max |out[r, :]|, tracked by the final column pass for
 * free; impb_peak_index accepts it as known_max and then skips its own max pass.  argmax_out (nullable, device
 * int64[batch], needs absmax_out) receives one index j with |out[r, j]| == absmax_out[r] (the smallest one found
 * among the per-tile candidates), for impb_view.hint_argmax.  blockmax_out (nullable, device double[batch, nblk]):
 * max |out[r, :]| per contiguous time block, geometry from impb_deconv_blockmax_info (nblk = 0: the plan cannot
 * provide them, pass NULL), for impb_view.blk_max.
 */
IMPB_API int32_t impb_deconv_plan_create(impb_ctx* ctx, const double* inv_filter_host, int64_t M, int64_t L,
                                int32_t precision, impb_plan** out);
IMPB_API void impb_deconv_plan_destroy(impb_plan* plan);
IMPB_API int64_t impb_deconv_workspace_bytes(const impb_plan* plan, int64_t batch);
IMPB_API int32_t impb_deconv_plan_info(const impb_plan* plan, int64_t* nfft, int64_t* n1, int64_t* n2);
IMPB_API int32_t impb_deconv_execute(impb_plan* plan, const void* rec, const int64_t* in_offsets, int64_t ld_in,
                            int32_t in_fmt, int64_t len, void* out, int64_t ld_out, int32_t out_f64,
                            int64_t out_offset, int64_t out_len, double* absmax_out, int64_t* argmax_out,
                            double* blockmax_out, int64_t batch, void* workspace, void* stream);
/* block p of blockmax_out covers output samples [blk_first + p * blk_len, blk_first + (p + 1) * blk_len) */
IMPB_API int32_t impb_deconv_blockmax_info(const impb_plan* plan, int64_t out_offset, int64_t* nblk, int64_t* blk_len,
                                  int64_t* blk_first);
/* name and launch count of the kernels one execute() issues (bench.py gpu_launches accounting) */
IMPB_API int32_t impb_deconv_launches_per_execute(const impb_plan* plan, int64_t batch);

/* ---- K2: first-peak search -----------------------------------------------------------------------
 * Replaces ImpulseResponse.peak_index(), core/impulse_response.py:32-70 and its twin
 * core/decay.py:12-41: m = max|x|; if m < 1e-20 -> start; first strict local maximum (plateau midpoint,
 * scipy _local_maxima_1d) of x/m or -x/m that is >= height; else argmax|x|.  Segment r is
 * x[offsets[r] : offsets[r] + lengths[r]]; peak_out[r] is relative to the segment start; max_len = the
 * longest segment (grid sizing only); height must be > 0; known_max (nullable, device double[batch]) = max|segment|
 * if the caller already has it (it must be exact: it is the normalisation divisor).  Bit-exact. */
IMPB_API int32_t impb_peak_index(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* lengths,
                        const impb_view* view, const double* known_max, double height, int64_t* peak_out,
                        int64_t batch, int64_t max_len, void* stream);

/* ---- K4: Lundeby decay parameters ----------------------------------------------------------------
 * Replaces decay.decay_params(), core/decay.py:44-260 (called by HRIR.crop_tails, core/hrir.py:627-636).
 * peak_out / knee_out are relative to the segment start. */
IMPB_API int32_t impb_lundeby(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* lengths,
                     const impb_view* view, int32_t fs, int64_t* peak_out, int64_t* knee_out, double* noise_db_out,
                     int64_t* win_out, int64_t batch, int64_t max_len, void* stream);

/* ---- K17: decay times (next row; feeds decay_adjustment_params of row a11) -------
 * Replaces decay.decay_times(), core/decay.py:263-352: Schroeder backward integral from the peak to the
 * knee, vertical offset against the moving average of the squared response, and the EDT / RT20 / RT30 /
 * RT60 regression lines.  peak / knee / noise_db / win are what impb_lundeby returned for the same rows
 * (knee absolute, like decay_params' second value).  out: [batch, 4] seconds, NaN where the reference
 * returns None.  workspace: impb_decay_times_workspace_bytes(batch, max_len) bytes of device memory. */
IMPB_API int64_t impb_decay_times_workspace_bytes(int64_t batch, int64_t max_len);
IMPB_API int32_t impb_decay_times(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* lengths,
                         const int64_t* peak, const int64_t* knee, const double* noise_db, const int64_t* win,
                         int32_t fs, double* out, void* workspace, int64_t batch, int64_t max_len, void* stream);

/* ---- K18: mean squares of index ranges -----------------------------------------------------------------
 * The arithmetic of HRIR.calculate_reflection_levels(), core/hrir.py:1003-1090: for row r and segment q,
 * out[r * nseg + q] = mean(x[offsets[r] + seg_lo : offsets[r] + seg_hi] ** 2), NaN for an empty range. */
IMPB_API int32_t impb_segment_energy(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* seg_lo,
                            const int64_t* seg_hi, int32_t nseg, double* out, int64_t batch, void* stream);

/* ---- K19 / K20: virtual-bass helpers (next row) --------------------------------------
 * impb_dft_bin_abs replaces _mag_at(), core/virtual_bass.py:47-58: out[r] = |sum_j x[r, j] exp(-2 pi i k j / n)|.
 * impb_add_delayed replaces _delay_signal() (:29-44) and the final add (:173-174):
 *   dst[r, j] += src[which[r], j - delay[r]]   for 0 <= j - delay[r] < src_len   (src: [n_src, src_ld]). */
IMPB_API int32_t impb_dft_bin_abs(impb_ctx* ctx, const double* x, int64_t ld, int64_t n, int64_t k, double* out,
                         int64_t batch, void* stream);
IMPB_API int32_t impb_add_delayed(impb_ctx* ctx, double* dst, int64_t ld, int64_t n, const double* src, int64_t src_ld,
                         int64_t src_len, int32_t n_src, const int32_t* which, const int64_t* delay, int64_t batch,
                         void* stream);

/* ---- K6: onset cross-correlation lag ---------------------------------------------------------------
 * Replaces the scipy.signal.correlate(..., 'full') + argmax of HRIR.align_ipsilateral_all,
 * core/hrir.py:932-948.  a, b: [batch] segments of x given by offsets/lengths (<= max_seg_len samples);
 * lag_out[r] = np.arange(-len(a)+1, len(a))[np.argmax(correlate(a, b, 'full'))], first maximum on ties.
 * fade_in (nullable): crop_heads' rising window, applied to the first fade_in_len samples of both segments. */
IMPB_API int32_t impb_xcorr_lag(impb_ctx* ctx, const double* x, const int64_t* a_offsets, const int64_t* a_lengths,
                       const int64_t* b_offsets, const int64_t* b_lengths, const double* fade_in,
                       int64_t fade_in_len, int64_t* lag_out, int64_t batch, int64_t max_seg_len, void* stream);

/* ---- K3+K5+K7+K16: fused gather / window --------------------------------------------------------------
 * Replaces the slicing, ImpulseResponse.shift (core/impulse_response.py:92-108), Hann fade-in of
 * HRIR.crop_heads (core/hrir.py:591-612), the tail crop + Hann fade-out of HRIR.crop_tails
 * (core/hrir.py:642-651) and the gain multiply of HRIR.normalize (core/hrir.py:527-544).
 *   dst[r, j] = gain * fade_out(j) * S(j),  j in [0, out_len)
 *   S(j)      = fade_in(j - shift[r]) * src[src_start[r] + j - shift[r]]   if valid_lo[r] <= j < valid_hi[r]
 *             = 0                                                            otherwise
 * fade_in: [fade_in_len] rising half window (hann(2*head)[:head], built by the host exactly as the
 * reference does) applied to head-relative index m = j - shift; fade_out: [fade_out_len] falling half
 * window laid over the last fade_out_len output samples.  shift / valid_lo / valid_hi may be NULL
 * (no shift, everything valid).  Each factor is applied as a separately rounded multiply, in the
 * reference's order, so results are bit-identical to numpy for identical src. */
IMPB_API int32_t impb_gather_window(impb_ctx* ctx, const double* src, const int64_t* src_start, const int64_t* shift,
                           const int64_t* valid_lo, const int64_t* valid_hi, int64_t out_len,
                           const double* fade_in, int64_t fade_in_len, const double* fade_out,
                           int64_t fade_out_len, double gain, double* dst, int64_t ld_dst, int64_t batch,
                           void* stream);

/* Ragged form of impb_gather_window: row r has its own length out_lens[r] (fade_out laid over ITS last fade_out_len
 * samples; a row shorter than the fade gets none, the caller rejects such rows as the reference does) and is written
 * at dst + dst_offsets[r]; max_out_len sizes the grid.  One launch for recording sets with different tail lengths
 * (HRIR.crop_tails per set, core/hrir.py:614-653). */
IMPB_API int32_t impb_gather_rows(impb_ctx* ctx, const double* src, const int64_t* src_start, const int64_t* shift,
                         const int64_t* valid_lo, const int64_t* valid_hi, const int64_t* out_lens,
                         const int64_t* dst_offsets, int64_t max_out_len, const double* fade_in, int64_t fade_in_len,
                         const double* fade_out, int64_t fade_out_len, double gain, double* dst, int64_t batch,
                         void* stream);

/* ---- generic batched complex FFT plan (used by K8 / K12 / K13 and the host-side shim) -------------- */
IMPB_API int32_t impb_fft_create(impb_ctx* ctx, int64_t n, int32_t precision, impb_fft** out);
IMPB_API void impb_fft_destroy(impb_fft* fft);

/* ---- K8: magnitude response -------------------------------------------------------------------------
 * Replaces audio_io.magnitude_response(), core/audio_io.py:100-113: 20*log10(|rfft(x)|) for the first
 * ceil(n/2) bins.  x: [batch, ld_x] rows of n samples; mag_out: [batch, ld_out]. */
IMPB_API int32_t impb_magnitude_db(impb_fft* fft, const double* x, int64_t ld_x, double* mag_out, int64_t ld_out,
                          int64_t batch, void* workspace, void* stream);
IMPB_API int64_t impb_fft_workspace_bytes(const impb_fft* fft, int64_t batch);
/* The same bins after the decimation of ImpulseResponse.frequency_response(), core/impulse_response.py:170-178
 * (`f[1::step]`, `m[1::step]`): mag_out[b][j] = 20*log10(|rfft(x_b)[first + step*j]|) for first + step*j < ceil(n/2).
 * Two rows ride one complex transform and only the kept bins are read back; workspace as for impb_magnitude_db.
 * row_offsets (nullable, device, [batch]): start of every row in elements from x, for rows that are slices of one
 * buffer (the cropped responses of many recording sets); ld_x is ignored then.
 * IMPB_EUNSUPPORTED for chirp-z plans (lengths with large prime factors): use impb_magnitude_db and slice. */
IMPB_API int32_t impb_magnitude_db_decimated(impb_fft* fft, const double* x, int64_t ld_x, const int64_t* row_offsets,
                                    int64_t first, int64_t step, double* mag_out, int64_t ld_out, int64_t batch,
                                    void* workspace, void* stream);
/* The same kept bins when `step` divides the row length: bins first + step*j of an (M*step)-point DFT are the M-point DFT
 * of the folded, modulated row y[m] = W^(first*m) * sum_r x[m + M*r] * exp(-2*pi*i*first*r/step), so the transform is
 * `step` times smaller.  `fft` is a plan of length M = n / step (impb_fft_create(ctx, M, ...)); rows hold M*step samples;
 * first < step <= 64; workspace: impb_fft_workspace_bytes(fft, batch).  IMPB_EUNSUPPORTED if M needs a chirp-z plan or
 * sub-FFT lengths outside the specialised menu (the paired transform of the full row is faster then). */
IMPB_API int32_t impb_magnitude_db_folded(impb_fft* fft, const double* x, int64_t ld_x, const int64_t* row_offsets,
                                 int64_t first, int64_t step, double* mag_out, int64_t ld_out, int64_t batch,
                                 void* workspace, void* stream);
/* Linear twin used by the interaural microphone-mismatch estimator (next row): replaces
 * np.abs(scipy.fft.rfft(seg, n=nfft)) in MicrophoneMatchingCorrector._windowed_power,
 * core/microphone_deviation_correction.py:130-133.  Rows hold n_valid <= n samples (zero-padded to the
 * plan length n); bins 0 .. n_bins-1 (n_bins <= n/2+1) are written. */
IMPB_API int32_t impb_magnitude_lin(impb_fft* fft, const double* x, int64_t ld_x, int64_t n_valid, double* mag_out,
                           int64_t ld_out, int64_t n_bins, int64_t batch, void* workspace, void* stream);

/* ---- K10: log-frequency linear interpolation ----------------------------------------------------------
 * Replaces InterpolatedUnivariateSpline(log10 f, y, k=1)(log10 f_out) of FrequencyResponse.interpolate,
 * autoeq/frequency_response.py:874-894 (shared f_in/f_out grids, y: [batch, n_in]). */
IMPB_API int32_t impb_loglerp(impb_ctx* ctx, const double* logf_in, const double* y_in, int64_t n_in, int64_t ld_in,
                     const double* logf_out, int64_t n_out, double* y_out, int64_t ld_out, int64_t batch,
                     void* stream);

/* The gain grid of FrequencyResponse.minimum_phase_impulse_response(), autoeq/frequency_response.py:653-676, for rows of
 * equalization curves (dB on the log10-frequency grid logf_in): k=1 interpolation onto logf_out (log10 of
 * linspace(0, fs//2, n_out) with 0 Hz moved to 0.001), the value at logf_min for the n_below leading points (f <= f_min),
 * [normalize: - max - 0.5 dB], * 2, 10^(./20), last point 0.  gain_out: [batch, ld_out] linear gains, the input of
 * impb_minphase_fir.  row_scratch: batch * 8 bytes (only read when normalize != 0). */
IMPB_API int32_t impb_minphase_gain_grid(impb_ctx* ctx, const double* logf_in, const double* eq_db, int64_t n_in, int64_t ld_in,
                                const double* logf_out, int64_t n_out, int64_t n_below, double logf_min, int32_t normalize,
                                double* gain_out, int64_t ld_out, int64_t batch, void* row_scratch, void* stream);

/* ---- K11: fractional-octave smoothing -------------------------------------------------------------------
 * Replaces FrequencyResponse._smoothen_fractional_octave, autoeq/frequency_response.py:1060-1105:
 * `it_normal` passes of savgol_filter(y, w_normal, 2) and `it_treble` passes of savgol_filter(y, w_treble, 2)
 * (mode='interp' edge fit), blended by k_treble[n] (the log-frequency sigmoid, computed by the host shim).
 * coef_*: [w] convolution weights; edge_*: [w/2, w] edge-fit weights (host: savgol_coeffs / polyfit). */
IMPB_API int32_t impb_savgol_blend(impb_ctx* ctx, const double* y, int64_t n, int64_t ld, int32_t w_normal,
                          int32_t it_normal, const double* coef_normal, const double* edge_normal,
                          int32_t w_treble, int32_t it_treble, const double* coef_treble,
                          const double* edge_treble, const double* k_treble, double* out, int64_t ld_out,
                          int64_t batch, void* stream);

/* ---- K21: equalization curve ---------------------------------------------------------------------------------
 * Replaces FrequencyResponse.equalize, autoeq/frequency_response.py:1241-1310, for rows of error curves on one
 * frequency grid: gain = -error * gain_k clipped to cap (both [n], the host's log-frequency sigmoids), kink windows
 * of +-half_window points dropped (not the last two points), k = 2 interpolating spline over log10 f re-evaluated on
 * the grid (smoothen = 0: clipped gain only).  status (nullable, device int32[batch]): 1 where fewer than 3 points
 * survive (the reference raises from dfitpack there).  n <= 2048. */
IMPB_API int32_t impb_equalize(impb_ctx* ctx, const double* error, int64_t ld_err, const double* logf, const double* cap,
                      const double* gain_k, int64_t n, int32_t half_window, int32_t smoothen, double* out,
                      int64_t ld_out, int32_t* status, int64_t batch, void* stream);

/* ---- K12+K13: minimum-phase FIR ---------------------------------------------------------------------------
 * Replaces scipy firwin2 + minimum_phase(method='homomorphic', half=True) as called by
 * FrequencyResponse.minimum_phase_impulse_response, autoeq/frequency_response.py:672-680.
 * gain: [batch, n] linear amplitude on linspace(0, fs//2, n) (last sample 0); fir_out: [batch, n].
 * fs (even) only enters through the rounding of firwin2's phase ramp, which is reproduced bit for bit
 * because the homomorphic log amplifies it ("min-phase tolerance"). */
typedef struct impb_minphase impb_minphase;
/* hamming_host: scipy.signal.get_window('hamming', 2n, fftbins=False), [2n] doubles on the HOST */
IMPB_API int32_t impb_minphase_create(impb_ctx* ctx, int64_t n, const double* hamming_host, impb_minphase** out);
IMPB_API void impb_minphase_destroy(impb_minphase* mp);
IMPB_API int64_t impb_minphase_workspace_bytes(const impb_minphase* mp, int64_t batch);
IMPB_API int32_t impb_minphase_fir(impb_minphase* mp, const double* gain, int64_t ld_gain, int32_t fs, double* fir_out,
                          int64_t ld_out, int64_t batch, void* workspace, void* stream);

/* ---- K15+K9: decay window + polyphase resampling ------------------------------------------------------------
 * Replaces decay.apply_decay_window (core/decay.py:383-403) followed by ImpulseResponse.resample,
 * core/impulse_response.py:121-124 (nnresample -> scipy.signal.resample_poly -> upfirdn).
 * taps: the resample_poly window already scaled by `up` ([ntaps], odd).  decay: [batch,4] doubles
 * (window_start, half_window, knee_index, window_level_dB) or NULL.  dst row length = ceil(n*up/down). */
IMPB_API int32_t impb_window_resample(impb_ctx* ctx, const double* src, int64_t ld_src, int64_t n, const double* decay,
                             const double* taps, int64_t ntaps, int32_t up, int32_t down, double* dst,
                             int64_t ld_dst, int64_t batch, void* stream);

/* Planned form of the same operation for callers that hold the taps on the HOST (the reference does: nnresample
 * designs the window with numpy and hands it to resample_poly, core/impulse_response.py:121-124).  The plan keeps
 * per-run tap tables that travel to the kernel in its parameter space (constant bank), see csrc/resample.cu;
 * ratios without long runs of unit-step outputs (and every up == 1 decimation) run the kernels of
 * impb_window_resample on a device copy of the taps.  Results are bit-identical between the two entry points. */
typedef struct impb_resample impb_resample;
IMPB_API int32_t impb_resample_plan_create(impb_ctx* ctx, const double* taps_host, int64_t ntaps, int32_t up, int32_t down,
                                  impb_resample** out);
IMPB_API void impb_resample_plan_destroy(impb_resample* plan);
IMPB_API int32_t impb_resample_execute(impb_resample* plan, const double* src, int64_t ld_src, int64_t n,
                              const double* decay, double* dst, int64_t ld_dst, int64_t batch, void* stream);

/* ---- K22: per-set integer bookkeeping on the device ------------------------------------------------------------
 * Replaces the per-recording-set Python arithmetic of HRIR.crop_heads (core/hrir.py:573-606: pairwise first peak
 * minus the speaker delay), align_ipsilateral_all (:936-955: which responses a lag shifts),
 * align_onset_groups_peak_leftref (:981-1001: group onset against FL-left) and crop_tails (:627-646:
 * tail = min(min len, next_fast_len(max knee))) for S sets of one layout at once, so that a batch needs no
 * device->host read between its kernels.  All arrays are device pointers; rows are (set, response) pairs, s * n + r. */
typedef struct impb_set_layout {
  int32_t n;                 /* responses per set */
  int32_t n_spk;             /* speakers: rows left[k] / right[k], crop delay[k] = round(SPEAKER_DELAYS * fs) + head */
  const int32_t* left;
  const int32_t* right;
  const int64_t* delay;
  const int32_t* spk_of_row; /* [n] speaker index of every row */
  int32_t n_pairs;           /* ipsilateral pairs: left row of a, right row of b */
  const int32_t* pa;
  const int32_t* pb;
  int32_t n_pos;             /* rows shifted by max(lag, 0) / max(-lag, 0) of pair pos_k / neg_k */
  const int32_t* pos_rows;
  const int32_t* pos_k;
  int32_t n_neg;
  const int32_t* neg_rows;
  const int32_t* neg_k;
  int32_t n_groups;          /* onset groups: leader row gl[g]; g_ref = the group led by FL-left */
  const int32_t* gl;
  int32_t g_ref;
  int32_t n_grows;           /* rows moved by the onset difference of group g_of_row[i] */
  const int32_t* g_rows;
  const int32_t* g_of_row;
} impb_set_layout;

/* crop, n1 = L - crop, off1 = row * L + crop, hint_arg = e_arg - crop per row; segment offsets / lengths (<= seg) of the
 * ipsilateral pairs for impb_xcorr_lag (xa_*, xb_*: [S, n_pairs], nullable) */
IMPB_API int32_t impb_set_heads(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* peaks, const int64_t* e_arg,
                       int64_t L, int64_t seg, int64_t S, int64_t* crop, int64_t* n1, int64_t* off1, int64_t* hint_arg,
                       int64_t* xa_off, int64_t* xa_len, int64_t* xb_off, int64_t* xb_len, void* stream);
/* shift / surviving range [lo, hi) per row from the lags ([S, n_pairs], nullable = no pairs); gathered views of the
 * group leaders for impb_peak_index (g_*: [S, n_groups], nullable) */
IMPB_API int32_t impb_set_shift(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* lags, const int64_t* n1,
                       const int64_t* off1, const int64_t* hint_arg, const double* e_max, int64_t S, int64_t* shift,
                       int64_t* lo, int64_t* hi, int64_t* g_off, int64_t* g_len, int64_t* g_shift, int64_t* g_lo,
                       int64_t* g_hi, double* g_max, int64_t* g_arg, void* stream);
/* composes the group-onset shift into shift / lo / hi (in place); group_peaks: [S, n_groups], nullable = none */
IMPB_API int32_t impb_set_shift2(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* group_peaks, const int64_t* n1,
                        int64_t S, int64_t* shift, int64_t* lo, int64_t* hi, void* stream);
/* tails[S], set_off[S + 1] (exclusive sum of tail * n), out_len / dst_off per row; smooth: ascending 5-smooth numbers
 * (next_fast_len table); flags[0] bits: 1 = a response shorter than the fade-in, 2 = a tail shorter than the fade-out,
 * 4 = a tail above tail_bound or a knee beyond the table (the caller falls back to an exactly sized gather) */
IMPB_API int32_t impb_set_tails(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* knee, const int64_t* n1, int64_t S,
                       const int64_t* smooth, int32_t n_smooth, int64_t head, int64_t fade_out_len, int64_t tail_bound,
                       int64_t* tails, int64_t* set_off, int64_t* out_len, int64_t* dst_off, int64_t* flags, void* stream);

/* ---- output packing (next-row component) -------------------------------------------------------
 * Replaces the np.vstack / transpose / libsndfile PCM conversion of HRIR.write_wav (core/hrir.py:426-455) and
 * audio_io.write_wav (core/audio_io.py:82-97).  row_offsets: device int64[n_tracks], element offset of each
 * track's n samples inside src, or -1 for a silent (missing) track.  dst: interleaved [n][n_tracks] frames,
 * int16 for bit_depth 16, int32 for 24 (value range +-2^23) and 32.  Rounding: lrint(x * 2^(bits-1)), clipped. */
IMPB_API int32_t impb_pack_tracks(impb_ctx* ctx, const double* src, const int64_t* row_offsets, int64_t n,
                         int32_t n_tracks, int32_t bit_depth, void* dst, void* stream);

/* ceil(n * up / down) after reducing up/down by their gcd: the row length impb_window_resample writes */
IMPB_API int64_t impb_resample_out_len(int64_t n, int32_t up, int32_t down);

/* ---- planner introspection (host only, needs no device) -------------------------------------------------
 * The N = n1 x n2 factorisation and radix sequences the library would use for a transform of at least
 * (exact = 0) or exactly (exact = 1) n_min points with elem_bytes-wide reals (8 / 4).
 * radices_col / radices_row: int32[6], zero-terminated. */
IMPB_API int32_t impb_plan_probe(int64_t n_min, int32_t exact, int32_t elem_bytes, int64_t* n1, int64_t* n2,
                        int32_t* radices_col, int32_t* radices_row);

#ifdef __cplusplus
}
#endif
#endif /* IMPB_H */
