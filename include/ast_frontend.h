/*
 * ast_frontend.h - C-ABI of the B200-native spectral front-end for Audio-Style-Transfer.
 *
 * The reference exposes no FFI: its boundary is the set of Python functions in
 * utilityFunctions.py / dataloader.py.  Each entry point below names the
 * reference function(s) it replaces (file:line relative to the reference tree).  The
 * Python mirror in audio-style-transfer_b200/ binds these with ctypes; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - every entry point returns int: AST_OK (0) or a negative ast_status; no exceptions
 *     cross the ABI; ast_last_error() gives the message of the calling thread's last error;
 *   - the CALLER owns every data buffer (device pointers unless a parameter is named
 *     host_*); the library allocates only plan-time constants;
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*), no
 *     internal synchronisation; a plan is immutable after creation and may be used from
 *     several host threads on different streams;
 *   - EXCEPT: at most ONE ast_features_forward / ast_cqt_forward / ast_stats_accumulate call
 *     may be in flight per device at a time (issue them on one stream, or order streams with
 *     events).  Their tensor-core kernels are persistent grids whose CTAs wait on completion
 *     counters written by other CTAs of the same call; two such grids dispatched interleaved
 *     from different streams could each be partly resident and starve one another (the
 *     bounded waits would then trap).  Copies and every other entry point may overlap freely;
 *   - a plan lives on ONE device (ast_config.device): ast_plan_create makes that device the calling
 *     thread's current device and leaves it so; every compute entry point launches on the calling
 *     thread's current device, which must be the plan's (and the stream's) device;
 *   - float32 data, int32 lengths, row-major contiguous tensors in the reference's layouts.
 *
 * Geometry (fixed by the reference's call sites, utilityFunctions.py:12,39,62):
 *   sample_rate 22050, n_fft 1024, hop 256, Hann periodic, 513 STFT bins, 84 CQT bins
 *   (12 per octave, fmin C1), F = 597.  Section window / overlap are parameters
 *   (287 / 96 in utilityFunctions.py:8-10, 287 / 86 in evaluation_style_transfer.py:27).
 */
#ifndef AST_FRONTEND_H
#define AST_FRONTEND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AST_N_FFT 1024
#define AST_HOP 256
#define AST_F_STFT 513
#define AST_F_CQT 84
#define AST_F_TOTAL 597
#define AST_N_OCTAVES 7
#define AST_CQT_NFFT 256
#define AST_SAMPLE_RATE 22050

typedef enum ast_status {
  AST_OK = 0,
  AST_ERR_INVALID_ARG = -1,  /* null pointer, negative size, unsupported geometry          */
  AST_ERR_TOO_SHORT = -2,    /* clip <= n_fft/2 samples (torch.stft reflect pad would raise) */
  AST_ERR_NO_SECTIONS = -3,  /* fewer than window/2 frames (torch.stack([]) in the reference) */
  AST_ERR_WORKSPACE = -4,    /* workspace too small                                          */
  AST_ERR_CUDA = -5,         /* a CUDA runtime call failed; see ast_last_error()             */
  AST_ERR_SHAPE = -6         /* tensor shapes disagree (ValueError in concat_stft_cqt)       */
} ast_status;

typedef enum ast_layout {
  AST_LAYOUT_FLAT = 0,     /* (B, 2, T, F)            - get_STFT / get_CQT / concat_stft_cqt  */
  AST_LAYOUT_SECTIONS = 1  /* (B, S, 2, window, F)    - get_overlap_windows + custom_collate_fn */
} ast_layout;

typedef struct ast_config {
  int32_t sample_rate;     /* must be 22050 */
  int32_t n_fft;           /* must be 1024  */
  int32_t hop;             /* must be 256   */
  int32_t n_bins;          /* must be 84    */
  int32_t window_size;     /* default 287, utilityFunctions.py:8  */
  int32_t overlap_frames;  /* default 96,  utilityFunctions.py:10 */
  int32_t device;          /* CUDA device ordinal the plan lives on */
} ast_config;

typedef struct ast_plan ast_plan;

/* ---- plan ---------------------------------------------------------------------------- */
int ast_default_config(ast_config* cfg);
int ast_plan_create(const ast_config* cfg, ast_plan** plan);
int ast_plan_destroy(ast_plan* plan);
/* message of the calling thread's last failing call (empty string if none) */
const char* ast_last_error(void);
const char* ast_version(void);

/* ---- host-side geometry helpers (no CUDA) ----------------------------------------------- */
/* T = 1 + L / hop                                         (torch.stft, utilityFunctions.py:26) */
int32_t ast_num_frames(int64_t n_samples);
/* number of sections get_overlap_windows emits for T frames (utilityFunctions.py:249-261)   */
int32_t ast_num_sections(int32_t n_frames, int32_t window_size, int32_t overlap_frames);
/* hop * (T - 1) samples out of torch.istft               (utilityFunctions.py:78-80)        */
int64_t ast_istft_length(int32_t n_frames);
/* plan constants, so a checker can compare them with its own derivation:
 *   decimator taps (double, n = 385), CQT time-domain kernel (12 x 256 complex, interleaved
 *   re/im doubles, top octave), per-bin filter lengths (84 doubles)                          */
int ast_host_decimator_taps(double* taps, int32_t capacity, int32_t* n_taps);
int ast_host_cqt_kernel(double* kernel_12x256x2);
int ast_host_cqt_lengths(double* lengths_84);

/* bytes of device scratch ast_cqt_forward / ast_features_forward need for a batch of B clips
 * of at most max_samples samples (decimated octave signals + the normalisation table)        */
size_t ast_workspace_bytes(const ast_plan* plan, int32_t batch, int64_t max_samples);
/* bytes ast_stats_accumulate needs (adds the un-normalised flat features of the batch and the
 * per-clip moments); ast_stats_accumulate_features needs only the per-clip moments and
 * accepts the same figure                                                                    */
size_t ast_stats_workspace_bytes(const ast_plan* plan, int32_t batch, int64_t max_samples);

/* ---- feature extraction ---------------------------------------------------------------- */
/*
 * replaces get_STFT (utilityFunctions.py:12-37), batched.
 *   wave     (B, wave_stride) f32, clip b holds lengths[b] valid samples (lengths == NULL:
 *            every clip has max_samples samples)
 *   out      (B, 2, t_out, 513) f32 contiguous (the reference returns a permuted view; values
 *            are identical); rows t >= 1 + lengths[b]/256 are zero-filled
 *   imag of bin 0 and bin 512 is exactly 0.0
 */
int ast_stft_forward(const ast_plan* plan, const float* wave, const int32_t* lengths, int32_t batch,
                     int64_t max_samples, int64_t wave_stride, float* out, int32_t t_out, void* stream);

/*
 * replaces get_CQT (utilityFunctions.py:39-60) = librosa.cqt(y, sr=22050, n_bins=84,
 * hop_length=256), batched.  out (B, 2, t_out, 84) f32.
 */
int ast_cqt_forward(const ast_plan* plan, const float* wave, const int32_t* lengths, int32_t batch,
                    int64_t max_samples, int64_t wave_stride, void* workspace, size_t workspace_bytes,
                    float* out, int32_t t_out, void* stream);

/*
 * replaces DualInstrumentDataset.__getitem__'s spectral part + custom_collate_fn's layout
 * (dataloader.py:100-112, :123-147): get_STFT + get_CQT -> normalize each -> concat on F ->
 * get_overlap_windows, for a whole batch in one call.
 *   mean, std   (2, 597) f32 (STFT stats then CQT stats) or both NULL = no normalisation
 *               (as evaluation_style_transfer.process_audio, :135-139);
 *               x -> (x - mean) / (std + eps)                        (dataloader.py:9-13)
 *   per-clip statistics: mean/std may instead be (B, 2, 597) when stats_per_clip != 0
 *               (piano rows use piano stats, violin rows violin stats, dataloader.py:106-109)
 *   layout FLAT      out (B, 2, dim1, 597), dim1 = t_out frames
 *   layout SECTIONS  out (B, dim1, 2, window, 597), dim1 = s_max sections; sections
 *               >= n_sections[b] and rows past the clip's last frame are zero-filled
 *   n_sections  (B) int32 device, nullable; number of sections (SECTIONS) or frames (FLAT)
 */
int ast_features_forward(const ast_plan* plan, const float* wave, const int32_t* lengths, int32_t batch,
                         int64_t max_samples, int64_t wave_stride, const float* mean, const float* std,
                         int32_t stats_per_clip, float eps, void* workspace, size_t workspace_bytes,
                         float* out, int32_t dim1, int32_t layout, int32_t* n_sections, void* stream);

/* ---- reconstruction -------------------------------------------------------------------- */
/*
 * replaces sections2spectrogram + inverse_STFT (utilityFunctions.py:265-283, :62-82) and
 * reconstruct_audio_from_sections (evaluation_reconstruction.py:161-189), batched.
 *   layout FLAT      spec (B, 2, dim1, f_in), dim1 = T frames
 *   layout SECTIONS  spec (B, dim1, 2, window, f_in), dim1 = S sections; merged with the
 *                    count-normalised overlap average, hop = window - overlap, then cropped
 *                    to original_size frames (<= 0: no crop)
 *   f_in >= 513; only the first 513 columns are read (597-wide feature tensors are accepted,
 *                as test_correctness.ipynb cell 11 slices [:, :, :513])
 *   wave_out (B, out_stride) f32; 256 * (T' - 1) samples per clip are written,
 *            T' = merged / cropped frame count.  imag of bins 0 and 512 is ignored (torch.istft).
 */
int ast_istft_forward(const ast_plan* plan, const float* spec, int32_t batch, int32_t dim1, int32_t f_in,
                      int32_t layout, int32_t overlap_frames, int32_t original_size, float* wave_out,
                      int64_t out_stride, void* stream);

/* ---- the small layout operators, one call each ------------------------------------------- */
/* normalize (dataloader.py:9-13): x (n_ch, T, F), mean/std (n_ch, F) -> out (n_ch, T, F)    */
int ast_normalize(const float* x, const float* mean, const float* std, float eps, int32_t n_ch,
                  int32_t n_time, int32_t n_freq, float* out, void* stream);
/* concat_stft_cqt (utilityFunctions.py:285-299): (n_ch, T, f1) + (n_ch, T, f2) -> (n_ch, T, f1+f2) */
int ast_concat(const float* a, const float* b, int32_t n_ch, int32_t n_time, int32_t f1, int32_t f2,
               float* out, void* stream);
/* get_overlap_windows (utilityFunctions.py:240-263): (n_ch, T, F) -> (S, n_ch, window, F)   */
int ast_overlap_windows(const float* spec, int32_t n_ch, int32_t n_time, int32_t n_freq, int32_t window_size,
                        int32_t overlap_frames, float* out, int32_t n_sections, void* stream);
/* sections2spectrogram (utilityFunctions.py:265-283): (S, n_ch, window, F) -> (n_ch, t_out, F),
 * t_out = min(hop * (S - 1) + window, original_size)                                        */
int ast_sections_merge(const float* sections, int32_t n_sections, int32_t n_ch, int32_t window_size,
                       int32_t n_freq, int32_t overlap_frames, int32_t t_out, float* out, void* stream);

/* ---- dataset statistics ------------------------------------------------------------------ */
/*
 * replaces compute_stats (Preprocessing_Dataset/compute_separated_stats.py:16-43,
 * compute_unified_stats.py:25-50).  For every clip: per-bin mean over T and UNBIASED variance
 * over T of the raw (2, T, 597) features; adds them (float64) into the group's running sums.
 *   group_ids  (B) int32 device, nullable (all group 0); e.g. 0 = piano, 1 = violin
 *   acc        (n_groups, 2, 2, 597) f64 device: [g][0] = sum of clip means, [g][1] = sum of
 *              clip variances;  counts (n_groups) f64 device: clips accumulated
 * Deterministic: clips are added in batch order by a single writer per element.  Partial
 * sums of several ranks are combined by the caller (one all-reduce of acc and counts).
 */
int ast_stats_accumulate(const ast_plan* plan, const float* wave, const int32_t* lengths,
                         const int32_t* group_ids, int32_t batch, int64_t max_samples, int64_t wave_stride,
                         void* workspace, size_t workspace_bytes, int32_t n_groups, double* acc,
                         double* counts, void* stream);
/* same accumulation from already-extracted flat features (B, 2, t_dim, f_dim) f32         */
int ast_stats_accumulate_features(const ast_plan* plan, const float* feats, const int32_t* n_frames,
                                  const int32_t* group_ids, int32_t batch, int32_t t_dim, int32_t f_dim,
                                  void* workspace, size_t workspace_bytes, int32_t n_groups, double* acc,
                                  double* counts, void* stream);
/* host: mean = sum_mean / N, std = sqrt(sum_var / N) -> (2, 597) f32 each                   */
int ast_stats_finalize(const double* host_acc_2x2x597, double count, float* host_mean, float* host_std);

/* ---- audio loading: the step before the hot path ------------------------- */
/*
 * The device part of load_audio (utilityFunctions.py:105-122): zero-pad / cut each clip to cut_samples
 * (= int(cut_time_seconds * orig_sr)), resample orig_sr -> new_sr exactly as
 * torchaudio.functional.resample(waveform, orig_sr, new_sr) does with its defaults (sinc_interp_hann,
 * lowpass_filter_width = 6, rolloff = 0.99), then average the two channels of a stereo clip
 * (torch.mean(waveform, dim=0, keepdim=True), :119-120).  File decoding (torchaudio.load) stays on the host.
 *
 * An ast_resampler holds the polyphase taps of one (orig_sr, new_sr) pair on one device; immutable after creation.
 *   wave        (B, C, in_stride) f32 device, C in {1, 2}; clip b holds lengths_in[b] valid samples per channel
 *               (lengths_in == NULL: in_stride), samples past min(lengths_in[b], cut_samples) read as zeros
 *   out         (B, out_stride) f32 device; ast_resample_length(cut_samples, ...) samples per clip are written
 */
typedef struct ast_resampler ast_resampler;
/* reduced rates (divided by their gcd) and the half width of the FIR: taps per phase = 2 * width + orig_reduced */
int ast_resample_geometry(int32_t orig_sr, int32_t new_sr, int32_t* orig_reduced, int32_t* new_reduced, int32_t* width);
/* ceil(new * n_in / orig), the output length of torchaudio.functional.resample; -1 on invalid arguments */
int64_t ast_resample_length(int64_t n_in, int32_t orig_sr, int32_t new_sr);
/* host copy of the taps, (new_reduced, 2 * width + orig_reduced) f32, so a checker can compare them */
int ast_host_resample_taps(int32_t orig_sr, int32_t new_sr, float* taps, int32_t capacity);
int ast_resampler_create(int32_t orig_sr, int32_t new_sr, int32_t device, ast_resampler** resampler);
int ast_resampler_destroy(ast_resampler* resampler);
int ast_load_audio_forward(const ast_resampler* resampler, const float* wave, const int32_t* lengths_in,
                           int32_t batch, int32_t channels, int64_t in_stride, int64_t cut_samples, float* out,
                           int64_t out_stride, void* stream);

/* ---- evaluation metric on the far side of the iSTFT ----------------------- */
/*
 * replaces mse_spectrogram (evaluation_reconstruction.py:105-118, evaluation_style_transfer.py:111-119):
 * mean over (513, T) of (|librosa.stft(a, n_fft=1024, hop_length=256)| - |librosa.stft(b, ...)|)^2 with
 * T = min(T_a, T_b); librosa.stft defaults: center=True with ZERO padding, periodic Hann window.
 *   a, b     mono signals on the device (n_a, n_b samples, both >= 1); result: 1 double on the device
 */
size_t ast_mse_workspace_bytes(const ast_plan* plan, int64_t n_a, int64_t n_b);
int ast_mse_spectrogram(const ast_plan* plan, const float* a, int64_t n_a, const float* b, int64_t n_b,
                        void* workspace, size_t workspace_bytes, double* result, void* stream);

/*
 * replaces instrumentation_similarity (evaluation_style_transfer.py:111-119): Pearson correlation between the
 * per-bin sums over time of |librosa.stft(a)| and |librosa.stft(b)| (librosa defaults: n_fft = 2048, hop = 512,
 * center=True with ZERO padding, periodic Hann -> 1025 bins); 0.0 where the reference's pearsonr returns NaN
 * (a constant energy profile, e.g. silence).
 *   a, b     mono signals on the device (n_a, n_b samples, both >= 1); result: 1 double on the device
 */
size_t ast_instrumentation_similarity_workspace_bytes(const ast_plan* plan, int64_t n_a, int64_t n_b);
int ast_instrumentation_similarity(const ast_plan* plan, const float* a, int64_t n_a, const float* b, int64_t n_b,
                                   void* workspace, size_t workspace_bytes, double* result, void* stream);

/* ---- synthetic workload (no counterpart in the reference) ---------------------------------- */
/*
 * Dataset-scale inputs for BASELINE.json configs[3] (100 000 clips): clip `first_clip_id + b` is written
 * to out + b * out_stride.  Every sample is a pure function of (clip id, sample index): piano-like decaying notes for
 * ids < violin_from_id, violin-like sustained notes with vibrato from there on, plus white noise, RMS about 0.07.
 * Counter-based (integer hash of 1000 + clip id), so any rank regenerates any clip bit for bit in any chunking.
 */
int ast_synth_clips(float* out, int64_t out_stride, int32_t n_clips, int64_t n_samples, int64_t first_clip_id,
                    int64_t violin_from_id, void* stream);

/* ---- diagnostics --------------------------------------------------------------------------- */
/*
 * Per-kernel device timing (no counterpart in the reference).  While enabled, every kernel launch
 * made by the entry points above is bracketed by CUDA events on the caller's stream;
 * ast_profile_collect() synchronises those events, sums the elapsed time per kernel name and
 * clears the record.  names: capacity x 32 chars (NUL-terminated), total_ms / launches: capacity.
 * Process-wide, not thread-safe; leave disabled on production paths.
 */
int ast_profile_enable(int32_t on);
int ast_profile_collect(char* names, float* total_ms, int32_t* launches, int32_t capacity, int32_t* n_kernels);

#ifdef __cplusplus
}
#endif
#endif /* AST_FRONTEND_H */
