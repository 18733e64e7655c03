// common.cuh - shared declarations of libast_frontend.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ast_frontend.h"
#include "fft_core.h"

// Diagnostic build only (-DAST_TIMELINE, scratch/timeline.py): every CTA of the three feature kernels stamps its start and
// end (%globaltimer) and its SM, per kernel, so the whole call can be drawn as a Gantt chart.
#ifdef AST_TIMELINE
#define AST_TIMELINE_DEFINE(name)                                                                          \
  __device__ unsigned long long g_tl_##name[8192][4];                                                      \
  extern "C" int ast_debug_timeline_##name(unsigned long long* host) {                                     \
    return (int)cudaMemcpyFromSymbol(host, g_tl_##name, sizeof(g_tl_##name));                              \
  }
#define AST_TIMELINE_STAMP(name, cta, k) AST_TIMELINE_STAMP_IF(threadIdx.x == 0, name, cta, k)
#define AST_TIMELINE_STAMP_IF(pred, name, cta, k)                                                          \
  do {                                                                                                     \
    if ((pred) && (cta) < 8192) {                                                                          \
      unsigned long long t_;                                                                               \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                                \
      g_tl_##name[cta][k] = t_;                                                                            \
      if ((k) == 0) {                                                                                      \
        unsigned s_;                                                                                       \
        asm volatile("mov.u32 %0, %smid;" : "=r"(s_));                                                     \
        g_tl_##name[cta][2] = s_;                                                                          \
      }                                                                                                    \
    }                                                                                                      \
  } while (0)
#else
#define AST_TIMELINE_DEFINE(name)
#define AST_TIMELINE_STAMP(name, cta, k) do {} while (0)
#define AST_TIMELINE_STAMP_IF(pred, name, cta, k) do {} while (0)
#endif

namespace ast {

constexpr int kNfft = AST_N_FFT;
constexpr int kHop = AST_HOP;
constexpr int kFStft = AST_F_STFT;
constexpr int kFCqt = AST_F_CQT;
constexpr int kFTotal = AST_F_TOTAL;
constexpr int kOctaves = AST_N_OCTAVES;
constexpr int kCqtNfft = AST_CQT_NFFT;
constexpr int kBinsPerOctave = 12;
constexpr int kCqtCols = 2 * kBinsPerOctave;  // 12 real + 12 imaginary outputs per octave
constexpr int kDecTaps = 385;
constexpr int kDecHalf = (kDecTaps - 1) / 2;  // 192

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define AST_CUDA_TRY(expr)                                                                      \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      return ::ast::fail(AST_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                         __FILE__, __LINE__);                                                   \
  } while (0)

// Diagnostic per-kernel timing (ast_profile_enable / ast_profile_collect, api.cu).
bool profile_on();
void profile_mark(const char* name, cudaStream_t st, bool begin);
struct ProfileSpan {
  const char* name;
  cudaStream_t st;
  bool on;
  ProfileSpan(const char* n, cudaStream_t s) : name(n), st(s), on(profile_on()) {
    if (on) profile_mark(name, st, true);
  }
  ~ProfileSpan() {
    if (on) profile_mark(name, st, false);
  }
};

#define AST_LAUNCH_CHECK(name)                                                                   \
  do {                                                                                          \
    cudaError_t err__ = cudaGetLastError();                                                     \
    if (err__ != cudaSuccess)                                                                   \
      return ::ast::fail(AST_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(err__)); \
  } while (0)

// Programmatic dependent launch: the kernel may start (prologue: shared-memory tables, TMEM allocation, barrier
// init) while the previous kernel of the stream is still draining; it must execute pdl_wait() before touching
// anything the previous kernel reads or writes.  Off while profiling (events between launches would serialise).
template <class Params>
inline cudaError_t launch_with_pdl(void (*kernel)(Params), dim3 grid, unsigned block, size_t smem, cudaStream_t st,
                                   const Params& params) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = profile_on() ? 0 : 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, params);
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---- host-side constant derivation (plan.cu; double precision, no CUDA) -------------------
void host_hann(double* w, int n);
void host_decimator_taps(double* taps);                 // kDecTaps, unit DC gain
void host_cqt_kernel(double* k_re, double* k_im);       // 12 x 256 each, top octave
void host_cqt_lengths(double* lengths);                 // 84

// ---- where one frame of one clip lands in the output tensor -------------------------------
// FLAT:     out[b][c][t][f]                rows t >= frames(b) are zero
// SECTIONS: out[b][s][c][tau][f], t = s * step + tau; rows of sections >= n_sections(b) or of
//           frames >= frames(b) are zero (get_overlap_windows pads AFTER normalisation).
struct OutSpec {
  float* out;
  int layout;          // ast_layout
  int dim1;            // t_out (FLAT) or s_max (SECTIONS)
  int f_row;           // floats per output row
  int f_off;           // first column this kernel writes
  int window, step;    // SECTIONS geometry
  const float2* stats; // (mean, 1/(std+eps)) pairs, [clip or 0][c][f_stats], or nullptr
  int stats_clip_stride;  // 0 (shared) or 2 * f_stats
  int f_stats;            // columns per stats row
  int stats_off;          // column in the stats row of this kernel's first bin
  float2* cqt_part;       // statistics mode of the CQT projection: [clip][octave][tile][quadrant][24] (mean, M2) of the
                          // quadrant's (<= 32) live frames; nothing else is stored when set
};

__host__ __device__ inline int num_frames(long long n_samples) { return 1 + (int)(n_samples / kHop); }

// get_overlap_windows' section count (utilityFunctions.py:249-261) in closed form
__host__ __device__ inline int num_sections(int n_frames, int window, int overlap) {
  const int step = window - overlap;
  if (n_frames <= 0 || step <= 0) return 0;
  int s_last = n_frames > window ? (n_frames - window + step - 1) / step : 0;  // first section reaching the end
  if (s_last * step >= n_frames) return s_last;  // (cannot happen for step <= window; kept for safety)
  return s_last + ((2 * (n_frames - s_last * step) >= window) ? 1 : 0);
}

// number of frame slots the feature kernels iterate over for one clip
__host__ __device__ inline int frame_slots(int layout, int dim1, int window, int step) {
  return layout == AST_LAYOUT_FLAT ? dim1 : (dim1 > 0 ? step * (dim1 - 1) + window : 0);
}

#if defined(__CUDACC__)
// Destination rows (channel-0 plane) of frame slot t of clip b; at most 2 (2 * overlap <= window).
struct RowDest {
  float* row[2];
  bool live[2];       // false -> the row must be written as zeros
  int n;
  long long plane;    // floats from the channel-0 row to the channel-1 row
};

__device__ __forceinline__ RowDest row_dest(const OutSpec& o, int b, int t, int frames_b, int sections_b) {
  RowDest d;
  d.n = 0;
  d.row[0] = d.row[1] = nullptr;
  d.live[0] = d.live[1] = false;
  if (o.layout == AST_LAYOUT_FLAT) {
    d.plane = (long long)o.dim1 * o.f_row;
    d.row[0] = o.out + ((long long)b * 2 * o.dim1 + t) * o.f_row + o.f_off;
    d.live[0] = t < frames_b;
    d.n = 1;
  } else {
    d.plane = (long long)o.window * o.f_row;
    int s_hi = t / o.step;
    if (s_hi > o.dim1 - 1) s_hi = o.dim1 - 1;
    for (int s = s_hi; s >= 0 && s >= s_hi - 1; --s) {
      const int tau = t - s * o.step;
      if (tau >= o.window) break;
      d.row[d.n] = o.out + (((long long)b * o.dim1 + s) * 2 * o.window + tau) * o.f_row + o.f_off;
      d.live[d.n] = (s < sections_b) && (t < frames_b);
      ++d.n;
    }
  }
  return d;
}
#endif

}  // namespace ast

struct ast_plan {
  ast_config cfg;
  int sm_count;
  float2* d_tw1;        // stage-1 twiddles [16][16]  (fft_core.h)
  float2* d_tw2;        // stage-2 twiddles [16][64]
  float2* d_tw32;       // [32][32] W_1024^(n2 k1): the one-warp 32 x 32 transform (stft.cu)
  float* d_hann;        // 1024 periodic Hann
  float* d_hann_inv_n;  // Hann / 1024 (iSTFT synthesis window with the irfft scale folded in)
  float* d_hann_sq;     // Hann^2 (iSTFT envelope)
  float* d_cqt_kernel;  // [256][24] base time-domain CQT kernel (12 re then 12 im columns)
  float* d_cqt_scale;   // [7][12] per octave / bin scale sqrt(2^i) / sqrt(length_k)
  float* d_cqt_tc_images;  // CQT kernel as TF32 hi / lo B-operand images per pass (cqt_tc.cu)
  float* d_dec_strip_hi;  // decimator Toeplitz strip, TF32 hi part (smem image, decimate.cu)
  float* d_dec_strip_lo;  // ... and the TF32 residual
  uint16_t* d_dec_strip_h_hi;  // the FP16-split kernel's strip (taps x 2^15 as FP16) and its FP16 residual
  uint16_t* d_dec_strip_h_lo;
};

namespace ast {
// kernels' host launchers (each returns an ast_status)
constexpr int kStat4Stride = 516;   // float4 per clip of the STFT kernel's per-bin statistics (513 bins, padded)
int launch_prep_stats(const float* mean, const float* std, float eps, int n, float2* table, float4* stat4, cudaStream_t st);
int launch_count_sections(const int32_t* lengths, int batch, long long max_samples, int layout, int dim1,
                          int window, int overlap, int32_t* n_out, cudaStream_t st);
// stat4: per-bin (-mean_re, -mean_im, rstd_re, rstd_im) of the 513 STFT bins, kStat4Stride float4 per clip (or one shared
// row), or nullptr; part / part_n: statistics mode (per-tile moments instead of any output), see stft.cu
int launch_stft(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
                long long wave_stride, const OutSpec& out, cudaStream_t st, int pad_zero = 0, bool pdl = false,
                unsigned int* tail_counter = nullptr, const float4* stat4 = nullptr, int stat4_clip_stride = 0,
                float2* part = nullptr, float* part_n = nullptr);
int stft_tiles_per_clip(const ast_plan* plan, int batch, int slots, bool stats_mode);   // grid.x of the launch above
int launch_features_prologue(const float* mean, const float* std_, float eps, int n_stats, float2* table, float4* stat4,
                             const int32_t* lengths, int batch, long long max_samples, int layout, int dim1, int window,
                             int overlap, int32_t* n_out, int* flags, int n_flags, cudaStream_t st);
// K6 fused: per-clip moments from the per-tile partials the STFT / CQT kernels leave in statistics mode
int launch_stats_finalize_clips(const float2* part_stft, const float* part_n, int stft_tiles, const float2* part_cqt,
                                int cqt_tiles, const int32_t* lengths, long long max_samples, int batch, double* clip_stats,
                                cudaStream_t st);
int launch_decimate_cascade(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch,
                            long long max_samples, long long wave_stride, float* ws, long long ws_clip_stride,
                            int* flags, cudaStream_t st, bool flags_zeroed = false);
int launch_cqt(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
               long long wave_stride, const float* ws, long long ws_clip_stride, const int* dec_flags, const OutSpec& out,
               cudaStream_t st, bool tile_queue = false);
int launch_istft(const ast_plan* plan, const float* spec, int batch, int dim1, int f_in, int layout, int window,
                 int overlap, int n_frames, float* wave_out, long long out_stride, cudaStream_t st);
int launch_clip_stats(const float* feats, const int32_t* n_frames, int batch, int t_dim, int f_dim, double* clip_stats,
                      cudaStream_t st);
int launch_stats_accumulate(const double* clip_stats, const int32_t* group_ids, int batch, int f_dim, int n_groups,
                            double* acc, double* counts, cudaStream_t st);
int upload_decimator_taps(const float* taps_scaled);
int stft_init();   // opt-in shared memory + occupancy query (once per device)
int decimate_init();
int cqt_tc_init();
int cqt_tc_image_floats();
void host_cqt_tc_images(const double* kmat_256x24, float* images);
int launch_cqt_tc(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
                  long long wave_stride, const float* ws, long long ws_clip_stride, const int* dec_flags, const OutSpec& out,
                  cudaStream_t st, bool tile_queue = false);
void set_tc_cqt(int on);
void set_overlap_streams(int on);
bool use_tc_cqt();
int decimator_strip_floats();
void host_decimator_strip(const double* taps_scaled, float* strip_hi, float* strip_lo);  // decimator_strip_floats() each
int decimator_strip_h_bytes();
void host_decimator_strip_h(const double* taps_scaled, uint16_t* strip_hi, uint16_t* strip_lo);  // decimator_strip_h_bytes() each
void set_decimator_half(int on);
void set_stft_second(int on);
size_t decimator_flag_bytes(int batch, long long max_samples);
int decimator_tile_outputs();                       // outputs per decimator tile (7424)
int decimator_tiles_stage0(long long max_samples);  // tiles per clip of the first stage = row stride of the flag array
bool use_tc_decimator();
long long decimator_stage_done_offset(int batch, long long max_samples);  // ints from the flags to the per-stage counters
int decimator_tiles_of_stage(long long max_samples, int stage);            // tiles per clip of a stage
int launch_decimate_cascade_tc(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch,
                               long long max_samples, long long wave_stride, float* ws, long long ws_clip_stride,
                               int* flags, cudaStream_t st, bool flags_zeroed = false);
void set_tc_decimator(int on);
int istft_init();  // into __constant__ memory of decimate.cu

// octave buffer layout inside the CQT workspace (floats, per clip): buffers 1..6, each padded
long long octave_len(long long n_samples, int octave);          // ceil(L / 2^octave)
long long octave_offset(long long max_samples, int octave);      // offset of buffer `octave` (>= 1)
long long cqt_ws_clip_stride(long long max_samples);
}  // namespace ast
