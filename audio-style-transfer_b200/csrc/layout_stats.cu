// layout_stats.cu - the small operators of the path:
//   prep_stats        (mean, std) -> (mean, 1 / (std + eps)) table consumed by the fused epilogues
//   count_sections    per-clip section / frame counts
//   normalize         dataloader.normalize                       (dataloader.py:9-13)
//   concat            concat_stft_cqt                            (utilityFunctions.py:285-299, dataloader.py:15-18)
//   overlap_windows   get_overlap_windows                        (utilityFunctions.py:240-263)
//   sections_merge    sections2spectrogram                       (utilityFunctions.py:265-283)
//   clip_stats / stats_accumulate   K6, compute_stats            (compute_separated_stats.py:16-43)
#include "common.cuh"

namespace ast {

// ------------------------------------------------------------------------------------------
// table[i] = (mean, 1 / (std + eps)) for the CQT epilogue; stat4[row][k] = (-mean_re, -mean_im, rstd_re, rstd_im) of
// STFT bin k: one 16-byte load per bin in the STFT kernel, already paired for its packed FP32x2 normalisation.  n = rows * 2 * 597.
__device__ __forceinline__ void prep_stats_element(const float* __restrict__ mean, const float* __restrict__ std_, float eps,
                                                   int i, float2* __restrict__ table, float4* __restrict__ stat4) {
  table[i] = make_float2(mean[i], 1.0f / (std_[i] + eps));
  const int row = i / (2 * kFTotal), r = i - row * 2 * kFTotal;
  if (stat4 && r < kFStft) {
    const int j = i + kFTotal;   // the channel-1 entry of the same bin
    stat4[row * kStat4Stride + r] = make_float4(-mean[i], -mean[j], 1.0f / (std_[i] + eps), 1.0f / (std_[j] + eps));
  }
}

__global__ void prep_stats_kernel(const float* __restrict__ mean, const float* __restrict__ std_, float eps, int n,
                                  float2* __restrict__ table, float4* __restrict__ stat4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) prep_stats_element(mean, std_, eps, i, table, stat4);
}

int launch_prep_stats(const float* mean, const float* std_, float eps, int n, float2* table, float4* stat4, cudaStream_t st) {
  if (n == 0) return AST_OK;
  prep_stats_kernel<<<(n + 255) / 256, 256, 0, st>>>(mean, std_, eps, n, table, stat4);
  AST_LAUNCH_CHECK("prep_stats_kernel");
  return AST_OK;
}

__global__ void count_sections_kernel(const int32_t* __restrict__ lengths, int batch, long long max_samples, int layout,
                                      int dim1, int window, int overlap, int32_t* __restrict__ n_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int frames = num_frames(lengths ? lengths[b] : max_samples);
  int v = layout == AST_LAYOUT_FLAT ? frames : num_sections(frames, window, overlap);
  if (v > dim1) v = dim1;
  n_out[b] = v;
}

int launch_count_sections(const int32_t* lengths, int batch, long long max_samples, int layout, int dim1, int window,
                          int overlap, int32_t* n_out, cudaStream_t st) {
  if (batch == 0) return AST_OK;
  count_sections_kernel<<<(batch + 127) / 128, 128, 0, st>>>(lengths, batch, max_samples, layout, dim1, window, overlap,
                                                             n_out);
  AST_LAUNCH_CHECK("count_sections_kernel");
  return AST_OK;
}

// One launch for everything the fused feature call needs before its first big kernel: the (mean, rstd) table, the
// per-clip section counts and the zeroing of the decimator's completion counters (otherwise a memset node).  The
// decimator that follows is a programmatic dependent of this kernel: its prologue (operand strips, TMEM, barriers)
// overlaps it, and its griddepcontrol.wait - executed BEFORE it releases its own dependents - orders the CQT
// projection and the STFT behind these writes.
struct FeaturesPrologueParams {
  const float* mean;
  const float* std_;
  float eps;
  int n_stats;
  float2* table;
  float4* stat4;
  const int32_t* lengths;
  int batch;
  long long max_samples;
  int layout, dim1, window, overlap;
  int32_t* n_out;
  int* flags;
  int n_flags;
};

__global__ void features_prologue_kernel(const FeaturesPrologueParams p) {
  // itself a programmatic dependent of whatever kernel precedes it on the stream (back-to-back feature calls: the
  // previous call's STFT, which still reads the table this kernel rewrites): launch latency overlaps, the work waits
  pdl_wait();
  pdl_launch_dependents();  // the decimator's prologue may start; it waits for this grid before it reads the counters
  const int stride = gridDim.x * blockDim.x;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = i0; i < p.n_flags; i += stride) p.flags[i] = 0;
  for (int i = i0; i < p.n_stats; i += stride) prep_stats_element(p.mean, p.std_, p.eps, i, p.table, p.stat4);
  if (p.n_out)
    for (int b = i0; b < p.batch; b += stride) {
      const int frames = num_frames(p.lengths ? p.lengths[b] : p.max_samples);
      int v = p.layout == AST_LAYOUT_FLAT ? frames : num_sections(frames, p.window, p.overlap);
      p.n_out[b] = v > p.dim1 ? p.dim1 : v;
    }
}

int launch_features_prologue(const float* mean, const float* std_, float eps, int n_stats, float2* table, float4* stat4,
                             const int32_t* lengths, int batch, long long max_samples, int layout, int dim1, int window,
                             int overlap, int32_t* n_out, int* flags, int n_flags, cudaStream_t st) {
  int work = n_stats > n_flags ? n_stats : n_flags;
  if (n_out && batch > work) work = batch;
  if (work == 0) return AST_OK;
  int ctas = (work + 255) / 256;
  if (ctas > 1184) ctas = 1184;
  const FeaturesPrologueParams p{mean, std_, eps, n_stats, table, stat4, lengths, batch, max_samples, layout, dim1, window,
                                 overlap, n_out, flags, n_flags};
  ProfileSpan span("features_prologue_kernel", st);
  AST_CUDA_TRY(launch_with_pdl(features_prologue_kernel, dim3((unsigned)ctas), 256, 0, st, p));
  return AST_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void normalize_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                 const float* __restrict__ std_, float eps, int n_time, int n_freq, long long total,
                                 float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % n_freq);
    const int c = (int)(i / ((long long)n_time * n_freq));
    const int s = c * n_freq + f;
    out[i] = (x[i] - __ldg(mean + s)) / (__ldg(std_ + s) + eps);  // exact reference formula
  }
}

__global__ void concat_kernel(const float* __restrict__ a, const float* __restrict__ b, long long rows, int f1, int f2,
                              float* __restrict__ out) {
  const int f = f1 + f2;
  const long long total = rows * f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / f;
    const int c = (int)(i - r * f);
    out[i] = c < f1 ? a[r * f1 + c] : b[r * f2 + (c - f1)];
  }
}

__global__ void overlap_windows_kernel(const float* __restrict__ spec, int n_ch, int n_time, int n_freq, int window,
                                       int step, int n_sections, float* __restrict__ out) {
  const long long total = (long long)n_sections * n_ch * window * n_freq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % n_freq);
    long long r = i / n_freq;
    const int tau = (int)(r % window);
    r /= window;
    const int c = (int)(r % n_ch);
    const int s = (int)(r / n_ch);
    const int t = s * step + tau;
    out[i] = t < n_time ? spec[((long long)c * n_time + t) * n_freq + f] : 0.f;
  }
}

__global__ void sections_merge_kernel(const float* __restrict__ sec, int n_sections, int n_ch, int window, int n_freq,
                                      int hop, int t_out, float* __restrict__ out) {
  const long long total = (long long)n_ch * t_out * n_freq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % n_freq);
    long long r = i / n_freq;
    const int t = (int)(r % t_out);
    const int c = (int)(r / t_out);
    float acc = 0.f, cnt = 0.f;
    int s_hi = t / hop;
    if (s_hi > n_sections - 1) s_hi = n_sections - 1;
    int s_lo = t >= window ? (t - window) / hop + 1 : 0;
    for (int s = s_lo; s <= s_hi; ++s) {  // ascending, like the reference's += loop
      const int tau = t - s * hop;
      acc += sec[(((long long)s * n_ch + c) * window + tau) * n_freq + f];
      cnt += 1.f;
    }
    out[i] = acc / fmaxf(cnt, 1.f);
  }
}

static unsigned grid_for(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// ------------------------------------------------------------------------------------------
// K6 part 1: per clip, per (channel, bin): mean over T and unbiased variance over T (two passes in
// float64; the second pass re-reads a 64-column strip that is L2 resident).  clip_stats[b][0] = mean,
// clip_stats[b][1] = variance, each (2, f_dim).
constexpr int kStatCols = 64;
constexpr int kStatSlices = 4;

__global__ void __launch_bounds__(kStatCols* kStatSlices) clip_stats_kernel(const float* __restrict__ feats,
                                                                             const int32_t* __restrict__ n_frames,
                                                                             int t_dim, int f_dim,
                                                                             double* __restrict__ clip_stats) {
  __shared__ double red[kStatSlices][kStatCols];
  const int b = blockIdx.z, c = blockIdx.y;
  const int col = threadIdx.x % kStatCols, slice = threadIdx.x / kStatCols;
  const int f = blockIdx.x * kStatCols + col;
  int T = n_frames ? n_frames[b] : t_dim;
  if (T > t_dim) T = t_dim;
  const float* __restrict__ x = feats + ((long long)(b * 2 + c) * t_dim) * f_dim + f;
  double s = 0.0;
  if (f < f_dim)
    for (int t = slice; t < T; t += kStatSlices) s += (double)x[(long long)t * f_dim];
  red[slice][col] = s;
  __syncthreads();
  double mean = 0.0;
  for (int i = 0; i < kStatSlices; ++i) mean += red[i][col];
  mean = T > 0 ? mean / T : 0.0;
  __syncthreads();
  double q = 0.0;
  if (f < f_dim)
    for (int t = slice; t < T; t += kStatSlices) {
      const double d = (double)x[(long long)t * f_dim] - mean;
      q += d * d;
    }
  red[slice][col] = q;
  __syncthreads();
  if (slice == 0 && f < f_dim) {
    double m2 = 0.0;
    for (int i = 0; i < kStatSlices; ++i) m2 += red[i][col];
    double* o = clip_stats + (long long)b * 4 * f_dim;
    o[c * f_dim + f] = mean;
    o[(2 + c) * f_dim + f] = T > 1 ? m2 / (T - 1) : 0.0;  // torch.std(dim=1) is unbiased
  }
}

int launch_clip_stats(const float* feats, const int32_t* n_frames, int batch, int t_dim, int f_dim, double* clip_stats,
                      cudaStream_t st) {
  if (batch == 0) return AST_OK;
  dim3 grid((unsigned)((f_dim + kStatCols - 1) / kStatCols), 2, (unsigned)batch);
  ProfileSpan span("clip_stats_kernel", st);
  clip_stats_kernel<<<grid, kStatCols * kStatSlices, 0, st>>>(feats, n_frames, t_dim, f_dim, clip_stats);
  AST_LAUNCH_CHECK("clip_stats_kernel");
  return AST_OK;
}

// K6 fused, part 1b: the feature kernels in statistics mode leave per-tile (mean, M2) partials instead of features
// (stft.cu: one per CTA tile of frames; cqt_tc.cu: one per 32-frame quadrant of a 128-frame tile).  One CTA per clip
// merges them in frame order with Chan's formula in float64 and writes the clip's mean and UNBIASED variance
// (compute_separated_stats.py:27-28) in the layout stats_accumulate_kernel adds up.
__global__ void __launch_bounds__(128) stats_finalize_clips_kernel(const float2* __restrict__ part_stft,
                                                                   const float* __restrict__ part_n, int stft_tiles,
                                                                   const float2* __restrict__ part_cqt, int cqt_tiles,
                                                                   const int32_t* __restrict__ lengths, long long max_samples,
                                                                   double* __restrict__ clip_stats) {
  const int b = blockIdx.x;
  const int frames = num_frames(lengths ? lengths[b] : max_samples);
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < 2 * kFTotal; idx += gridDim.y * blockDim.x) {
    const int c = idx / kFTotal, f = idx - c * kFTotal;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    auto merge = [&](double nb, float2 pm) {
      if (nb <= 0.0) return;
      const double nn = n + nb, d = (double)pm.x - mean;
      mean += d * nb / nn;
      m2 += (double)pm.y + d * d * n * nb / nn;
      n = nn;
    };
    if (f < kFStft) {
      for (int i = 0; i < stft_tiles; ++i) {
        const long long tile = (long long)b * stft_tiles + i;
        merge((double)part_n[tile], part_stft[(tile * 2 + c) * kFStft + f]);
      }
    } else {
      const int j = f - kFStft, oct = kOctaves - 1 - j / kBinsPerOctave;
      const int col = j % kBinsPerOctave + c * kBinsPerOctave;
      for (int i = 0; i < cqt_tiles; ++i)
        for (int q = 0; q < 4; ++q) {
          int nb = frames - (i * 128 + q * 32);
          nb = nb < 0 ? 0 : nb > 32 ? 32 : nb;
          merge((double)nb, part_cqt[((((long long)b * kOctaves + oct) * cqt_tiles + i) * 4 + q) * kCqtCols + col]);
        }
    }
    double* o = clip_stats + (long long)b * 4 * kFTotal;
    o[c * kFTotal + f] = mean;
    o[(2 + c) * kFTotal + f] = frames > 1 ? m2 / (frames - 1) : 0.0;   // torch.std(dim=1) is unbiased
  }
}

int launch_stats_finalize_clips(const float2* part_stft, const float* part_n, int stft_tiles, const float2* part_cqt,
                                int cqt_tiles, const int32_t* lengths, long long max_samples, int batch, double* clip_stats,
                                cudaStream_t st) {
  if (batch == 0) return AST_OK;
  ProfileSpan span("stats_finalize_clips_kernel", st);
  // one thread per (clip, channel, bin): each walks its ~ 7 - 28 partials serially (latency bound), so spread them wide
  stats_finalize_clips_kernel<<<dim3((unsigned)batch, (2 * kFTotal + 127) / 128), 128, 0, st>>>(
      part_stft, part_n, stft_tiles, part_cqt, cqt_tiles, lengths, max_samples, clip_stats);
  AST_LAUNCH_CHECK("stats_finalize_clips_kernel");
  return AST_OK;
}

// K6 part 2: add the clips of this batch into the running group sums, in batch order, one writer per
// element (deterministic).  acc[g][0] += clip mean, acc[g][1] += clip variance, counts[g] += 1.
__global__ void stats_accumulate_kernel(const double* __restrict__ clip_stats, const int32_t* __restrict__ group_ids,
                                        int batch, int f_dim, int n_groups, double* __restrict__ acc,
                                        double* __restrict__ counts) {
  const int per_group = 4 * f_dim;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_groups * per_group) {
    const int g = i / per_group, e = i - g * per_group;
    double s = acc[i];
    for (int b = 0; b < batch; ++b) {
      const int gb = group_ids ? group_ids[b] : 0;
      if (gb == g) s += clip_stats[(long long)b * per_group + e];
    }
    acc[i] = s;
  }
  if (i < n_groups) {
    double n = counts[i];
    for (int b = 0; b < batch; ++b) n += ((group_ids ? group_ids[b] : 0) == i) ? 1.0 : 0.0;
    counts[i] = n;
  }
}

int launch_stats_accumulate(const double* clip_stats, const int32_t* group_ids, int batch, int f_dim, int n_groups,
                            double* acc, double* counts, cudaStream_t st) {
  const int total = n_groups * 4 * f_dim;
  stats_accumulate_kernel<<<(total + 127) / 128, 128, 0, st>>>(clip_stats, group_ids, batch, f_dim, n_groups, acc,
                                                               counts);
  AST_LAUNCH_CHECK("stats_accumulate_kernel");
  return AST_OK;
}

}  // namespace ast

using namespace ast;

extern "C" {

int ast_normalize(const float* x, const float* mean, const float* std_, float eps, int32_t n_ch, int32_t n_time,
                  int32_t n_freq, float* out, void* stream) {
  if (!x || !mean || !std_ || !out || n_ch < 0 || n_time < 0 || n_freq <= 0)
    return fail(AST_ERR_INVALID_ARG, "ast_normalize: bad argument");
  const long long total = (long long)n_ch * n_time * n_freq;
  if (total == 0) return AST_OK;
  normalize_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, mean, std_, eps, n_time, n_freq, total, out);
  AST_LAUNCH_CHECK("normalize_kernel");
  return AST_OK;
}

int ast_concat(const float* a, const float* b, int32_t n_ch, int32_t n_time, int32_t f1, int32_t f2, float* out,
               void* stream) {
  if (!a || !b || !out || n_ch < 0 || n_time < 0 || f1 < 0 || f2 < 0) return fail(AST_ERR_INVALID_ARG, "ast_concat: bad argument");
  const long long rows = (long long)n_ch * n_time;
  if (rows * (f1 + f2) == 0) return AST_OK;
  concat_kernel<<<grid_for(rows * (f1 + f2)), 256, 0, (cudaStream_t)stream>>>(a, b, rows, f1, f2, out);
  AST_LAUNCH_CHECK("concat_kernel");
  return AST_OK;
}

int ast_overlap_windows(const float* spec, int32_t n_ch, int32_t n_time, int32_t n_freq, int32_t window_size,
                        int32_t overlap_frames, float* out, int32_t n_sections, void* stream) {
  if (!spec || !out || n_ch <= 0 || n_freq <= 0 || window_size <= 0 || overlap_frames < 0 || overlap_frames >= window_size)
    return fail(AST_ERR_INVALID_ARG, "ast_overlap_windows: bad argument");
  const int expect = num_sections(n_time, window_size, overlap_frames);
  if (expect == 0) return fail(AST_ERR_NO_SECTIONS, "%d frames give no section of %d frames", n_time, window_size);
  if (n_sections != expect) return fail(AST_ERR_SHAPE, "expected %d sections for %d frames, got %d", expect, n_time, n_sections);
  const long long total = (long long)n_sections * n_ch * window_size * n_freq;
  overlap_windows_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(spec, n_ch, n_time, n_freq, window_size,
                                                                           window_size - overlap_frames, n_sections, out);
  AST_LAUNCH_CHECK("overlap_windows_kernel");
  return AST_OK;
}

int ast_sections_merge(const float* sections, int32_t n_sections, int32_t n_ch, int32_t window_size, int32_t n_freq,
                       int32_t overlap_frames, int32_t t_out, float* out, void* stream) {
  if (!sections || !out || n_sections <= 0 || n_ch <= 0 || n_freq <= 0 || window_size <= 0 || overlap_frames < 0 ||
      overlap_frames >= window_size)
    return fail(AST_ERR_INVALID_ARG, "ast_sections_merge: bad argument");
  const int hop = window_size - overlap_frames;
  const int n_time = hop * (n_sections - 1) + window_size;
  if (t_out < 0 || t_out > n_time) return fail(AST_ERR_SHAPE, "t_out %d exceeds the %d merged frames", t_out, n_time);
  const long long total = (long long)n_ch * t_out * n_freq;
  if (total == 0) return AST_OK;
  sections_merge_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(sections, n_sections, n_ch, window_size,
                                                                          n_freq, hop, t_out, out);
  AST_LAUNCH_CHECK("sections_merge_kernel");
  return AST_OK;
}

int ast_stats_finalize(const double* acc, double count, float* mean, float* std_) {
  if (!acc || !mean || !std_ || !(count > 0)) return fail(AST_ERR_INVALID_ARG, "ast_stats_finalize: bad argument");
  for (int i = 0; i < 2 * kFTotal; ++i) {
    mean[i] = (float)(acc[i] / count);
    std_[i] = (float)std::sqrt(acc[2 * kFTotal + i] / count);
  }
  return AST_OK;
}

}  // extern "C"
