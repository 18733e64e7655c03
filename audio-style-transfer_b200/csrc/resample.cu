// resample.cu - the device part of load_audio (utilityFunctions.py:105-122), the step right before the hot path
// (SURVEY.md 8f-1): zero-pad / cut every clip to cut_samples, resample orig_sr -> new_sr the way
// torchaudio.functional.resample does with its defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99:
// a polyphase FIR, y[n * new + i] = sum_k K[i][k] xpad[n * orig + k], xpad = x shifted by `width` zeros), and
// average the two channels of a stereo file (torch.mean(dim=0, keepdim=True), applied AFTER the resample).
//
// Two kernels: the dataset's case 44.1 kHz -> 22.05 kHz (orig : new = 2 : 1, 28 taps) keeps the even / odd
// polyphase components of a tile in shared memory and slides an 8-register window per thread, so one LDS feeds
// 8 FMAs and the kernel is HBM-bound (4.4 MB per 10 s stereo clip); every other ratio takes a plain
// one-output-per-thread loop over its phase's taps.
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

struct ast_resampler {
  int orig, neu, width, n_taps;  // reduced rates, half width, taps per phase = 2 * width + orig
  int device;
  float* d_taps;                 // [neu][n_taps]
};

namespace ast {

static int gcd_int(int a, int b) {
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// torchaudio.functional._get_sinc_resample_kernel with dtype=None: float64 arithmetic, float32 result
static void host_resample_taps(int orig, int neu, int width, std::vector<float>& taps) {
  const double kPi = 3.14159265358979323846;
  const double lowpass = 6.0, rolloff = 0.99;
  const double base = (double)(orig < neu ? orig : neu) * rolloff;
  const int n_taps = 2 * width + orig;
  taps.assign((size_t)neu * n_taps, 0.f);
  for (int i = 0; i < neu; ++i)
    for (int k = 0; k < n_taps; ++k) {
      double t = (-(double)i / neu + (double)(k - width) / orig) * base;
      if (t < -lowpass) t = -lowpass;
      if (t > lowpass) t = lowpass;
      const double c = std::cos(t * kPi / lowpass / 2.0);
      const double window = c * c;
      t *= kPi;
      const double sinc = t == 0.0 ? 1.0 : std::sin(t) / t;
      taps[(size_t)i * n_taps + k] = (float)(sinc * window * (base / orig));
    }
}

static int resample_geometry(int orig_sr, int new_sr, int& orig, int& neu, int& width) {
  if (orig_sr <= 0 || new_sr <= 0) return fail(AST_ERR_INVALID_ARG, "sample rates must be positive");
  const int g = gcd_int(orig_sr, new_sr);
  orig = orig_sr / g;
  neu = new_sr / g;
  const double base = (double)(orig < neu ? orig : neu) * 0.99;
  width = (int)std::ceil(6.0 * orig / base);
  return AST_OK;
}

constexpr int kRsThreads = 128;
constexpr int kRsPerThread = 8;
constexpr int kRsTile = kRsThreads * kRsPerThread;  // 1024 outputs per CTA
constexpr int kRsHalfTaps = 14;                     // 28 taps = 14 on the even + 14 on the odd input samples
constexpr int kRsPhaseLen = kRsTile + kRsHalfTaps;  // entries per polyphase component (needs tile + 13)
constexpr int kRsPhasePad = kRsPhaseLen + kRsPhaseLen / 8 + 8;

struct ResampleParams {
  const float* wave;
  const int32_t* lengths_in;
  long long in_stride;   // floats between channels (clip b, channel c at wave + (b * C + c) * in_stride)
  long long cut_samples;
  float* out;
  long long out_stride, len_out;
  const float* taps;
  int orig, neu, width, n_taps, channels;
};

__device__ __forceinline__ int rs_pad(int i) { return i + (i >> 3); }

// orig : new = 2 : 1, width 13:  y[m] = sum_{k < 28} K[k] x[2 m - 13 + k]
//   odd k = 2 u + 1 -> even sample x[2 (m - 6 + u)];  even k = 2 u -> odd sample x[2 (m - 7 + u) + 1]
template <int C>
__global__ void __launch_bounds__(kRsThreads) resample_2to1_kernel(const ResampleParams p) {
  __shared__ float xe[C][kRsPhasePad];
  __shared__ float xo[C][kRsPhasePad];
  __shared__ float taps[2 * kRsHalfTaps];
  const int b = blockIdx.y;
  const long long m_blk = (long long)blockIdx.x * kRsTile;
  if (m_blk >= p.len_out) return;
  long long valid = p.lengths_in ? p.lengths_in[b] : p.in_stride;
  if (valid > p.cut_samples) valid = p.cut_samples;
  if (threadIdx.x < 2 * kRsHalfTaps) taps[threadIdx.x] = __ldg(p.taps + threadIdx.x);
  // phase entry r <-> input samples 2 (m_blk - 7 + r) and + 1
  const long long i0 = 2 * (m_blk - 7);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* __restrict__ x = p.wave + ((long long)b * C + c) * p.in_stride;
    for (int r = threadIdx.x; r < kRsPhaseLen; r += kRsThreads) {
      const long long i = i0 + 2LL * r;
      float e = 0.f, o = 0.f;
      if (i >= 0 && i < valid) e = __ldg(x + i);
      if (i + 1 >= 0 && i + 1 < valid) o = __ldg(x + i + 1);
      xe[c][rs_pad(r)] = e;
      xo[c][rs_pad(r)] = o;
    }
  }
  __syncthreads();

  const int r0 = threadIdx.x * kRsPerThread;
  float y[kRsPerThread];
#pragma unroll
  for (int m = 0; m < kRsPerThread; ++m) y[m] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float acc[kRsPerThread];
#pragma unroll
    for (int m = 0; m < kRsPerThread; ++m) acc[m] = 0.f;
    // Output m (local index r0 + j) reads xo[r0 + j + u] (tap 2 u) and xe[r0 + j + 1 + u] (tap 2 u + 1), u = 0..13,
    // accumulated in tap order k = 0, 1, 2, ... like a direct convolution.
    float wo[8], we[8];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      wo[j] = xo[c][rs_pad(r0 + j)];
      we[j] = xe[c][rs_pad(r0 + 1 + j)];
    }
#pragma unroll
    for (int u = 0; u < kRsHalfTaps; ++u) {
      wo[(u + 7) & 7] = xo[c][rs_pad(r0 + u + 7)];
      we[(u + 7) & 7] = xe[c][rs_pad(r0 + 1 + u + 7)];
      const float g0 = taps[2 * u], g1 = taps[2 * u + 1];
#pragma unroll
      for (int m = 0; m < kRsPerThread; ++m) {
        acc[m] = fmaf(g0, wo[(u + m) & 7], acc[m]);
        acc[m] = fmaf(g1, we[(u + m) & 7], acc[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < kRsPerThread; ++m) y[m] = c == 0 ? acc[m] : (y[m] + acc[m]) * 0.5f;  // torch.mean over 2 channels
  }
  float* __restrict__ out = p.out + (long long)b * p.out_stride;
  const long long m0 = m_blk + r0;
#pragma unroll
  for (int m = 0; m < kRsPerThread; ++m)
    if (m0 + m < p.len_out) out[m0 + m] = y[m];
}

// any ratio: one output per thread
__global__ void __launch_bounds__(256) resample_generic_kernel(const ResampleParams p) {
  const int b = blockIdx.y;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.len_out) return;
  long long valid = p.lengths_in ? p.lengths_in[b] : p.in_stride;
  if (valid > p.cut_samples) valid = p.cut_samples;
  const long long n = m / p.neu;
  const int i = (int)(m - n * p.neu);
  const float* __restrict__ taps = p.taps + (long long)i * p.n_taps;
  const long long j0 = n * p.orig - p.width;
  float y = 0.f;
  for (int c = 0; c < p.channels; ++c) {
    const float* __restrict__ x = p.wave + ((long long)b * p.channels + c) * p.in_stride;
    float acc = 0.f;
    for (int k = 0; k < p.n_taps; ++k) {
      const long long j = j0 + k;
      if (j >= 0 && j < valid) acc = fmaf(__ldg(taps + k), __ldg(x + j), acc);
    }
    y = c == 0 ? acc : (y + acc) * 0.5f;
  }
  p.out[(long long)b * p.out_stride + m] = y;
}

// orig == new: pad / cut + channel mean only (the reference skips the resample, utilityFunctions.py:116)
__global__ void __launch_bounds__(256) pad_cut_mix_kernel(const ResampleParams p) {
  const int b = blockIdx.y;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.len_out) return;
  long long valid = p.lengths_in ? p.lengths_in[b] : p.in_stride;
  if (valid > p.cut_samples) valid = p.cut_samples;
  float y = 0.f;
  for (int c = 0; c < p.channels; ++c) {
    const float v = m < valid ? __ldg(p.wave + ((long long)b * p.channels + c) * p.in_stride + m) : 0.f;
    y = c == 0 ? v : (y + v) * 0.5f;
  }
  p.out[(long long)b * p.out_stride + m] = y;
}

}  // namespace ast

using namespace ast;

extern "C" {

int ast_resample_geometry(int32_t orig_sr, int32_t new_sr, int32_t* orig_reduced, int32_t* new_reduced, int32_t* width) {
  int o, n, w;
  const int rc = resample_geometry(orig_sr, new_sr, o, n, w);
  if (rc != AST_OK) return rc;
  if (orig_reduced) *orig_reduced = o;
  if (new_reduced) *new_reduced = n;
  if (width) *width = w;
  return AST_OK;
}

int64_t ast_resample_length(int64_t n_in, int32_t orig_sr, int32_t new_sr) {
  int o, n, w;
  if (n_in < 0 || resample_geometry(orig_sr, new_sr, o, n, w) != AST_OK) return -1;
  return (n_in * n + o - 1) / o;  // ceil(new * length / orig), torchaudio _apply_sinc_resample_kernel
}

int ast_host_resample_taps(int32_t orig_sr, int32_t new_sr, float* taps, int32_t capacity) {
  int o, n, w;
  const int rc = resample_geometry(orig_sr, new_sr, o, n, w);
  if (rc != AST_OK) return rc;
  const int need = n * (2 * w + o);
  if (!taps || capacity < need) return fail(AST_ERR_INVALID_ARG, "tap buffer too small: need %d floats", need);
  std::vector<float> t;
  host_resample_taps(o, n, w, t);
  for (int i = 0; i < need; ++i) taps[i] = t[i];
  return AST_OK;
}

int ast_resampler_create(int32_t orig_sr, int32_t new_sr, int32_t device, ast_resampler** out) {
  if (!out) return fail(AST_ERR_INVALID_ARG, "null output pointer");
  *out = nullptr;
  int o, n, w;
  const int rc = resample_geometry(orig_sr, new_sr, o, n, w);
  if (rc != AST_OK) return rc;
  AST_CUDA_TRY(cudaSetDevice(device));
  ast_resampler* r = new ast_resampler{o, n, w, 2 * w + o, device, nullptr};
  std::vector<float> t;
  host_resample_taps(o, n, w, t);
  cudaError_t e = cudaMalloc((void**)&r->d_taps, sizeof(float) * t.size());
  if (e == cudaSuccess) e = cudaMemcpy(r->d_taps, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(r->d_taps);
    delete r;
    return fail(AST_ERR_CUDA, "resampler upload failed: %s", cudaGetErrorString(e));
  }
  *out = r;
  return AST_OK;
}

int ast_resampler_destroy(ast_resampler* r) {
  if (!r) return AST_OK;
  cudaFree(r->d_taps);
  delete r;
  return AST_OK;
}

int ast_load_audio_forward(const ast_resampler* r, const float* wave, const int32_t* lengths_in, int32_t batch,
                           int32_t channels, int64_t in_stride, int64_t cut_samples, float* out, int64_t out_stride,
                           void* stream) {
  if (!r || !wave || !out) return fail(AST_ERR_INVALID_ARG, "null pointer");
  if (batch < 0 || in_stride < 0 || cut_samples < 0) return fail(AST_ERR_INVALID_ARG, "negative size");
  if (channels != 1 && channels != 2)
    return fail(AST_ERR_INVALID_ARG, "load_audio mixes down stereo only (utilityFunctions.py:119); got %d channels", channels);
  const long long len_out = (cut_samples * r->neu + r->orig - 1) / r->orig;
  if (out_stride < len_out) return fail(AST_ERR_SHAPE, "output rows hold %lld samples, need %lld", (long long)out_stride, len_out);
  if (batch == 0 || len_out == 0) return AST_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ResampleParams p;
  p.wave = wave;
  p.lengths_in = lengths_in;
  p.in_stride = in_stride;
  p.cut_samples = cut_samples;
  p.out = out;
  p.out_stride = out_stride;
  p.len_out = len_out;
  p.taps = r->d_taps;
  p.orig = r->orig;
  p.neu = r->neu;
  p.width = r->width;
  p.n_taps = r->n_taps;
  p.channels = channels;
  if (r->orig == r->neu) {
    dim3 grid((unsigned)((len_out + 255) / 256), (unsigned)batch);
    ProfileSpan span("pad_cut_mix_kernel", st);
    pad_cut_mix_kernel<<<grid, 256, 0, st>>>(p);
    AST_LAUNCH_CHECK("pad_cut_mix_kernel");
  } else if (r->orig == 2 && r->neu == 1 && r->width == 13) {
    dim3 grid((unsigned)((len_out + kRsTile - 1) / kRsTile), (unsigned)batch);
    ProfileSpan span("resample_2to1_kernel", st);
    if (channels == 2)
      resample_2to1_kernel<2><<<grid, kRsThreads, 0, st>>>(p);
    else
      resample_2to1_kernel<1><<<grid, kRsThreads, 0, st>>>(p);
    AST_LAUNCH_CHECK("resample_2to1_kernel");
  } else {
    dim3 grid((unsigned)((len_out + 255) / 256), (unsigned)batch);
    ProfileSpan span("resample_generic_kernel", st);
    resample_generic_kernel<<<grid, 256, 0, st>>>(p);
    AST_LAUNCH_CHECK("resample_generic_kernel");
  }
  return AST_OK;
}

}  // extern "C"
