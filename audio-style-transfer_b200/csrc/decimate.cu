// decimate.cu - K2: the 2:1 decimation cascade librosa.cqt runs between octaves
// (audio.resample(my_y, orig_sr=2, target_sr=1, res_type="soxr_hq", scale=True), called from
// librosa.core.constantq.vqt, reached through utilityFunctions.py:52).
//
//   y_out[j] = sum_{k=0}^{384} g[k] * y_in[2 j + k - 192],  g = sqrt(2) * h,  zero extension,
//   len_out = ceil(len_in / 2).
//
// FMA-pipe formulation: the input tile is split into its even / odd polyphase components in
// shared memory (padded one word per eight so that lanes 8 samples apart hit distinct banks);
// each thread produces 8 consecutive outputs with a sliding 8-register window, so one LDS feeds
// 8 FMAs.  Taps live in __constant__ memory (uniform index -> constant-cache broadcast).
#include "common.cuh"

namespace ast {

__constant__ float c_dec_taps[kDecTaps + 7];

constexpr int kDecThreads = 128;
constexpr int kDecPerThread = 8;
constexpr int kDecTile = kDecThreads * kDecPerThread;   // 1024 outputs per CTA
constexpr int kDecPhaseHalf = kDecHalf / 2;             // 96: out[j] needs xe/xo[j - 96 .. j + 96]
constexpr int kDecPhaseLen = kDecTile + 2 * kDecPhaseHalf;  // 1216 entries per phase
constexpr int kDecPhasePad = kDecPhaseLen + kDecPhaseLen / 8 + 8;

__device__ __forceinline__ int dec_pad(int i) { return i + (i >> 3); }

struct DecimateParams {
  const float* in;        // clip b at in + b * in_stride
  long long in_stride;
  float* out;             // clip b at out + b * out_stride
  long long out_stride;
  const int32_t* lengths; // original clip lengths (samples at octave 0) or nullptr
  long long max_samples;
  int in_octave;          // input is the signal decimated in_octave times
};

__global__ void __launch_bounds__(kDecThreads) decimate2_kernel(const DecimateParams p) {
  __shared__ float xe[kDecPhasePad];
  __shared__ float xo[kDecPhasePad];
  const int b = blockIdx.y;
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  const long long len_in = (len0 + (1LL << p.in_octave) - 1) >> p.in_octave;
  const long long len_out = (len_in + 1) >> 1;
  const long long j_blk = (long long)blockIdx.x * kDecTile;
  if (j_blk >= len_out) return;
  const float* __restrict__ x = p.in + (long long)b * p.in_stride;

  // stage: phase index r in [0, kDecPhaseLen) <-> input samples 2 (j_blk - 96 + r) and + 1
  const long long i0 = 2 * (j_blk - kDecPhaseHalf);
  for (int r = threadIdx.x; r < kDecPhaseLen; r += kDecThreads) {
    const long long i = i0 + 2LL * r;
    float e = 0.f, o = 0.f;
    if (i >= 0 && i + 1 < len_in) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(x + i));  // i even, rows 8-byte aligned
      e = v.x;
      o = v.y;
    } else {
      if (i >= 0 && i < len_in) e = __ldg(x + i);
      if (i + 1 >= 0 && i + 1 < len_in) o = __ldg(x + i + 1);
    }
    xe[dec_pad(r)] = e;
    xo[dec_pad(r)] = o;
  }
  __syncthreads();

  // out[j] = sum_u g[2u] xe[j + u - 96] + sum_u g[2u + 1] xo[j + u - 96];  local r = (j - j_blk) + u
  const int r0 = threadIdx.x * kDecPerThread;
  float acc[kDecPerThread];
#pragma unroll
  for (int m = 0; m < kDecPerThread; ++m) acc[m] = 0.f;

  // even phase: u = 0 .. 192 (193 taps), odd phase: u = 0 .. 191 (192 taps)
#pragma unroll
  for (int phase = 0; phase < 2; ++phase) {
    const float* __restrict__ xs = phase == 0 ? xe : xo;
    const int n_taps = phase == 0 ? kDecHalf + 1 : kDecHalf;
    float win[8];
#pragma unroll
    for (int m = 0; m < 7; ++m) win[m] = xs[dec_pad(r0 + m)];
    int u = 0;
    for (; u + 8 <= n_taps; u += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        win[(i + 7) & 7] = xs[dec_pad(r0 + u + i + 7)];
        const float g = c_dec_taps[2 * (u + i) + phase];
#pragma unroll
        for (int m = 0; m < kDecPerThread; ++m) acc[m] = fmaf(g, win[(i + m) & 7], acc[m]);
      }
    }
    // remainder (even phase: one tap, u = 192; odd phase: none)
    for (; u < n_taps; ++u) {
      const float g = c_dec_taps[2 * u + phase];
#pragma unroll
      for (int m = 0; m < kDecPerThread; ++m) acc[m] = fmaf(g, xs[dec_pad(r0 + u + m)], acc[m]);
    }
  }

  float* __restrict__ y = p.out + (long long)b * p.out_stride;
  const long long j0 = j_blk + r0;
  if (j0 + kDecPerThread <= len_out) {
    reinterpret_cast<float4*>(y + j0)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(y + j0)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
#pragma unroll
    for (int m = 0; m < kDecPerThread; ++m)
      if (j0 + m < len_out) y[j0 + m] = acc[m];
  }
}


static int g_use_tc_decimator = 1;
void set_tc_decimator(int on) { g_use_tc_decimator = on; }
bool use_tc_decimator() { return g_use_tc_decimator != 0; }

int upload_decimator_taps(const float* taps_scaled) {
  float padded[kDecTaps + 7] = {0};
  for (int i = 0; i < kDecTaps; ++i) padded[i] = taps_scaled[i];
  AST_CUDA_TRY(cudaMemcpyToSymbol(c_dec_taps, padded, sizeof(padded)));
  return AST_OK;
}

// Runs the six stages: octave buffer i (1..6) of clip b lives at ws + b * ws_clip_stride + octave_offset(i).
int launch_decimate_cascade(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch,
                            long long max_samples, long long wave_stride, float* ws, long long ws_clip_stride,
                            int* flags, cudaStream_t st, bool flags_zeroed) {
  if ((batch > 1 && (wave_stride & 1)) || (reinterpret_cast<uintptr_t>(wave) & 7))
    return fail(AST_ERR_INVALID_ARG, "wave rows must be 8-byte aligned (even stride) for the decimator");
  if (g_use_tc_decimator)
    return launch_decimate_cascade_tc(plan, wave, lengths, batch, max_samples, wave_stride, ws, ws_clip_stride, flags, st, flags_zeroed);
  for (int i = 0; i < kOctaves - 1; ++i) {
    const float* in = i == 0 ? wave : ws + octave_offset(max_samples, i);
    const long long in_stride = i == 0 ? wave_stride : ws_clip_stride;
    float* out = ws + octave_offset(max_samples, i + 1);
    const long long len_out = octave_len(max_samples, i + 1);
    {
      DecimateParams p;
      p.in = in;
      p.in_stride = in_stride;
      p.out = out;
      p.out_stride = ws_clip_stride;
      p.lengths = lengths;
      p.max_samples = max_samples;
      p.in_octave = i;
      dim3 grid((unsigned)((len_out + kDecTile - 1) / kDecTile), (unsigned)batch);
      ProfileSpan span("decimate2_kernel", st);
      decimate2_kernel<<<grid, kDecThreads, 0, st>>>(p);
      AST_LAUNCH_CHECK("decimate2_kernel");
    }
  }
  return AST_OK;
}

}  // namespace ast
