// istft.cu - K5: section merge + C2R inverse FFT-1024 + Hann synthesis window + gather overlap-add
// + envelope normalisation + centre trim, in one kernel.  Replaces sections2spectrogram
// (utilityFunctions.py:265-283) and torch.istft as called by inverse_STFT (utilityFunctions.py:62-82;
// the device-aware copy in style_transfer_inference_test.ipynb cell 1) and by
// reconstruct_audio_from_sections (evaluation_reconstruction.py:161-189).
//
// torch.istft semantics: x_t = irfft(X[:, t], n = 1024) (1/N scale, imaginary parts of bins 0 and
// 512 ignored); output sample n in [0, 256 (T - 1)) sits at p = n + 512 of the untrimmed signal and is
//     y[n] = sum_t w[p - 256 t] x_t[p - 256 t] / sum_t w[p - 256 t]^2,
// t over the (at most 4) frames with 0 <= p - 256 t < 1024.  The sum is a GATHER over frames kept in
// shared memory: no atomics, deterministic.
//
// A CTA owns 13 consecutive 256-sample output segments of one clip and therefore needs 16 frames
// (one before, two after).  Two frames share one complex FFT (see fft_core.h).  Shared memory:
// 8 KB twiddles + 4 x 16.1 KB exchange + 64 KB windowed frames = 138 KB.
#include "common.cuh"

namespace ast {

constexpr int kIstftGroups = 4;
constexpr int kIstftThreads = kIstftGroups * kFftThreads;
constexpr int kIstftFrames = 16;
constexpr int kIstftSegs = 13;
constexpr size_t kIstftSmem =
    sizeof(float2) * (kFftN + kIstftGroups * (kBuf1Size + kBuf2Size)) + sizeof(float) * kIstftFrames * kFftN;

struct IstftParams {
  const float* spec;
  int batch, dim1, f_in, layout;
  int window, sec_hop;   // SECTIONS: rows per section, frames between section starts
  int n_frames;          // T' after merge / crop
  long long clip_stride; // floats per clip of spec
  float* out;
  long long out_stride;
  const float2* tw;
  const float* hann_inv_n;
  const float* hann_sq;
};

// complex bin kk of merged frame t (count-normalised average of the sections covering it)
__device__ __forceinline__ float2 load_bin(const IstftParams& p, const float* __restrict__ clip, int t, int kk) {
  if (p.layout == AST_LAYOUT_FLAT) {
    const float* r = clip + (long long)t * p.f_in + kk;
    return make_float2(__ldg(r), __ldg(r + (long long)p.dim1 * p.f_in));
  }
  const long long plane = (long long)p.window * p.f_in;
  int s_hi = t / p.sec_hop;
  if (s_hi > p.dim1 - 1) s_hi = p.dim1 - 1;
  float re = 0.f, im = 0.f, cnt = 0.f;
  for (int s = s_hi; s >= 0 && s >= s_hi - 1; --s) {
    const int tau = t - s * p.sec_hop;
    if (tau >= p.window) break;
    const float* r = clip + ((long long)s * 2 * p.window + tau) * p.f_in + kk;
    re += __ldg(r);
    im += __ldg(r + plane);
    cnt += 1.f;
  }
  cnt = fmaxf(cnt, 1.f);  // count.clamp(min=1.0), utilityFunctions.py:282
  return make_float2(re / cnt, im / cnt);
}

struct IstftEmit {
  float* fa;  // windowed frame A in shared memory (1024 floats)
  float* fb;
  const float* w;
  __device__ __forceinline__ void operator()(int n, float2 z) const {
    const float wn = __ldg(w + n);
    fa[n] = z.x * wn;
    fb[n] = -z.y * wn;
  }
};

__global__ void __launch_bounds__(kIstftThreads, 1) istft_kernel(const IstftParams p) {
  extern __shared__ __align__(16) float2 smem[];
  float2* tw = smem;
  const int group = threadIdx.x >> 6, tid = threadIdx.x & 63;
  float2* buf1 = smem + kFftN + group * (kBuf1Size + kBuf2Size);
  float2* buf2 = buf1 + kBuf1Size;
  float* frames = reinterpret_cast<float*>(smem + kFftN + kIstftGroups * (kBuf1Size + kBuf2Size));
  for (int i = threadIdx.x; i < kFftN; i += kIstftThreads) tw[i] = p.tw[i];

  const int b = blockIdx.y;
  const int seg0 = blockIdx.x * kIstftSegs;
  const int t_first = seg0 - 1;
  const float* __restrict__ clip = p.spec + (long long)b * p.clip_stride;
  __syncthreads();

  for (int round = 0; round < kIstftFrames / 2 / kIstftGroups; ++round) {
    const int pair = round * kIstftGroups + group;
    const int ta = t_first + 2 * pair, tb = ta + 1;
    const bool live_a = ta >= 0 && ta < p.n_frames, live_b = tb >= 0 && tb < p.n_frames;
    float* fa = frames + (2 * pair) * kFftN;
    float* fb = fa + kFftN;
    if (live_a || live_b) {
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int m = 64 * n1 + tid;
        const int kk = m <= 512 ? m : kFftN - m;
        const float2 xa = live_a ? load_bin(p, clip, ta, kk) : make_float2(0.f, 0.f);
        const float2 xb = live_b ? load_bin(p, clip, tb, kk) : make_float2(0.f, 0.f);
        v[n1] = pack_conj_hermitian_pair(m, xa, xb);
      }
      fft1024_stage1(v, tid, tw, buf1);
    }
    __syncthreads();
    if (live_a || live_b) fft1024_stage2(tid, tw, buf1, buf2);
    __syncthreads();
    if (live_a || live_b) {
      IstftEmit emit{fa, fb, p.hann_inv_n};
      fft1024_stage3_complex(tid, buf2, emit);
    }
  }
  __syncthreads();

  // gather overlap-add over the CTA's 13 segments
  const long long out_len = (long long)kHop * (p.n_frames - 1);
  float* __restrict__ y = p.out + (long long)b * p.out_stride;
  for (int idx = threadIdx.x; idx < kIstftSegs * kHop; idx += kIstftThreads) {
    const long long n = (long long)seg0 * kHop + idx;
    if (n >= out_len) break;
    const int pos = (int)n + kNfft / 2;
    int t_hi = pos >> 8;
    if (t_hi > p.n_frames - 1) t_hi = p.n_frames - 1;
    int t_lo = (pos - (kNfft - kHop)) >> 8;  // floor((pos - 768) / 256) == ceil((pos - 1023) / 256)
    if (t_lo < 0) t_lo = 0;
    float acc = 0.f, env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
      const int off = pos - (t << 8);
      acc += frames[(t - t_first) * kFftN + off];
      env += __ldg(p.hann_sq + off);
    }
    y[n] = acc / env;
  }
}

int istft_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIstftSmem));
  return AST_OK;
}

int launch_istft(const ast_plan* plan, const float* spec, int batch, int dim1, int f_in, int layout, int window,
                 int overlap, int n_frames, float* wave_out, long long out_stride, cudaStream_t st) {
  if (n_frames < 2 || batch == 0) return AST_OK;  // 256 * (T - 1) = 0 samples
  IstftParams p;
  p.spec = spec;
  p.batch = batch;
  p.dim1 = dim1;
  p.f_in = f_in;
  p.layout = layout;
  p.window = window;
  p.sec_hop = window - overlap;
  p.n_frames = n_frames;
  p.clip_stride = layout == AST_LAYOUT_FLAT ? 2LL * dim1 * f_in : 2LL * dim1 * window * f_in;
  p.out = wave_out;
  p.out_stride = out_stride;
  p.tw = plan->d_tw;
  p.hann_inv_n = plan->d_hann_inv_n;
  p.hann_sq = plan->d_hann_sq;
  const int segs = n_frames - 1;
  dim3 grid((unsigned)((segs + kIstftSegs - 1) / kIstftSegs), (unsigned)batch);
  ProfileSpan span("istft_kernel", st);
  istft_kernel<<<grid, kIstftThreads, kIstftSmem, st>>>(p);
  AST_LAUNCH_CHECK("istft_kernel");
  return AST_OK;
}

}  // namespace ast
