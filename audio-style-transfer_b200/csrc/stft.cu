// stft.cu - K1: framing + reflect padding + Hann window + real FFT-1024 of two frames per complex
// transform, with the normalise / real-imag plane split / section scatter epilogue fused into the
// stores.  Replaces torch.stft (utilityFunctions.py:26-28), the real/imag stacking (:31-35),
// dataloader.normalize (dataloader.py:9-13) and get_overlap_windows (utilityFunctions.py:240-263)
// for the STFT columns of the feature tensor.
//
// One group of 64 threads owns one frame pair; a 256-thread CTA runs four groups in lock step over
// a grid-stride list of (clip, frame pair) items.  Shared memory per CTA: 8 KB twiddles +
// 4 x (8320 B + 8192 B) exchange buffers = 74 240 B -> 3 CTAs / SM.
#include "common.cuh"

namespace ast {

constexpr int kStftGroups = 4;
constexpr int kStftThreads = kStftGroups * kFftThreads;
constexpr size_t kStftSmem = sizeof(float2) * (kFftN + kStftGroups * (kBuf1Size + kBuf2Size));

struct StftParams {
  const float* wave;
  const int32_t* lengths;
  long long max_samples, wave_stride;
  int batch;
  int slots;            // frame slots per clip (rows of the output the kernel must cover)
  int pairs_per_clip;   // ceil(slots / 2)
  long long items;      // batch * pairs_per_clip
  const float2* tw;
  const float* hann;
  int overlap;
  OutSpec out;
};

struct StftEmit {
  RowDest da, db;          // destinations of frame A / B
  const float2* st0;       // stats row of channel 0 (mean, rstd) or nullptr
  const float2* st1;       // stats row of channel 1
  __device__ __forceinline__ void operator()(int k, float are, float aim, float bre, float bim) const {
    if (st0) {
      const float2 m0 = __ldg(st0 + k), m1 = __ldg(st1 + k);
      are = (are - m0.x) * m0.y;
      bre = (bre - m0.x) * m0.y;
      aim = (aim - m1.x) * m1.y;
      bim = (bim - m1.x) * m1.y;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i < da.n) {
        da.row[i][k] = da.live[i] ? are : 0.f;
        da.row[i][da.plane + k] = da.live[i] ? aim : 0.f;
      }
      if (i < db.n) {
        db.row[i][k] = db.live[i] ? bre : 0.f;
        db.row[i][db.plane + k] = db.live[i] ? bim : 0.f;
      }
    }
  }
};

__device__ __forceinline__ float load_reflect(const float* __restrict__ x, long long i, long long len) {
  // torch.stft(center=True, pad_mode="reflect"): one reflection suffices because len > n_fft / 2
  if (i < 0) i = -i;
  if (i >= len) i = 2 * (len - 1) - i;
  return __ldg(x + i);
}

__global__ void __launch_bounds__(kStftThreads, 2) stft_kernel(const StftParams p) {
  extern __shared__ __align__(16) float2 smem[];
  float2* tw = smem;
  const int group = threadIdx.x >> 6, tid = threadIdx.x & 63;
  float2* buf1 = smem + kFftN + group * (kBuf1Size + kBuf2Size);
  float2* buf2 = buf1 + kBuf1Size;
  for (int i = threadIdx.x; i < kFftN; i += kStftThreads) tw[i] = p.tw[i];
  float win[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) win[n1] = __ldg(p.hann + 64 * n1 + tid);
  __syncthreads();

  const long long stride = (long long)gridDim.x * kStftGroups;
  const long long rounds = (p.items + stride - 1) / stride;
  for (long long r = 0; r < rounds; ++r) {
    const long long item = r * stride + (long long)blockIdx.x * kStftGroups + group;
    const bool active = item < p.items;
    int b = 0, ta = 0, frames_b = 0, sections_b = 0;
    bool any_live = false;
    if (active) {
      b = (int)(item / p.pairs_per_clip);
      ta = 2 * (int)(item % p.pairs_per_clip);
      const long long len = p.lengths ? p.lengths[b] : p.max_samples;
      frames_b = num_frames(len);
      sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
      any_live = ta < frames_b;
      if (any_live) {
        const float* x = p.wave + (long long)b * p.wave_stride;
        const bool live_b = ta + 1 < frames_b;
        const long long base = (long long)ta * kHop - kNfft / 2 + tid;
        float2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const long long i = base + 64 * n1;
          const float xa = load_reflect(x, i, len);
          const float xb = live_b ? load_reflect(x, i + kHop, len) : 0.f;
          v[n1] = make_float2(xa * win[n1], xb * win[n1]);
        }
        fft1024_stage1(v, tid, tw, buf1);
      }
    }
    __syncthreads();
    if (any_live) fft1024_stage2(tid, tw, buf1, buf2);
    __syncthreads();
    if (active) {
      StftEmit emit;
      emit.da = row_dest(p.out, b, ta, frames_b, sections_b);
      if (ta + 1 < p.slots) {
        emit.db = row_dest(p.out, b, ta + 1, frames_b, sections_b);
      } else {
        emit.db.n = 0;
        emit.db.plane = 0;
      }
      emit.st0 = nullptr;
      emit.st1 = nullptr;
      if (p.out.stats) {
        emit.st0 = p.out.stats + (long long)b * p.out.stats_clip_stride + p.out.stats_off;
        emit.st1 = emit.st0 + p.out.f_stats;
      }
      if (any_live) {
        fft1024_stage3_real_pair(tid, buf2, emit);
      } else {
        // both frames lie past the clip: the rows exist in the output and must be zeros
        emit.da.live[0] = emit.da.live[1] = emit.db.live[0] = emit.db.live[1] = false;
        emit.st0 = nullptr;
        for (int k = tid; k < kFStft; k += kFftThreads) emit(k, 0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

static int g_stft_ctas_per_sm = 2;

int stft_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftSmem));
  int n = 0;
  AST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_kernel, kStftThreads, kStftSmem));
  g_stft_ctas_per_sm = n > 0 ? n : 1;
  return AST_OK;
}

int launch_stft(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
                long long wave_stride, const OutSpec& out, cudaStream_t st) {
  StftParams p;
  p.wave = wave;
  p.lengths = lengths;
  p.max_samples = max_samples;
  p.wave_stride = wave_stride;
  p.batch = batch;
  p.slots = frame_slots(out.layout, out.dim1, out.window, out.step);
  p.pairs_per_clip = (p.slots + 1) / 2;
  p.items = (long long)batch * p.pairs_per_clip;
  p.tw = plan->d_tw;
  p.hann = plan->d_hann;
  p.overlap = out.window - out.step;
  p.out = out;
  if (p.items == 0) return AST_OK;
  long long blocks = (p.items + kStftGroups - 1) / kStftGroups;
  const long long cap = (long long)plan->sm_count * g_stft_ctas_per_sm;  // persistent: every CTA resident
  if (blocks > cap) blocks = cap;
  ProfileSpan span("stft_kernel", st);
  stft_kernel<<<(unsigned)blocks, kStftThreads, kStftSmem, st>>>(p);
  AST_LAUNCH_CHECK("stft_kernel");
  return AST_OK;
}

}  // namespace ast
