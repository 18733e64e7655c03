// stft.cu - K1: framing + reflect padding + Hann window + real FFT-1024 of two frames per complex
// transform, with the normalise / real-imag plane split / section scatter epilogue fused into the
// stores.  Replaces torch.stft (utilityFunctions.py:26-28), the real/imag stacking (:31-35),
// dataloader.normalize (dataloader.py:9-13) and get_overlap_windows (utilityFunctions.py:240-263)
// for the STFT columns of the feature tensor.
//
// One group of 64 threads owns one frame pair (frames 2p and 2p+1 share one complex FFT, and
// share 3/4 of their samples: 20 strided loads feed both).  A 128-thread CTA runs two groups in
// lock step over `iters` consecutive groups of pairs of ONE clip (grid = pair tiles x clips), so all
// per-clip bookkeeping is CTA-uniform.  Interior frames take a check-free load path; rows whose
// destinations are all live take a select-free store path.  The (mean, rstd) pairs of the bins a thread
// emits are fetched before the barrier that precedes stage 3, so their L2 round trip overlaps it.
// Shared memory per CTA: 10 KB twiddle tables + 2 x (8320 B + 8192 B) exchange buffers = 43 264 B;
// four CTAs (16 warps, 128 registers per thread, no spills) per SM.  The shape is the best of a sweep, ms per
// 64 clips (groups per CTA x CTAs per SM): 4x3 at 80 registers without the early statistics fetch 0.149 (with it:
// spills, 0.176), 4x2 0.157, 5x2 0.159, 3x3 0.156, 1x8 0.151, 2x5 at 96 registers 0.144, 2x4 at 128 registers 0.140
// (0.143 without the early fetch; prefetching the next pair's samples on top spills again, 0.151).
#include <cstdlib>

#include "common.cuh"

namespace ast {

#ifndef AST_STFT_GROUPS
#define AST_STFT_GROUPS 2   // frame-pair groups (of 64 threads) per CTA
#endif
constexpr int kStftGroups = AST_STFT_GROUPS;
constexpr int kStftThreads = kStftGroups * kFftThreads;
constexpr size_t kStftSmem = sizeof(float2) * (kTw1Size + kTw2Size + kStftGroups * (kBuf1Size + kBuf2Size));

struct StftParams {
  const float* wave;
  const int32_t* lengths;
  long long max_samples, wave_stride;
  int slots;            // frame slots per clip (rows of the output the kernel must cover)
  int pairs_per_clip;   // ceil(slots / 2)
  int iters;            // pair groups per CTA
  const float2* t1;
  const float2* t2;
  const float* hann;
  int overlap;
  int debug;            // diagnostic (AST_STFT_DEBUG): 1 stores suppressed (only k == 1000000 would store)
  unsigned int* tail_counter;  // chained feature call: finished-CTA counter (zeroed by the prologue kernel), else nullptr
  int pad_zero;         // 0: reflect padding (torch.stft, get_STFT); 1: zero padding (librosa.stft default, mse_spectrogram)
  OutSpec out;
};

// Destination of the two frames of a pair: up to two rows each (a frame inside the overlap of two
// sections is stored twice).  kAllLive: every present row receives data (no zero padding involved).
// compile-time switches kept for the shape sweep (scratch/run_variants.sh)
#ifndef AST_STFT_PRELOAD
#define AST_STFT_PRELOAD 1  // 1: per-bin statistics fetched ahead of the stage-3 barrier (36 live registers)
#endif
#ifndef AST_STFT_CTAS
#define AST_STFT_CTAS 4     // resident CTAs per SM the register allocation is sized for
#endif
#ifndef AST_STFT_PROLOGUE
#define AST_STFT_PROLOGUE 1 // 1: 16-byte twiddle copy with all loads in flight + first pair's samples prefetched before it
#endif
#ifndef AST_STFT_PREFETCH
#define AST_STFT_PREFETCH 2 // 1: the next interior pair's 20 samples are loaded right after stage 1 (20 more live registers, spills)
                            // 2: the 512 samples of the next pair that this iteration has not touched are prefetched into the L1
#endif

template <bool kAllLive>
struct StftEmit {
  float* a0;
  float* a1;
  float* b0;
  float* b1;            // channel-0 row bases (nullptr: absent)
  bool la0, la1, lb0, lb1;
  long long plane;      // floats from the channel-0 row to the channel-1 row
#if AST_STFT_PRELOAD
  // (mean, rstd) of the bins this thread emits, in emit order (fft1024_stage3_bin), fetched before the barrier that
  // precedes stage 3 so that their L2 round trip overlaps it
  const float2 (&m0)[9];
  const float2 (&m1)[9];
  bool has_stats;
  int idx;              // emit() calls so far: compile-time after inlining, the arrays stay in registers
#else
  const float2* st0;    // (mean, rstd) of channel 0, or nullptr
  const float2* st1;
#endif
  // values arrive WITHOUT the factor 1/2 of the Hermitian separation; x = 0.5 s is exact, so
  // fmaf(s, 0.5, -mean) rounds once, exactly like the reference's (x - mean).
  __device__ __forceinline__ void operator()(int k, float are, float aim, float bre, float bim) {
#if AST_STFT_PRELOAD
    if (has_stats) {
      const float2 s0 = m0[idx], s1 = m1[idx];
#else
    if (st0) {
      const float2 s0 = __ldg(st0 + k), s1 = __ldg(st1 + k);
#endif
      are = fmaf(are, 0.5f, -s0.x) * s0.y;
      bre = fmaf(bre, 0.5f, -s0.x) * s0.y;
      aim = fmaf(aim, 0.5f, -s1.x) * s1.y;
      bim = fmaf(bim, 0.5f, -s1.x) * s1.y;
    } else {
      are *= 0.5f, aim *= 0.5f, bre *= 0.5f, bim *= 0.5f;
    }
#if AST_STFT_PRELOAD
    ++idx;
#endif
    if (kAllLive) {
      a0[k] = are;
      a0[plane + k] = aim;
      if (a1) a1[k] = are, a1[plane + k] = aim;
      if (b0) b0[k] = bre, b0[plane + k] = bim;
      if (b1) b1[k] = bre, b1[plane + k] = bim;
    } else {
      if (a0) a0[k] = la0 ? are : 0.f, a0[plane + k] = la0 ? aim : 0.f;
      if (a1) a1[k] = la1 ? are : 0.f, a1[plane + k] = la1 ? aim : 0.f;
      if (b0) b0[k] = lb0 ? bre : 0.f, b0[plane + k] = lb0 ? bim : 0.f;
      if (b1) b1[k] = lb1 ? bre : 0.f, b1[plane + k] = lb1 ? bim : 0.f;
    }
  }
};

__device__ __forceinline__ float load_reflect(const float* __restrict__ x, int i, int len, int pad_zero) {
  if (pad_zero) return (i >= 0 && i < len) ? __ldg(x + i) : 0.f;  // librosa.stft(center=True, pad_mode="constant")
  // torch.stft(center=True, pad_mode="reflect"): one reflection suffices because len > n_fft / 2
  if (i < 0) i = -i;
  if (i >= len) i = 2 * (len - 1) - i;
  return __ldg(x + i);
}

// rows of frame slot t (SECTIONS: the section starting at or before t, and the previous one if t is
// still inside it; FLAT: one row)
__device__ __forceinline__ void frame_rows(const OutSpec& o, int b, int t, int frames_b, int sections_b, float*& r0,
                                           float*& r1, bool& l0, bool& l1) {
  r0 = r1 = nullptr;
  l0 = l1 = false;
  if (o.layout == AST_LAYOUT_FLAT) {
    r0 = o.out + ((long long)b * 2 * o.dim1 + t) * o.f_row + o.f_off;
    l0 = t < frames_b;
    return;
  }
  int s_hi = t / o.step;
  if (s_hi > o.dim1 - 1) s_hi = o.dim1 - 1;
  const int tau = t - s_hi * o.step;
  if (tau < o.window) {
    r0 = o.out + (((long long)b * o.dim1 + s_hi) * 2 * o.window + tau) * o.f_row + o.f_off;
    l0 = (s_hi < sections_b) && (t < frames_b);
  }
  if (s_hi >= 1 && tau + o.step < o.window) {
    r1 = o.out + (((long long)b * o.dim1 + s_hi - 1) * 2 * o.window + tau + o.step) * o.f_row + o.f_off;
    l1 = (s_hi - 1 < sections_b) && (t < frames_b);
  }
  if (!r0) {  // keep r0 the primary row
    r0 = r1, l0 = l1;
    r1 = nullptr, l1 = false;
  }
}

__global__ void __launch_bounds__(kStftThreads, AST_STFT_CTAS) stft_kernel(const StftParams p) {
  extern __shared__ __align__(16) float2 smem[];
  float2* t1 = smem;
  float2* t2 = smem + kTw1Size;
  pdl_launch_dependents();
  const int group = threadIdx.x >> 6, tid = threadIdx.x & 63;
  float2* buf1 = smem + kTw1Size + kTw2Size + group * (kBuf1Size + kBuf2Size);
  float2* buf2 = buf1 + kBuf1Size;
  const int b = blockIdx.y;
  const int len = (int)(p.lengths ? p.lengths[b] : p.max_samples);
#if AST_STFT_PROLOGUE
  {
    // the first pair's sample lines start their trip to the L1 before the twiddle tables are copied
    const int ta0 = 2 * (blockIdx.x * p.iters * kStftGroups + group);
    if (ta0 >= 2 && (ta0 + 1) * kHop + kNfft / 2 <= len) {
      const float* xp = p.wave + (long long)b * p.wave_stride + (ta0 * kHop - kNfft / 2 + tid);
#pragma unroll
      for (int j = 0; j < 20; ++j) asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + 64 * j));
    }
  }
  {
    // 10 KB of twiddles as 16-byte loads, all in flight before the first store (the tables are 16-byte aligned)
    static_assert((kTw1Size + kTw2Size) % (2 * kStftThreads) == 0, "twiddle copy assumes whole rounds");
    constexpr int kRounds = (kTw1Size + kTw2Size) / (2 * kStftThreads);
    const float4* __restrict__ src1 = reinterpret_cast<const float4*>(p.t1);
    const float4* __restrict__ src2 = reinterpret_cast<const float4*>(p.t2);
    float4 tw[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int i = threadIdx.x + r * kStftThreads;   // float4 index into [t1 | t2]
      tw[r] = i < kTw1Size / 2 ? __ldg(src1 + i) : __ldg(src2 + (i - kTw1Size / 2));
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) reinterpret_cast<float4*>(smem)[threadIdx.x + r * kStftThreads] = tw[r];
  }
#else
  for (int i = threadIdx.x; i < kTw1Size; i += kStftThreads) t1[i] = p.t1[i];
  for (int i = threadIdx.x; i < kTw2Size; i += kStftThreads) t2[i] = p.t2[i];
#endif
  float win[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) win[n1] = __ldg(p.hann + 64 * n1 + tid);

  const int frames_b = num_frames(len);
  const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
  const float* __restrict__ x = p.wave + (long long)b * p.wave_stride;
  const float2* st0 = nullptr;
  const float2* st1 = nullptr;
  if (p.out.stats) {
    st0 = p.out.stats + (long long)b * p.out.stats_clip_stride + p.out.stats_off;
    st1 = st0 + p.out.f_stats;
  }
  const long long plane = (long long)(p.out.layout == AST_LAYOUT_FLAT ? p.out.dim1 : p.out.window) * p.out.f_row;
  const bool no_store = p.debug & 1;
  __syncthreads();

#if AST_STFT_PREFETCH == 1
  float xv[20];          // samples of the next interior pair, in flight across stages 2 and 3 of the current one
  bool have_xv = false;
#endif
  for (int it = 0; it < p.iters; ++it) {
    const int pair = (blockIdx.x * p.iters + it) * kStftGroups + group;
    const bool active = pair < p.pairs_per_clip;
    const int ta = 2 * pair;
    const bool any_live = active && ta < frames_b;
    if (any_live) {
      float2 v[16];
      const int base = ta * kHop - kNfft / 2 + tid;
      const bool interior = ta >= 2 && (ta + 1) * kHop + kNfft / 2 <= len;
      if (interior) {
        // frame B sample n is frame A sample n + 256: x[base + 64 j], j = 0..19, feeds both
#if AST_STFT_PREFETCH == 1
        if (!have_xv) {
          const float* __restrict__ xp = x + base;
#pragma unroll
          for (int j = 0; j < 20; ++j) xv[j] = __ldg(xp + 64 * j);
        }
#else
        const float* __restrict__ xp = x + base;
        float xv[20];
#pragma unroll
        for (int j = 0; j < 20; ++j) xv[j] = __ldg(xp + 64 * j);
#endif
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = make_float2(xv[n1] * win[n1], xv[n1 + 4] * win[n1]);
      } else {
        const bool live_b = ta + 1 < frames_b;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int i = base + 64 * n1;
          const float xa = load_reflect(x, i, len, p.pad_zero);
          const float xb = live_b ? load_reflect(x, i + kHop, len, p.pad_zero) : 0.f;
          v[n1] = make_float2(xa * win[n1], xb * win[n1]);
        }
      }
      fft1024_stage1(v, tid, t1, buf1);
    }
#if AST_STFT_PREFETCH == 2
    {
      // the 512 samples of this group's next pair that no group of this iteration has touched: lines into L1
      const int tn = ta + 2 * kStftGroups;
      if (it + 1 < p.iters && tn >= 2 && (tn + 1) * kHop + kNfft / 2 <= len) {
        const float* __restrict__ xp = x + (tn * kHop - kNfft / 2 + tid);
#pragma unroll
        for (int j = 12; j < 20; ++j) asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + 64 * j));
      }
    }
#elif AST_STFT_PREFETCH
    {
      const int tn = ta + 2 * kStftGroups;   // this group's pair of the next iteration
      have_xv = it + 1 < p.iters && tn >= 2 && (tn + 1) * kHop + kNfft / 2 <= len;
      if (have_xv) {
        const float* __restrict__ xp = x + (tn * kHop - kNfft / 2 + tid);
#pragma unroll
        for (int j = 0; j < 20; ++j) xv[j] = __ldg(xp + 64 * j);
      }
    }
#endif
    __syncthreads();
#if AST_STFT_PRELOAD
    float2 m0[9], m1[9];
#endif
    if (any_live) {
      fft1024_stage2(tid, t2, buf1, buf2);
#if AST_STFT_PRELOAD
      if (st0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          if (i < 8 || tid == 0) {
            const int k = fft1024_stage3_bin(tid, i);
            m0[i] = __ldg(st0 + k), m1[i] = __ldg(st1 + k);
          }
        }
      }
#endif
    }
    __syncthreads();
    if (active) {
      float *a0, *a1, *b0 = nullptr, *b1 = nullptr;
      bool la0, la1, lb0 = false, lb1 = false;
      frame_rows(p.out, b, ta, frames_b, sections_b, a0, a1, la0, la1);
      if (ta + 1 < p.slots) frame_rows(p.out, b, ta + 1, frames_b, sections_b, b0, b1, lb0, lb1);
      if (no_store) a0 = a1 = b0 = b1 = nullptr, la0 = false;  // diagnostic: all stores predicated off (not-all-live path)
      const bool all_live = la0 && (la1 || !a1) && (lb0 || !b0) && (lb1 || !b1);
      if (any_live && all_live) {
#if AST_STFT_PRELOAD
        StftEmit<true> emit{a0, a1, b0, b1, true, true, true, true, plane, m0, m1, st0 != nullptr, 0};
#else
        StftEmit<true> emit{a0, a1, b0, b1, true, true, true, true, plane, st0, st1};
#endif
        fft1024_stage3_real_pair<false>(tid, buf2, emit);
      } else if (any_live) {
#if AST_STFT_PRELOAD
        StftEmit<false> emit{a0, a1, b0, b1, la0, la1, lb0, lb1, plane, m0, m1, st0 != nullptr, 0};
#else
        StftEmit<false> emit{a0, a1, b0, b1, la0, la1, lb0, lb1, plane, st0, st1};
#endif
        fft1024_stage3_real_pair<false>(tid, buf2, emit);
      } else {
        // both frames lie past the clip: their rows exist in the output and must be zeros
#if AST_STFT_PRELOAD
        StftEmit<false> emit{a0, a1, b0, b1, false, false, false, false, plane, m0, m1, false, 0};
#else
        StftEmit<false> emit{a0, a1, b0, b1, false, false, false, false, plane, nullptr, nullptr};
#endif
        for (int k = tid; k < kFStft; k += kFftThreads) emit(k, 0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (p.tail_counter) {
    // This grid never waited for its programmatic primary (the CQT projection) and may finish before it.  The last CTA
    // to get here waits for that grid, so the STFT - the feature call's last kernel - cannot COMPLETE before the call's
    // other kernels have: a following programmatic dependent that waits for "the previous kernel" gets the whole call.
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int done = atomicAdd(p.tail_counter, 1u);
      if (done == gridDim.x * gridDim.y - 1) pdl_wait();
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
                long long wave_stride, const OutSpec& out, cudaStream_t st, int pad_zero, bool pdl, unsigned int* tail_counter) {
  StftParams p;
  p.pad_zero = pad_zero;
  p.tail_counter = pdl ? tail_counter : nullptr;
  {
    const char* env = getenv("AST_STFT_DEBUG");
    p.debug = env ? atoi(env) : 0;
  }
  p.wave = wave;
  p.lengths = lengths;
  p.max_samples = max_samples;
  p.wave_stride = wave_stride;
  p.slots = frame_slots(out.layout, out.dim1, out.window, out.step);
  p.pairs_per_clip = (p.slots + 1) / 2;
  p.t1 = plan->d_tw1;
  p.t2 = plan->d_tw2;
  p.hann = plan->d_hann;
  p.overlap = out.window - out.step;
  p.out = out;
  if (p.pairs_per_clip == 0 || batch == 0) return AST_OK;
  if (max_samples >= (1LL << 30)) return fail(AST_ERR_INVALID_ARG, "clips longer than 2^30 samples are not supported");
  // pair groups per CTA: aim at ~6 CTAs per resident slot over the whole grid, at most 8 groups per CTA (measured at
  // 64 clips: 3 - 4 groups per CTA 0.1316 ms, 2: 0.1335, 6: 0.1332, 8: 0.1347, 1: 0.155)
  const long long groups_per_clip = (p.pairs_per_clip + kStftGroups - 1) / kStftGroups;
  const long long slots_total = (long long)plan->sm_count * g_stft_ctas_per_sm * 6;
  long long iters = (groups_per_clip * batch + slots_total - 1) / slots_total;
  if (iters < 1) iters = 1;
  if (iters > 8) iters = 8;
  if (const char* env = getenv("AST_STFT_ITERS")) {  // diagnostic override
    const int v = atoi(env);
    if (v >= 1 && v <= 64) iters = v;
  }
  p.iters = (int)iters;
  dim3 grid((unsigned)((groups_per_clip + iters - 1) / iters), (unsigned)batch);
  ProfileSpan span("stft_kernel", st);
  if (pdl) {
    // programmatic dependent of the CQT projection launched just before it on the same stream: the kernel never
    // waits for it (disjoint output columns), so its CTAs fill the SMs as the persistent CQT CTAs retire
    AST_CUDA_TRY(launch_with_pdl(stft_kernel, grid, kStftThreads, kStftSmem, st, p));
  } else {
    stft_kernel<<<grid, kStftThreads, kStftSmem, st>>>(p);
    AST_LAUNCH_CHECK("stft_kernel");
  }
  return AST_OK;
}

}  // namespace ast
