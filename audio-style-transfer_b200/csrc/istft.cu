// istft.cu - K5: section merge + C2R inverse FFT-1024 + Hann synthesis window + overlap-add
// + envelope normalisation + centre trim, in one kernel.  Replaces sections2spectrogram
// (utilityFunctions.py:265-283) and torch.istft as called by inverse_STFT (utilityFunctions.py:62-82;
// the device-aware copy in style_transfer_inference_test.ipynb cell 1) and by
// reconstruct_audio_from_sections (evaluation_reconstruction.py:161-189).
//
// torch.istft semantics: x_t = irfft(X[:, t], n = 1024) (1/N scale, imaginary parts of bins 0 and
// 512 ignored); output sample n in [0, 256 (T - 1)) sits at p = n + 512 of the untrimmed signal and is
//     y[n] = sum_t w[p - 256 t] x_t[p - 256 t] / sum_t w[p - 256 t]^2,
// t over the (at most 4) frames with 0 <= p - 256 t < 1024.
//
// Formulation: a 64-thread CTA walks a RUN of consecutive frames of one clip, two frames per complex
// FFT.  After stage 3 thread j holds, for each of its four columns q, the samples q + 256 a (a = 0..3)
// of both frames - i.e. the same offset q of four consecutive 256-sample output segments.  Every
// contribution to output position (segment g, offset q) is therefore produced by the SAME thread, so the
// overlap-add runs in registers: no atomics, no shared-memory frame store, deterministic order
// (frames ascending, like torch's fold).  A run of R segments needs 4 halo frames.
// Shared memory per CTA: 16.1 KB exchange buffers (twiddles come from global memory through L1); 8 CTAs / SM.
#include <cstdlib>

#include "common.cuh"

namespace ast {

constexpr int kIstftThreads = kFftThreads;  // 64
constexpr size_t kIstftSmem = sizeof(float2) * (kBuf1Size + kBuf2Size);  // twiddle tables are read through L1

struct IstftParams {
  const float* spec;
  int dim1, f_in, layout;
  int window, sec_hop;   // SECTIONS: rows per section, frames between section starts
  int n_frames;          // T' after merge / crop
  int run_segs;          // R: output segments per CTA (even)
  long long clip_stride; // floats per clip of spec
  long long plane;       // floats between the real and the imaginary plane of a row
  float* out;
  long long out_stride;
  const float2* t1;
  const float2* t2;
  const float* hann_inv_n;
};

// Row pointers (real plane) of merged frame t: one row, or two section rows to be averaged
// (count-normalised overlap average of sections2spectrogram, utilityFunctions.py:276-282).
struct FrameRows {
  const float* r0;
  const float* r1;  // nullptr when a single row covers the frame
};

__device__ __forceinline__ FrameRows merged_rows(const IstftParams& p, const float* __restrict__ clip, int t) {
  FrameRows f;
  f.r1 = nullptr;
  if (p.layout == AST_LAYOUT_FLAT) {
    f.r0 = clip + (long long)t * p.f_in;
    return f;
  }
  int s_hi = t / p.sec_hop;
  if (s_hi > p.dim1 - 1) s_hi = p.dim1 - 1;
  const int tau = t - s_hi * p.sec_hop;
  f.r0 = clip + ((long long)s_hi * 2 * p.window + tau) * p.f_in;
  if (s_hi >= 1 && tau + p.sec_hop < p.window)
    f.r1 = clip + ((long long)(s_hi - 1) * 2 * p.window + tau + p.sec_hop) * p.f_in;
  return f;
}

__device__ __forceinline__ float2 load_bin(const FrameRows& f, long long plane, int kk, bool live) {
  if (!live) return make_float2(0.f, 0.f);
  float re = __ldg(f.r0 + kk), im = __ldg(f.r0 + plane + kk);
  if (f.r1) {
    // ascending section order like the reference's += loop, then / count (= 2)
    re = (__ldg(f.r1 + kk) + re) * 0.5f;
    im = (__ldg(f.r1 + plane + kk) + im) * 0.5f;
  }
  return make_float2(re, im);
}

// The 16 inputs of a thread: all loads of a batch are issued before the first value is used (ncu: the kernel's
// dominant stall is the long scoreboard on these loads, one round trip per batch).  Two batches of eight inputs per
// pair: 32 loads each, 64 when a frame lies inside a section overlap (second section row to be averaged in).
// Row pointers of one frame as the loads want them: bin 64 n + tid (n < 8) is lo[64 n], bin 64 (16 - n) - tid (n >= 8) is
// hi[64 (16 - n)] - every load is [pointer + immediate].  (Round 1 formed each address as base + 64-bit index: 700 of the
// kernel's 2 600 warp instructions per frame pair were integer address arithmetic.)
struct RowPtrs {
  const float* re_lo;
  const float* re_hi;
  const float* im_lo;
  const float* im_hi;
};
__device__ __forceinline__ RowPtrs row_ptrs(const float* row, long long plane, int tid) {
  RowPtrs r;
  r.re_lo = row + tid;
  r.re_hi = row - tid;
  r.im_lo = r.re_lo + plane;
  r.im_hi = r.re_hi + plane;
  return r;
}
template <int N>
__device__ __forceinline__ float load_re(const RowPtrs& r) { return N < 8 ? __ldg(r.re_lo + 64 * N) : __ldg(r.re_hi + 64 * (16 - N)); }
template <int N>
__device__ __forceinline__ float load_im(const RowPtrs& r) { return N < 8 ? __ldg(r.im_lo + 64 * N) : __ldg(r.im_hi + 64 * (16 - N)); }

template <int N0, int I, int CNT>
struct LoadRow {   // compile-time loop over the CNT inputs N0 .. N0 + CNT - 1 of a batch
  static __device__ __forceinline__ void run(const RowPtrs& r, bool on, float (&re)[CNT], float (&im)[CNT]) {
    re[I] = on ? load_re<N0 + I>(r) : 0.f;
    im[I] = on ? load_im<N0 + I>(r) : 0.f;
    if constexpr (I + 1 < CNT) LoadRow<N0, I + 1, CNT>::run(r, on, re, im);
  }
};

template <int N0, int CNT, bool kDup>
__device__ __forceinline__ void load_inputs(float2 (&v)[16], int tid, const FrameRows& fa, const FrameRows& fb,
                                            long long plane, bool live_b) {
  float are[CNT], aim[CNT], bre[CNT], bim[CNT];
  const RowPtrs pa = row_ptrs(fa.r0, plane, tid), pb = row_ptrs(fb.r0, plane, tid);
  LoadRow<N0, 0, CNT>::run(pa, true, are, aim);
  LoadRow<N0, 0, CNT>::run(pb, live_b, bre, bim);
  if (kDup) {
    // the second section row of an overlap frame is loaded in the same batch (predicated off elsewhere), not after it
    const bool dup_a = fa.r1 != nullptr, dup_b = live_b && fb.r1 != nullptr;
    float are1[CNT], aim1[CNT], bre1[CNT], bim1[CNT];
    const RowPtrs pa1 = row_ptrs(dup_a ? fa.r1 : fa.r0, plane, tid), pb1 = row_ptrs(dup_b ? fb.r1 : fb.r0, plane, tid);
    LoadRow<N0, 0, CNT>::run(pa1, dup_a, are1, aim1);
    LoadRow<N0, 0, CNT>::run(pb1, dup_b, bre1, bim1);
    if (dup_a) {  // ascending section order like the reference's += loop, then / count (= 2)
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
        are[i] = (are1[i] + are[i]) * 0.5f;
        aim[i] = (aim1[i] + aim[i]) * 0.5f;
      }
    }
    if (dup_b) {
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
        bre[i] = (bre1[i] + bre[i]) * 0.5f;
        bim[i] = (bim1[i] + bim[i]) * 0.5f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    float2 xa = make_float2(are[i], aim[i]), xb = make_float2(bre[i], bim[i]);
    if ((N0 + i == 0 || N0 + i == 8) && tid == 0) xa.y = 0.f, xb.y = 0.f;  // self-conjugate bins 0 / 512
    v[N0 + i] = (N0 + i) < 8 ? make_float2(xa.x - xb.y, -(xa.y + xb.x)) : make_float2(xa.x + xb.y, -(xb.x - xa.y));
  }
}

// Register-resident overlap-add state of one thread: for each of its 4 columns, partial sums of the
// five segments t .. t+4 touched by the frame pair (t, t+1).
struct OlaEmit {
  float acc[4][5];
  const float* __restrict__ w;   // hann[n] / 1024; read through L1 (4 KB, always resident) instead of 16 registers per
                                 // thread: the registers go to keeping the next pair's input loads in flight
  template <int S>
  __device__ __forceinline__ void col(int q, const float2 (&z)[4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float wn = __ldg(w + q + 256 * a);
      acc[S][a] = fmaf(z[a].x, wn, acc[S][a]);           // frame A (real lane)  -> segment t + a
      acc[S][a + 1] = fmaf(-z[a].y, wn, acc[S][a + 1]);  // frame B (-imag lane) -> segment t + 1 + a
    }
  }
};

// writes segments t and t + 1 (if they belong to this run) from acc[.][0] and acc[.][1]
__device__ __forceinline__ void emit_segments(const IstftParams& p, const OlaEmit& ola, const float (&renv)[4],
                                              const int (&qs)[4], float* __restrict__ y, int t, int g0, int g1) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int g = t + c;
    if (g >= g0 && g < g1) {
      float* __restrict__ yo = y + (long long)(g - 2) * kHop;
      if (g >= 3 && g < p.n_frames) {
        // interior: all four frames exist, the envelope is a per-column constant
#pragma unroll
        for (int s = 0; s < 4; ++s) yo[qs[s]] = ola.acc[s][c] * renv[s];
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          // envelope: sum of w^2 over the live frames g - a, a = 3..0 (frame-ascending order)
          float env = 0.f;
#pragma unroll
          for (int a = 3; a >= 0; --a)
            if (g - a >= 0 && g - a < p.n_frames) {
              const float w = __ldg(p.hann_inv_n + qs[s] + 256 * a) * 1024.f;
              env += w * w;
            }
          yo[qs[s]] = ola.acc[s][c] / env;
        }
      }
    }
  }
}

// resident CTAs per SM the register allocation is sized for: 8 -> 128 registers, no spills, 0.1355 ms per 64 clips;
// 10 -> 102 registers, 148 B of spills, 0.1415; 9 -> 0.1405; 12 -> 80 registers, 0.215.  (An L1 prefetch of the next
// pair's rows after stage 1 made it slower at every occupancy: 0.146 - 0.152.)
// 1: a pair outside the section overlaps loads its 16 inputs in ONE batch of 64 loads instead of two of 32 (measured:
// 0.139 vs 0.1357 ms per 64 clips at 8 CTAs per SM - the longer dependency-free prologue costs more than the saved trip)
#ifndef AST_ISTFT_ONE_BATCH
#define AST_ISTFT_ONE_BATCH 0
#endif
#ifndef AST_ISTFT_CTAS
#define AST_ISTFT_CTAS 8
#endif
__global__ void __launch_bounds__(kIstftThreads, AST_ISTFT_CTAS) istft_kernel(const IstftParams p) {
  extern __shared__ __align__(16) float2 smem[];
  const float2* __restrict__ t1 = p.t1;  // 10 KB of twiddles stay L1-resident; shared memory is kept for the
  const float2* __restrict__ t2 = p.t2;  // exchange buffers
  float2* buf1 = smem;
  float2* buf2 = buf1 + kBuf1Size;
  const int tid = threadIdx.x;
  pdl_launch_dependents();  // a programmatic dependent launched next may run its prologue; an ordinary launch still waits

  const int b = blockIdx.y;
  const float* __restrict__ clip = p.spec + (long long)b * p.clip_stride;
  float* __restrict__ y = p.out + (long long)b * p.out_stride;
  // untrimmed segments g (256 samples each); outputs exist for g in [2, n_frames + 1)
  const int g0 = 2 + blockIdx.x * p.run_segs;
  const int g1 = min(g0 + p.run_segs, p.n_frames + 1);
  const bool first = tid == 0;
  const int qs[4] = {first ? 0 : tid, first ? 128 : 256 - tid, first ? 64 : 128 - tid, first ? 192 : 128 + tid};

  OlaEmit ola;
  ola.w = p.hann_inv_n;
  float renv[4];  // 1 / (w^2 summed over four frames, frame-ascending), the interior envelope
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float wsq[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float w = __ldg(p.hann_inv_n + qs[s] + 256 * a) * 1024.f;
      wsq[a] = w * w;
    }
    renv[s] = 1.0f / (((wsq[3] + wsq[2]) + wsq[1]) + wsq[0]);
  }
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int c = 0; c < 5; ++c) ola.acc[s][c] = 0.f;
  // launched with programmatic stream serialisation: everything above (plan constants only) overlaps the tail of the
  // previous kernel on the stream; the spectrogram it may have produced and the output buffer are touched only below
  pdl_wait();
  __syncthreads();

  // frames g0 - 3 .. g1 - 1 contribute; pairs start on the even frame g0 - 4 (g0 is even)
  int t_next = max(g0 - 4, 0);
  for (int t = t_next; t < g1 && t < p.n_frames; t += 2) {
    const bool live_a = true, live_b = t + 1 < p.n_frames;
    {
      const FrameRows fa = merged_rows(p, clip, t);
      const FrameRows fb = live_b ? merged_rows(p, clip, t + 1) : fa;
      float2 v[16];
#if AST_ISTFT_ONE_BATCH
      if (fa.r1 == nullptr && (!live_b || fb.r1 == nullptr)) {
        load_inputs<0, 16, false>(v, tid, fa, fb, p.plane, live_b);
      } else
#endif
      {
        load_inputs<0, 8, true>(v, tid, fa, fb, p.plane, live_b);
        load_inputs<8, 8, true>(v, tid, fa, fb, p.plane, live_b);
      }
      fft1024_stage1(v, tid, t1, buf1);
    }
    __syncthreads();
    fft1024_stage2(tid, t2, buf1, buf2);
    __syncthreads();
    fft1024_stage3_columns(tid, buf2, ola);
    // segments t and t + 1 are complete now (no later frame reaches them)
    emit_segments(p, ola, renv, qs, y, t, g0, g1);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      ola.acc[s][0] = ola.acc[s][2];
      ola.acc[s][1] = ola.acc[s][3];
      ola.acc[s][2] = ola.acc[s][4];
      ola.acc[s][3] = 0.f;
      ola.acc[s][4] = 0.f;
    }
    t_next = t + 2;
  }
  // when the clip ends on an even frame count the last segment (g = n_frames) is still pending
  emit_segments(p, ola, renv, qs, y, t_next, g0, g1);
}

static int g_istft_ctas_per_sm = 12;

int istft_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIstftSmem));
  int n = 0;
  AST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, istft_kernel, kIstftThreads, kIstftSmem));
  g_istft_ctas_per_sm = n > 0 ? n : 1;
  return AST_OK;
}

int launch_istft(const ast_plan* plan, const float* spec, int batch, int dim1, int f_in, int layout, int window,
                 int overlap, int n_frames, float* wave_out, long long out_stride, cudaStream_t st) {
  if (n_frames < 2 || batch == 0) return AST_OK;  // 256 * (T - 1) = 0 samples
  IstftParams p;
  p.spec = spec;
  p.dim1 = dim1;
  p.f_in = f_in;
  p.layout = layout;
  p.window = window;
  p.sec_hop = window - overlap;
  p.n_frames = n_frames;
  p.clip_stride = layout == AST_LAYOUT_FLAT ? 2LL * dim1 * f_in : 2LL * dim1 * window * f_in;
  p.plane = (long long)(layout == AST_LAYOUT_FLAT ? dim1 : window) * f_in;
  p.out = wave_out;
  p.out_stride = out_stride;
  p.t1 = plan->d_tw1;
  p.t2 = plan->d_tw2;
  p.hann_inv_n = plan->d_hann_inv_n;
  // run length R (even): long runs amortise the 4 halo frames, short runs fill the machine.  Pick the R
  // in [16, 128] that minimises (frames transformed) x (wave quantisation) for this grid.
  const int segs = n_frames - 1;
  const long long slots = (long long)plan->sm_count * g_istft_ctas_per_sm;
  double best = 1e300;
  int best_r = 32;
  for (int r = 16; r <= 128; r += 2) {
    const long long runs = (segs + r - 1) / r;
    const long long ctas = runs * batch;
    const double waves = (double)((ctas + slots - 1) / slots);
    const double per_cta = r + 4;                       // frames per run incl. halo
    const double cost = waves * per_cta;                // time ~ waves x work per CTA
    if (cost < best * 0.999) best = cost, best_r = r;
  }
  if (const char* env = getenv("AST_ISTFT_RUN")) {  // diagnostic override of the run length
    const int r = atoi(env);
    if (r >= 2 && r <= 4096) best_r = r & ~1;
  }
  p.run_segs = best_r;
  dim3 grid((unsigned)((segs + best_r - 1) / best_r), (unsigned)batch);
  ProfileSpan span("istft_kernel", st);
  AST_CUDA_TRY(launch_with_pdl(istft_kernel, grid, kIstftThreads, kIstftSmem, st, p));
  return AST_OK;
}

}  // namespace ast
