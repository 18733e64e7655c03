// stft.cu - K1: framing + reflect padding + Hann window + real FFT-1024 of two frames per complex
// transform, with the normalise / real-imag plane split / section scatter epilogue fused into the
// stores.  Replaces torch.stft (utilityFunctions.py:26-28), the real/imag stacking (:31-35),
// dataloader.normalize (dataloader.py:9-13) and get_overlap_windows (utilityFunctions.py:240-263)
// for the STFT columns of the feature tensor; in statistics mode it replaces the per-clip reductions of
// compute_stats (compute_separated_stats.py:27-28) for those columns and stores nothing else.
//
// ONE WARP owns one frame pair (frames 2p and 2p+1 share one complex FFT and 3/4 of their samples: 40 coalesced
// loads feed both).  1024 = 32 x 32 (fft_core.h): every lane transforms 32 points in registers, the warp exchanges
// them once through its private 8.4 KB shared-memory tile, every lane transforms 32 points again and holds bins
// k1 + 32 k2 (k1 = lane).  Bin 1024 - k lives in lane (32 - k1) % 32, so the Hermitian separation of the two frames is
// one shuffle per bin, and the warp's stores run along output rows (128 contiguous bytes per instruction).  There is
// no block-level barrier and no table in shared memory: twiddles, window and the per-bin (mean, rstd) table are 20 KB
// of read-only data that stay in the L1 (16 warps x 8.4 KB of tiles leave it ~ 85 KB).  Round 1's kernel (64 threads
// per pair, 16 x 16 x 4, two exchanges, three block barriers per pair) issued ~ 2 800 warp instructions per pair and
// was issue / latency bound at 0.131 ms per 64 clips; this formulation issues ~ 1 100.
#include <cstdlib>

#include "common.cuh"

namespace ast {

#ifndef AST_STFT_WARPS
#define AST_STFT_WARPS 4    // warps (= frame pairs in flight) per CTA
#endif
#ifndef AST_STFT_CTAS
#define AST_STFT_CTAS 4     // resident CTAs per SM the register allocation is sized for (16 warps, 128 registers)
#endif
constexpr int kStftWarps = AST_STFT_WARPS;
constexpr int kStftThreads = kStftWarps * 32;
constexpr int kAccStride = 516;   // float4 per warp of statistics accumulators (513 bins, padded)
constexpr int kStftStatsIters = 64 / kStftWarps;   // statistics mode: 64 pairs = 128 frames per CTA tile
constexpr size_t kStftSmem = sizeof(float2) * kStftWarps * kTileSize;
constexpr size_t kStftStatsSmem = kStftSmem + sizeof(float4) * kStftWarps * kAccStride + sizeof(float) * kStftWarps;

struct StftParams {
  const float* wave;
  const int32_t* lengths;
  long long max_samples, wave_stride;
  int slots;            // frame slots per clip (rows of the output the kernel must cover)
  int pairs_per_clip;   // ceil(slots / 2)
  int iters;            // pairs per warp (consecutive warps of a CTA take consecutive pairs)
  const float2* tw32;   // [32][32] W_1024^(n2 k1)
  const float* hann;
  const float4* stat4;  // (-mean_re, -mean_im, rstd_re, rstd_im) per bin, [clip or 0][516]; nullptr: no normalisation
  int stat4_clip_stride;
  int overlap;
  unsigned int* tail_counter;  // chained feature call: finished-CTA counter (zeroed by the prologue kernel), else nullptr
  int pad_zero;         // 0: reflect padding (torch.stft, get_STFT); 1: zero padding (librosa.stft default, mse_spectrogram)
  OutSpec out;
  float2* part;         // statistics mode: [clip][tile][2][513] (mean, M2) of the tile's frames
  float* part_n;        //                  [clip][tile] frames in the tile
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

// Destination of the two frames of a pair: up to two rows each (a frame inside the overlap of two sections is stored
// twice).  Pointers are pre-offset by the lane, so bin lane + 32 k2 is element 32 k2; re / im row pointers are both
// precomputed so that every store is [pointer + immediate].
//   kRows = 1: frames A and B have one live row each (the common case)      4 stores per bin
//   kRows = 2: both frames sit inside a section overlap, all four rows live 8 stores per bin
//   kRows = 0: anything else (clip end, zero padding, a pair straddling an overlap boundary): predicated
// Normalisation runs on the packed FP32x2 pipe: (re, im) of a frame against (-mean_re, -mean_im) and (rstd_re, rstd_im).
// Values arrive WITHOUT the factor 1/2 of the Hermitian separation; x = 0.5 s is exact, so fma(s, 0.5, -mean) rounds
// once, exactly like the reference's (x - mean).
template <int kRows>
struct StftStore {
  float *a0r, *a0i, *a1r, *a1i, *b0r, *b0i, *b1r, *b1i;   // nullptr: absent
  bool la0, la1, lb0, lb1;
  const float4* st;     // per-bin (-mean_re, -mean_im, rstd_re, rstd_im), pre-offset by the lane; nullptr: raw
  __device__ __forceinline__ void operator()(int k, float2 a, float2 b) const {
    const float2 half = make_float2(0.5f, 0.5f);
#if defined(__CUDA_ARCH__)
    if (st) {
      const float4 s = __ldg(st + k);
      a = __fmul2_rn(__ffma2_rn(a, half, make_float2(s.x, s.y)), make_float2(s.z, s.w));
      b = __fmul2_rn(__ffma2_rn(b, half, make_float2(s.x, s.y)), make_float2(s.z, s.w));
    } else {
      a = __fmul2_rn(a, half);
      b = __fmul2_rn(b, half);
    }
#endif
    if (kRows == 1) {
      a0r[k] = a.x, a0i[k] = a.y, b0r[k] = b.x, b0i[k] = b.y;
    } else if (kRows == 2) {
      a0r[k] = a.x, a0i[k] = a.y, b0r[k] = b.x, b0i[k] = b.y;
      a1r[k] = a.x, a1i[k] = a.y, b1r[k] = b.x, b1i[k] = b.y;
    } else {
      if (a0r) a0r[k] = la0 ? a.x : 0.f, a0i[k] = la0 ? a.y : 0.f;
      if (a1r) a1r[k] = la1 ? a.x : 0.f, a1i[k] = la1 ? a.y : 0.f;
      if (b0r) b0r[k] = lb0 ? b.x : 0.f, b0i[k] = lb0 ? b.y : 0.f;
      if (b1r) b1r[k] = lb1 ? b.x : 0.f, b1i[k] = lb1 ? b.y : 0.f;
    }
  }
};

// Statistics mode: running (mean, M2) per bin and channel over the frames this warp has seen, in the warp's own
// shared-memory array (each lane owns its bins: no conflicts, no atomics).  A pair enters as its own two-sample
// (mean, M2) and is merged with Chan's formula; w1 = cnt / n, w2 = n_prev cnt / n are warp-uniform.
struct StftMoments {
  float4* acc;   // [516] (mean_re, M2_re, mean_im, M2_im), pre-offset by the lane
  float w1, w2;
  bool two;      // both frames of the pair are live
  __device__ __forceinline__ void operator()(int k, float2 a, float2 b) const {
    const float are = 0.5f * a.x, aim = 0.5f * a.y, bre = 0.5f * b.x, bim = 0.5f * b.y;
    float4 s = acc[k];
    float mre = are, mim = aim, qre = 0.f, qim = 0.f;
    if (two) {
      mre = 0.5f * (are + bre), mim = 0.5f * (aim + bim);
      const float dre = are - bre, dim = aim - bim;
      qre = 0.5f * dre * dre, qim = 0.5f * dim * dim;
    }
    const float ere = mre - s.x, eim = mim - s.z;
    s.x = fmaf(ere, w1, s.x);
    s.y += fmaf(ere * ere, w2, qre);
    s.z = fmaf(eim, w1, s.z);
    s.w += fmaf(eim * eim, w2, qim);
    acc[k] = s;
  }
};

// all 513 bins of both frames from the transform's registers: lane k1 holds Z[k1 + 32 k2] in v[k2].
// emit(k, (A_re, A_im), (B_re, B_im)), both without the factor 1/2:
//   A[k] = Z[k] + conj Z[N - k]          B[k] = -i (Z[k] - conj Z[N - k])
template <class Emit>
__device__ __forceinline__ void separate_and_emit(const float2 (&v)[32], int lane, const Emit& emit) {
  const int src = (32 - lane) & 31;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    // Z[1024 - k]: lane (32 - k1) % 32, register 31 - k2; lane 0 pairs with its own register (32 - k2) % 32
    float2 zp;
    zp.x = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
    zp.y = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
    if (lane == 0) zp = v[(32 - k2) & 31];
    const float2 a = cadd(v[k2], make_float2(zp.x, -zp.y));
    const float2 d = csub(v[k2], make_float2(zp.x, -zp.y));
    emit(32 * k2, a, make_float2(d.y, -d.x));
  }
  if (lane == 0) {  // bin 512 is its own conjugate partner: imaginary parts exactly 0
    const float2 zk = v[16];
    emit(512, make_float2(zk.x + zk.x, zk.y - zk.y), make_float2(zk.y + zk.y, zk.x - zk.x));
  }
}

template <int kMode>   // 0: features (normalise + store), 1: statistics (per-bin moments, nothing stored)
__global__ void __launch_bounds__(kStftThreads, kMode == 0 ? AST_STFT_CTAS : 2) stft_kernel(const StftParams p) {
  extern __shared__ __align__(16) unsigned char stft_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* tile = reinterpret_cast<float2*>(stft_smem) + warp * kTileSize;
  float4* acc = reinterpret_cast<float4*>(stft_smem + kStftSmem) + warp * kAccStride;   // statistics mode only
  float* warp_n = reinterpret_cast<float*>(stft_smem + kStftSmem + sizeof(float4) * kStftWarps * kAccStride);
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int len = (int)(p.lengths ? p.lengths[b] : p.max_samples);
  const int frames_b = num_frames(len);
  const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
  const float* __restrict__ x = p.wave + (long long)b * p.wave_stride;
  const float2* __restrict__ tw = p.tw32;
  const float* __restrict__ win = p.hann + lane;
  const float4* st = p.stat4 ? p.stat4 + (long long)b * p.stat4_clip_stride + lane : nullptr;
  const long long plane = (long long)(p.out.layout == AST_LAYOUT_FLAT ? p.out.dim1 : p.out.window) * p.out.f_row;
  float n_acc = 0.f;
  if (kMode == 1)
    for (int k = lane; k < kAccStride; k += 32) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int it = 0; it < p.iters; ++it) {
    const int pair = (blockIdx.x * p.iters + it) * kStftWarps + warp;
    if (pair >= p.pairs_per_clip) break;   // warp-uniform; nothing below synchronises across warps
    const int ta = 2 * pair;
    const bool any_live = ta < frames_b, live_b = ta + 1 < frames_b;
    float2 v[32];
    if (any_live) {
      const int base = ta * kHop - kNfft / 2 + lane;
      if (ta >= 2 && (ta + 1) * kHop + kNfft / 2 <= len) {
        // interior: frame B sample n is frame A sample n + 256, x[base + 32 i], i = 0..39, feeds both
        const float* __restrict__ xp = x + base;
        float xv[40];
#pragma unroll
        for (int i = 0; i < 40; ++i) xv[i] = __ldg(xp + 32 * i);
        // this warp's next pair starts kStftWarps x 512 samples further; its first 768 samples are being read by the
        // CTA's other warps right now, the last 512 (16 lines) are new: one line per lane into the L1
        if (it + 1 < p.iters && lane < 16 && (ta + 2 * kStftWarps + 1) * kHop + kNfft / 2 <= len)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + 2 * kStftWarps * kHop + 768 - lane + 32 * lane));
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float w = __ldg(win + 32 * i);
#if defined(__CUDA_ARCH__)
          v[i] = __fmul2_rn(make_float2(xv[i], xv[i + 8]), make_float2(w, w));
#endif
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int idx = base + 32 * i;
          const float w = __ldg(win + 32 * i);
          const float xa = load_reflect(x, idx, len, p.pad_zero);
          const float xb = live_b ? load_reflect(x, idx + kHop, len, p.pad_zero) : 0.f;
          v[i] = make_float2(xa * w, xb * w);
        }
      }
      fft32(v);
      tile[lane] = v[0];
#pragma unroll
      for (int k1 = 1; k1 < 32; ++k1) tile[k1 * kTileStride + lane] = cmul(v[k1], __ldg(tw + k1 * 32 + lane));
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) v[n2] = tile[lane * kTileStride + n2];
      __syncwarp();   // the tile may be rewritten by the next pair
      fft32(v);
    }
    if (kMode == 1) {
      if (any_live) {
        const float cnt = live_b ? 2.f : 1.f, n_new = n_acc + cnt;
        const StftMoments emit{acc + lane, cnt / n_new, n_acc * cnt / n_new, live_b};
        separate_and_emit(v, lane, emit);
        n_acc = n_new;
      }
      continue;
    }
    float *a0, *a1, *b0 = nullptr, *b1 = nullptr;
    bool la0, la1, lb0 = false, lb1 = false;
    frame_rows(p.out, b, ta, frames_b, sections_b, a0, a1, la0, la1);
    if (ta + 1 < p.slots) frame_rows(p.out, b, ta + 1, frames_b, sections_b, b0, b1, lb0, lb1);
    auto re = [&](float* r) { return r ? r + lane : nullptr; };
    auto im = [&](float* r) { return r ? r + lane + plane : nullptr; };
    if (any_live && la0 && lb0 && !a1 && !b1) {
      const StftStore<1> emit{re(a0), im(a0), nullptr, nullptr, re(b0), im(b0), nullptr, nullptr, true, false, true, false, st};
      separate_and_emit(v, lane, emit);
    } else if (any_live && la0 && lb0 && la1 && lb1) {
      const StftStore<2> emit{re(a0), im(a0), re(a1), im(a1), re(b0), im(b0), re(b1), im(b1), true, true, true, true, st};
      separate_and_emit(v, lane, emit);
    } else if (any_live) {
      const StftStore<0> emit{re(a0), im(a0), re(a1), im(a1), re(b0), im(b0), re(b1), im(b1), la0, la1, lb0, lb1, st};
      separate_and_emit(v, lane, emit);
    } else {
      // both frames lie past the clip: their rows exist in the output and must be zeros
      const StftStore<0> emit{re(a0), im(a0), re(a1), im(a1), re(b0), im(b0), re(b1), im(b1), false, false, false, false, nullptr};
      for (int k = 0; k + lane < kFStft; k += 32) emit(k, make_float2(0.f, 0.f), make_float2(0.f, 0.f));
    }
  }

  if (kMode == 1) {
    // merge the CTA's warps (frame-ascending interleave does not matter to Chan's formula) in double and write the
    // tile's partial moments; the finalise kernel merges tiles in order
    if (lane == 0) warp_n[warp] = n_acc;
    __syncthreads();
    const float4* accs = reinterpret_cast<const float4*>(stft_smem + kStftSmem);
    const long long tile_id = (long long)b * gridDim.x + blockIdx.x;
    for (int k = threadIdx.x; k < kFStft; k += kStftThreads) {
      double n = 0.0, mre = 0.0, qre = 0.0, mim = 0.0, qim = 0.0;
#pragma unroll
      for (int w = 0; w < kStftWarps; ++w) {
        const double nw = warp_n[w];
        if (nw > 0.0) {
          const float4 s = accs[w * kAccStride + k];
          const double nn = n + nw, dre = (double)s.x - mre, dim = (double)s.z - mim;
          mre += dre * nw / nn;
          qre += (double)s.y + dre * dre * n * nw / nn;
          mim += dim * nw / nn;
          qim += (double)s.w + dim * dim * n * nw / nn;
          n = nn;
        }
      }
      p.part[(tile_id * 2 + 0) * kFStft + k] = make_float2((float)mre, (float)qre);
      p.part[(tile_id * 2 + 1) * kFStft + k] = make_float2((float)mim, (float)qim);
      if (k == 0) p.part_n[tile_id] = (float)n;
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

static int g_stft_ctas_per_sm = AST_STFT_CTAS;

int stft_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(stft_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftSmem));
  AST_CUDA_TRY(cudaFuncSetAttribute(stft_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftStatsSmem));
  int n = 0;
  AST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_kernel<0>, kStftThreads, kStftSmem));
  g_stft_ctas_per_sm = n > 0 ? n : 1;
  return AST_OK;
}

// pairs per warp: the count that minimises (waves of CTAs) x (pairs per warp) for this grid, i.e. the wall time of a
// kernel whose CTAs all take the same time, among counts that keep a CTA's pairs contiguous
static int pick_iters(long long groups_per_clip, int batch, long long slots, int lo, int hi) {
  double best = 1e300;
  int best_it = lo;
  for (int it = lo; it <= hi; ++it) {
    const long long ctas = (groups_per_clip + it - 1) / it * batch;
    const double waves = (double)((ctas + slots - 1) / slots);
    const double cost = waves * (it + 0.35);   // + a CTA's fixed cost (launch, first loads) in units of one pair
    if (cost < best * 0.999) best = cost, best_it = it;
  }
  return best_it;
}

int stft_tiles_per_clip(const ast_plan* plan, int batch, int slots, bool stats_mode) {
  const long long pairs = (slots + 1) / 2, groups = (pairs + kStftWarps - 1) / kStftWarps;
  if (groups == 0 || batch == 0) return 0;
  // statistics mode: FIXED tiles of kStftStatsIters pairs per warp (128 frames), whatever the batch - a clip's partial
  // moments, and so its float32 rounding, must not depend on which clips share its launch (determinism across ranks)
  const int it = stats_mode ? kStftStatsIters : pick_iters(groups, batch, (long long)plan->sm_count * g_stft_ctas_per_sm, 1, 12);
  return (int)((groups + it - 1) / it);
}

int launch_stft(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
                long long wave_stride, const OutSpec& out, cudaStream_t st, int pad_zero, bool pdl, unsigned int* tail_counter,
                const float4* stat4, int stat4_clip_stride, float2* part, float* part_n) {
  StftParams p;
  p.pad_zero = pad_zero;
  p.tail_counter = pdl ? tail_counter : nullptr;
  p.wave = wave;
  p.lengths = lengths;
  p.max_samples = max_samples;
  p.wave_stride = wave_stride;
  p.slots = frame_slots(out.layout, out.dim1, out.window, out.step);
  p.pairs_per_clip = (p.slots + 1) / 2;
  p.tw32 = plan->d_tw32;
  p.hann = plan->d_hann;
  p.stat4 = stat4;
  p.stat4_clip_stride = stat4_clip_stride;
  p.overlap = out.window - out.step;
  p.out = out;
  p.part = part;
  p.part_n = part_n;
  if (p.pairs_per_clip == 0 || batch == 0) return AST_OK;
  if (max_samples >= (1LL << 30)) return fail(AST_ERR_INVALID_ARG, "clips longer than 2^30 samples are not supported");
  const bool stats_mode = part != nullptr;
  const long long groups_per_clip = (p.pairs_per_clip + kStftWarps - 1) / kStftWarps;
  const int tiles = stft_tiles_per_clip(plan, batch, p.slots, stats_mode);
  long long iters = stats_mode ? kStftStatsIters : (groups_per_clip + tiles - 1) / tiles;
  if (!stats_mode)
    if (const char* env = getenv("AST_STFT_ITERS")) {  // diagnostic override
      const int v = atoi(env);
      if (v >= 1 && v <= 64) iters = v;
    }
  p.iters = (int)iters;
  dim3 grid((unsigned)((groups_per_clip + iters - 1) / iters), (unsigned)batch);
  ProfileSpan span("stft_kernel", st);
  if (stats_mode) {
    if ((int)grid.x != tiles) return fail(AST_ERR_INVALID_ARG, "internal: statistics tile count mismatch");
    stft_kernel<1><<<grid, kStftThreads, kStftStatsSmem, st>>>(p);
    AST_LAUNCH_CHECK("stft_kernel<stats>");
  } else if (pdl) {
    // programmatic dependent of the CQT projection launched just before it on the same stream: the kernel never
    // waits for it (disjoint output columns), so its CTAs fill the SMs as the persistent CQT CTAs retire
    AST_CUDA_TRY(launch_with_pdl(stft_kernel<0>, grid, kStftThreads, kStftSmem, st, p));
  } else {
    stft_kernel<0><<<grid, kStftThreads, kStftSmem, st>>>(p);
    AST_LAUNCH_CHECK("stft_kernel");
  }
  return AST_OK;
}

}  // namespace ast
