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
#include <cstring>

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
// AST_STFT_SMEM_PF: the samples of a warp's NEXT interior pair (1280 floats) are fetched by one cp.async.bulk into the
// warp's own 5 KB buffer while it works on the current pair (feature mode; completion on a per-warp mbarrier)
#ifndef AST_STFT_SMEM_PF
#define AST_STFT_SMEM_PF 0
#endif
constexpr int kPfFloats = 40 * 32;   // one pair's window: frames 2p and 2p + 1
constexpr size_t kStftFeatSmem = kStftSmem + (AST_STFT_SMEM_PF ? sizeof(float) * kStftWarps * kPfFloats + 8 * kStftWarps : 0);

struct StftParams {
  const float* wave;
  const int32_t* lengths;
  long long max_samples, wave_stride;
  int slots;            // frame slots per clip (rows of the output the kernel must cover)
  int pairs_per_clip;   // ceil(slots / 2)
  int iters;            // pairs per warp (consecutive warps of a CTA take consecutive pairs)
  int taper_from;       // feature mode: grid rows >= taper_from are HALF rows of the last clips (iters / 2 pairs per warp, two
                        // rows per clip), and after taper_half_rows of those QUARTER rows (iters / 4, four rows per clip): the
  int taper_half_rows;  // kernel's last CTAs are shorter, so the SMs drain sooner; taper_from >= rows of the grid: off
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
// (-mean_re, -mean_im, rstd_re, rstd_im) of one bin against both frames of the pair
__device__ __forceinline__ void normalise_pair(float2& a, float2& b, const float4& s) {
#if defined(__CUDA_ARCH__)
  const float2 half = make_float2(0.5f, 0.5f);
  a = __fmul2_rn(__ffma2_rn(a, half, make_float2(s.x, s.y)), make_float2(s.z, s.w));
  b = __fmul2_rn(__ffma2_rn(b, half, make_float2(s.x, s.y)), make_float2(s.z, s.w));
#endif
}
struct StftStats {   // the per-bin table of a clip, or none (raw features: x = 0.5 s)
  const float4* st;  // pre-offset by the lane; nullptr: raw
  __device__ __forceinline__ float4 operator()(int k) const {
    return st ? __ldg(st + k) : make_float4(0.f, 0.f, 1.f, 1.f);
  }
};

template <int kRows>
struct StftStore {
  float *a0r, *a0i, *a1r, *a1i, *b0r, *b0i, *b1r, *b1i;   // nullptr: absent
  bool la0, la1, lb0, lb1;
  StftStats stat;
  __device__ __forceinline__ void operator()(int k, float2 a, float2 b, const float4& s) const {
    normalise_pair(a, b, s);
#ifdef AST_STFT_NOSTORE   // diagnostic build (scratch/build_variant.sh): the arithmetic stays, (almost) nothing is stored
    if (a.x != 1.2345e30f) return;
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

// The common case goes through shared memory and the bulk-copy (TMA) engine: output rows start on 4-byte boundaries
// only, so a warp-wide scalar store covers 128 bytes across two lines and five sectors, two of them partial, and the
// 91 such stores of a pair were the kernel's largest single cost (35 of 124 us per 64 clips, measured by predicating them
// off).  Instead the four row-planes of a pair (A re, A im, B re, B im; 513 floats each) are staged in the warp's
// exchange tile, each shifted by its destination's phase (address / 4 mod 4) so that the 16-byte-aligned body of the
// row is 16-byte aligned in shared memory too, and written by ONE cp.async.bulk per row-plane (2 032 - 2 048 bytes);
// the <= 3 floats of head and tail go out as scalar stores.  Second rows of frames inside a section overlap have a
// different phase and are stored the plain way (kDup).
constexpr int kStageRow = 516;   // floats per staged row-plane: 513 + up to 3 of phase shift
template <bool kDup>
struct StftStage {
  float *s0, *s1, *s2, *s3;          // staged rows A re, A im, B re, B im; pre-offset by phase + lane
  float *a1r, *a1i, *b1r, *b1i;      // kDup: the frames' second rows, pre-offset by the lane
  StftStats stat;
  __device__ __forceinline__ void operator()(int k, float2 a, float2 b, const float4& s) const {
    normalise_pair(a, b, s);
    s0[k] = a.x, s1[k] = a.y, s2[k] = b.x, s3[k] = b.y;
    if (kDup) a1r[k] = a.x, a1i[k] = a.y, b1r[k] = b.x, b1i[k] = b.y;
  }
};

// AST_STFT_EVICT_FIRST: the feature rows are written once and not read again inside the call; marking them evict-first
// in the L2 leaves the octave signals (113 MB per 64 clips, written by the decimator, read next by the CQT projection)
// a better chance to stay resident behind the STFT's 300 MB of output.
#ifndef AST_STFT_EVICT_FIRST
#define AST_STFT_EVICT_FIRST 0
#endif
__device__ __forceinline__ void bulk_store(float* gmem, const float* smem, uint32_t bytes) {
#if AST_STFT_EVICT_FIRST
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem),
               "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(bytes), "l"(policy)
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem),
               "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// The four staged row-planes -> their destination rows.  Lane j < 4 issues the bulk copy of row-plane j (one
// instruction, four copies); lanes 8 j + i, i < 6, carry the <= 3 head and <= 3 tail floats of row-plane j.
__device__ __forceinline__ void flush_rows(float* const (&rows)[4], const float* __restrict__ stage, const int (&ph)[4], int lane) {
  const int j = (lane >> 3) & 3, i = lane & 7;
  float* row = j == 0 ? rows[0] : j == 1 ? rows[1] : j == 2 ? rows[2] : rows[3];
  const int phase = j == 0 ? ph[0] : j == 1 ? ph[1] : j == 2 ? ph[2] : ph[3];
  const float* staged = stage + j * kStageRow + phase;
  const int head = (4 - phase) & 3;                  // floats before the first 16-byte boundary
  const int body = (kFStft - head) & ~3;             // 508 or 512 floats
  const int tail = kFStft - head - body;
  const int e = i < 3 ? i : head + body + (i - 3);
  if (i < 6 && (i < 3 ? i < head : i - 3 < tail)) row[e] = staged[e];
  if (i == 7) bulk_store(row + head, staged + head, (uint32_t)body * 4u);
}

// Statistics mode: running (mean, M2) per bin and channel over the frames this warp has seen, in the warp's own
// shared-memory array (each lane owns its bins: no conflicts, no atomics).  A pair enters as its own two-sample
// (mean, M2) and is merged with Chan's formula; w1 = cnt / n, w2 = n_prev cnt / n are warp-uniform.
struct StftMoments {
  float4* acc;   // [516] (mean_re, M2_re, mean_im, M2_im), pre-offset by the lane
  float w1, w2;
  bool two;      // both frames of the pair are live
  StftStats stat;
  __device__ __forceinline__ void operator()(int k, float2 a, float2 b, const float4&) const {
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
__device__ __forceinline__ void separate_and_emit(float2 (&v)[32], int lane, const Emit& emit) {
  // Z[1024 - k] of bin k = lane + 32 k2 lives in lane (32 - lane) % 32, register 31 - k2 (lane 0: its own register
  // (32 - k2) % 32).  (AST_STFT_EMIT_BATCH: all 32 shuffles first, in place, statistics fetched four bins at a time -
  // measured slower: the longer live ranges spill at the 128-register cap.)
  const int src = (32 - lane) & 31;
#if defined(AST_STFT_EMIT_BATCH)
#pragma unroll
  for (int j = 16; j < 32; ++j) {
    v[j].x = __shfl_sync(0xffffffffu, v[j].x, src);
    v[j].y = __shfl_sync(0xffffffffu, v[j].y, src);
  }
  // bins in batches of four: the batch's per-bin statistics are fetched together, ahead of its arithmetic
#pragma unroll
  for (int kb = 0; kb < 16; kb += 4) {
    float4 st[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) st[j] = emit.stat(32 * (kb + j));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k2 = kb + j;
      float2 zp = v[31 - k2];
      if (lane == 0) zp = v[(32 - k2) & 31];   // (lane 0 shuffled with itself: registers unchanged)
      const float2 a = cadd(v[k2], make_float2(zp.x, -zp.y));
      const float2 d = csub(v[k2], make_float2(zp.x, -zp.y));
      emit(32 * k2, a, make_float2(d.y, -d.x), st[j]);
    }
  }
#else
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    float2 zp;
    zp.x = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
    zp.y = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
    if (lane == 0) zp = v[(32 - k2) & 31];
    const float2 a = cadd(v[k2], make_float2(zp.x, -zp.y));
    const float2 d = csub(v[k2], make_float2(zp.x, -zp.y));
    emit(32 * k2, a, make_float2(d.y, -d.x), emit.stat(32 * k2));
  }
#endif
  if (lane == 0) {  // bin 512 is its own conjugate partner: imaginary parts exactly 0
    const float2 zk = v[16];
    emit(512, make_float2(zk.x + zk.x, zk.y - zk.y), make_float2(zk.y + zk.y, zk.x - zk.x), emit.stat(512));
  }
}

// periodic Hann w[32 i + lane] from the lane's (cos theta, sin theta), theta = 2 pi lane / 1024; i is a compile-time
// constant after unrolling, so the two coefficients are immediates: two FFMA, no load
__device__ __forceinline__ float hann_at(int i, float cos_t, float sin_t) {
  // -0.5 cos(2 pi i / 32) and 0.5 sin(2 pi i / 32), i = 0..31 (literals: the device compiler does not fold cos())
  constexpr float kC[32] = {
      -0.5f, -0.49039264f, -0.461939766f, -0.415734806f,
      -0.353553391f, -0.277785117f, -0.191341716f, -0.097545161f,
      0.f, 0.097545161f, 0.191341716f, 0.277785117f,
      0.353553391f, 0.415734806f, 0.461939766f, 0.49039264f,
      0.5f, 0.49039264f, 0.461939766f, 0.415734806f,
      0.353553391f, 0.277785117f, 0.191341716f, 0.097545161f,
      0.f, -0.097545161f, -0.191341716f, -0.277785117f,
      -0.353553391f, -0.415734806f, -0.461939766f, -0.49039264f,
  };
  constexpr float kS[32] = {
      0.f, 0.097545161f, 0.191341716f, 0.277785117f,
      0.353553391f, 0.415734806f, 0.461939766f, 0.49039264f,
      0.5f, 0.49039264f, 0.461939766f, 0.415734806f,
      0.353553391f, 0.277785117f, 0.191341716f, 0.097545161f,
      0.f, -0.097545161f, -0.191341716f, -0.277785117f,
      -0.353553391f, -0.415734806f, -0.461939766f, -0.49039264f,
      -0.5f, -0.49039264f, -0.461939766f, -0.415734806f,
      -0.353553391f, -0.277785117f, -0.191341716f, -0.097545161f,
  };
  return fmaf(kC[i], cos_t, fmaf(kS[i], sin_t, 0.5f));
}

#ifdef AST_STFT_TRACE
// diagnostic build only (scratch/trace_stft.py): cycles per phase of a pair, summed per warp of the first CTAs
__device__ long long g_stft_trace[64][8];
#define STFT_STAMP(k)                                                                 \
  do {                                                                                \
    if (kMode == 0 && lane == 0 && blockIdx.y == 0 && blockIdx.x < 16) {              \
      const long long now_ = clock64();                                               \
      g_stft_trace[blockIdx.x * kStftWarps + warp][k] += now_ - t_prev_;              \
      t_prev_ = now_;                                                                 \
    }                                                                                 \
  } while (0)
__device__ unsigned long long g_stft_cta_ns[4096][3];   // start, end (globaltimer), SM id per CTA (linear block id)
extern "C" int ast_debug_stft_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, g_stft_trace, sizeof(long long) * 64 * 8);
}
extern "C" int ast_debug_stft_cta_times(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, g_stft_cta_ns, sizeof(unsigned long long) * 4096 * 3);
}
__device__ __forceinline__ unsigned long long stft_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#else
#define STFT_STAMP(k) do {} while (0)
#endif

AST_TIMELINE_DEFINE(stft)

template <int kMode>   // 0: features (normalise + store), 1: statistics (per-bin moments, nothing stored)
__global__ void __launch_bounds__(kStftThreads, kMode == 0 ? AST_STFT_CTAS : 3) stft_kernel(const StftParams p) {
  extern __shared__ __align__(16) unsigned char stft_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* tile = reinterpret_cast<float2*>(stft_smem) + warp * kTileSize;
  float4* acc = reinterpret_cast<float4*>(stft_smem + kStftSmem) + warp * kAccStride;   // statistics mode only
  float* warp_n = reinterpret_cast<float*>(stft_smem + kStftSmem + sizeof(float4) * kStftWarps * kAccStride);
  if (kMode == 0) AST_TIMELINE_STAMP(stft, blockIdx.x + gridDim.x * blockIdx.y, 0);
  pdl_launch_dependents();
  int b = blockIdx.y, bx = blockIdx.x, iters = p.iters;
  if (kMode == 0 && b >= p.taper_from) {   // a half or quarter row of one of the last clips (launch_stft)
    int r = b - p.taper_from;
    if (r < p.taper_half_rows) {
      b = p.taper_from + (r >> 1);
      iters >>= 1;
      bx += (r & 1) * gridDim.x;
    } else {
      r -= p.taper_half_rows;
      b = p.taper_from + (p.taper_half_rows >> 1) + (r >> 2);
      iters >>= 2;
      bx += (r & 3) * gridDim.x;
    }
  }
  const int len = (int)(p.lengths ? p.lengths[b] : p.max_samples);
  const int frames_b = num_frames(len);
  const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
  const float* __restrict__ x = p.wave + (long long)b * p.wave_stride;
  const float2* __restrict__ tw = p.tw32;
  const float4* st = p.stat4 ? p.stat4 + (long long)b * p.stat4_clip_stride + lane : nullptr;
  const long long plane = (long long)(p.out.layout == AST_LAYOUT_FLAT ? p.out.dim1 : p.out.window) * p.out.f_row;
  float n_acc = 0.f;
  if (kMode == 1)
    for (int k = lane; k < kAccStride; k += 32) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#if AST_STFT_SMEM_PF
  float* pf = reinterpret_cast<float*>(stft_smem + kStftSmem) + warp * kPfFloats;
  uint64_t* pf_bar = reinterpret_cast<uint64_t*>(stft_smem + kStftSmem + sizeof(float) * kStftWarps * kPfFloats) + warp;
  const bool pf_ok = kMode == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  uint32_t pf_phase = 0;
  bool pf_pending = false;
  // interior pair (no reflection, both frames live): its 1280 samples are one contiguous, 1 KB-aligned run of the clip
  auto pf_interior = [&](int pr) { return pr < p.pairs_per_clip && 2 * pr >= 2 && (2 * pr + 1) * kHop + kNfft / 2 <= len; };
  auto pf_issue = [&](int pr) {
    if (lane == 0) {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(pf_bar), dst = (uint32_t)__cvta_generic_to_shared(pf);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(kPfFloats * 4)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(x + 2 * pr * kHop - kNfft / 2), "r"((uint32_t)(kPfFloats * 4)), "r"(bar) : "memory");
    }
  };
  if (pf_ok) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(pf_bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int first = bx * iters * kStftWarps + warp;
    if (iters > 0 && pf_interior(first)) pf_issue(first), pf_pending = true;
  }
#endif

  // Hann from the lane's rotation: w[32 i + lane] = 0.5 - 0.5 cos(2 pi i / 32 + theta), theta = 2 pi lane / 1024, expanded
  // with the compile-time cos / sin of 2 pi i / 32 against (cos theta, sin theta) = conj(W_1024^lane): two FFMA, no load
  const float2 w1 = __ldg(tw + 32 + lane);
  const float cos_t = w1.x, sin_t = -w1.y;
#ifdef AST_STFT_TRACE
  long long t_prev_ = clock64();
  const unsigned cta_lin_ = blockIdx.x + gridDim.x * blockIdx.y;
  if (kMode == 0 && threadIdx.x == 0 && cta_lin_ < 4096) {
    unsigned smid_;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid_));
    g_stft_cta_ns[cta_lin_][0] = stft_global_ns();
    g_stft_cta_ns[cta_lin_][2] = smid_;
  }
#endif
  for (int it = 0; it < iters; ++it) {
    STFT_STAMP(0);   // loop overhead / previous pair's tail
    const int pair = (bx * iters + it) * kStftWarps + warp;
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
#if AST_STFT_SMEM_PF
        if (pf_pending) {
          const uint32_t bar = (uint32_t)__cvta_generic_to_shared(pf_bar);
          uint32_t done = 0;
          while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(pf_phase) : "memory");
          pf_phase ^= 1;
#pragma unroll
          for (int i = 0; i < 40; ++i) xv[i] = pf[32 * i + lane];
          __syncwarp();   // every lane has its samples: the buffer may be refilled
          pf_pending = false;
        } else
#endif
        {
#pragma unroll
          for (int i = 0; i < 40; ++i) xv[i] = __ldg(xp + 32 * i);
        }
#if AST_STFT_SMEM_PF
        if (pf_ok && it + 1 < iters && pf_interior(pair + kStftWarps)) pf_issue(pair + kStftWarps), pf_pending = true;
#endif
#ifndef AST_STFT_PF
#define AST_STFT_PF (AST_STFT_SMEM_PF ? 0 : 1)
#endif
#if AST_STFT_PF == 1
        // this warp's next pair starts kStftWarps x 512 samples further; its first 768 samples are being read by the
        // CTA's other warps right now, the last 512 (16 lines) are new: one line per lane into the L1
        if (it + 1 < iters && lane < 16 && (ta + 2 * kStftWarps + 1) * kHop + kNfft / 2 <= len)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + 2 * kStftWarps * kHop + 768 - lane + 32 * lane));
#elif AST_STFT_PF == 2
        if (it + 1 < iters && (ta + 2 * kStftWarps + 1) * kHop + kNfft / 2 <= len) {
          const float* xn = xp - lane + 2 * kStftWarps * kHop;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(xn + 32 * lane));
          if (lane < 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(xn + 32 * (32 + lane)));
        }
#endif
#pragma unroll
        for (int i = 0; i < 32; ++i) {
#if !defined(AST_STFT_WIN_COMPUTED)
          const float w = __ldg(p.hann + lane + 32 * i);
#else
          const float w = hann_at(i, cos_t, sin_t);
#endif
#if defined(__CUDA_ARCH__)
          v[i] = __fmul2_rn(make_float2(xv[i], xv[i + 8]), make_float2(w, w));
#endif
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int idx = base + 32 * i;
#if !defined(AST_STFT_WIN_COMPUTED)
          const float w = __ldg(p.hann + lane + 32 * i);
#else
          const float w = hann_at(i, cos_t, sin_t);
#endif
          const float xa = load_reflect(x, idx, len, p.pad_zero);
          const float xb = live_b ? load_reflect(x, idx + kHop, len, p.pad_zero) : 0.f;
          v[i] = make_float2(xa * w, xb * w);
        }
      }
      STFT_STAMP(1);   // sample loads + window
#ifndef AST_STFT_NOFFT    // diagnostic build: loads, exchange and stores without the two in-register transforms
      fft32(v);
#endif
      STFT_STAMP(2);   // first transform
      if (kMode == 0) {   // the previous pair's bulk stores have read the tile they were staged in
        bulk_wait_read();   // (every lane: bulk groups are per thread, lanes 7 / 15 / 23 / 31 issued the copies)
        __syncwarp();
      }
      tile[lane] = v[0];
      {
        // the twiddle W_1024^(lane k1) that follows pass 1, from ten table entries instead of 31:
        // wb[b] = W^(lane b), b = 1..7; wa[a] = W^(8 a lane), a = 1..3; W^(lane k1) = wa[k1 / 8] wb[k1 % 8]
        // (one multiplication of two correctly rounded entries: one more rounding than a direct table entry)
#if !defined(AST_STFT_TW_LADDER)
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) tile[k1 * kTileStride + lane] = cmul(v[k1], __ldg(tw + k1 * 32 + lane));
#else
        float2 wb[8], wa[4];
#pragma unroll
        for (int i = 1; i < 8; ++i) wb[i] = __ldg(tw + i * 32 + lane);
#pragma unroll
        for (int i = 1; i < 4; ++i) wa[i] = __ldg(tw + 8 * i * 32 + lane);
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) {
          const float2 w = (k1 & 7) == 0 ? wa[k1 >> 3] : k1 < 8 ? wb[k1] : cmul(wa[k1 >> 3], wb[k1 & 7]);
          tile[k1 * kTileStride + lane] = cmul(v[k1], w);
        }
#endif
      }
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) v[n2] = tile[lane * kTileStride + n2];
      __syncwarp();   // the tile may be rewritten by the next pair
      STFT_STAMP(3);   // twiddle + exchange
#ifndef AST_STFT_NOFFT
      fft32(v);
#endif
      STFT_STAMP(4);   // second transform
    }
    if (kMode == 1) {
      if (any_live) {
        const float cnt = live_b ? 2.f : 1.f, n_new = n_acc + cnt;
        const StftMoments emit{acc + lane, cnt / n_new, n_acc * cnt / n_new, live_b, StftStats{nullptr}};
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
#ifndef AST_STFT_PLAIN_STORES
    const bool single = !a1 && !b1, dup = la1 && lb1;
    if (any_live && la0 && lb0 && (single || dup)) {
      float* stage = reinterpret_cast<float*>(tile);   // free: pass 2 has read it (and the warp has synchronised)
      float* rows[4] = {a0, a0 + plane, b0, b0 + plane};
      int ph[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) ph[j] = (int)((reinterpret_cast<uintptr_t>(rows[j]) >> 2) & 3);
      float* s0 = stage + ph[0] + lane;
      float* s1 = stage + kStageRow + ph[1] + lane;
      float* s2 = stage + 2 * kStageRow + ph[2] + lane;
      float* s3 = stage + 3 * kStageRow + ph[3] + lane;
      if (single) {
        const StftStage<false> emit{s0, s1, s2, s3, nullptr, nullptr, nullptr, nullptr, StftStats{st}};
        separate_and_emit(v, lane, emit);
      } else {
        const StftStage<true> emit{s0, s1, s2, s3, re(a1), im(a1), re(b1), im(b1), StftStats{st}};
        separate_and_emit(v, lane, emit);
      }
      STFT_STAMP(5);   // separation + normalisation + staging
#ifdef AST_STFT_VECTOR_FLUSH   // diagnostic A/B: the staged rows leave as 16-byte vector stores instead of bulk copies
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* staged = stage + j * kStageRow + ph[j];
        const int head = (4 - ph[j]) & 3, body = (kFStft - head) & ~3, tail = kFStft - head - body;
        if (lane < head) rows[j][lane] = staged[lane];
        if (lane < tail) rows[j][head + body + lane] = staged[head + body + lane];
        const float4* s4 = reinterpret_cast<const float4*>(staged + head);
        float4* d4 = reinterpret_cast<float4*>(rows[j] + head);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (lane + 32 * i < body / 4) d4[lane + 32 * i] = s4[lane + 32 * i];
      }
      __syncwarp();
#else
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the bulk-copy engine
      __syncwarp();
      flush_rows(rows, stage, ph, lane);
      bulk_commit();   // (per thread: lanes 7, 15, 23, 31 have one copy each in their group, the others an empty one)
#endif
      STFT_STAMP(6);   // fence + flush
    } else
#endif
    if (any_live && la0 && lb0 && !a1 && !b1) {
      const StftStore<1> emit{re(a0), im(a0), nullptr, nullptr, re(b0), im(b0), nullptr, nullptr, true, false, true, false, StftStats{st}};
      separate_and_emit(v, lane, emit);
    } else if (any_live && la0 && lb0 && la1 && lb1) {
      const StftStore<2> emit{re(a0), im(a0), re(a1), im(a1), re(b0), im(b0), re(b1), im(b1), true, true, true, true, StftStats{st}};
      separate_and_emit(v, lane, emit);
    } else if (any_live) {
      const StftStore<0> emit{re(a0), im(a0), re(a1), im(a1), re(b0), im(b0), re(b1), im(b1), la0, la1, lb0, lb1, StftStats{st}};
      separate_and_emit(v, lane, emit);
    } else {
      // both frames lie past the clip: their rows exist in the output and must be zeros
      const StftStore<0> emit{re(a0), im(a0), re(a1), im(a1), re(b0), im(b0), re(b1), im(b1), false, false, false, false, StftStats{nullptr}};
      for (int k = 0; k + lane < kFStft; k += 32) emit(k, make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float4(0.f, 0.f, 1.f, 1.f));
    }
  }

  if (kMode == 0) bulk_wait_all();   // shared memory must outlive the bulk copies that read it
#ifdef AST_STFT_TRACE
  if (kMode == 0 && threadIdx.x == 0 && cta_lin_ < 4096) g_stft_cta_ns[cta_lin_][1] = stft_global_ns();
#endif
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
  if (kMode == 0) AST_TIMELINE_STAMP(stft, blockIdx.x + gridDim.x * blockIdx.y, 1);   // (warp 0's end)
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
  AST_CUDA_TRY(cudaFuncSetAttribute(stft_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftFeatSmem));
  if (const char* env = getenv("AST_STFT_CARVEOUT"))   // diagnostic: shared-memory share of the unified L1 / shared array, %
    AST_CUDA_TRY(cudaFuncSetAttribute(stft_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env)));
  AST_CUDA_TRY(cudaFuncSetAttribute(stft_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftStatsSmem));
  int n = 0;
  AST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_kernel<0>, kStftThreads, kStftFeatSmem));
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

// clips whose rows are dealt as half / quarter rows at the end of the grid (launch_stft): none unless the grid runs over
// several waves of resident CTAs.  Measured on the chained feature call (64 clips x 10 s, 27 CTAs per clip, 592 resident):
// no taper 0.2563 ms, 8 half 0.2542, 16 half 0.2542, 32 half 0.2555, 8 half + 4 quarter 0.2538, 4 + 2 0.2531, 3 + 2 0.2530,
// 2 + 1 0.2537 (profiles/r2_v10_stft_taper_ab.txt): the quarter clips' CTAs amount to ~ 0.09 of a wave, the half clips' to ~ 0.18
static void stft_default_taper(const ast_plan* plan, int batch, int ctas_per_clip, int iters, int* half, int* quarter) {
  const long long slots = (long long)plan->sm_count * g_stft_ctas_per_sm;
  *half = *quarter = 0;
  if ((long long)ctas_per_clip * batch <= slots || ctas_per_clip <= 0) return;
  const double wave_clips = (double)slots / ctas_per_clip;   // clips whose CTAs fill the machine once
  int q = iters % 4 == 0 ? (int)(0.09 * wave_clips + 0.5) : 0;
  int h = (int)(0.18 * wave_clips + 0.5);
  if (q > batch / 8) q = batch / 8;
  if (h > batch / 4) h = batch / 4;
  *half = h;
  *quarter = q;
}

int stft_tiles_per_clip(const ast_plan* plan, int batch, int slots, bool stats_mode) {
  const long long pairs = (slots + 1) / 2, groups = (pairs + kStftWarps - 1) / kStftWarps;
  if (groups == 0 || batch == 0) return 0;
  // statistics mode: FIXED tiles of kStftStatsIters pairs per warp (128 frames), whatever the batch - a clip's partial
  // moments, and so its float32 rounding, must not depend on which clips share its launch (determinism across ranks)
  // feature mode: at most 4 pairs per warp - finer CTAs let the STFT fill the SMs sooner as the persistent CQT CTAs of
  // the chained call retire, and even out the 2x spread of CTA durations (measured: 12 pairs per warp 0.2993 ms per
  // step, 4: 0.2952, 3: 0.2942)
  const int it = stats_mode ? kStftStatsIters : pick_iters(groups, batch, (long long)plan->sm_count * g_stft_ctas_per_sm, 1, 4);
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
  // Tapered tail (feature mode): the last `taper` clips are dealt as two half rows each, with half the pairs per warp.
  // CTAs are dispatched in grid order, so these are the kernel's last CTAs: the SMs drain in half the time, and whatever
  // follows on the stream (the chained call's CQT projection, one persistent CTA per SM that needs the SM empty) starts
  // sooner - at the fixed cost of a CTA for the extra rows only.
  int half = 0, quarter = 0;   // clips dealt as half rows, then clips dealt as quarter rows
  if (!stats_mode && iters >= 2 && iters % 2 == 0) {
    stft_default_taper(plan, batch, (int)grid.x, (int)iters, &half, &quarter);
    if (const char* env = getenv("AST_STFT_TAPER")) {   // diagnostic override: "half[,quarter]" clips
      half = atoi(env), quarter = 0;
      if (const char* c = strchr(env, ',')) quarter = atoi(c + 1);
    }
    if (iters % 4 != 0) half += quarter, quarter = 0;
    if (half < 0) half = 0;
    if (quarter < 0) quarter = 0;
    if (quarter > batch) quarter = batch;
    if (half + quarter > batch) half = batch - quarter;
    while (batch + half + 3 * quarter > 65535) {   // rows of the grid
      if (quarter > 0) --quarter; else --half;
    }
  }
  p.taper_from = batch - half - quarter;
  p.taper_half_rows = 2 * half;
  grid.y = (unsigned)(batch + half + 3 * quarter);
  ProfileSpan span("stft_kernel", st);
  if (stats_mode) {
    if ((int)grid.x != tiles) return fail(AST_ERR_INVALID_ARG, "internal: statistics tile count mismatch");
    if (pdl) {   // behind the CQT projection of the statistics call, like the feature call's STFT
      AST_CUDA_TRY(launch_with_pdl(stft_kernel<1>, grid, kStftThreads, kStftStatsSmem, st, p));
    } else {
      stft_kernel<1><<<grid, kStftThreads, kStftStatsSmem, st>>>(p);
      AST_LAUNCH_CHECK("stft_kernel<stats>");
    }
  } else if (pdl) {
    // programmatic dependent of the CQT projection launched just before it on the same stream: the kernel never
    // waits for it (disjoint output columns), so its CTAs fill the SMs as the persistent CQT CTAs retire
    AST_CUDA_TRY(launch_with_pdl(stft_kernel<0>, grid, kStftThreads, kStftFeatSmem, st, p));
  } else {
    stft_kernel<0><<<grid, kStftThreads, kStftFeatSmem, st>>>(p);
    AST_LAUNCH_CHECK("stft_kernel");
  }
  return AST_OK;
}

}  // namespace ast
