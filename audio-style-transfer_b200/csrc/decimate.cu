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
#include "umma.cuh"
#include <cstring>

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


// ------------------------------------------------------------------------------------------------
// Tensor-core formulation (tcgen05, kind::tf32 with a 3-term split for FP32-level accuracy).
//
// 32 consecutive outputs form one GEMM row:  y[32 r + u] = sum_{n < 448} x[64 r - 192 + n] * g[n - 2 u],
// i.e. Y[128 x 32] = A[128 x 448] * G[448 x 32] per tile, with A's rows overlapping windows (hop 64) of the
// signal.  Neither operand is materialised:
//  * A: the signal segment is stored once, as 16-byte chunks in "chunk-column" order smem[e][R] (e = chunk
//    within a 64-sample row, R = row).  In the no-swizzle K-major canonical layout rows are 16 B apart
//    (SBO = 128 B) and the two K chunks of an MMA are LBO = RT * 16 B apart, so the window shift by d rows
//    (n = 64 d + 4 e + k) is just +16 d bytes on the descriptor's start address: 7 x 8 K-steps, no im2col.
//  * G: Toeplitz, G_ks[u][k] = g[8 ks + k - 2 u] = T[u - 4 ks][k] with ONE strip T[jj][k] = g[k - 2 jj]
//    (252 rows, 8 KB), so B_ks is the strip at start address + (220 - 4 ks) * 16 B.
// FP32 accuracy: x = hi + lo (hi = TF32 truncation), g = hi + lo (precomputed from the double taps);
// D += hi*hi + lo*hi + hi*lo, FP32 accumulation in TMEM; the dropped lo*lo term is 2^-22 relative.
// Accuracy note (measured on B200): the tensor core's FP32 accumulate rounds TOWARD ZERO, so a chain of n
// accumulating MMAs shrinks the result by ~n * 1.4e-8 (168 steps: -2.35e-6 gain per stage, -1.4e-5 after
// six stages).  The hi*hi terms therefore go to 4 interleaved accumulators (14 steps each) and the small
// cross terms to a fifth; the epilogue adds the five in registers with round-to-nearest.
namespace tc {
constexpr int kM = 128;                       // rows (of 32 outputs) per tile
constexpr int kNB = 32;
constexpr int kRowHop = 2 * kNB;              // 64 input samples between rows
constexpr int kCh = kRowHop / 4;              // 16 chunks per row
constexpr int kShifts = 7;                    // window = 7 rows = 448 samples >= 385 + 2 * 31
constexpr int kRowsUsed = kM + kShifts - 1;   // 134
constexpr int kRT = 135;                      // rows per chunk column (odd multiple -> conflict-free transposed stores)
constexpr int kKSteps = kShifts * kCh / 2;    // 56 MMAs of K = 8 per operand pair
constexpr int kStripRows = 256;
constexpr int kStripRow0 = 4 * (kKSteps - 1); // 220: strip row of (u = 0, ks = 0)
constexpr int kThreads = 256;
constexpr int kAFloats = kCh * kRT * 4;
constexpr int kTFloats = 2 * kStripRows * 4;
constexpr size_t kSmem = sizeof(float) * (2 * kAFloats + 2 * kTFloats) + 64;
constexpr int kMainAcc = 4;                   // hi*hi accumulators (K-steps interleaved), see the accuracy note below
constexpr int kTmemCols = 256;                // (kMainAcc + 1) * 32 = 160 columns used
constexpr int kLoadIters = (kRowsUsed * kCh + kThreads - 1) / kThreads;  // 17 chunks per thread
}  // namespace tc

struct DecimateTcParams {
  const float* in;
  long long in_stride;
  float* out;
  long long out_stride;
  const int32_t* lengths;
  long long max_samples;
  int in_octave;
  int batch;
  int tiles_per_clip;
  bool vec_ok;
  const float* strip_hi;   // [2][256][4] smem image of the Toeplitz strip (TF32-exact values)
  const float* strip_lo;
};

// what one tile reads / writes, and this thread's share of the staging: chunk i is read at s0 + i * 1024 samples
// and stored at slot0 + 16 i (u = tid + 256 i -> row R = (tid >> 4) + 16 i, chunk e = tid & 15)
struct DecTile {
  const float* x;
  float* y;
  int len_in, len_out, row0;
  int s0, slot0, n;
  bool live;      // false: the tile lies past the clip's end (ragged batch), nothing to do
  bool interior;  // whole segment inside [0, len_in): no bounds checks
};

__device__ __forceinline__ DecTile decode_dec_tile(const DecimateTcParams& p, int tile, int tid) {
  using namespace tc;
  DecTile t;
  const int b = tile / p.tiles_per_clip;
  t.row0 = (tile - b * p.tiles_per_clip) * kM;
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  t.len_in = (int)((len0 + (1LL << p.in_octave) - 1) >> p.in_octave);
  t.len_out = (t.len_in + 1) >> 1;
  t.live = t.row0 * kNB < t.len_out;
  t.x = p.in + (long long)b * p.in_stride;
  t.y = p.out + (long long)b * p.out_stride;
  const int s_base = kRowHop * t.row0 - kDecHalf;
  t.s0 = s_base + kRowHop * (tid >> 4) + 4 * (tid & 15);
  t.slot0 = (tid & 15) * kRT + (tid >> 4);
  t.n = tid < kRowsUsed * kCh ? (kRowsUsed * kCh - tid + kThreads - 1) / kThreads : 0;
  t.interior = s_base >= 0 && s_base + kRowsUsed * kRowHop <= t.len_in && p.vec_ok;
  return t;
}

__device__ __forceinline__ void prefetch_dec_tile(const DecTile& t, bool vec_ok, float4 (&v)[tc::kLoadIters]) {
  using namespace tc;
  if (!t.live) return;
  if (t.interior) {
#pragma unroll
    for (int i = 0; i < kLoadIters; ++i)
      if (i < t.n) v[i] = __ldg(reinterpret_cast<const float4*>(t.x + t.s0 + i * (16 * kRowHop)));
  } else {
#pragma unroll
    for (int i = 0; i < kLoadIters; ++i)
      if (i < t.n) v[i] = umma::load4_zero_ext(t.x, t.s0 + i * (16 * kRowHop), t.len_in, vec_ok);
  }
}

__global__ void __launch_bounds__(tc::kThreads, 2) decimate2_tc_kernel(const DecimateTcParams p) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* a_hi = reinterpret_cast<float*>(smem_raw);
  float* a_lo = a_hi + kAFloats;
  float* t_hi = a_lo + kAFloats;
  float* t_lo = t_hi + kTFloats;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(t_lo + kTFloats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kTFloats / 4; i += kThreads) {
    reinterpret_cast<float4*>(t_hi)[i] = __ldg(reinterpret_cast<const float4*>(p.strip_hi) + i);
    reinterpret_cast<float4*>(t_lo)[i] = __ldg(reinterpret_cast<const float4*>(p.strip_lo) + i);
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) umma::mbar_init(mbar, 1);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma::instr_desc_tf32(kM, kNB);
  const uint32_t a_hi_addr = umma::smem_u32(a_hi), a_lo_addr = umma::smem_u32(a_lo);
  const uint32_t t_hi_addr = umma::smem_u32(t_hi), t_lo_addr = umma::smem_u32(t_lo);

  uint32_t phase = 0;
  const int total = p.tiles_per_clip * p.batch;  // gridDim.x <= total
  // software pipeline: the next tile's global loads are in flight while the tensor core works on this one
  float4 v[kLoadIters];
  DecTile cur = decode_dec_tile(p, blockIdx.x, tid);
  prefetch_dec_tile(cur, p.vec_ok, v);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const DecTile t = cur;
    // ---- split into TF32 hi / lo and store in chunk-column order: slot [e][R]
    if (t.live) {
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i)
        if (i < t.n) {
          float4 h, l;
          umma::split_tf32(v[i], h, l);
          reinterpret_cast<float4*>(a_hi)[t.slot0 + 16 * i] = h;
          reinterpret_cast<float4*>(a_lo)[t.slot0 + 16 * i] = l;
        }
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();

    // ---- 56 K-steps x 3 split terms.  Warp 0 runs the issue code converged (descriptors are warp-uniform,
    // offsets compile-time constants after unrolling); one elected lane issues.
    if (warp == 0 && t.live) {
      umma::fence_after_thread_sync();
      if (umma::elect_one_sync()) {
        const uint64_t da_hi0 = umma::smem_desc(a_hi_addr, kRT * 16, 128);
        const uint64_t da_lo0 = umma::smem_desc(a_lo_addr, kRT * 16, 128);
        const uint64_t db_hi0 = umma::smem_desc(t_hi_addr, kStripRows * 16, 128);
        const uint64_t db_lo0 = umma::smem_desc(t_lo_addr, kStripRows * 16, 128);
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks) {
          const int d = ks >> 3, e = (ks & 7) * 2;
          const uint64_t a_off = (uint64_t)((16 * d + e * kRT * 16) >> 4);  // start-address field is in 16-byte units
          const uint64_t t_off = (uint64_t)(kStripRow0 - 4 * ks);
          umma::mma_tf32(tmem_base + 32u * (ks & (kMainAcc - 1)), da_hi0 + a_off, db_hi0 + t_off, idesc,
                         ks >= kMainAcc ? 1u : 0u);
          umma::mma_tf32(tmem_base + 32u * kMainAcc, da_lo0 + a_off, db_hi0 + t_off, idesc, ks > 0 ? 1u : 0u);
          umma::mma_tf32(tmem_base + 32u * kMainAcc, da_hi0 + a_off, db_lo0 + t_off, idesc, 1u);
        }
        umma::commit(mbar);
      }
    }
    __syncwarp();

    if (tile + (int)gridDim.x < total) {
      cur = decode_dec_tile(p, tile + gridDim.x, tid);
      prefetch_dec_tile(cur, p.vec_ok, v);
    }
    if (!t.live) continue;  // CTA-uniform

    umma::mbar_wait(mbar, phase);
    phase ^= 1;
    umma::fence_after_thread_sync();

    // ---- epilogue: GEMM row 32 (warp & 3) + lane = 32 consecutive outputs; warps 0-3 take columns 0..15,
    //      warps 4-7 columns 16..31
    const int part = warp >> 2;
    float acc[16];
    {
      const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 16u * part;
      float m1[16], m2[16];
      umma::tmem_ld_32x16(lane_base, acc);
      umma::tmem_ld_32x16(lane_base + 32, m1);
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] += m1[c];
      umma::tmem_ld_32x16(lane_base + 64, m2);
      umma::tmem_ld_32x16(lane_base + 96, m1);
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] += m2[c] + m1[c];
      umma::tmem_ld_32x16(lane_base + 128, m1);  // cross terms
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] += m1[c];
    }
    umma::fence_before_thread_sync();
    const int j0 = (t.row0 + (warp & 3) * 32 + lane) * kNB + 16 * part;
    float* __restrict__ y = t.y + j0;
    if (j0 + 16 <= t.len_out) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        reinterpret_cast<float4*>(y)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (j0 + c < t.len_out) y[c] = acc[c];
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, kTmemCols);
}

static int g_use_tc_decimator = 1;
void set_tc_decimator(int on) { g_use_tc_decimator = on; }

int decimate_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(decimate2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmem));
  return AST_OK;
}

// host: the Toeplitz strip images, values exactly representable in TF32 (hi) / TF32-truncated residual (lo)
void host_decimator_strip(const double* taps_scaled, float* strip_hi, float* strip_lo) {
  using namespace tc;
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < kStripRows; ++i)
      for (int kk = 0; kk < 4; ++kk) {
        const int k = 4 * c + kk;
        const int tap = k - 2 * (i - kStripRow0);
        double g = (tap >= 0 && tap < kDecTaps && i <= kStripRow0 + kNB - 1) ? taps_scaled[tap] : 0.0;
        float gf = (float)g;
        uint32_t hb;
        memcpy(&hb, &gf, 4);
        hb = umma::tf32_trunc_bits(hb);
        float hi;
        memcpy(&hi, &hb, 4);
        float lo = (float)(g - (double)hi);
        uint32_t lb;
        memcpy(&lb, &lo, 4);
        lb = umma::tf32_trunc_bits(lb);
        memcpy(&lo, &lb, 4);
        strip_hi[(c * kStripRows + i) * 4 + kk] = hi;
        strip_lo[(c * kStripRows + i) * 4 + kk] = lo;
      }
}

int upload_decimator_taps(const float* taps_scaled) {
  float padded[kDecTaps + 7] = {0};
  for (int i = 0; i < kDecTaps; ++i) padded[i] = taps_scaled[i];
  AST_CUDA_TRY(cudaMemcpyToSymbol(c_dec_taps, padded, sizeof(padded)));
  return AST_OK;
}

// Runs the six stages: octave buffer i (1..6) of clip b lives at ws + b * ws_clip_stride + octave_offset(i).
int launch_decimate_cascade(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch,
                            long long max_samples, long long wave_stride, float* ws, long long ws_clip_stride,
                            cudaStream_t st) {
  if ((batch > 1 && (wave_stride & 1)) || (reinterpret_cast<uintptr_t>(wave) & 7))
    return fail(AST_ERR_INVALID_ARG, "wave rows must be 8-byte aligned (even stride) for the decimator");
  for (int i = 0; i < kOctaves - 1; ++i) {
    const float* in = i == 0 ? wave : ws + octave_offset(max_samples, i);
    const long long in_stride = i == 0 ? wave_stride : ws_clip_stride;
    float* out = ws + octave_offset(max_samples, i + 1);
    const long long len_out = octave_len(max_samples, i + 1);
    if (g_use_tc_decimator) {
      DecimateTcParams p;
      p.in = in;
      p.in_stride = in_stride;
      p.out = out;
      p.out_stride = ws_clip_stride;
      p.lengths = lengths;
      p.max_samples = max_samples;
      p.in_octave = i;
      p.batch = batch;
      const long long rows = (len_out + tc::kNB - 1) / tc::kNB;
      p.tiles_per_clip = (int)((rows + tc::kM - 1) / tc::kM);
      p.vec_ok = i > 0 || ((wave_stride % 4 == 0 || batch == 1) && (reinterpret_cast<uintptr_t>(wave) & 15) == 0);
      p.strip_hi = plan->d_dec_strip_hi;
      p.strip_lo = plan->d_dec_strip_lo;
      long long ctas = (long long)p.tiles_per_clip * batch;
      const long long cap = 2LL * plan->sm_count;
      if (ctas > cap) ctas = cap;
      ProfileSpan span("decimate2_tc_kernel", st);
      decimate2_tc_kernel<<<(unsigned)ctas, tc::kThreads, tc::kSmem, st>>>(p);
      AST_LAUNCH_CHECK("decimate2_tc_kernel");
    } else {
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
