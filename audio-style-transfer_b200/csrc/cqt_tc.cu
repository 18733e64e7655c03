// cqt_tc.cu - K3 on the tensor cores: the constant-Q projection (the one dense contraction of the path)
// as tcgen05.mma kind::tf32 with a split for FP32-level accuracy, warp-specialised and software-pipelined.
//
//     resp_i[t, j] = sum_{n < 256} x_i[t h_i - 128 + n] K[n][j],   h_i = 256 >> i,  j < 24 (12 re, 12 im)
//
// (derivation in cqt.cu / plan.cu; replaces librosa.cqt's per-octave STFT + sparse basis product reached
// through utilityFunctions.py:52).  A tile is 128 frames of one (clip, octave): D[128 x 32] = A[128 x 256] *
// B[256 x 32], run as four passes over 64-sample slices of the window so that a staged A slice is at most
// 37 KB per split term whatever the hop:
//   * hop >= 64 (octaves 0-2): the slice rows do not overlap; slot [chunk c'][row r].
//   * hop <  64 (octaves 3-6): the slice is one contiguous run of the decimated signal, stored as rows of
//     m = hop / 4 chunks in chunk-column order [e][R]; frame r's chunk c' = m d + e sits at row r + d, and in
//     the no-swizzle K-major layout (rows 16 B apart) that shift is +16 d bytes on the descriptor start
//     address - the overlapping frames are never materialised.  For hop 4 the raw signal IS the operand.
//
// Split precision: x = hi + lo (hi = TF32 truncation, lo exact residual), K = hi + lo (host, from doubles).
// B is stored as [B_hi | B_lo] (N = 64) so ONE MMA yields hi*hi (columns 0..31) and hi*lo (columns 32..63) for
// one A fetch - measured on B200 an M128 K8 TF32 MMA costs ~64 cycles of A-operand fetch whatever N <= 128 - and a
// second MMA adds lo*hi into columns 32..63.  The tensor core accumulates with round-toward-zero (measured
// -4.2e-8 relative per accumulating MMA), so K-steps rotate over four accumulators (8 steps each); the epilogue
// sums them in registers with round-to-nearest, applies the per-bin scale and (x - mean) * rstd, and scatters to
// the flat / section layout (columns 513..596).
//
// One persistent CTA per SM, 13 warps:
//   warps 0-7   producers: global -> registers -> hi / lo split -> shared A stage (2 stages, mbarrier full / empty)
//   warps 8-11  epilogue : TMEM -> registers -> global, one TMEM lane quadrant each (2 accumulator sets, ping-pong)
//   warp  12    MMA issue: one elected lane, tcgen05.commit releases A stages and publishes accumulator sets
#include <cstring>
#include <vector>

#include "common.cuh"
#include "umma.cuh"

namespace ast {
namespace cqt_tc {
constexpr int kM = 128;                 // frames per tile
constexpr int kN = 32;                  // 24 outputs padded to 32
constexpr int kPasses = 4;              // 64-sample slices of the 256-sample window
constexpr int kPassChunks = 16;         // 16-byte chunks per slice
constexpr int kKStepsPerPass = 8;
constexpr int kKSteps = kPasses * kKStepsPerPass;  // 32
constexpr int kRT = 145;                // rows per chunk column (>= 143, = 1 mod 8: conflict-free transposed stores)
constexpr int kProducers = 256;         // threads of warps 0-7
constexpr int kEpilogueWarp0 = 8;
constexpr int kMmaWarp = 12;
constexpr int kThreads = 13 * 32;
constexpr int kAFloats = kPassChunks * kRT * 4;           // 9280 floats = 37 120 B per split term
constexpr int kStageFloats = 2 * kAFloats;                // hi + lo
constexpr int kBStepFloats = 2 * 2 * kN * 4;              // one K-step of [B_hi | B_lo]: [c 2][j 64][4] = 512 floats
constexpr int kBFloats = kKSteps * kBStepFloats;          // 16384 floats = 64 KB
constexpr int kMainAcc = 4;
constexpr int kSetCols = kMainAcc * 2 * kN;               // 256 TMEM columns per accumulator set
constexpr int kTmemCols = 512;
constexpr int kStage = kM * kPassChunks / kProducers;     // 8 chunks per producer thread per pass at most
constexpr int kEpiStride = 25;            // floats per staged row (24 values + 1: conflict-free row-per-lane writes)
constexpr int kEpiFloats = 4 * 32 * kEpiStride;  // one [32 rows][25] transpose buffer per epilogue warp
constexpr size_t kSmem = sizeof(float) * (2 * kStageFloats + kBFloats + kEpiFloats) + 128;
}  // namespace cqt_tc

struct CqtTcParams {
  const float* wave;
  long long wave_stride;
  const float* ws;
  long long ws_clip_stride;
  long long oct_off[kOctaves];
  const int32_t* lengths;
  long long max_samples;
  int batch, slots, tiles_per_clip_oct, overlap;
  const float* bmat;   // [32 ks][2 c][64 j: hi then lo][4] smem image
  const float* scale;  // [7][12]
  bool vec_ok;
  OutSpec out;
};

// One producer thread's share of a (tile, pass) slice: chunk i (0..n-1) is read at s0 + i * src_step (samples)
// and stored at slot0 + i * slot_step (16-byte units).
struct StagePlan {
  const float* x;
  int len;
  int s0, src_step;
  int slot0, slot_step;
  int n;
  int s_base, span;  // first sample and extent of the slice (for the interior test of the following passes)
  bool vec_ok;
  bool interior;  // whole slice inside [0, len): no bounds checks
};

__device__ __forceinline__ void decode_tile(const CqtTcParams& p, int tile, int& b, int& oct, int& t0) {
  const int tiles_per_clip = p.tiles_per_clip_oct * kOctaves;
  b = tile / tiles_per_clip;
  const int rem = tile - b * tiles_per_clip;
  oct = rem / p.tiles_per_clip_oct;
  t0 = (rem - oct * p.tiles_per_clip_oct) * cqt_tc::kM;
}

__device__ __forceinline__ StagePlan plan_stage(const CqtTcParams& p, int tile, int pass, int tid) {
  using namespace cqt_tc;
  StagePlan s;
  int b, oct, t0;
  decode_tile(p, tile, b, oct, t0);
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  const int hop = kHop >> oct, m = hop >> 2;
  s.len = (int)((len0 + (1LL << oct) - 1) >> oct);
  s.x = oct == 0 ? p.wave + (long long)b * p.wave_stride : p.ws + (long long)b * p.ws_clip_stride + p.oct_off[oct];
  s.vec_ok = oct == 0 ? p.vec_ok : true;
  const int s_base = t0 * hop - kCqtNfft / 2 + 64 * pass;
  int span;  // samples covered by the slice
  if (m >= kPassChunks) {
    // chunk u = tid + 256 i -> row r = (tid >> 4) + 16 i, chunk c' = tid & 15: sample r hop + 4 c', slot [c'][r]
    s.s0 = s_base + (tid >> 4) * hop + 4 * (tid & 15);
    s.src_step = 16 * hop;
    s.slot0 = (tid & 15) * kRT + (tid >> 4);
    s.slot_step = 16;
    s.n = kStage;
    span = (kM - 1) * hop + 64;
  } else {
    // contiguous chunk u = tid + 256 i -> row R = u / m, column e = u % m (m divides 256): slot [e][R]
    const int lg = 6 - oct;  // log2(m)
    s.s0 = s_base + 4 * tid;
    s.src_step = 4 * kProducers;
    s.slot0 = (tid & (m - 1)) * kRT + (tid >> lg);
    s.slot_step = kProducers >> lg;
    const int n_chunks = (kM - 1) * m + kPassChunks;
    s.n = tid < n_chunks ? (n_chunks - tid + kProducers - 1) / kProducers : 0;
    span = 4 * n_chunks;
  }
  s.s_base = s_base;
  s.span = span;
  s.interior = s_base >= 0 && s_base + span <= s.len && s.vec_ok;
  return s;
}

// The 8 K-steps of one pass for octave OCT, issued by one elected lane.  With OCT a template parameter every
// descriptor offset is a compile-time constant (tight UIADD3 + UTCHMMA sequences, no address arithmetic at run time).
template <int OCT>
__device__ __forceinline__ void issue_pass(uint32_t a_hi_addr, uint32_t b_addr, uint32_t acc_set, int pass,
                                           uint32_t idesc64, uint32_t idesc32) {
  using namespace cqt_tc;
  constexpr int m = (kHop >> OCT) >> 2;
  constexpr uint32_t lbo = m == 1 ? 16u : (uint32_t)kRT * 16u;
  const uint64_t da_hi0 = umma::smem_desc(a_hi_addr, lbo, 128);
  const uint64_t da_lo0 = umma::smem_desc(a_hi_addr + kAFloats * 4, lbo, 128);
  const uint64_t db0 = umma::smem_desc(b_addr, 2 * kN * 16, 128) + (uint64_t)(pass * kKStepsPerPass * (kBStepFloats / 4));
  const int g0 = pass * kKStepsPerPass;
  // Back-to-back MMAs on the same TMEM columns serialise on the accumulate dependency (measured: ~230 cycles per
  // small MMA when dependent), so the 16 MMAs of a pass are ordered to keep dependent ones four issues apart:
  // first the eight hi * [hi | lo] products rotating over the four accumulators, then the eight lo * hi products.
#pragma unroll
  for (int term = 0; term < 2; ++term) {
#pragma unroll
    for (int ks = 0; ks < kKStepsPerPass; ++ks) {
      const int c = 2 * ks;  // first window chunk of the K-step inside this slice
      int a_units;           // start-address offset in 16-byte units
      if (m >= kPassChunks) {
        a_units = c * kRT;
      } else if (m == 1) {
        a_units = c;
      } else {
        a_units = (c / m) + (c % m) * kRT;
      }
      const uint64_t a_off = (uint64_t)a_units, b_off = (uint64_t)(ks * (kBStepFloats / 4));
      // K-steps rotate over the four accumulators; kKStepsPerPass is a multiple of 4 so the slot is ks & 3
      const uint32_t acc = acc_set + (uint32_t)((ks & (kMainAcc - 1)) * 2 * kN);
      if (term == 0)  // hi * [hi | lo] -> columns 0..63 of the accumulator
        umma::mma_tf32(acc, da_hi0 + a_off, db0 + b_off, idesc64, (g0 + ks) >= kMainAcc ? 1u : 0u);
      else            // lo * hi -> columns 32..63
        umma::mma_tf32(acc + kN, da_lo0 + a_off, db0 + b_off, idesc32, 1u);
    }
  }
}

__global__ void __launch_bounds__(cqt_tc::kThreads, 1) cqt_tc_kernel(const CqtTcParams p) {
  using namespace cqt_tc;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* a_stage = reinterpret_cast<float*>(smem_raw);            // [2 stages][hi | lo]
  float* b_img = a_stage + 2 * kStageFloats;                      // 64 KB
  float* epi_buf = b_img + kBFloats;                              // [4 warps][32][25] epilogue transpose
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kEpiFloats);
  uint64_t* full = bars;            // [2] producers -> MMA   (8 arrivals: one per producer warp)
  uint64_t* empty = bars + 2;       // [2] MMA -> producers   (tcgen05.commit)
  uint64_t* acc_full = bars + 4;    // [2] MMA -> epilogue    (tcgen05.commit)
  uint64_t* acc_empty = bars + 6;   // [2] epilogue -> MMA    (4 arrivals: one per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // B images: resident for the CTA's lifetime
  for (int i = tid; i < kBFloats / 4; i += kThreads)
    reinterpret_cast<float4*>(b_img)[i] = __ldg(reinterpret_cast<const float4*>(p.bmat) + i);
  if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, kProducers / 32);
      umma::mbar_init(empty + i, 1);
      umma::mbar_init(acc_full + i, 1);
      umma::mbar_init(acc_empty + i, 4);
    }
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int total = p.tiles_per_clip_oct * kOctaves * p.batch;  // gridDim.x <= total

  if (warp < kProducers / 32) {
    // ================================================================= producers
    // The loads of item k + 1 are issued before item k is split and stored, so global-memory latency hides
    // behind the store phase and the wait for the stage to be released.
    float4 v[kStage], vn[kStage];
    auto issue_loads = [&](const StagePlan& sp, float4 (&dst)[kStage]) {
      if (sp.interior) {
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) dst[i] = __ldg(reinterpret_cast<const float4*>(sp.x + sp.s0 + i * sp.src_step));
      } else {
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) dst[i] = umma::load4_zero_ext(sp.x, sp.s0 + i * sp.src_step, sp.len, sp.vec_ok);
      }
    };
    int item = 0;
    StagePlan sp = plan_stage(p, blockIdx.x, 0, tid);
    issue_loads(sp, v);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int pass = 0; pass < kPasses; ++pass, ++item) {
        const int s = item & 1;
        // next item: same tile, next 64-sample slice (addresses advance by 64 samples) or the next tile's first
        StagePlan spn = sp;
        bool have_next = true;
        if (pass + 1 < kPasses) {
          spn.s0 += 64;
          const int s_base_next = sp.s_base + 64;
          spn.s_base = s_base_next;
          spn.interior = s_base_next >= 0 && s_base_next + sp.span <= sp.len && sp.vec_ok;
        } else if (tile + (int)gridDim.x < total) {
          spn = plan_stage(p, tile + gridDim.x, 0, tid);
        } else {
          have_next = false;
        }
        if (have_next) issue_loads(spn, vn);
        umma::mbar_wait(empty + s, ((item >> 1) & 1) ^ 1);  // the MMAs that read this stage two items ago are done
        float* a_hi = a_stage + s * kStageFloats;
        float* a_lo = a_hi + kAFloats;
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) {
            float4 h, l;
            umma::split_tf32(v[i], h, l);
            reinterpret_cast<float4*>(a_hi)[sp.slot0 + i * sp.slot_step] = h;
            reinterpret_cast<float4*>(a_lo)[sp.slot0 + i * sp.slot_step] = l;
          }
        umma::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(full + s);
        sp = spn;
#pragma unroll
        for (int i = 0; i < kStage; ++i) v[i] = vn[i];
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issue
    const uint32_t idesc64 = umma::instr_desc_tf32(kM, 2 * kN), idesc32 = umma::instr_desc_tf32(kM, kN);
    const uint32_t b_addr = umma::smem_u32(b_img);
    int item = 0, n_tile = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++n_tile) {
      const int q = n_tile & 1;
      int b, oct, t0;
      decode_tile(p, tile, b, oct, t0);
      umma::mbar_wait(acc_empty + q, ((n_tile >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator set
      umma::fence_after_thread_sync();
      for (int pass = 0; pass < kPasses; ++pass, ++item) {
        const int s = item & 1;
        umma::mbar_wait(full + s, (item >> 1) & 1);
        umma::fence_after_thread_sync();
        if (umma::elect_one_sync()) {
          const uint32_t a_hi_addr = umma::smem_u32(a_stage + s * kStageFloats);
          const uint32_t acc_set = tmem_base + (uint32_t)(q * kSetCols);
          switch (oct) {
            case 0: issue_pass<0>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
            case 1: issue_pass<1>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
            case 2: issue_pass<2>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
            case 3: issue_pass<3>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
            case 4: issue_pass<4>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
            case 5: issue_pass<5>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
            default: issue_pass<6>(a_hi_addr, b_addr, acc_set, pass, idesc64, idesc32); break;
          }
          umma::commit(empty + s);                          // stage s may be overwritten once these MMAs finish
          if (pass == kPasses - 1) umma::commit(acc_full + q);  // ... and the accumulator set is complete
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= epilogue (warps 8-11)
    // TMEM holds one frame per lane; written that way every store instruction would touch 32 output rows
    // (32 L1 wavefronts for 128 B).  The 32 x 24 block is therefore transposed through shared memory and
    // stored with consecutive lanes on consecutive columns of a row (12-float runs, ~3 rows per instruction).
    const int quad = warp - kEpilogueWarp0;  // == warp % 4: the TMEM lane quadrant this warp may read
    float* stg = epi_buf + quad * 32 * kEpiStride;
    const long long clip_floats = p.out.layout == AST_LAYOUT_FLAT ? 2LL * p.out.dim1 * p.out.f_row
                                                                  : 2LL * p.out.dim1 * p.out.window * p.out.f_row;
    const long long plane = (long long)(p.out.layout == AST_LAYOUT_FLAT ? p.out.dim1 : p.out.window) * p.out.f_row;
    int n_tile = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++n_tile) {
      const int q = n_tile & 1;
      int b, oct, t0;
      decode_tile(p, tile, b, oct, t0);
      const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
      const int frames_b = num_frames(len0);
      const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
      const int col0 = kFCqt - kBinsPerOctave * (oct + 1);
      // lane c < 24 keeps the constants of output column c (12 re then 12 im): scale, mean, 1 / (std + eps)
      float sc_l = 0.f, mean_l = 0.f, rstd_l = 1.f;
      if (lane < kCqtCols) {
        const int j = lane < kBinsPerOctave ? lane : lane - kBinsPerOctave;
        sc_l = __ldg(p.scale + oct * kBinsPerOctave + j);
        if (p.out.stats) {
          const float2 m = __ldg(p.out.stats + (long long)b * p.out.stats_clip_stride + p.out.stats_off + col0 + j +
                                 (lane < kBinsPerOctave ? 0 : p.out.f_stats));
          mean_l = m.x, rstd_l = m.y;
        }
      }
      // destination rows of this lane's frame, as offsets from the clip's first output float
      float* const clip_out = p.out.out + (long long)b * clip_floats;
      const int t = t0 + quad * 32 + lane;
      long long off0 = 0, off1 = 0;
      int flags = 0;  // bit 0 / 1: row 0 / 1 live (data, else zeros); bit 2 / 3: row 0 / 1 present
      if (t < p.slots) {
        const RowDest d = row_dest(p.out, b, t, frames_b, sections_b);
        if (d.n > 0) off0 = d.row[0] - clip_out + col0, flags |= 4 | (d.live[0] ? 1 : 0);
        if (d.n > 1) off1 = d.row[1] - clip_out + col0, flags |= 8 | (d.live[1] ? 2 : 0);
      }

      umma::mbar_wait(acc_full + q, (n_tile >> 1) & 1);
      umma::fence_after_thread_sync();
      float acc[32], tmp[32];
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(q * kSetCols);
      umma::tmem_ld_32x32(lane_base, acc);
      umma::tmem_ld_32x32(lane_base + 2 * kN, tmp);
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += tmp[c];
      float acc2[32];
      umma::tmem_ld_32x32(lane_base + 4 * kN, acc2);
      umma::tmem_ld_32x32(lane_base + 6 * kN, tmp);
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += acc2[c] + tmp[c];
      // cross terms (hi*lo + lo*hi), the same four accumulators, columns 32..63
      umma::tmem_ld_32x32(lane_base + kN, acc2);
      umma::tmem_ld_32x32(lane_base + 3 * kN, tmp);
#pragma unroll
      for (int c = 0; c < 32; ++c) acc2[c] += tmp[c];
      umma::tmem_ld_32x32(lane_base + 5 * kN, tmp);
#pragma unroll
      for (int c = 0; c < 32; ++c) acc2[c] += tmp[c];
      umma::tmem_ld_32x32(lane_base + 7 * kN, tmp);
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += acc2[c] + tmp[c];
      umma::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(acc_empty + q);  // this warp's quadrant of the set is drained

      // scale + normalise (column constants broadcast from their lane), row-per-lane into the staging buffer
#pragma unroll
      for (int c = 0; c < kCqtCols; ++c) {
        const float sc = __shfl_sync(0xffffffffu, sc_l, c);
        const float mu = __shfl_sync(0xffffffffu, mean_l, c);
        const float rs = __shfl_sync(0xffffffffu, rstd_l, c);
        float v = acc[c] * sc;
        if (p.out.stats) v = (v - mu) * rs;
        stg[lane * kEpiStride + c] = v;
      }
      __syncwarp();
      // 32 rows x 12 columns per plane = 12 store rounds; lane l of round i owns element 32 i + l
#pragma unroll
      for (int i = 0; i < kBinsPerOctave; ++i) {
        const int idx = lane + 32 * i;
        const int r = idx / kBinsPerOctave, j = idx - r * kBinsPerOctave;
        const float re = stg[r * kEpiStride + j], im = stg[r * kEpiStride + kBinsPerOctave + j];
        const long long o0 = __shfl_sync(0xffffffffu, off0, r), o1 = __shfl_sync(0xffffffffu, off1, r);
        const int fl = __shfl_sync(0xffffffffu, flags, r);
        if (fl & 4) {
          clip_out[o0 + j] = (fl & 1) ? re : 0.f;
          clip_out[o0 + plane + j] = (fl & 1) ? im : 0.f;
        }
        if (fl & 8) {
          clip_out[o1 + j] = (fl & 2) ? re : 0.f;
          clip_out[o1 + plane + j] = (fl & 2) ? im : 0.f;
        }
      }
      __syncwarp();  // the staging buffer is rewritten by the next tile
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_base, kTmemCols);
}

// host: B image.  kmat[n][24] (double) -> [ks 32][c 2][j 64: hi 0..31 then lo 32..63][kk 4], n = 8 ks + 4 c + kk
void host_cqt_tc_images(const double* kmat_256x24, float* images) {
  using namespace cqt_tc;
  for (int ks = 0; ks < kKSteps; ++ks)
    for (int c = 0; c < 2; ++c)
      for (int j = 0; j < kN; ++j)
        for (int kk = 0; kk < 4; ++kk) {
          const int n = 8 * ks + 4 * c + kk;
          const double g = j < kCqtCols ? kmat_256x24[n * kCqtCols + j] : 0.0;
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
          const size_t base = ((size_t)ks * 2 + c) * (2 * kN) * 4;
          images[base + (size_t)j * 4 + kk] = hi;
          images[base + (size_t)(kN + j) * 4 + kk] = lo;
        }
}

int cqt_tc_image_floats() { return cqt_tc::kBFloats; }

int cqt_tc_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(cqt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cqt_tc::kSmem));
  return AST_OK;
}

int launch_cqt_tc(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
                  long long wave_stride, const float* ws, long long ws_clip_stride, const OutSpec& out, cudaStream_t st) {
  CqtTcParams p;
  p.wave = wave;
  p.wave_stride = wave_stride;
  p.ws = ws;
  p.ws_clip_stride = ws_clip_stride;
  for (int i = 0; i < kOctaves; ++i) p.oct_off[i] = i == 0 ? 0 : octave_offset(max_samples, i);
  p.lengths = lengths;
  p.max_samples = max_samples;
  p.batch = batch;
  p.slots = frame_slots(out.layout, out.dim1, out.window, out.step);
  p.tiles_per_clip_oct = (p.slots + cqt_tc::kM - 1) / cqt_tc::kM;
  p.overlap = out.window - out.step;
  p.bmat = plan->d_cqt_tc_images;
  p.scale = plan->d_cqt_scale;
  p.vec_ok = (wave_stride % 4 == 0 || batch == 1) && ((reinterpret_cast<uintptr_t>(wave) & 15) == 0);
  p.out = out;
  if (p.slots == 0 || batch == 0) return AST_OK;
  long long ctas = (long long)p.tiles_per_clip_oct * kOctaves * batch;
  if (ctas > plan->sm_count) ctas = plan->sm_count;  // persistent: one CTA per SM
  ProfileSpan span("cqt_tc_kernel", st);
  cqt_tc_kernel<<<(unsigned)ctas, cqt_tc::kThreads, cqt_tc::kSmem, st>>>(p);
  AST_LAUNCH_CHECK("cqt_tc_kernel");
  return AST_OK;
}

}  // namespace ast
