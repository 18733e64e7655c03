// cqt_tc.cu - K3 on the tensor cores: the constant-Q projection (the one dense contraction of the path)
// as tcgen05.mma kind::tf32 with a 3-term split for FP32-level accuracy.
//
//     resp_i[t, j] = sum_{n < 256} x_i[t h_i - 128 + n] K[n][j],   h_i = 256 >> i,  j < 24 (12 re, 12 im)
//
// (derivation in cqt.cu / plan.cu; replaces librosa.cqt's per-octave STFT + sparse basis product reached
// through utilityFunctions.py:52).  A tile is 128 frames of one (clip, octave): D[128 x 32] = A[128 x 256] *
// B[256 x 32], run as four passes over 64-sample slices of the window so that the staged A slice is at most
// 32 KB per split term whatever the hop:
//   * hop >= 64 (octaves 0-2): the slice rows do not overlap; slot [chunk c'][row r].
//   * hop <  64 (octaves 3-6): the slice is one contiguous run of the decimated signal, stored as rows of
//     m = hop / 4 chunks in chunk-column order [e][R]; frame r's chunk c' = m d + e sits at row r + d, and in
//     the no-swizzle K-major layout (rows 16 B apart) that shift is +16 d bytes on the descriptor start
//     address - the overlapping frames are never materialised.  For hop 4 the raw signal IS the operand.
// B (the CQT kernel, padded to 32 columns) is pre-split on the host into TF32 hi / lo images per pass.
// The tensor core accumulates with round-toward-zero (measured: -4.2e-8 relative per accumulating MMA), so
// the hi*hi terms are spread over four accumulators (8 steps each) and the cross terms go to a fifth;
// the epilogue sums them in registers, applies the per-bin scale and (x - mean) * rstd and scatters to the
// flat / section layout (columns 513..596).
//
// 256 threads: all eight warps stage (8 chunks per thread per pass, addresses advance by constant steps),
// warp 0 issues the MMAs, and in the epilogue warps 0-3 write the real plane and warps 4-7 the imaginary
// plane of their TMEM lane quadrant.  The global loads of the next (tile, pass) are issued before waiting
// for the current MMAs (register prefetch); B images are double-buffered with cp.async.
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
constexpr int kRT = 145;                // rows per chunk column (>= 143, = 1 mod 8: conflict-free transposed stores)
constexpr int kThreads = 256;
constexpr int kAFloats = kPassChunks * kRT * 4;           // 9280 floats = 37 120 B per split term
constexpr int kBFloats = kKStepsPerPass * 2 * kN * 4;     // 2048 floats = 8 KB per split term per pass
constexpr int kMainAcc = 4;
constexpr int kTmemCols = 256;
constexpr int kStage = kM * kPassChunks / kThreads;       // 8 chunks per thread per pass at most
constexpr size_t kSmem = sizeof(float) * (2 * kAFloats + 4 * kBFloats) + 64;  // A hi/lo + double-buffered B hi/lo
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
  const float* bmat;   // [4 passes][hi, lo][8 ks][2 c][32 j][4] smem images
  const float* scale;  // [7][12]
  bool vec_ok;
  OutSpec out;
};

// One thread's share of a (tile, pass) slice: chunk i (0..n-1) is read at src + i * src_step (samples) and
// stored at slot + i * slot_step (16-byte units).
struct StagePlan {
  const float* x;
  int len;
  int s0, src_step;
  int slot0, slot_step;
  int n;          // chunks this thread stages
  int oct, m;
  bool vec_ok;
  bool interior;  // whole slice inside [0, len): no bounds checks
};

__device__ __forceinline__ StagePlan plan_stage(const CqtTcParams& p, int tile, int pass, int tid) {
  using namespace cqt_tc;
  StagePlan s;
  const int tiles_per_clip = p.tiles_per_clip_oct * kOctaves;
  const int b = tile / tiles_per_clip;
  const int rem = tile - b * tiles_per_clip;
  s.oct = rem / p.tiles_per_clip_oct;
  const int t0 = (rem - s.oct * p.tiles_per_clip_oct) * kM;
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  const int hop = kHop >> s.oct;
  s.m = hop >> 2;
  s.len = (int)((len0 + (1LL << s.oct) - 1) >> s.oct);
  s.x = s.oct == 0 ? p.wave + (long long)b * p.wave_stride : p.ws + (long long)b * p.ws_clip_stride + p.oct_off[s.oct];
  s.vec_ok = s.oct == 0 ? p.vec_ok : true;
  const int s_base = t0 * hop - kCqtNfft / 2 + 64 * pass;
  int span;  // samples covered by the slice
  if (s.m >= kPassChunks) {
    // chunk u = tid + 256 i -> row r = (tid >> 4) + 16 i, chunk c' = tid & 15: sample r hop + 4 c', slot [c'][r]
    s.s0 = s_base + (tid >> 4) * hop + 4 * (tid & 15);
    s.src_step = 16 * hop;
    s.slot0 = (tid & 15) * kRT + (tid >> 4);
    s.slot_step = 16;
    s.n = kStage;
    span = (kM - 1) * hop + 64;
  } else {
    // contiguous chunk u = tid + 256 i -> row R = u / m, column e = u % m (m divides 256): slot [e][R]
    const int lg = 6 - s.oct;  // log2(m)
    s.s0 = s_base + 4 * tid;
    s.src_step = 4 * kThreads;
    s.slot0 = (tid & (s.m - 1)) * kRT + (tid >> lg);
    s.slot_step = kThreads >> lg;
    const int n_chunks = (kM - 1) * s.m + kPassChunks;
    s.n = tid < n_chunks ? (n_chunks - tid + kThreads - 1) / kThreads : 0;
    span = 4 * n_chunks;
  }
  s.interior = s_base >= 0 && s_base + span <= s.len && s.vec_ok;
  return s;
}

__device__ __forceinline__ void prefetch_slice(const StagePlan& s, float4 (&v)[cqt_tc::kStage]) {
  using namespace cqt_tc;
  if (s.interior) {
#pragma unroll
    for (int i = 0; i < kStage; ++i)
      if (i < s.n) v[i] = __ldg(reinterpret_cast<const float4*>(s.x + s.s0 + i * s.src_step));
  } else {
#pragma unroll
    for (int i = 0; i < kStage; ++i)
      if (i < s.n) v[i] = umma::load4_zero_ext(s.x, s.s0 + i * s.src_step, s.len, s.vec_ok);
  }
}

__device__ __forceinline__ void store_slice(const StagePlan& s, const float4 (&v)[cqt_tc::kStage], float* a_hi, float* a_lo) {
  using namespace cqt_tc;
#pragma unroll
  for (int i = 0; i < kStage; ++i)
    if (i < s.n) {
      float4 h, l;
      umma::split_tf32(v[i], h, l);
      reinterpret_cast<float4*>(a_hi)[s.slot0 + i * s.slot_step] = h;
      reinterpret_cast<float4*>(a_lo)[s.slot0 + i * s.slot_step] = l;
    }
}

// asynchronous copy of one pass's B images (hi then lo, 16 KB) into a B buffer
__device__ __forceinline__ void copy_b_async(const CqtTcParams& p, int pass, int tid, float* b_buf) {
  using namespace cqt_tc;
  const float4* src = reinterpret_cast<const float4*>(p.bmat) + (size_t)pass * 2 * (kBFloats / 4);
#pragma unroll
  for (int i = 0; i < 2 * kBFloats / 4 / kThreads; ++i)
    umma::cp_async_16(reinterpret_cast<float4*>(b_buf) + tid + i * kThreads, src + tid + i * kThreads);
}

__global__ void __launch_bounds__(cqt_tc::kThreads, 2) cqt_tc_kernel(const CqtTcParams p) {
  using namespace cqt_tc;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* a_hi = reinterpret_cast<float*>(smem_raw);
  float* a_lo = a_hi + kAFloats;
  float* b_buf0 = a_lo + kAFloats;             // [hi 8 KB][lo 8 KB], double buffered
  float* b_buf1 = b_buf0 + 2 * kBFloats;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_buf1 + 2 * kBFloats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) umma::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) umma::mbar_init(mbar, 1);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma::instr_desc_tf32(kM, kN);
  const uint32_t a_hi_addr = umma::smem_u32(a_hi), a_lo_addr = umma::smem_u32(a_lo);

  uint32_t phase = 0;
  const int tiles_per_clip = p.tiles_per_clip_oct * kOctaves;
  const int total = tiles_per_clip * p.batch;  // gridDim.x <= total

  // software pipeline: the global loads of item (tile, pass) + 1 are in flight while item (tile, pass) runs
  float4 v[kStage];
  StagePlan cur = plan_stage(p, blockIdx.x, 0, tid);
  prefetch_slice(cur, v);
  copy_b_async(p, 0, tid, b_buf0);
  int item = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    for (int pass = 0; pass < kPasses; ++pass, ++item) {
      float* b_cur = (item & 1) ? b_buf1 : b_buf0;
      float* b_nxt = (item & 1) ? b_buf0 : b_buf1;
      const int oct = cur.oct, m = cur.m;
      store_slice(cur, v, a_hi, a_lo);        // waits for this item's loads
      umma::cp_async_wait_all();              // this item's B images have landed
      umma::fence_proxy_async_smem();
      umma::fence_before_thread_sync();
      __syncthreads();

      if (warp == 0) {
        umma::fence_after_thread_sync();
        if (umma::elect_one_sync()) {
          const uint32_t b_hi_addr = umma::smem_u32(b_cur), b_lo_addr = umma::smem_u32(b_cur + kBFloats);
          const uint32_t lbo = m == 1 ? 16u : (uint32_t)kRT * 16u;
          const uint64_t da_hi0 = umma::smem_desc(a_hi_addr, lbo, 128);
          const uint64_t da_lo0 = umma::smem_desc(a_lo_addr, lbo, 128);
          const uint64_t db_hi0 = umma::smem_desc(b_hi_addr, kN * 16, 128);
          const uint64_t db_lo0 = umma::smem_desc(b_lo_addr, kN * 16, 128);
#pragma unroll
          for (int ks = 0; ks < kKStepsPerPass; ++ks) {
            const int c = 2 * ks;  // first window chunk of the K-step inside this slice
            int a_units;           // start-address offset in 16-byte units
            if (m >= kPassChunks) {
              a_units = c * kRT;
            } else if (m == 1) {
              a_units = c;
            } else {
              const int d = c >> (6 - oct), e = c & (m - 1);  // m = 64 >> oct is a power of two
              a_units = d + e * kRT;
            }
            const uint64_t a_off = (uint64_t)a_units, b_off = (uint64_t)(ks * 2 * kN);
            const int g = pass * kKStepsPerPass + ks;  // K-step 0..31 of the tile
            umma::mma_tf32(tmem_base + 32u * (g & (kMainAcc - 1)), da_hi0 + a_off, db_hi0 + b_off, idesc,
                           g >= kMainAcc ? 1u : 0u);
            umma::mma_tf32(tmem_base + 32u * kMainAcc, da_lo0 + a_off, db_hi0 + b_off, idesc, g > 0 ? 1u : 0u);
            umma::mma_tf32(tmem_base + 32u * kMainAcc, da_hi0 + a_off, db_lo0 + b_off, idesc, 1u);
          }
          umma::commit(mbar);
        }
      }
      __syncwarp();

      // prefetch the next work item while the tensor core runs
      {
        int ntile = tile, npass = pass + 1;
        if (npass == kPasses) npass = 0, ntile = tile + gridDim.x;
        if (ntile < total) {
          cur = plan_stage(p, ntile, npass, tid);
          prefetch_slice(cur, v);
          copy_b_async(p, npass, tid, b_nxt);
        }
      }

      // the MMAs read this item's smem: wait before the next store_slice overwrites it / before the epilogue
      umma::mbar_wait(mbar, phase);
      phase ^= 1;
      umma::fence_after_thread_sync();
    }

    // ---- epilogue: warps 0-3 write the real plane, warps 4-7 the imaginary plane; thread owns frame
    //      t = t0 + 32 (warp & 3) + lane of this tile
    const int b = tile / tiles_per_clip;
    const int rem = tile - b * tiles_per_clip;
    const int oct = rem / p.tiles_per_clip_oct;
    const int t0 = (rem - oct * p.tiles_per_clip_oct) * kM;
    const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
    const int frames_b = num_frames(len0);
    const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
    const int part = warp >> 2;  // 0: real (columns 0..11), 1: imaginary (columns 12..23)
    float acc[16];
    {
      const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 12u * part;
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
      umma::fence_before_thread_sync();
    }
    const int t = t0 + (warp & 3) * 32 + lane;
    if (t < p.slots) {
      const RowDest d = row_dest(p.out, b, t, frames_b, sections_b);
      const float2* st = nullptr;
      const int col0 = kFCqt - kBinsPerOctave * (oct + 1);
      if (p.out.stats)
        st = p.out.stats + (long long)b * p.out.stats_clip_stride + p.out.stats_off + part * p.out.f_stats + col0;
      const long long plane_off = part ? d.plane : 0;
#pragma unroll
      for (int j = 0; j < kBinsPerOctave; ++j) {
        float val = acc[j] * __ldg(p.scale + oct * kBinsPerOctave + j);
        if (st) {
          const float2 ms = __ldg(st + j);
          val = (val - ms.x) * ms.y;
        }
        if (d.n > 0) d.row[0][plane_off + col0 + j] = d.live[0] ? val : 0.f;
        if (d.n > 1) d.row[1][plane_off + col0 + j] = d.live[1] ? val : 0.f;
      }
    }
    // every thread's TMEM reads are complete (wait::ld) before the next tile's first MMA can be issued:
    // that MMA is behind the next __syncthreads() of the pass loop, and the fence above orders the reads
  }
  umma::cp_async_wait_all();
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, kTmemCols);
}

// host: B images.  kmat[n][24] (double) -> [pass][hi/lo][ks][c][j][kk], n = 64 pass + 8 ks + 4 c + kk
void host_cqt_tc_images(const double* kmat_256x24, float* images) {
  using namespace cqt_tc;
  for (int pass = 0; pass < kPasses; ++pass)
    for (int ks = 0; ks < kKStepsPerPass; ++ks)
      for (int c = 0; c < 2; ++c)
        for (int j = 0; j < kN; ++j)
          for (int kk = 0; kk < 4; ++kk) {
            const int n = 64 * pass + 8 * ks + 4 * c + kk;
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
            const size_t idx = (((size_t)ks * 2 + c) * kN + j) * 4 + kk;
            images[(size_t)pass * 2 * kBFloats + idx] = hi;
            images[(size_t)pass * 2 * kBFloats + kBFloats + idx] = lo;
          }
}

int cqt_tc_image_floats() { return cqt_tc::kPasses * 2 * cqt_tc::kBFloats; }

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
  const long long cap = 2LL * plan->sm_count;
  if (ctas > cap) ctas = cap;
  ProfileSpan span("cqt_tc_kernel", st);
  cqt_tc_kernel<<<(unsigned)ctas, cqt_tc::kThreads, cqt_tc::kSmem, st>>>(p);
  AST_LAUNCH_CHECK("cqt_tc_kernel");
  return AST_OK;
}

}  // namespace ast
