// decimate_tc.cu - K2 on the tensor cores: one 2:1 stage of librosa.cqt's decimation cascade
// (audio.resample(y, orig_sr=2, target_sr=1, res_type="soxr_hq", scale=True), reached through
// utilityFunctions.py:52) as a block-Toeplitz GEMM at the full tcgen05 rate.
//
//   y[j] = sum_{k=0}^{384} g[k] x[2 j + k - 192]      (zero extension, len_out = ceil(len_in / 2))
//
// Measured on B200 (scratch/mma_bench2.cu): an M128 K8 kind::tf32 MMA with both operands in shared memory costs
// max(~40, N / 2) cycles, so only N = 256 runs the tensor pipe at its peak.  The signal is therefore cut into
// NON-overlapping rows of 128 samples (row R = x[128 R - 192 ...]) and the 511-sample window of the 64 outputs
// y[64 R ... 64 R + 63] into D = 4 row blocks:
//
//   Y[R][u] = sum_{d < 4} Z[R + d][64 (3 - d) + u],      Z[rho][64 (3 - d) + u] = sum_{i < 128} X[rho][i] g[128 d + i - 2 u]
//
// Z = X * G is ONE GEMM with N = 256: A = X (128 rows x 128 samples, plain rows, no im2col, no shifted
// descriptors) and, because the blocks are laid out in reverse order along N, B is a single Toeplitz strip
// T[jj][kk] = g[kk - 2 jj]: K-step gs (8 samples) reads strip rows (60 - 4 gs) ... + 255 - just a start-address
// shift in the no-swizzle K-major layout (rows 16 B apart).  The row shift R + d is applied in the epilogue with
// warp shuffles; each warp owns 32 consecutive rows and emits the 29 complete ones, so the four row groups of a
// tile overlap by three rows (116 output rows = 7424 outputs per tile).
//
// FP32 accuracy: x = hi + lo (hi = what the tensor core keeps of an FP32 operand: truncation to TF32, measured;
// so the RAW samples are the hi operand, no masking needed) and
// g = hi + lo (host, from the double taps).  The tensor core accumulates with round-toward-zero (about -1.4e-8
// relative per accumulating MMA at full magnitude), so per tile the 32 small cross-term MMAs (lo*hi, hi*lo) are
// issued FIRST - their truncation error scales with the still tiny accumulator - and the 16 hi*hi MMAs last.
//
// All six stages of the cascade run in ONE persistent launch: tiles are numbered stage by stage (stage, clip, tile in
// clip) and dealt round-robin to the CTAs; a tile of stage s + 1 waits for the (at most three) stage-s tiles that
// produce its input through per-tile completion counters in the workspace (release: __threadfence + atomicAdd by
// each of the four epilogue warps; acquire: ld.acquire.gpu polling by one lane per producer warp, bounded).  Every
// dependency points to an earlier tile number and each CTA walks its tiles in order, so the earliest unfinished
// tile can always run; stage boundaries cost nothing (the tail of stage s overlaps the head of stage s + 1) and five
// launches disappear.  Inputs produced by this launch are read with ld.global.cg (L2), never through L1 / the
// non-coherent path.
//
// One persistent CTA per SM, 14 warps:
//   warps 0-7   producers, two groups of four warps: group g stages slices 2 g, 2 g + 1 of each tile (raw -> A_hi, two
//               whole tiles resident; x - trunc(x) -> A_lo, 2-stage ring of 32-sample slices), then loads the same
//               slices of the NEXT tile into registers (16 x LDG.128 per thread) - fence.proxy.async waits for a
//               thread's outstanding loads, so each group fences only when its prefetch has had ~2 us to land.  (cp.async was tried first: LDGSTS with a scattered destination costs
//               one shared-memory wavefront per thread, 8x the LDG + conflict-free STS.128 path - ncu, profiles/.)
//   warps 8-11  epilogue : TMEM -> registers, shifted sum by shuffles, stores (two accumulators, ping-pong)
//   warp  12    MMA issue
//   warp  13    publisher: completion counters of finished tiles
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <cmath>

#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace ast {

namespace dtc {
constexpr int kM = 128;                          // staged rows per tile (TMEM lanes)
constexpr int kGroupValid = 29;                  // complete output rows per 32-row group (rows R, R+1, R+2, R+3 needed)
constexpr int kGroups = 4;
constexpr int kRowsOut = kGroups * kGroupValid;  // 116
constexpr int kP = 64;                           // outputs per row
constexpr int kRS = 2 * kP;                      // 128 input samples per row
constexpr int kD = 4;                            // row blocks per window: 4 * 128 = 512 >= 2 * 63 + 385
constexpr int kN = kD * kP;                      // 256
constexpr int kSlices = 4;                       // 32-sample K slices per tile
constexpr int kSliceChunks = 8;                  // 16-byte chunks per row per slice
constexpr int kKStepsPerSlice = 4;
constexpr int kRT = 129;                         // rows per chunk column (odd: conflict-free 16-byte stores)
constexpr int kSliceFloats = kSliceChunks * kRT * 4;   // 4128 floats = 16 512 B
constexpr int kTileFloats = kSlices * kSliceFloats;    // A_hi of one tile
constexpr int kStripRows = 320;                  // 316 used: jj = -252 ... 63
constexpr int kStripRow0 = 60;                   // strip row of (n = 0, gs = 0): jj = -192
constexpr int kStripFloats = 2 * kStripRows * 4; // [chunk 2][row 320][4]
constexpr int kProducers = 256;
constexpr int kEpilogueWarp0 = 8;
constexpr int kMmaWarp = 12;
constexpr int kPublishWarp = 13;                 // sets the tiles' completion counters, off the epilogue's critical path
constexpr int kThreads = 14 * 32;
constexpr int kTmemCols = 512;                   // two 256-column accumulators
constexpr int kProducerGroup = 128;               // producer threads per slice (two groups, two slices each)
constexpr int kChunksPerThread = kM * kSliceChunks / kProducerGroup;  // 8
constexpr int kEpiStride = kP + 4;               // floats per staged output row (conflict-free 16-byte accesses)
constexpr int kEpiFloats = 4 * 32 * kEpiStride;  // one [32 rows][68] transpose buffer per epilogue warp
constexpr size_t kSmem = sizeof(float) * (2 * kTileFloats + 2 * kSliceFloats + 2 * kStripFloats + kEpiFloats) + 256;
}  // namespace dtc

constexpr int kDecStages = kOctaves - 1;  // 6

#ifdef AST_TRACE
#ifndef AST_TRACE_CTA
#define AST_TRACE_CTA 0   // 147: a CTA that runs a clip's chain stages
#endif
// diagnostic build only (scratch/trace_dec.py): clock64 stamps of one CTA's pipeline roles, per tile
__device__ long long g_dec_trace[4][64][8];
#define DTC_STAMP(role, idx, k) \
  do { if (blockIdx.x == (AST_TRACE_CTA) && (threadIdx.x & 31) == 0 && (idx) < 64) g_dec_trace[role][idx][k] = clock64(); } while (0)
#else
#define DTC_STAMP(role, idx, k) do {} while (0)
#endif

AST_TIMELINE_DEFINE(dec)

struct DecimateTcParams {
  const float* wave;       // octave 0: clip b at wave + b * wave_stride
  long long wave_stride;
  float* ws;               // octaves 1..6: clip b, octave i at ws + b * ws_clip_stride + oct_off[i]
  long long ws_clip_stride;
  long long oct_off[kOctaves];
  const int32_t* lengths;
  long long max_samples;
  int batch;
  int tiles_per_clip[kDecStages];
  int tile_prefix[kDecStages + 1];   // first tile number of each stage; [6] = total
  int tail_stage;                    // first stage whose tiles are dealt to the last tail_ctas CTAs only (6: none)
  int tail_ctas;
  bool vec_ok0;                      // 16-byte loads legal on the octave-0 rows
  int* flags;                        // [stage][clip][tile of stage 0's count]: epilogue warps that finished the tile (4 = done)
  const float* strip_hi;   // [2][320][4] smem image of the Toeplitz strip (TF32-exact values)
  const float* strip_lo;
  const uint4* strip_h_hi; // FP16-split kernel: [2][320][8 halves] image of the Toeplitz strip x 2^15, and its residual
  const uint4* strip_h_lo;
  int debug;               // diagnostic bit mask (AST_DEC_DEBUG): 1 no epilogue shuffles / stores, 2 no producer loads, 4 no MMAs
};

struct DtcTile {
  const float* x;
  float* y;
  int len_in, len_out;
  int row0;        // first row of the tile (rows advance by 116 per tile)
  int stage, clip, k;
  bool live;       // false: the tile lies past the clip's end (ragged batch)
  bool interior;   // every staged sample lies inside [0, len_in) and 16-byte loads are legal
  bool vec_ok;
  int debug;
};

__device__ __forceinline__ DtcTile dtc_decode(const DecimateTcParams& p, int tile) {
  using namespace dtc;
  DtcTile t;
  int s = 0;
#pragma unroll
  for (int i = 1; i < kDecStages; ++i) s += tile >= p.tile_prefix[i] ? 1 : 0;
  const int r = tile - p.tile_prefix[s];
  const int b = r / p.tiles_per_clip[s];
  t.stage = s;
  t.clip = b;
  t.k = r - b * p.tiles_per_clip[s];
  t.row0 = t.k * kRowsOut;
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  t.len_in = (int)((len0 + (1LL << s) - 1) >> s);
  t.len_out = (t.len_in + 1) >> 1;
  t.live = t.row0 * kP < t.len_out;
  t.x = s == 0 ? p.wave + (long long)b * p.wave_stride : p.ws + (long long)b * p.ws_clip_stride + p.oct_off[s];
  t.y = p.ws + (long long)b * p.ws_clip_stride + p.oct_off[s + 1];
  t.vec_ok = s > 0 || p.vec_ok0;
  t.debug = p.debug;
  const int first = kRS * t.row0 - kDecHalf;
  const int last = kRS * (t.row0 + kGroupValid * (kGroups - 1) + 31) - kDecHalf + kRS;  // one past the last staged sample
  t.interior = first >= 0 && last <= t.len_in && t.vec_ok;
  return t;
}

// Dealing of the tiles to the CTAs.  Tiles below tile_prefix[tail_stage] ("early": the stages with many tiles per clip)
// go round-robin to every CTA.  The stages with at most two tiles per clip form a dependent chain per clip that no
// amount of parallelism shortens; spread over the whole grid they would keep (nearly) every SM occupied by a mostly
// idle CTA until the chain's end.  They are dealt to the LAST tail_ctas CTAs only (those have one early tile fewer),
// CTA j taking the clips j, j + tail_ctas, ...: stage by stage, clip by clip, so its list is still ascending in tile
// number (the progress argument in the header) and a clip's chain stays on one CTA.  The other CTAs exit early and their
// SMs go to the dependent launches (the CQT projection's octave-0 tiles, then the STFT).
__device__ __forceinline__ int dtc_tail_first(const DecimateTcParams& p, int total) {
  const int j = (int)blockIdx.x - ((int)gridDim.x - p.tail_ctas);
  if (p.tail_stage >= kDecStages || j < 0 || j >= p.batch) return total;
  return p.tile_prefix[p.tail_stage] + j * p.tiles_per_clip[p.tail_stage];
}
__device__ __forceinline__ int dtc_first(const DecimateTcParams& p, int total) {
  const int early = p.tile_prefix[p.tail_stage];
  return (int)blockIdx.x < early ? (int)blockIdx.x : dtc_tail_first(p, total);
}
__device__ __forceinline__ int dtc_step(const DecimateTcParams& p, int tile, int total) {   // total: end of the list
  const int early = p.tile_prefix[p.tail_stage];
  if (tile < early) {
    tile += gridDim.x;
    return tile < early ? tile : dtc_tail_first(p, total);
  }
  int s = p.tail_stage;
  for (int i = p.tail_stage + 1; i < kDecStages; ++i) s += tile >= p.tile_prefix[i] ? 1 : 0;
  const int r = tile - p.tile_prefix[s];
  int b = r / p.tiles_per_clip[s];
  if (r - b * p.tiles_per_clip[s] + 1 < p.tiles_per_clip[s]) return tile + 1;
  b += p.tail_ctas;
  if (b < p.batch) return p.tile_prefix[s] + b * p.tiles_per_clip[s];
  if (++s == kDecStages) return total;
  return p.tile_prefix[s] + ((int)blockIdx.x - ((int)gridDim.x - p.tail_ctas)) * p.tiles_per_clip[s];
}
__device__ __forceinline__ int dtc_next_live(const DecimateTcParams& p, int tile, int total) {
  for (tile = dtc_step(p, tile, total); tile < total; tile = dtc_step(p, tile, total))
    if (dtc_decode(p, tile).live) return tile;
  return -1;
}

__device__ __forceinline__ int* dtc_flag(const DecimateTcParams& p, int stage, int clip, int k) {
  return p.flags + ((long long)stage * p.batch + clip) * p.tiles_per_clip[0] + k;
}

// The stage s - 1 tiles whose outputs tile t (stage s >= 1) stages: output indices [lo, hi) -> tiles lo / 7424 ...
__device__ __forceinline__ void dtc_dep_range(const DtcTile& t, int& k_lo, int& k_hi) {
  using namespace dtc;
  int lo = kRS * t.row0 - kDecHalf, hi = kRS * (t.row0 + kGroupValid * (kGroups - 1) + 32) - kDecHalf;
  if (lo < 0) lo = 0;
  if (hi > t.len_in) hi = t.len_in;
  k_lo = lo / (kRowsOut * kP);
  k_hi = hi > lo ? (hi - 1) / (kRowsOut * kP) : k_lo - 1;
}

// relaxed poll (an acquire load would invalidate L1 on every poll: CCTL.IVALL); the acquire fence follows success
__device__ __forceinline__ int ld_relaxed_gpu(const int* ptr) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// true when every producer tile of t has been completed by its four epilogue warps (one lane polls for the warp)
__device__ __forceinline__ bool dtc_deps_ready(const DecimateTcParams& p, const DtcTile& t, int lane, unsigned& stages_complete) {
  if (t.stage == 0 || ((stages_complete >> (t.stage - 1)) & 1)) return true;
  {  // the whole previous stage finished? then nothing of this stage needs to poll (or fence) again
    int full = 0;
    if (lane == 0) {
      const int* stage_done = p.flags + (long long)kDecStages * p.batch * p.tiles_per_clip[0];
      full = ld_relaxed_gpu(stage_done + t.stage - 1) >= p.tiles_per_clip[t.stage - 1] * p.batch ? 1 : 0;
      if (full) __threadfence();
    }
    if (__shfl_sync(0xffffffffu, full, 0)) {
      stages_complete |= 1u << (t.stage - 1);
      return true;
    }
  }
  int k_lo, k_hi;
  dtc_dep_range(t, k_lo, k_hi);
  int ok = 1;
  if (lane == 0) {
    for (int k = k_lo; k <= k_hi; ++k) ok &= ld_relaxed_gpu(dtc_flag(p, t.stage - 1, t.clip, k)) >= 4 ? 1 : 0;
    if (ok) __threadfence();  // acquire: the producers' stores (released by their fence + atomicAdd) are visible from here on
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok != 0;
}

// Bounded wait: a broken dependency chain traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void dtc_deps_wait(const DecimateTcParams& p, const DtcTile& t, int lane, unsigned& stages_complete) {
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !dtc_deps_ready(p, t, lane, stages_complete); ++spin) {
    __nanosleep(200);
    if ((spin & 1023) == 1023) {
      const unsigned long long now = umma::global_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > umma::kPollTimeoutNs) __trap();
    }
  }
}

__device__ __forceinline__ float4 ld_cg_f4(const float* ptr) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cg_f(const float* ptr) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(ptr) : "memory");
  return v;
}
// x[s .. s + 3] with zeros outside [0, len), through L2 (the data may have been written by this launch)
static __device__ __noinline__ float4 dtc_load4_partial(const float* x, int s, int len) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s >= 0 && s < len) v.x = ld_cg_f(x + s);
  if (s + 1 >= 0 && s + 1 < len) v.y = ld_cg_f(x + s + 1);
  if (s + 2 >= 0 && s + 2 < len) v.z = ld_cg_f(x + s + 2);
  if (s + 3 >= 0 && s + 3 < len) v.w = ld_cg_f(x + s + 3);
  return v;
}
__device__ __forceinline__ float4 dtc_load4(const float* x, int s, int len, bool vec_ok) {
  if (s >= 0 && s + 3 < len && vec_ok) return ld_cg_f4(x + s);
  if (s + 3 < 0 || s >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
  return dtc_load4_partial(x, s, len);
}

// One 32-sample slice of a tile for a producer thread.  The producers work as two groups of 128 threads: group g
// stages slices 2 g and 2 g + 1 of every tile, so a thread owns chunk c = tg & 7 of the eight rows
// rho = (tg >> 3) + 16 i, i = 0..7 (tg = thread within the group), i.e. row group rho >> 5, lane row rho & 31,
// read at sample 128 (row0 + 29 (rho >> 5) + (rho & 31)) - 192 + 32 s + 4 c.  J = 0 / 1: the group's first / second slice.
template <int J>
__device__ __forceinline__ void dtc_load_slice(const DtcTile& t, int tg, int s, float4 (&v)[16]) {
  using namespace dtc;
  const int a0 = kRS * (t.row0 + (tg >> 3)) - kDecHalf + 32 * s + 4 * (tg & 7);
  if (t.debug & 2) {
#pragma unroll
    for (int i = 0; i < kChunksPerThread; ++i) v[kChunksPerThread * J + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
#pragma unroll
  for (int i = 0; i < kChunksPerThread; ++i) {
    const int a = a0 + kRS * (kGroupValid * (i >> 1) + 16 * (i & 1));
    v[kChunksPerThread * J + i] = t.interior ? ld_cg_f4(t.x + a) : dtc_load4(t.x, a, t.len_in, t.vec_ok);
  }
}

__global__ void __launch_bounds__(dtc::kThreads, 1) decimate2_tc_kernel(const DecimateTcParams p) {
  using namespace dtc;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AST_TIMELINE_STAMP(dec, blockIdx.x, 0);
  float* a_hi = reinterpret_cast<float*>(smem_raw);      // [2 tiles][4 slices][8 chunk columns][129 rows][4]
  float* a_lo = a_hi + 2 * kTileFloats;                  // [2 stages][8][129][4]
  float* t_hi = a_lo + 2 * kSliceFloats;
  float* t_lo = t_hi + kStripFloats;
  float* epi_buf = t_lo + kStripFloats;                  // [4 warps][32 rows][68] epilogue transpose
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kEpiFloats);
  // One barrier pair PER SLICE of a tile (not per ring stage): each is completed once per tile and waited once per
  // tile by one party, so the two producer groups never skip a phase of a barrier (parity waits alias otherwise).
  uint64_t* l_full = bars;          // [4] producers -> MMA: slice staged (raw + lo written); 4 warp arrivals (one group)
  uint64_t* l_empty = bars + 4;     // [4] MMA -> producers: the cross-term MMAs of the slice are done (lo stage free)
  uint64_t* h_empty = bars + 8;     // [2] MMA -> producers: all MMAs of the tile in this A_hi buffer are done
  uint64_t* acc_full = bars + 10;   // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 12;  // [2] epilogue -> MMA; 4 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  // epilogue -> publisher: epilogue warps that have stored their rows, summed over all tiles (a monotonic counter
  // cannot alias the way a lapped mbarrier parity would)
  unsigned int* stored_count = reinterpret_cast<unsigned int*>(bars + 20);   // [4]: one per epilogue warp
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kStripFloats / 4; i += kThreads) {
    reinterpret_cast<float4*>(t_hi)[i] = __ldg(reinterpret_cast<const float4*>(p.strip_hi) + i);
    reinterpret_cast<float4*>(t_lo)[i] = __ldg(reinterpret_cast<const float4*>(p.strip_lo) + i);
  }
  if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      umma::mbar_init(l_full + i, kProducerGroup / 32);
      umma::mbar_init(l_empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(h_empty + i, 1);
      umma::mbar_init(acc_full + i, 1);
      umma::mbar_init(acc_empty + i, 4);
    }
    stored_count[0] = stored_count[1] = stored_count[2] = stored_count[3] = 0;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  // The prologue above overlaps the previous kernel of the stream; everything below needs its writes (the zeroed
  // completion counters).  Dependents (the CQT projection, then the STFT) are released only AFTER the wait, so that
  // they too start behind the feature call's prologue kernel (statistics table) without waiting for anything themselves.
  pdl_wait();
  pdl_launch_dependents();
  const int total = p.tile_prefix[kDecStages];
  int tile = dtc_first(p, total);
  if (tile < total && !dtc_decode(p, tile).live) tile = dtc_next_live(p, tile, total);

  if (warp < kProducers / 32) {
    // ================================================================= producers
    // fence.proxy.async waits for every outstanding load of its thread, so a thread must not fence while the next
    // tile's loads are in flight: group g stages slices 2 g, 2 g + 1 of tile n, then loads the same two slices of
    // tile n + 1 and has the other group's two slices plus the hi*hi MMAs (~2 us) before its next fence.
    const int grp = warp >> 2, tg = tid & (kProducerGroup - 1);
    float4 v[16];  // v[8 j + i]: chunk i of the group's j-th slice
    unsigned stages_complete = 0;  // bit s: stage s has been seen complete
    const int slot0 = (tg & 7) * kRT + (tg >> 3);
    DtcTile cur;
    if (tile >= 0 && tile < total) {
      cur = dtc_decode(p, tile);
      dtc_deps_wait(p, cur, lane, stages_complete);
      dtc_load_slice<0>(cur, tg, 2 * grp, v);
      dtc_load_slice<1>(cur, tg, 2 * grp + 1, v);
    }
    for (int n = 0; tile >= 0 && tile < total; ++n) {
      const int next = dtc_next_live(p, tile, total);
      if ((warp & 3) == 0) DTC_STAMP(grp, n, 0);
      // this A_hi buffer is free once the MMAs of tile n - 2 have completed
      umma::mbar_wait(h_empty + (n & 1), ((n >> 1) & 1) ^ 1);
      if ((warp & 3) == 0) DTC_STAMP(grp, n, 1);
      float4* hi_tile = reinterpret_cast<float4*>(a_hi + (n & 1) * kTileFloats);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int sl = 2 * grp + j;                 // slice of the tile
        const int st = sl & 1;                      // lo ring stage
        // the cross MMAs that read this lo stage two slices ago are done
        if (sl >= 2) umma::mbar_wait(l_empty + sl - 2, n & 1);
        else if (n > 0) umma::mbar_wait(l_empty + sl + 2, (n - 1) & 1);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 2 + 2 * j);
        float4* hi = hi_tile + sl * (kSliceFloats / 4);
        float4* lo = reinterpret_cast<float4*>(a_lo + st * kSliceFloats);
#pragma unroll
        for (int i = 0; i < kChunksPerThread; ++i) {
          float4 h, l;
          umma::split_tf32(v[kChunksPerThread * j + i], h, l);
          hi[slot0 + 16 * i] = v[kChunksPerThread * j + i];   // raw: the tensor core truncates to TF32 itself
          lo[slot0 + 16 * i] = l;
        }
        umma::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(l_full + sl);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 3 + 2 * j);
      }
      if (next >= 0) {
        cur = dtc_decode(p, next);
        dtc_deps_wait(p, cur, lane, stages_complete);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 6);   // stage s + 1 tiles: the stage-s tiles that produce these samples have finished
        dtc_load_slice<0>(cur, tg, 2 * grp, v);
        dtc_load_slice<1>(cur, tg, 2 * grp + 1, v);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 7);
      }
      tile = next;
    }
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issue
    const uint32_t idesc = umma::instr_desc_tf32(kM, kN);
    const uint64_t db_hi0 = umma::smem_desc(umma::smem_u32(t_hi), kStripRows * 16, 128) + (uint64_t)kStripRow0;
    const uint64_t db_lo0 = umma::smem_desc(umma::smem_u32(t_lo), kStripRows * 16, 128) + (uint64_t)kStripRow0;
    for (int n = 0; tile >= 0 && tile < total; ++n, tile = dtc_next_live(p, tile, total)) {
      const int q = n & 1;
      const uint32_t acc = tmem_base + (uint32_t)(q * kN);
      const uint32_t hi_addr = umma::smem_u32(a_hi + q * kTileFloats);
      DTC_STAMP(2, n, 0);
      umma::mbar_wait(acc_empty + q, ((n >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
      umma::fence_after_thread_sync();
      DTC_STAMP(2, n, 1);
      // cross terms first (see the accuracy note in the header)
      for (int s = 0; s < kSlices; ++s) {
        const int st = s & 1;
        umma::mbar_wait(l_full + s, n & 1);
        umma::fence_after_thread_sync();
        DTC_STAMP(2, n, 2 + s);
        if (umma::elect_one_sync()) {
          const uint64_t da_hi = umma::smem_desc(hi_addr + (uint32_t)(s * kSliceFloats * 4), kRT * 16, 128);
          const uint64_t da_lo = umma::smem_desc(umma::smem_u32(a_lo + st * kSliceFloats), kRT * 16, 128);
#pragma unroll
          for (int k = 0; k < kKStepsPerSlice; ++k) {
            if (p.debug & 4) break;
            const uint64_t a_off = (uint64_t)(2 * k * kRT);              // start-address field is in 16-byte units
            const uint64_t b_off = (uint64_t)(4 * (kKStepsPerSlice * s + k));  // strip row 60 - 4 gs
            umma::mma_tf32(acc, da_lo + a_off, db_hi0 - b_off, idesc, (s | k) ? 1u : 0u);
            umma::mma_tf32(acc, da_hi + a_off, db_lo0 - b_off, idesc, 1u);
          }
          umma::commit(l_empty + s);
        }
        __syncwarp();
      }
      if (umma::elect_one_sync()) {
#pragma unroll
        for (int gs = 0; gs < kSlices * kKStepsPerSlice; ++gs) {
          if (p.debug & 4) break;
          const int s = gs >> 2, k = gs & 3;
          const uint64_t da_hi = umma::smem_desc(hi_addr + (uint32_t)(s * kSliceFloats * 4), kRT * 16, 128);
          umma::mma_tf32(acc, da_hi + (uint64_t)(2 * k * kRT), db_hi0 - (uint64_t)(4 * gs), idesc, 1u);
        }
        umma::commit(h_empty + q);    // this A_hi buffer may be refilled
        umma::commit(acc_full + q);   // ... and the accumulator is complete
      }
      __syncwarp();
      DTC_STAMP(2, n, 6);
    }
  } else if (warp < kPublishWarp) {
    // ================================================================= epilogue (warps 8-11)
    const int quad = warp - kEpilogueWarp0;  // == warp % 4: the TMEM lane quadrant this warp may read
    float* stg = epi_buf + quad * 32 * kEpiStride;
    for (int n = 0; tile >= 0 && tile < total; ++n, tile = dtc_next_live(p, tile, total)) {
      const int q = n & 1;
      const DtcTile t = dtc_decode(p, tile);
      if (quad == 0) DTC_STAMP(3, n, 0);
      umma::mbar_wait(acc_full + q, (n >> 1) & 1);
      umma::fence_after_thread_sync();
      if (quad == 0) DTC_STAMP(3, n, 1);
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(q * kN);
      // lane l of row group `quad` owns output row R = row0 + 29 quad + l (complete for l < 29)
#pragma unroll
      for (int cq = 0; cq < kP / 16; ++cq) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        umma::tmem_ld_32x16_nowait(lane_base + 3 * kP + 16 * cq, r0);   // d = 0: this row
        umma::tmem_ld_32x16_nowait(lane_base + 2 * kP + 16 * cq, r1);   // d = 1: wanted by the row above (lane - 1)
        umma::tmem_ld_32x16_nowait(lane_base + 1 * kP + 16 * cq, r2);
        umma::tmem_ld_32x16_nowait(lane_base + 16 * cq, r3);
        umma::tmem_wait_ld();                                            // one TMEM round trip for the four loads
        float v0[16], v1[16], v2[16], v3[16];
#pragma unroll
        for (int c = 0; c < 16; ++c)
          v0[c] = __uint_as_float(r0[c]), v1[c] = __uint_as_float(r1[c]), v2[c] = __uint_as_float(r2[c]), v3[c] = __uint_as_float(r3[c]);
        if (cq == kP / 16 - 1) {  // last TMEM read of the tile: hand the accumulator back
          umma::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(acc_empty + q);
          if (quad == 0) DTC_STAMP(3, n, 2);
        }
        if (p.debug & 1) continue;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float s1 = __shfl_down_sync(0xffffffffu, v1[c], 1);
          const float s2 = __shfl_down_sync(0xffffffffu, v2[c], 2);
          const float s3 = __shfl_down_sync(0xffffffffu, v3[c], 3);
          v0[c] = (v0[c] + s1) + (s2 + s3);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w)
          *reinterpret_cast<float4*>(stg + lane * kEpiStride + 16 * cq + 4 * w) =
              make_float4(v0[4 * w], v0[4 * w + 1], v0[4 * w + 2], v0[4 * w + 3]);
      }
      __syncwarp();
      // write-out along rows: 16 lanes cover one 256-byte row, so an instruction touches 4 lines instead of 29
      const int j_grp = (t.row0 + kGroupValid * quad) * kP;
      for (int idx = lane; idx < kGroupValid * (kP / 4) && !(p.debug & 1); idx += 32) {
        const int r = idx >> 4, c4 = idx & 15;
        const int j = j_grp + r * kP + 4 * c4;
        const float4 val = *reinterpret_cast<const float4*>(stg + r * kEpiStride + 4 * c4);
        if (j + 4 <= t.len_out) {
          *reinterpret_cast<float4*>(t.y + j) = val;
        } else {
          if (j < t.len_out) t.y[j] = val.x;
          if (j + 1 < t.len_out) t.y[j + 1] = val.y;
          if (j + 2 < t.len_out) t.y[j + 2] = val.z;
        }
      }
      // publish: this warp's rows of the tile are in global memory (stages 0..4 feed another stage of this launch)
      __syncwarp();  // also: the staging buffer is rewritten by the next tile
      if (quad == 0) DTC_STAMP(3, n, 3);
      if (lane == 0) {   // release (CTA scope): the publisher warp makes it visible GPU-wide
        __threadfence_block();
        atomicAdd(stored_count + quad, 1u);
      }
    }
  } else if (warp == kPublishWarp) {
    // ================================================================= publisher
    // Waiting for a tile's stores to be acknowledged (__threadfence) costs ~2000 cycles; done here it does not delay
    // the epilogue's next tile.  stores (epilogue warps) -> mbarrier arrive / wait (CTA scope) -> fence (GPU scope,
    // cumulative) -> counter: the consumers' acquire side is dtc_deps_ready.
    // Every tile of this CTA's list, live or not, also bumps its stage's finished-tile counter: a consumer that has
    // seen a stage complete (counter == tiles of the stage) stops polling per-tile counters for it.
    int* stage_done = p.flags + (long long)kDecStages * p.batch * p.tiles_per_clip[0];
    int n = 0;
    for (int g = dtc_first(p, total); g < total; g = dtc_step(p, g, total)) {
      const DtcTile t = dtc_decode(p, g);
      if (t.live) {   // every epilogue warp has released its rows of the CTA's n-th live tile (one counter per warp:
        ++n;          // a sum over the warps could be reached by three warps running a tile ahead of the fourth)
        unsigned long long t0 = 0;
        for (uint32_t spin = 0;; ++spin) {
          const bool ok = lane >= 4 || reinterpret_cast<volatile unsigned int*>(stored_count)[lane] >= (unsigned)n;
          if (__all_sync(0xffffffffu, ok)) break;
          __nanosleep(100);
          if ((spin & 1023) == 1023) {
            const unsigned long long now = umma::global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > umma::kPollTimeoutNs) __trap();
          }
        }
      }
      if (lane == 0) {
        if (t.live) {
          __threadfence();   // every stage: the CQT projection that follows polls the same counters
          atomicAdd(dtc_flag(p, t.stage, t.clip, t.k), 4);
        }
        atomicAdd(stage_done + t.stage, 1);
      }
      __syncwarp();
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  AST_TIMELINE_STAMP_IF(warp == kMmaWarp && lane == 0, dec, blockIdx.x, 1);
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_base, kTmemCols);
  AST_TIMELINE_STAMP_IF(warp == kMmaWarp && lane == 0, dec, blockIdx.x, 3);   // (after the TMEM release)
}


// =====================================================================================================================
// FP16-split variant (default).  The same GEMM, tiles, dependency protocol and epilogue; the operands are split into
// FP16 pairs instead of TF32 pairs: FP16 and TF32 carry the same 11 significand bits, so  x = hi + lo,  g = hi + lo  with
// three products (lo*hi, hi*lo, hi*hi) has the accuracy of the TF32 scheme, but a kind::f16 MMA contracts K = 16 per
// instruction at the cost of a K = 8 kind::tf32 one: 24 MMAs per tile instead of 48, and half the operand bytes in
// shared memory (138 KB per CTA instead of 220 KB).  FP16 has 5 exponent bits, so operands are scaled by powers of two
// (exact) into its range: the taps by 2^15 on the host, the samples PER TILE by 2^k with k chosen from the tile's own
// largest magnitude (producers: register max -> redux -> shared atomicMax -> one named barrier per tile) so that it lies
// in [2^8, 2^9); residuals that fall into FP16's subnormal range are then below 2^-33 of the tile's largest sample.  The
// epilogue multiplies the accumulator by 2^(-k-15).
// Co-residency with the STFT (diagnostic, -DAST_DEC_H_REGS=96 and AST_DEC_CARVEOUT=100): the register file is split over
// the four schedulers, 16 K registers each, and the fullest of them carries four of this kernel's 14 warps, so one 4-warp
// STFT CTA (128 registers) fits next to this kernel only at <= 96 registers per thread here (104 leaves 16 K registers free
// in total and still never co-schedules), and only if the whole unified array is configured as shared memory (the driver's
// choice for one 188 KB CTA is 196 KB).  Measured on B200 with the STFT launched right behind this kernel: 148 + 148 CTAs
// are resident from the start, each STFT CTA then takes 46 us instead of 30, this kernel's bulk phase 50 -> 62 us and its
// chain 85 -> 116 us (28 KB of L1, shared LSU / issue slots): feature step 0.2743 -> 0.2823 ms.  Kept off.
#ifdef AST_DEC_H_REGS
#define AST_DEC_H_BOUNDS __maxnreg__(AST_DEC_H_REGS)
#else
#define AST_DEC_H_BOUNDS __launch_bounds__(dtc::kThreads, 1)
#endif
namespace dth {
using dtc::kM; using dtc::kN; using dtc::kP; using dtc::kRS; using dtc::kGroupValid; using dtc::kGroups; using dtc::kRowsOut;
using dtc::kSlices; using dtc::kProducers; using dtc::kEpilogueWarp0; using dtc::kMmaWarp; using dtc::kPublishWarp;
using dtc::kThreads; using dtc::kTmemCols; using dtc::kProducerGroup; using dtc::kEpiStride; using dtc::kEpiFloats;
constexpr int kRT = 130;                       // 16-byte rows per chunk column (2 mod 8: the producers' 4 chunks x 8 rows stores are conflict-free)
constexpr int kSliceChunks = 4;                // 16-byte chunks (8 halves) per row per 32-sample slice
constexpr int kKStepsPerSlice = 2;             // K = 16 per MMA
constexpr int kSliceBytes = kSliceChunks * kRT * 16;   // 8 320
constexpr int kTileBytes = kSlices * kSliceBytes;      // 33 280: A_hi of one tile
constexpr int kStripRows = 320;                // 312 used: jj = -248 ... 63
constexpr int kStripRow0 = 56;                 // strip row of (n = 0, gs = 0): jj = -192
constexpr int kStripBytes = 2 * kStripRows * 16;
constexpr int kTapShift = 15;                  // taps are staged as g * 2^15 (largest tap ~ 0.65)
constexpr int kChunksPerThread = kM * kSliceChunks / kProducerGroup;   // 4 shared-memory chunks (8 samples each) per slice
constexpr size_t kSmem = 4 * kTileBytes + 2 * kStripBytes + sizeof(float) * kEpiFloats + 256 + 64 * 48;   // 191 744 B   // + the tile table
}  // namespace dth

__host__ __device__ constexpr uint32_t instr_desc_f16(int m, int n) {   // kind::f16, A = B = FP16, FP32 accumulate, K-major
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The CTA's tile list by position (the closed form of dtc_first / dtc_step): early tiles cta, cta + grid, ..., then - for
// the last tail_ctas CTAs - the chain stages, stage by stage, over the clips j, j + tail_ctas, ...
struct DthList {
  int n_early, n_total;   // positions [0, n_early) are early tiles, [n_early, n_total) chain tiles
  int j, per_cta;         // chain: this CTA's first clip and its number of clips
};
__device__ __forceinline__ DthList dth_list(const DecimateTcParams& p) {
  DthList l;
  const int early = p.tile_prefix[p.tail_stage], cta = (int)blockIdx.x, grid = (int)gridDim.x;
  l.n_early = early > cta ? (early - cta + grid - 1) / grid : 0;
  l.j = cta - (grid - p.tail_ctas);
  l.per_cta = 0;
  int chain = 0;
  if (p.tail_stage < kDecStages && l.j >= 0 && l.j < p.batch) {
    l.per_cta = (p.batch - l.j + p.tail_ctas - 1) / p.tail_ctas;
    for (int s = p.tail_stage; s < kDecStages; ++s) chain += l.per_cta * p.tiles_per_clip[s];
  }
  l.n_total = l.n_early + chain;
  return l;
}
__device__ __forceinline__ int dth_tile_at(const DecimateTcParams& p, const DthList& l, int i) {
  if (i < l.n_early) return (int)blockIdx.x + i * (int)gridDim.x;
  int ii = i - l.n_early, s = p.tail_stage;
  while (s < kDecStages - 1 && ii >= l.per_cta * p.tiles_per_clip[s]) ii -= l.per_cta * p.tiles_per_clip[s], ++s;
  const int ord = ii / p.tiles_per_clip[s], k = ii - ord * p.tiles_per_clip[s];
  return p.tile_prefix[s] + (l.j + ord * p.tail_ctas) * p.tiles_per_clip[s] + k;
}
constexpr int kDthTableCap = 64;    // decoded tiles kept in shared memory (3 KB); longer lists decode the rest on the fly
struct DthWalk {
  const DecimateTcParams& p;
  DthList l;
  const DtcTile* table;
  __device__ __forceinline__ DtcTile get(int i) const { return i < kDthTableCap ? table[i] : dtc_decode(p, dth_tile_at(p, l, i)); }
  __device__ __forceinline__ int next_live(int i) const {   // first live position >= i, or n_total
    while (i < l.n_total && !get(i).live) ++i;
    return i;
  }
};

// One 32-sample slice of a tile for a producer thread: 16 bytes (four samples, f = tg & 7) of the eight rows
// rho = (tg >> 3) + 16 i - eight lanes cover a row's 128 contiguous bytes, whole sectors (a thread that loads the two halves
// of a shared-memory chunk in two instructions requests every sector twice: .cg loads are not kept in the L1).
// v[8 J + i]: row i of the group's J-th slice.
template <int J>
__device__ __forceinline__ void dth_load_slice(const DtcTile& t, int tg, int s, float4 (&v)[16]) {
  using namespace dth;
  const int a0 = kRS * (t.row0 + (tg >> 3)) - kDecHalf + 32 * s + 4 * (tg & 7);
  if (t.debug & 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 * J + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = a0 + kRS * (kGroupValid * (i >> 1) + 16 * (i & 1));
    v[8 * J + i] = t.interior ? ld_cg_f4(t.x + a) : dtc_load4(t.x, a, t.len_in, t.vec_ok);
  }
}

// 4 samples x scale -> FP16 hi and FP16 residual images (half a 16-byte chunk each)
__device__ __forceinline__ void dth_split4(const float4& a, float scale, uint2& hi, uint2& lo) {
  const float x[4] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale};
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}

__global__ void AST_DEC_H_BOUNDS decimate2_tc_h_kernel(const DecimateTcParams p) {
  using namespace dth;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AST_TIMELINE_STAMP(dec, blockIdx.x, 0);
  unsigned char* a_hi = smem_raw;                            // [2 tiles][4 slices][4 chunk columns][130 rows][8 halves]
  unsigned char* a_lo = a_hi + 2 * kTileBytes;               // the residual image, same shape: with 227 KB per SM there is
  unsigned char* t_hi = a_lo + 2 * kTileBytes;               // room for whole tiles, so no producer waits for the cross-term MMAs
  unsigned char* t_lo = t_hi + kStripBytes;
  float* epi_buf = reinterpret_cast<float*>(t_lo + kStripBytes);   // [4 warps][32 rows][68] epilogue transpose
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kEpiFloats);
  // l_full[tile parity][slice]: the producers are held back by h_empty only (MMAs of tile n - 2 done), so they may arrive
  // for tile n + 1 before the MMA warp has waited for tile n: one barrier set per tile parity keeps every barrier at most
  // one phase ahead of its waiter
  uint64_t* l_full = bars;          // [2][4] producers -> MMA: slice staged (both images written); 4 warp arrivals
  uint64_t* h_empty = bars + 8;
  uint64_t* acc_full = bars + 10;
  uint64_t* acc_empty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  unsigned int* stored_count = reinterpret_cast<unsigned int*>(bars + 20);   // [4]: one per epilogue warp
  unsigned int* tile_max = reinterpret_cast<unsigned int*>(bars + 16);   // [3] bits of the largest |sample| of tiles n, n + 1, n + 2
  float* tile_inv = reinterpret_cast<float*>(bars + 18);                 // [4] 2^(-k - 15) of the last four tiles
  DtcTile* table = reinterpret_cast<DtcTile*>(bars + 32);                // [64] the CTA's first tiles, decoded once
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kStripBytes / 16; i += kThreads) {
    reinterpret_cast<uint4*>(t_hi)[i] = __ldg(p.strip_h_hi + i);
    reinterpret_cast<uint4*>(t_lo)[i] = __ldg(p.strip_h_lo + i);
  }
  // The CTA's tile list, decoded once (every role used to re-derive stage / clip / pointers of every tile: ~ 1 000 cycles
  // per tile on each role's critical path).  Only kernel parameters and `lengths` (an input of the call) are read.
  const DthWalk walk{p, dth_list(p), table};
  for (int i = tid; i < walk.l.n_total && i < kDthTableCap; i += kThreads) table[i] = dtc_decode(p, dth_tile_at(p, walk.l, i));
  if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      umma::mbar_init(l_full + i, kProducerGroup / 32);
      umma::mbar_init(l_full + 4 + i, kProducerGroup / 32);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(h_empty + i, 1);
      umma::mbar_init(acc_full + i, 1);
      umma::mbar_init(acc_empty + i, 4);
    }
    stored_count[0] = stored_count[1] = stored_count[2] = stored_count[3] = 0;
    tile_max[0] = tile_max[1] = tile_max[2] = 0u;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // (see decimate2_tc_kernel)
  pdl_launch_dependents();
  const int total = walk.l.n_total;   // positions in the CTA's list
  int tile = walk.next_live(0);

  if (warp < kProducers / 32) {
    // ================================================================= producers
    const int grp = warp >> 2, tg = tid & (kProducerGroup - 1);
    float4 v[16];
    unsigned stages_complete = 0;
    // 8-byte units within a slice: half (tg & 1) of chunk column (tg & 7) >> 1, row tg >> 3 (+ 16 i); with kRT = 2 mod 8
    // the 64-bit stores of a half-warp (8 pieces x 2 rows) fall into 32 different banks
    const int slot0 = 2 * (((tg & 7) >> 1) * kRT + (tg >> 3)) + (tg & 1);
    DtcTile cur;
    if (tile < total) {
      cur = walk.get(tile);
      dtc_deps_wait(p, cur, lane, stages_complete);
      dth_load_slice<0>(cur, tg, 2 * grp, v);
      dth_load_slice<1>(cur, tg, 2 * grp + 1, v);
    }
    for (int n = 0; tile < total; ++n) {
      const int next = walk.next_live(tile + 1);
      if ((warp & 3) == 0) DTC_STAMP(grp, n, 0);
      // the tile's scale: 2^k brings its largest magnitude into [2^8, 2^9)
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
      const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
      if (lane == 0) atomicMax(tile_max + n % 3, wmax);
      asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");
      int e = (int)(tile_max[n % 3] >> 23) - 127;      // floor(log2(max)); -127 for zero / subnormal tiles
      int k = 8 - e;
      k = k > 100 ? 100 : k;
      const float scale = __uint_as_float((uint32_t)(k + 127) << 23);
      if (tid == 0) {
        tile_inv[n & 3] = __uint_as_float((uint32_t)(127 - k - kTapShift) << 23);
        tile_max[(n + 2) % 3] = 0u;   // tile n - 1's slot: every producer read it before this barrier; next used by tile n + 2
      }
      // this A_hi buffer is free once the MMAs of tile n - 2 have completed
      umma::mbar_wait(h_empty + (n & 1), ((n >> 1) & 1) ^ 1);
      if ((warp & 3) == 0) DTC_STAMP(grp, n, 1);
      uint4* hi_tile = reinterpret_cast<uint4*>(a_hi + (n & 1) * kTileBytes);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int sl = 2 * grp + j;
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 2 + 2 * j);
        uint2* hi = reinterpret_cast<uint2*>(hi_tile + sl * (kSliceBytes / 16));
        uint2* lo = reinterpret_cast<uint2*>(a_lo + (n & 1) * kTileBytes + sl * kSliceBytes);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint2 h, l;
          dth_split4(v[8 * j + i], scale, h, l);
          hi[slot0 + 32 * i] = h;   // row + 16 i: 32 eight-byte units further
          lo[slot0 + 32 * i] = l;
        }
        umma::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(l_full + 4 * (n & 1) + sl);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 3 + 2 * j);
      }
      if (next < total) {
        cur = walk.get(next);
        dtc_deps_wait(p, cur, lane, stages_complete);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 6);
        dth_load_slice<0>(cur, tg, 2 * grp, v);
        dth_load_slice<1>(cur, tg, 2 * grp + 1, v);
        if ((warp & 3) == 0) DTC_STAMP(grp, n, 7);
      }
      tile = next;
    }
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issue
    const uint32_t idesc = instr_desc_f16(kM, kN);
    const uint64_t db_hi0 = umma::smem_desc(umma::smem_u32(t_hi), kStripRows * 16, 128) + (uint64_t)kStripRow0;
    const uint64_t db_lo0 = umma::smem_desc(umma::smem_u32(t_lo), kStripRows * 16, 128) + (uint64_t)kStripRow0;
    for (int n = 0; tile < total; ++n, tile = walk.next_live(tile + 1)) {
      const int q = n & 1;
      const uint32_t acc = tmem_base + (uint32_t)(q * kN);
      const uint32_t hi_addr = umma::smem_u32(a_hi + q * kTileBytes);
      DTC_STAMP(2, n, 0);
      umma::mbar_wait(acc_empty + q, ((n >> 1) & 1) ^ 1);
      umma::fence_after_thread_sync();
      DTC_STAMP(2, n, 1);
      for (int s = 0; s < kSlices; ++s) {     // cross terms first (see the accuracy note in the header)
        umma::mbar_wait(l_full + 4 * q + s, (n >> 1) & 1);
        umma::fence_after_thread_sync();
        DTC_STAMP(2, n, 2 + s);
        if (umma::elect_one_sync()) {
          const uint64_t da_hi = umma::smem_desc(hi_addr + (uint32_t)(s * kSliceBytes), kRT * 16, 128);
          const uint64_t da_lo = umma::smem_desc(umma::smem_u32(a_lo + q * kTileBytes + s * kSliceBytes), kRT * 16, 128);
#pragma unroll
          for (int k = 0; k < kKStepsPerSlice; ++k) {
            if (p.debug & 4) break;
            const uint64_t a_off = (uint64_t)(2 * k * kRT);                    // 16-byte units: chunk columns 2 k, 2 k + 1
            const uint64_t b_off = (uint64_t)(8 * (kKStepsPerSlice * s + k));  // strip row 56 - 8 gs
            mma_f16(acc, da_lo + a_off, db_hi0 - b_off, idesc, (s | k) ? 1u : 0u);
            mma_f16(acc, da_hi + a_off, db_lo0 - b_off, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (umma::elect_one_sync()) {
#pragma unroll
        for (int gs = 0; gs < kSlices * kKStepsPerSlice; ++gs) {
          if (p.debug & 4) break;
          const int s = gs >> 1, k = gs & 1;
          const uint64_t da_hi = umma::smem_desc(hi_addr + (uint32_t)(s * kSliceBytes), kRT * 16, 128);
          mma_f16(acc, da_hi + (uint64_t)(2 * k * kRT), db_hi0 - (uint64_t)(8 * gs), idesc, 1u);
        }
        umma::commit(h_empty + q);    // both images of this tile's buffer may be refilled
        umma::commit(acc_full + q);
      }
      __syncwarp();
      DTC_STAMP(2, n, 6);
    }
  } else if (warp < kPublishWarp) {
    // ================================================================= epilogue (warps 8-11)
    const int quad = warp - kEpilogueWarp0;
    float* stg = epi_buf + quad * 32 * kEpiStride;
    for (int n = 0; tile < total; ++n, tile = walk.next_live(tile + 1)) {
      const int q = n & 1;
      const DtcTile t = walk.get(tile);
      if (quad == 0) DTC_STAMP(3, n, 0);
      umma::mbar_wait(acc_full + q, (n >> 1) & 1);
      umma::fence_after_thread_sync();
      if (quad == 0) DTC_STAMP(3, n, 1);
      const float inv = *reinterpret_cast<volatile float*>(tile_inv + (n & 3));
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(q * kN);
      // Eight groups of eight columns, two register sets: the loads of group g + 1 are issued before group g is summed
      // and staged, so one TMEM round trip is exposed per tile instead of four (tcgen05.wait::ld waits for every
      // outstanding load, hence "wait, issue the next, work on this one").
      uint32_t ra[4][8], rb[4][8];
      auto issue = [&](uint32_t (&r)[4][8], int g) {
        umma::tmem_ld_32x8_nowait(lane_base + 3 * kP + 8 * g, r[0]);   // d = 0: this row
        umma::tmem_ld_32x8_nowait(lane_base + 2 * kP + 8 * g, r[1]);   // d = 1: wanted by the row above (lane - 1)
        umma::tmem_ld_32x8_nowait(lane_base + 1 * kP + 8 * g, r[2]);
        umma::tmem_ld_32x8_nowait(lane_base + 8 * g, r[3]);
      };
      auto finish = [&](const uint32_t (&r)[4][8], int g) {
        if (p.debug & 1) return;
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float s1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[1][c]), 1);
          const float s2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c]), 2);
          const float s3 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[3][c]), 3);
          o[c] = ((__uint_as_float(r[0][c]) + s1) + (s2 + s3)) * inv;
        }
        *reinterpret_cast<float4*>(stg + lane * kEpiStride + 8 * g) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(stg + lane * kEpiStride + 8 * g + 4) = make_float4(o[4], o[5], o[6], o[7]);
      };
      issue(ra, 0);
#pragma unroll
      for (int g = 0; g < kP / 8; g += 2) {
        umma::tmem_wait_ld();
        issue(rb, g + 1);
        finish(ra, g);
        umma::tmem_wait_ld();
        if (g + 2 < kP / 8) {
          issue(ra, g + 2);
        } else {   // last TMEM read of the tile: hand the accumulator back
          umma::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(acc_empty + q);
          if (quad == 0) DTC_STAMP(3, n, 2);
        }
        finish(rb, g + 1);
      }
      __syncwarp();
      const int j_grp = (t.row0 + kGroupValid * quad) * kP;
      if (p.debug & 1) {
      } else if (j_grp + kGroupValid * kP <= t.len_out) {
        // all 29 rows inside the signal: 14.5 float4 per lane, loads first, then the stores (16 lanes cover a 256-byte row)
        float4 val[15];
#pragma unroll
        for (int it = 0; it < 15; ++it) {
          const int idx = lane + 32 * it;
          if (it < 14 || lane < 16) val[it] = *reinterpret_cast<const float4*>(stg + (idx >> 4) * kEpiStride + 4 * (idx & 15));
        }
        float4* dst = reinterpret_cast<float4*>(t.y + j_grp) + lane;
#pragma unroll
        for (int it = 0; it < 15; ++it)
          if (it < 14 || lane < 16) dst[32 * it] = val[it];
      } else {
        for (int idx = lane; idx < kGroupValid * (kP / 4); idx += 32) {
          const int r = idx >> 4, c4 = idx & 15;
          const int j = j_grp + r * kP + 4 * c4;
          const float4 val = *reinterpret_cast<const float4*>(stg + r * kEpiStride + 4 * c4);
          if (j + 4 <= t.len_out) {
            *reinterpret_cast<float4*>(t.y + j) = val;
          } else {
            if (j < t.len_out) t.y[j] = val.x;
            if (j + 1 < t.len_out) t.y[j + 1] = val.y;
            if (j + 2 < t.len_out) t.y[j + 2] = val.z;
          }
        }
      }
      __syncwarp();
      if (quad == 0) DTC_STAMP(3, n, 3);
      // release to the publisher.  (Deferring it until the next tile has been drained - the fence waits ~ 1 500 cycles for
      // the stores' acknowledgements - was measured: the later completion flags cost the dependent tiles of other CTAs
      // more than the epilogue gains, 0.0990 -> 0.1023 ms.)
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(stored_count + quad, 1u);
      }
    }
  } else if (warp == kPublishWarp) {
    // ================================================================= publisher (see decimate2_tc_kernel)
    int* stage_done = p.flags + (long long)kDecStages * p.batch * p.tiles_per_clip[0];
    int n = 0;
    // (An L2 prefetch of the stage-0 tiles two to four tiles ahead - cp.async.bulk.prefetch.L2 issued from the producers or
    // from this warp - was measured: no gain, 0.0900 vs 0.0892 ms.)
    for (int g = 0; g < total; ++g) {
      const DtcTile t = walk.get(g);
      if (t.live) {   // every epilogue warp has released its rows of the CTA's n-th live tile (one counter per warp:
        ++n;          // a sum over the warps could be reached by three warps running a tile ahead of the fourth)
        unsigned long long t0 = 0;
        for (uint32_t spin = 0;; ++spin) {
          const bool ok = lane >= 4 || reinterpret_cast<volatile unsigned int*>(stored_count)[lane] >= (unsigned)n;
          if (__all_sync(0xffffffffu, ok)) break;
          __nanosleep(100);
          if ((spin & 1023) == 1023) {
            const unsigned long long now = umma::global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > umma::kPollTimeoutNs) __trap();
          }
        }
      }
      if (lane == 0) {
        if (t.live) {
          __threadfence();
          atomicAdd(dtc_flag(p, t.stage, t.clip, t.k), 4);
        }
        atomicAdd(stage_done + t.stage, 1);
      }
      __syncwarp();
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  AST_TIMELINE_STAMP_IF(warp == kMmaWarp && lane == 0, dec, blockIdx.x, 1);
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_base, kTmemCols);
  AST_TIMELINE_STAMP_IF(warp == kMmaWarp && lane == 0, dec, blockIdx.x, 3);
}

static int g_dec_half = 1;   // AST_DECIMATOR=tf32 selects the TF32-split kernel
void set_decimator_half(int on) { g_dec_half = on; }

int decimate_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(decimate2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dtc::kSmem));
  AST_CUDA_TRY(cudaFuncSetAttribute(decimate2_tc_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dth::kSmem));
  if (const char* env = getenv("AST_DEC_CARVEOUT"))   // diagnostic (co-residency experiment above): shared-memory share, %
    AST_CUDA_TRY(cudaFuncSetAttribute(decimate2_tc_h_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env)));
  return AST_OK;
}

int decimator_strip_floats() { return dtc::kStripFloats; }

#ifdef AST_TRACE
extern "C" int ast_debug_dec_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, g_dec_trace, sizeof(long long) * 4 * 64 * 8);
}
#endif

// host: the Toeplitz strip images T[jj][kk] = g[kk - 2 jj], jj = row - 252, as [chunk c][row][4] with kk = 4 c + e;
// hi values are exactly representable in TF32, lo is the TF32-truncated residual of the double tap.
void host_decimator_strip(const double* taps_scaled, float* strip_hi, float* strip_lo) {
  using namespace dtc;
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < kStripRows; ++i)
      for (int e = 0; e < 4; ++e) {
        const int jj = i - (kStripRow0 + 3 * kP);   // row 252 <-> jj = 0
        const int tap = 4 * c + e - 2 * jj;
        const double g = (tap >= 0 && tap < kDecTaps) ? taps_scaled[tap] : 0.0;
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
        strip_hi[(c * kStripRows + i) * 4 + e] = hi;
        strip_lo[(c * kStripRows + i) * 4 + e] = lo;
      }
}

// host: the FP16-split kernel's strip images T[jj][kk] = 2^15 g[kk - 2 jj], jj = row - 248, kk = 8 c + e < 16, as
// [chunk c][row][8 halves]: hi = FP16(value), lo = FP16(value - hi)
int decimator_strip_h_bytes() { return dth::kStripBytes; }
void host_decimator_strip_h(const double* taps_scaled, uint16_t* strip_hi, uint16_t* strip_lo) {
  using namespace dth;
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < kStripRows; ++i)
      for (int e = 0; e < 8; ++e) {
        const int jj = i - (kStripRow0 + 3 * kP);   // row 248 <-> jj = 0
        const int tap = 8 * c + e - 2 * jj;
        const double g = (tap >= 0 && tap < kDecTaps) ? std::ldexp(taps_scaled[tap], kTapShift) : 0.0;
        const __half hi = __float2half_rn((float)g);
        const __half lo = __float2half_rn((float)(g - (double)__half2float(hi)));
        memcpy(&strip_hi[(c * kStripRows + i) * 8 + e], &hi, 2);
        memcpy(&strip_lo[(c * kStripRows + i) * 8 + e], &lo, 2);
      }
}

int decimator_tile_outputs() { return dtc::kRowsOut * dtc::kP; }
int decimator_tiles_stage0(long long max_samples) {
  const long long rows = (octave_len(max_samples, 1) + dtc::kP - 1) / dtc::kP;
  return (int)((rows + dtc::kRowsOut - 1) / dtc::kRowsOut);
}

int decimator_tiles_of_stage(long long max_samples, int stage) {
  const long long rows = (octave_len(max_samples, stage + 1) + dtc::kP - 1) / dtc::kP;
  return (int)((rows + dtc::kRowsOut - 1) / dtc::kRowsOut);
}

size_t decimator_flag_bytes(int batch, long long max_samples) {
  const long long rows = (octave_len(max_samples, 1) + dtc::kP - 1) / dtc::kP;
  const long long tiles0 = (rows + dtc::kRowsOut - 1) / dtc::kRowsOut;
  // per-tile counters [stage][clip][tile], then one finished-tile counter per stage
  // ... and one more int: the CQT projection's tile queue (cqt_tc.cu), zeroed with the rest
  return sizeof(int) * ((size_t)kDecStages * (size_t)(batch > 0 ? batch : 1) * (size_t)(tiles0 > 0 ? tiles0 : 1) + kDecStages + 1);
}
long long decimator_stage_done_offset(int batch, long long max_samples) {
  return (long long)kDecStages * (batch > 0 ? batch : 1) * decimator_tiles_stage0(max_samples);
}

// the whole cascade: octave buffer i (1..6) of clip b lives at ws + b * ws_clip_stride + octave_offset(i)
int launch_decimate_cascade_tc(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch,
                               long long max_samples, long long wave_stride, float* ws, long long ws_clip_stride,
                               int* flags, cudaStream_t st, bool flags_zeroed) {
  if (batch == 0) return AST_OK;
  if (!flags) return fail(AST_ERR_WORKSPACE, "the decimator needs its completion-flag region of the workspace");
  DecimateTcParams p;
  p.wave = wave;
  p.wave_stride = wave_stride;
  p.ws = ws;
  p.ws_clip_stride = ws_clip_stride;
  for (int i = 0; i < kOctaves; ++i) p.oct_off[i] = i == 0 ? 0 : octave_offset(max_samples, i);
  p.lengths = lengths;
  p.max_samples = max_samples;
  p.batch = batch;
  long long total = 0;
  for (int s = 0; s < kDecStages; ++s) {
    const long long rows = (octave_len(max_samples, s + 1) + dtc::kP - 1) / dtc::kP;
    p.tiles_per_clip[s] = (int)((rows + dtc::kRowsOut - 1) / dtc::kRowsOut);
    p.tile_prefix[s] = (int)total;
    total += (long long)p.tiles_per_clip[s] * batch;
  }
  if (total >= (1LL << 31)) return fail(AST_ERR_INVALID_ARG, "too many decimator tiles for one call");
  p.tile_prefix[kDecStages] = (int)total;
  p.vec_ok0 = (wave_stride % 4 == 0 || batch == 1) && (reinterpret_cast<uintptr_t>(wave) & 15) == 0;
  p.flags = flags;
  p.strip_hi = plan->d_dec_strip_hi;
  p.strip_lo = plan->d_dec_strip_lo;
  p.strip_h_hi = reinterpret_cast<const uint4*>(plan->d_dec_strip_h_hi);
  p.strip_h_lo = reinterpret_cast<const uint4*>(plan->d_dec_strip_h_lo);
  {
    const char* env = getenv("AST_DEC_DEBUG");
    p.debug = env ? atoi(env) : 0;
  }
  // the fused feature call zeroes the counters in its prologue kernel (this kernel's programmatic primary) instead
  if (!flags_zeroed) AST_CUDA_TRY(cudaMemsetAsync(flags, 0, decimator_flag_bytes(batch, max_samples), st));
  long long ctas = total;
  if (ctas > plan->sm_count) ctas = plan->sm_count;  // persistent and co-resident: one CTA per SM (tiles wait on each other)
  // the chain stages (<= 2 tiles per clip) on one CTA per clip, when that leaves at least half of the grid free to go
  p.tail_stage = kDecStages;
  p.tail_ctas = 0;
  {
    // measured at 64 clips (scratch/tail_sweep.sh, FP16 kernel): off 0.2834 ms per feature step, 64 CTAs 0.2800, 48 0.2774,
    // 32 (two clips per CTA) 0.2750, 24 0.2738 but a slower statistics call (0.303 vs 0.293 ms), 16 0.2850
    int want = g_dec_half ? 32 : 64, clips_per_cta = g_dec_half ? 2 : 1;   // (the TF32 kernel's longer levels: 64 CTAs, 0.2980 vs 0.3015 ms)
    if (const char* env = getenv("AST_DEC_TAIL_CTAS")) want = atoi(env), clips_per_cta = 16;   // diagnostic sweep
    int first_chain = kDecStages;
    while (first_chain > 0 && p.tiles_per_clip[first_chain - 1] <= 2) --first_chain;
    const int k = batch < want ? batch : want;
    if (k > 0 && first_chain < kDecStages && first_chain > 0 && 2 * k <= ctas && batch <= clips_per_cta * k) {
      p.tail_stage = first_chain;
      p.tail_ctas = k;
    }
  }
  ProfileSpan span("decimate2_tc_kernel", st);
  if (g_dec_half)
    AST_CUDA_TRY(launch_with_pdl(decimate2_tc_h_kernel, dim3((unsigned)ctas), dtc::kThreads, dth::kSmem, st, p));
  else
    AST_CUDA_TRY(launch_with_pdl(decimate2_tc_kernel, dim3((unsigned)ctas), dtc::kThreads, dtc::kSmem, st, p));
  return AST_OK;
}

}  // namespace ast
