// cqt_tc.cu - K3 on the tensor cores: the constant-Q projection (the one dense contraction of the path)
// as tcgen05.mma kind::tf32 with a split for FP32-level accuracy, warp-specialised and software-pipelined.
//
//     resp_i[t, j] = sum_{n < 256} x_i[t h_i - 128 + n] K[n][j],   h_i = 256 >> i,  j < 24 (12 re, 12 im)
//
// (derivation in cqt.cu / plan.cu; replaces librosa.cqt's per-octave STFT + sparse basis product reached
// through utilityFunctions.py:52).  A tile is 128 frames of one (clip, octave): D[128 x 32] = A[128 x 256] *
// B[256 x 32].  A's rows are overlapping windows of the octave signal (hop h_i <= window), and they are never
// materialised: the signal is staged ONCE per tile as "blocks" of 16-byte chunks in chunk-column order
// smem[e][R] (e = chunk within a row stride of m = h_i / 4 chunks, R = row), and in the no-swizzle K-major
// layout (rows 16 B apart) frame r's chunk c' = m d + e sits at row r + d: a window shift is +16 d bytes on the
// descriptor start address.
// A block is eight chunk columns (32 samples of every row it holds):
//   octave 0 (m = 64): eight blocks (chunk columns 8 j .. 8 j + 7), 4 K-steps each, no shift needed
//   octave 1 (m = 32): four blocks of 129 rows, each serving the K-steps of d = 0 and d = 1
//   octave 2 (m = 16): two blocks of 131 rows, d = 0..3
//   octaves 3-6 (m = 8, 4, 2, 1): one block = one contiguous run of the signal, 135 .. 191 rows
// so a clip's seven octaves cost 1.76 MB of staging instead of the 3.5 MB of a per-pass slice scheme (the kernel
// is bound by the shared-memory / LSU data path, not by the tensor pipe: profiles/).
//
// Split precision: x = hi + lo (hi = TF32 truncation, lo exact residual), K = hi + lo (host, from doubles).
// B is stored as [B_hi | B_lo] (N = 64) so ONE MMA yields hi*hi (columns 0..31) and hi*lo (columns 32..63) for
// one A fetch - measured on B200 (scratch/mma_bench2.cu) an M128 K8 TF32 MMA with shared-memory operands costs
// max(~40, N / 2) cycles - and a second MMA adds lo*hi into columns 32..63.  The tensor core accumulates with
// round-toward-zero (measured -4.2e-8 relative per accumulating MMA), so K-steps alternate between two accumulators
// (16 steps each); the epilogue sums them in registers with round-to-nearest, applies the per-bin scale and
// (x - mean) * rstd, and stores to the flat / section layout (columns 513..596).
//
// One persistent CTA per SM, 16 warps:
//   warp  6     TMA: one thread walks the CTA's block sequence; per block it waits for the decimator tiles the block
//               reads (completion counters) and for a free stage, then issues one cp.async.bulk.tensor.3d PER CHUNK
//               COLUMN - a box of (4 samples, 128..192 rows, 1 clip) out of a per-octave tensor map whose row stride
//               is the octave's hop.  Rows 16 bytes apart are exactly the K-major no-swizzle operand layout, so the
//               boxes land where the MMA reads them (the hi image: the tensor core truncates raw FP32 to TF32 itself);
//               rows before the clip and past the map arrive as zeros (librosa's zero padding)
//   warps 0-5   splitters: lo[u] = x[u] - trunc_tf32(x[u]) at the same offset of the stage's second half, an
//               elementwise pass over shared memory (blocks that reach the clip's end re-read their tail chunks from
//               global with the bounds applied and patch both images).  Round 1's producers fetched the samples
//               themselves (LDG -> registers -> transposing stores): a thread's fence.proxy.async and releasing
//               mbarrier.arrive wait for its outstanding loads, which exposed the load latency, and the per-block
//               index algebra was the other half of their time.  (Inputs whose rows are not 16-byte aligned cannot be
//               described by a tensor map: they keep that register path.)
//   warp  7     MMA issue: one elected lane; the hi * [hi | lo] products of a block go out as soon as its boxes have
//               landed, the lo * hi products when the splitters are done; tcgen05.commit releases stages and publishes
//               accumulator sets
//   warps 8-15  epilogue : TMEM -> registers -> shared transpose -> global; two groups of four warps (one TMEM lane
//               quadrant each) take alternate tiles, i.e. one accumulator set each - the epilogue, not the tensor
//               pipe, is the longest stage of a tile (scratch/trace_cqt.py, scratch/dbg_cqt.sh)
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "umma.cuh"

namespace ast {
namespace cqt_tc {
constexpr int kM = 128;                 // frames per tile
constexpr int kN = 32;                  // 24 outputs padded to 32
constexpr int kKSteps = 32;             // 256-sample window / 8
constexpr int kGroupThreads = 96;       // producer group: three warps (a multiple of 8 and of every m < 8)
constexpr int kProducerWarps = 6;       // two groups, warps 0-2 and 3-5, alternate blocks
constexpr int kTmaWarp = 6;             // one thread: tensor-map loads into the raw ring
constexpr int kMmaWarp = 7;
constexpr int kEpilogueWarp0 = 8;       // register path: warps 8-11 even tiles, warps 12-15 odd tiles; TMA path: warps 8-11
constexpr int kThreads = 16 * 32;       // register path: 4 warps per scheduler, 128 registers per thread
// AST_CQT_LEAN (diagnostic): the TMA-path kernel with 12 warps (one epilogue group), 3 operand stages and <= 120
// registers, so that ONE 4-warp STFT CTA (34.8 KB, 16 K registers) fits next to it on every SM.  Measured on B200: it
// works (the STFT's first CTAs start under the projection), but the step gets SLOWER (0.298 -> 0.317 ms per 64 clips):
// this kernel is bound by shared-memory operand traffic and the instruction caches, which the STFT's warps also load,
// and the programmatic chain's tail overlap (21 us) shrinks to 8 us.  Kept off.
#ifdef AST_CQT_LEAN
constexpr int kThreadsTma = 12 * 32;
constexpr int kStagesTma = 3;
constexpr int kEpiGroupsTma = 1;
#else
constexpr int kThreadsTma = kThreads;
constexpr int kStagesTma = 4;
constexpr int kEpiGroupsTma = 2;
#endif
constexpr int kBlockCols = 8;                             // chunk columns per block (32 samples of every row)
constexpr int kStages = 4;                                // ring of staged blocks (hi + lo)
constexpr int kMaxBlockChunks = 8 * 136;                  // octaves 0-3: the largest blocks, 1088 chunks
constexpr int kAFloats = kMaxBlockChunks * 4;             // 4352 floats = 17 408 B per split term
constexpr int kStageFloats = 2 * kAFloats;                // hi + lo
constexpr int kBStepFloats = 2 * 2 * kN * 4;              // one K-step of [B_hi | B_lo]: [c 2][j 64][4] = 512 floats
constexpr int kBFloats = kKSteps * kBStepFloats;          // 16384 floats = 64 KB
constexpr int kMainAcc = 2;
constexpr int kSetCols = kMainAcc * 2 * kN;               // 128 TMEM columns per accumulator set
constexpr int kTmemCols = 256;
constexpr int kStage = (kMaxBlockChunks + kGroupThreads - 1) / kGroupThreads;  // 12 chunks per producer thread per block at most
constexpr int kEpiStride = 25;            // floats per staged row (24 values + 1: conflict-free row-per-lane writes)
constexpr int kEpiFloats = 8 * 32 * kEpiStride;  // one [32 rows][25] transpose buffer per epilogue warp
constexpr size_t kSmem = sizeof(float) * (kStages * kStageFloats + kBFloats + kEpiFloats) + 512 + 1024;   // barriers + tile ring, alignment slack
constexpr size_t kSmemTma = sizeof(float) * (kStagesTma * kStageFloats + kBFloats + kEpiFloats / 2 * kEpiGroupsTma) + 512 + 1024;

// rows per chunk column of an octave's blocks: 128 frames + window / hop - 1 shifts, rounded up to a multiple of 8
// (chunk columns are TMA destinations: 128-byte aligned)
__host__ __device__ constexpr int block_rows(int oct) {
  return oct == 0 ? 128 : oct == 1 ? 129 : 127 + (256 >> (8 - oct));  // 128, 129, 131, 135, 143, 159, 191
}
__host__ __device__ constexpr int block_rt(int oct) { return (block_rows(oct) + 7) & ~7; }  // 128, 136, 136, 136, 144, 160, 192
__host__ __device__ constexpr int block_cols(int oct) { return oct <= 3 ? 8 : 8 >> (oct - 3); }  // chunk columns per block
__host__ __device__ constexpr int blocks_per_tile(int oct) { return oct == 0 ? 8 : oct == 1 ? 4 : oct == 2 ? 2 : 1; }
}  // namespace cqt_tc

#ifdef AST_TRACE
// diagnostic build only (scratch/trace_cqt.py): clock64 stamps of CTA 0's pipeline roles
__device__ long long g_cqt_trace[3][512][6];
#define AST_STAMP(role, idx, k) \
  do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (idx) < 512) g_cqt_trace[role][idx][k] = clock64(); } while (0)
#else
#define AST_STAMP(role, idx, k) do {} while (0)
#endif

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
  const int* flags;    // the decimator's per-tile completion counters ([stage][clip][tile], 4 = done) or nullptr
  int flag_tiles0;     // tiles per clip of decimator stage 0 (row stride of the counters)
  const int* stage_done;          // finished-tile counter per decimator stage
  int* queue;                     // TMA path: the launch's next unassigned tile (zeroed with the counters), or nullptr: tiles dealt round-robin
  int stage_tiles[kOctaves - 1];  // tiles of each stage (all clips): stage_done[s] == stage_tiles[s] <=> stage complete
  int dec_tile_outputs;
  int debug;   // diagnostic bit mask (AST_CQT_DEBUG): 1 no epilogue stores, 2 no producer loads, 4 no MMAs, 8 no L2 prefetch
  OutSpec out;
  int use_tma;                 // 1: blocks arrive through the tensor maps below; 0: register path (unaligned input rows)
  int tma_end[kOctaves];       // samples of a clip's octave signal the map covers (whole rows of `hop` samples)
  CUtensorMap maps[kOctaves];  // (sample in row < hop, row, clip) over each octave's signals, box = one block
};

// One producer thread's share of a block: chunk i (0..n-1) is read at s0 + i * src_step (samples) and stored at
// slot0 + i * slot_step (16-byte units).
struct BlockPlan {
  const float* x;
  int len;
  int s0, src_step;
  int slot0, slot_step;
  int n;
  bool vec_ok;
  bool interior;  // whole block inside [0, len) and 16-byte loads legal: no bounds checks
                  // (TMA path: no sample at or past ok_end - rows before the clip arrive as zeros by themselves)
  int ok_end;     // TMA path: min(len, samples the tensor map covers); chunks reaching past it are re-read from global
  bool coherent;  // octave >= 1: written by the decimator launch this kernel overlaps with -> loads go through L2
  const int* dep; // first completion counter this block waits for (nullptr: none), dep_n of them
  int dep_n;
  int dep_stage;
};

// Tiles are numbered octave by octave (octave, clip, tile in clip): octave 0 needs nothing from the decimator and
// octave i only its stage i - 1, so the kernel can start while the decimator launch before it is still draining.
__device__ __forceinline__ void decode_tile(const CqtTcParams& p, int tile, int& b, int& oct, int& t0) {
  const int tiles_per_oct = p.tiles_per_clip_oct * p.batch;
  oct = tile / tiles_per_oct;
  const int rem = tile - oct * tiles_per_oct;
  b = rem / p.tiles_per_clip_oct;
  t0 = (rem - b * p.tiles_per_clip_oct) * cqt_tc::kM;
}

// A tile none of whose 128 frames exists in its clip (ragged batch: the clip is shorter than the padded length).  Every
// role skips it - no boxes, no MMAs, no accumulator hand-over - except the epilogue, which writes its rows as zeros.
__device__ __forceinline__ bool tile_dead(const CqtTcParams& p, int b, int t0) {
  return p.lengths != nullptr && t0 >= num_frames(p.lengths[b]);
}

// tid: thread index within its producer group (0..95)
__device__ __forceinline__ BlockPlan plan_block(const CqtTcParams& p, int b, int oct, int t0, int j, int tid,
                                                unsigned stages_complete) {
  using namespace cqt_tc;
  BlockPlan s;
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  const int hop = kHop >> oct, m = hop >> 2;
  const int rt = block_rt(oct), rows = block_rows(oct);
  s.len = (int)((len0 + (1LL << oct) - 1) >> oct);
  s.x = oct == 0 ? p.wave + (long long)b * p.wave_stride : p.ws + (long long)b * p.ws_clip_stride + p.oct_off[oct];
  s.vec_ok = oct == 0 ? p.vec_ok : true;
  const int first = t0 * hop - kCqtNfft / 2 + 32 * j;  // first sample of the block (j > 0 only for octaves 0..2)
  int n_chunks, last;
  if (m >= kBlockCols) {
    // eight chunk columns: chunk u = tg + 96 i -> row R = (tg >> 3) + 12 i, column e = tg & 7
    s.s0 = first + (tid >> 3) * hop + 4 * (tid & 7);
    s.src_step = (kGroupThreads / kBlockCols) * hop;
    s.slot0 = (tid & 7) * rt + (tid >> 3);
    s.slot_step = kGroupThreads / kBlockCols;
    n_chunks = kBlockCols * rows;
    last = first + (rows - 1) * hop + 32;
  } else {
    // one contiguous run: chunk u = tg + 96 i -> row R = u / m, column e = u % m (m divides 96)
    const int lg = 6 - oct;  // log2(m)
    s.s0 = first + 4 * tid;
    s.src_step = 4 * kGroupThreads;
    s.slot0 = (tid & (m - 1)) * rt + (tid >> lg);
    s.slot_step = kGroupThreads >> lg;
    n_chunks = m * rows;
    last = first + 4 * n_chunks;
  }
  s.n = tid < n_chunks ? (n_chunks - tid + kGroupThreads - 1) / kGroupThreads : 0;
  s.interior = first >= 0 && last <= s.len && s.vec_ok;
  s.ok_end = s.len;
  if (p.use_tma) {
    s.ok_end = s.len < p.tma_end[oct] ? s.len : p.tma_end[oct];
    s.interior = last <= s.ok_end;
  }
  s.coherent = oct > 0;
  s.dep = nullptr;
  s.dep_n = 0;
  if (oct > 0 && p.flags && !p.use_tma && !((stages_complete >> (oct - 1)) & 1)) {  // (a finished stage needs no bookkeeping;
                                                                                   // TMA path: the TMA thread waits)
    // decimator stage oct - 1 produced samples [lo, hi) of this octave in tiles lo / 7424 ... (hi - 1) / 7424
    const int lo = first > 0 ? first : 0, hi = last < s.len ? last : s.len;
    if (hi > lo) {
      const int k_lo = lo / p.dec_tile_outputs, k_hi = (hi - 1) / p.dec_tile_outputs;
      s.dep = p.flags + ((long long)(oct - 1) * p.batch + b) * p.flag_tiles0 + k_lo;
      s.dep_n = k_hi - k_lo + 1;
      s.dep_stage = oct - 1;
    }
  }
  return s;
}

// The MMAs of one staged block of octave OCT, issued by one elected lane.  With OCT a template parameter every
// A-descriptor offset is a compile-time constant.  K-step ks uses window chunks c' = 2 ks, 2 ks + 1.
template <int OCT>
__device__ __forceinline__ void issue_block(uint32_t a_hi_addr, uint32_t b_addr, uint32_t acc_set, int j,
                                            uint32_t idesc64, uint32_t idesc32, int term_lo, int term_hi, bool sw128) {
  using namespace cqt_tc;
  constexpr int m = (kHop >> OCT) >> 2;
  constexpr int rt = block_rt(OCT);
  constexpr uint32_t lbo = m == 1 ? 16u : (uint32_t)rt * 16u;
  constexpr int n_steps = kKSteps / blocks_per_tile(OCT);  // 4, 8, 16 or 32 K-steps per block
  // TMA path, octaves 0-3: the block is ONE box of 128-byte rows in the SWIZZLE_128B layout; a window shift by d rows is
  // + 128 d bytes on the start address, a K-step inside the row + 32 bytes.  The tensor core applies the swizzle to the
  // ADDRESS bits (chunk bits 4..6 ^= row bits 7..9), so a start that is not 1024-byte aligned needs no base offset in the
  // descriptor (measured on B200: base offset d mod 8 gives wrong rows, 0 matches the oracle to 1e-6).
  // Octaves 4-6 (rows shorter than 128 bytes) and the register path: chunk columns, no swizzle.
  const bool sw = sw128 && OCT <= 3;
  const uint64_t da_hi0 = sw ? umma::smem_desc_sw128(a_hi_addr) : umma::smem_desc(a_hi_addr, lbo, 128);
  const uint64_t da_lo0 = sw ? umma::smem_desc_sw128(a_hi_addr + kAFloats * 4) : umma::smem_desc(a_hi_addr + kAFloats * 4, lbo, 128);
  const uint64_t db0 = umma::smem_desc(b_addr, 2 * kN * 16, 128);
  // The products of one K-step go to accumulator ks & 1: first every hi * [hi | lo] of the block, then every
  // lo * hi, so that MMAs on the same TMEM columns stay several issues apart.
#pragma unroll
  for (int term = 0; term < 2; ++term) {
    if (term < term_lo || term > term_hi) continue;   // (TMA path: the hi terms go out before the lo image exists)
#pragma unroll
    for (int i = 0; i < n_steps; ++i) {
      int ks_c;        // K-step for j == 0 (compile time); the block index adds 4 j (octaves 0..2)
      int a_units;     // A start-address offset in 16-byte units
      if (sw) {
        const int d = i >> 2, k = i & 3;            // row r + d of the box, K-step k of its 128 bytes
        ks_c = (m / 2) * d + k;
        a_units = 8 * d + 2 * k;
      } else if (OCT <= 2) {
        const int d = i >> 2, k = i & 3;            // block j holds chunks m d + 8 j + e, e < 8, at row r + d
        ks_c = (m / 2) * d + k;
        a_units = 2 * k * rt + d;
      } else {
        const int c = 2 * i;
        ks_c = i;
        a_units = m == 1 ? c : (c / m) + (c % m) * rt;
      }
      const uint64_t b_off = (uint64_t)((ks_c + (OCT <= 2 ? 4 * j : 0)) * (kBStepFloats / 4));
      const uint32_t acc = acc_set + (uint32_t)((ks_c & (kMainAcc - 1)) * 2 * kN);
      if (term == 0)  // hi * [hi | lo] -> columns 0..63 of the accumulator; the tile's first two K-steps overwrite
        umma::mma_tf32(acc, da_hi0 + (uint64_t)a_units, db0 + b_off, idesc64, (i >= kMainAcc || j > 0) ? 1u : 0u);
      else            // lo * hi -> columns 32..63
        umma::mma_tf32(acc + kN, da_lo0 + (uint64_t)a_units, db0 + b_off, idesc32, 1u);
    }
  }
}

// One out-of-line copy of the seven unrolled per-octave MMA sequences (every descriptor offset an immediate: ~4
// instructions per MMA; a runtime loop over the same formulas cost ~280 cycles per MMA on the single issuing lane and
// made the MMA warp the kernel's bottleneck).  Out of line because the kernel's roles share the SM's instruction caches:
// inlined at both call sites of the TMA path (hi terms early, lo terms late) the sequences doubled, and a third of all
// stall samples of that version were instruction fetches (ncu: stall_no_inst).
__device__ __noinline__ void issue_block_any(int oct, uint32_t a_hi_addr, uint32_t b_addr, uint32_t acc_set, int j,
                                             uint32_t idesc64, uint32_t idesc32, int term_lo, int term_hi, bool sw128) {
  switch (oct) {
    case 0: issue_block<0>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
    case 1: issue_block<1>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
    case 2: issue_block<2>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
    case 3: issue_block<3>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
    case 4: issue_block<4>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
    case 5: issue_block<5>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
    default: issue_block<6>(a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, sw128); break;
  }
}

// The CTA's k-th tile.  Round-robin (tile = CTA + k grid) makes every CTA's work equal, but the persistent CTAs of this
// kernel do not start together - each needs a whole SM and gets it when the kernel before it leaves one (the feature
// call's STFT drains over ~ 30 us) - so equal lists end as staggered as they start.  With a queue the TMA thread draws
// the next tile number from a global counter (tiles are numbered heaviest octave first) and publishes it in a small
// shared-memory ring the other roles read; total marks the end.
constexpr int kTileRing = 32;   // entries; the TMA thread is at most ~ 7 tiles ahead of the epilogue (4 stages + 2 accumulator sets)
// An entry carries its own validity: (lap + 1) << 24 | tile, lap = k / kTileRing - ONE store publishes it (a separate
// "entries published" word would need a fence between the two stores, and a fence in the TMA thread waits for whatever
// that thread has in flight).  Tile numbers stay below 2^24 (65 535 clips x 49 tiles).
template <bool kQueue>
struct TileSeq {
  volatile int* ring;   // [kTileRing] entries (zero: never valid), [kTileRing + 1 + w]: positions epilogue warp w has read
  int total;
  static constexpr bool dynamic = kQueue;   // compile time: the kernel is sensitive to its code size (instruction fetches)
  __device__ __forceinline__ int get(int k) const {
    if (!dynamic) {
      const long long t = (long long)blockIdx.x + (long long)k * gridDim.x;
      return t < total ? (int)t : total;
    }
    const int tag = (k / kTileRing + 1) & 0x7F;
    int e;
    while (((e = ring[k & (kTileRing - 1)]) >> 24) != tag) __nanosleep(20);
    return e & 0xFFFFFF;
  }
};
template <bool kQueue>
struct TileFetcher {   // the TMA thread only
  TileSeq<kQueue> seq;
  int* queue;
  int fetched;
  bool ended;
  int n_epi;   // epilogue warps: ring[kTileRing + 1 + w] = positions warp w has read
  // One draw in flight (draw_ahead): the atomic's round trip (~ 1 000 cycles; 21 tiles per CTA: 10 us of this thread's
  // time when each draw was awaited on the spot) overlaps this thread's work on the tile before it.
  int inflight;
  bool have_inflight;
  __device__ __forceinline__ void ensure(int upto) {   // entries 0 .. upto are published (or the end marker is)
    while (kQueue && fetched <= upto && !ended) {
      // Dead tiles (ragged batch) cost this thread nothing but the epilogue a tile of zero rows, so it could run any
      // distance ahead: an entry is overwritten only when every epilogue warp - the last readers - is past it.  (The
      // epilogue never waits for a tile this thread has not issued yet: it is behind, so this cannot deadlock.)
      for (;;) {
        int slowest = fetched;
        for (int w = 0; w < n_epi; ++w) {
          const int r = seq.ring[kTileRing + 1 + w];
          slowest = r < slowest ? r : slowest;
        }
        if (fetched - slowest < kTileRing - 1) break;
        __nanosleep(100);
      }
      int t = have_inflight ? inflight : atomicAdd(queue, 1);
      have_inflight = false;
      if (t >= seq.total) t = seq.total, ended = true;
      seq.ring[fetched & (kTileRing - 1)] = (((fetched / kTileRing + 1) & 0x7F) << 24) | t;
      ++fetched;
    }
  }
  __device__ __forceinline__ void draw_ahead() {   // start the next draw (its result is awaited by the next ensure)
    if (kQueue && !ended && !have_inflight) inflight = atomicAdd(queue, 1), have_inflight = true;
  }
  __device__ __forceinline__ int peek(int k) {   // tile k if there is one, else total
    if (!kQueue) return seq.get(k);
    ensure(k);
    return k < fetched ? (seq.ring[k & (kTileRing - 1)] & 0xFFFFFF) : seq.total;
  }
};

AST_TIMELINE_DEFINE(cqt)

template <bool kTma, bool kQueue>
__global__ void __launch_bounds__(kTma ? cqt_tc::kThreadsTma : cqt_tc::kThreads, 1)
    cqt_tc_kernel(const __grid_constant__ CqtTcParams p) {
  using namespace cqt_tc;
  constexpr int kNStages = kTma ? kStagesTma : kStages;   // operand stages of this variant
  constexpr int kEpiGroups = kTma ? kEpiGroupsTma : 2;
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  // stages hold 128-byte-swizzled TMA boxes: 1024-byte aligned (the swizzle pattern repeats every 8 rows of 128 bytes)
  unsigned char* smem_raw = smem_dyn + ((1024u - (umma::smem_u32(smem_dyn) & 1023u)) & 1023u);
  float* a_stage = reinterpret_cast<float*>(smem_raw);            // [4 stages][hi | lo]
  float* b_img = a_stage + kNStages * kStageFloats;               // 64 KB
  float* epi_buf = b_img + kBFloats;                              // [8 warps][32][25] epilogue transpose
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kEpiFloats / 2 * kEpiGroups);
  // stage s = block number % 4 always belongs to producer group s % 2, so every barrier is completed and waited in
  // strict phase order by one party on each side
  uint64_t* full = bars;            // [4] producers -> MMA   (TMA path: 6 arrivals, the splitter warps: lo image written;
                                    //                         register path: 3 arrivals, the warps of one producer group)
  uint64_t* empty = bars + 4;       // [4] MMA -> producers / TMA   (tcgen05.commit)
  uint64_t* acc_full = bars + 8;    // [2] MMA -> epilogue    (tcgen05.commit)
  uint64_t* acc_empty = bars + 10;  // [2] epilogue -> MMA    (4 arrivals: one per epilogue warp)
  uint64_t* hi_full = bars + 12;    // [4] TMA -> splitters + MMA   (1 arrival + the boxes' bytes: hi image landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint64_t* b_full = bars + 17;     // TMA variants: the B images have landed
  volatile int* tile_ring = reinterpret_cast<volatile int*>(bars + 18);   // [32] entries, [1] unused, [8] read by epilogue warp w
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  AST_TIMELINE_STAMP(cqt, blockIdx.x, 0);
  pdl_launch_dependents();
  // B images: resident for the CTA's lifetime.  TMA variants: one 64 KB bulk copy, started below as soon as its barrier
  // exists, lands while the TMEM is allocated and the first boxes are requested; the MMA warp waits for it once.
  if (!kTma)
    for (int i = tid; i < kBFloats / 4; i += (int)blockDim.x)
      reinterpret_cast<float4*>(b_img)[i] = __ldg(reinterpret_cast<const float4*>(p.bmat) + i);
  if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < kNStages; ++i) {
      umma::mbar_init(full + i, kTma ? kProducerWarps : kGroupThreads / 32);
      umma::mbar_init(empty + i, 1);
      umma::mbar_init(hi_full + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(acc_full + i, 1);
      umma::mbar_init(acc_empty + i, 4);
    }
    for (int i = 0; i < kTileRing + 9; ++i) tile_ring[i] = 0;
    if (kTma) {
      umma::mbar_init(b_full, 1);
      umma::mbar_arrive_expect_tx(b_full, (uint32_t)(kBFloats * 4));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(umma::smem_u32(b_img)), "l"(p.bmat), "r"((uint32_t)(kBFloats * 4)), "r"(umma::smem_u32(b_full)) : "memory");
    }
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Without completion counters (FMA decimator) wait for the whole previous launch; with them the kernel runs into the
  // decimator's tail and every block waits only for the decimator tiles it reads (the counters are cleared by a
  // memset BEFORE the decimator launch, which also orders this kernel after the previous call's kernels).
  if (!p.flags) pdl_wait();
  const int total = p.tiles_per_clip_oct * kOctaves * p.batch;  // gridDim.x <= total
  const TileSeq<kQueue> seq{tile_ring, total};

  if (warp < kProducerWarps) {
    // ================================================================= producers
    // The loads of block k + 1 are issued right AFTER block k has been published: fence.proxy.async waits for every
    // outstanding load of the thread (measured: prefetching before the fence made the kernel 40 % slower), so the
    // latency hides behind the wait for the next stage to be released instead.
    float4 v[kStage];
    // a block of octave >= 1 waits for the decimator tiles that produce its samples (one lane polls for the warp;
    // bounded: a broken chain traps instead of hanging the GPU)
    unsigned stages_complete = 0;  // bit s: decimator stage s has been seen complete (warp-uniform)
    auto wait_deps = [&](const BlockPlan& sp) {
      if (!sp.dep || ((stages_complete >> sp.dep_stage) & 1)) return;
      {  // one look at the stage's finished-tile counter: once it is full, no block of this octave polls again
        int full = 0;
        if (lane == 0) {
          int f;
          asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.stage_done + sp.dep_stage) : "memory");
          full = f >= p.stage_tiles[sp.dep_stage] ? 1 : 0;
          if (full) __threadfence();
        }
        if (__shfl_sync(0xffffffffu, full, 0)) {
          stages_complete |= 1u << sp.dep_stage;
          return;
        }
      }
      unsigned long long t_first = 0;
      for (uint32_t spin = 0;; ++spin) {
        int ok = 1;
        if (lane == 0) {
          for (int k = 0; k < sp.dep_n; ++k) {
            int f;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(sp.dep + k) : "memory");
            ok &= f >= 4 ? 1 : 0;
          }
          if (ok) __threadfence();
        }
        if (__shfl_sync(0xffffffffu, ok, 0)) return;
        __nanosleep(200);
        if ((spin & 1023) == 1023) {
          const unsigned long long now = umma::global_ns();
          if (t_first == 0) t_first = now;
          if (now - t_first > umma::kPollTimeoutNs) __trap();
        }
      }
    };
    auto issue_loads = [&](const BlockPlan& sp) {
      wait_deps(sp);
      if (p.debug & 2) {
#pragma unroll
        for (int i = 0; i < kStage; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (sp.interior && !sp.coherent) {
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) v[i] = __ldg(reinterpret_cast<const float4*>(sp.x + sp.s0 + i * sp.src_step));
      } else if (sp.interior) {
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) v[i] = umma::ld_cg_f4(sp.x + sp.s0 + i * sp.src_step);
      } else if (!sp.coherent) {
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) v[i] = umma::load4_zero_ext(sp.x, sp.s0 + i * sp.src_step, sp.len, sp.vec_ok);
      } else {
#pragma unroll
        for (int i = 0; i < kStage; ++i)
          if (i < sp.n) v[i] = umma::load4_zero_ext_cg(sp.x, sp.s0 + i * sp.src_step, sp.len);
      }
    };
    // group g takes blocks g, g + 2, g + 4, ... of this CTA's block sequence
    const int grp = warp / 3, tg = tid - grp * kGroupThreads;
    int tile = blockIdx.x, j = 0;
    int b, oct, t0;
    if (tile < total) decode_tile(p, tile, b, oct, t0);
    auto advance = [&]() {  // next block: the same tile's next chunk columns, or the next tile's first block
      if (++j == (oct == 0 ? 8 : oct == 1 ? 4 : oct == 2 ? 2 : 1)) {
        tile += gridDim.x;
        j = 0;
        if (tile < total) decode_tile(p, tile, b, oct, t0);
        // one thread asks L2 for the signal span of the tile after that one, so its loads find it there
        if (tid == 0 && !kTma && tile + (int)gridDim.x < total && !(p.debug & 8)) {
          int b2, oct2, t2;
          decode_tile(p, tile + gridDim.x, b2, oct2, t2);
          const long long len2 = ((p.lengths ? p.lengths[b2] : p.max_samples) + (1LL << oct2) - 1) >> oct2;
          const float* x2 = oct2 == 0 ? p.wave + (long long)b2 * p.wave_stride
                                      : p.ws + (long long)b2 * p.ws_clip_stride + p.oct_off[oct2];
          const int hop2 = kHop >> oct2;
          long long lo = (long long)t2 * hop2 - kCqtNfft / 2, hi = lo + (long long)(kM - 1) * hop2 + kCqtNfft;
          if (lo < 0) lo = 0;
          if (hi > len2) hi = len2;
          const uintptr_t a0 = (reinterpret_cast<uintptr_t>(x2 + lo) + 15) & ~(uintptr_t)15;
          const uintptr_t a1 = reinterpret_cast<uintptr_t>(x2 + hi) & ~(uintptr_t)15;
          if (a1 > a0) umma::prefetch_l2_bulk(reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0));
        }
      }
    };
    if (grp == 1 && tile < total) advance();   // group 1 starts at block 1
    if (kTma) {
      // ------------------------------------------------------------- splitters (blocks arrive through TMA)
      // The TMA thread has written the block's hi image in place (chunk-column operand layout).  All six warps make
      // the lo image: lo[u] = x[u] - trunc_tf32(x[u]) at the SAME offset of the stage's second half - an elementwise
      // pass, no index algebra.  Only a block that reaches the clip's end (or the end of what the tensor map covers)
      // needs its tail chunks re-read with the bounds applied; those are patched in the hi image as well.
      int item = 0;
      for (int k = 0, tile; (tile = seq.get(k)) < total; ++k) {
        int b, oct, t0;
        decode_tile(p, tile, b, oct, t0);
        if (tile_dead(p, b, t0)) continue;
        const int hop = kHop >> oct, rt = block_rt(oct), cols = block_cols(oct);
        const int n_units = cols * rt;
        const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
        const int len = (int)((len0 + (1LL << oct) - 1) >> oct);
        const int ok_end = len < p.tma_end[oct] ? len : p.tma_end[oct];
        const float* x = oct == 0 ? p.wave + (long long)b * p.wave_stride : p.ws + (long long)b * p.ws_clip_stride + p.oct_off[oct];
        for (int j = 0; j < blocks_per_tile(oct); ++j, ++item) {
          const int s = item % kNStages;
          const int first = t0 * hop - kCqtNfft / 2 + 32 * j;
          const bool interior = first + (rt - 1) * hop + 4 * cols <= ok_end;
          float4* hi = reinterpret_cast<float4*>(a_stage + s * kStageFloats);
          float4* lo = hi + kAFloats / 4;
          if (warp == 0) AST_STAMP(0, item, 0);
          umma::mbar_wait(hi_full + s, (item / kNStages) & 1);   // the boxes have landed (and the stage's old MMAs are done)
          if (warp == 0) AST_STAMP(0, item, 2);
          if (interior) {
            for (int u = tid; u < n_units; u += kProducerWarps * 32) {
              const float4 x4 = hi[u];
              float4 h, l;
              umma::split_tf32(x4, h, l);
              lo[u] = l;
            }
          } else {
            for (int u = tid; u < n_units; u += kProducerWarps * 32) {
              float4 x4 = hi[u];
              // octaves 0-3: 128-byte rows, chunk e of row r at position e ^ (r & 7); octaves 4-6: chunk columns
              const int r = oct <= 3 ? u >> 3 : u % rt, e = oct <= 3 ? (u & 7) ^ (r & 7) : u / rt;
              const int smp = first + r * hop + 4 * e;
              if (smp + 3 >= ok_end) {
                x4 = oct > 0 ? umma::load4_zero_ext_cg(x, smp, len) : umma::load4_zero_ext(x, smp, len, true);
                hi[u] = x4;
              }
              float4 h, l;
              umma::split_tf32(x4, h, l);
              lo[u] = l;
            }
          }
          if (warp == 0) AST_STAMP(0, item, 3);
          umma::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(full + s);
          if (warp == 0) AST_STAMP(0, item, 4);
        }
      }
    } else if (!kTma) {
    BlockPlan sp;
    if (tile < total) {
      sp = plan_block(p, b, oct, t0, j, tg, stages_complete);
      issue_loads(sp);
    }
    for (int item = grp; tile < total; item += 2) {
      const int s = item % kStages;
      if (warp == 0) AST_STAMP(0, item, 0);
      // the MMAs that read this stage kStages blocks ago are done
      umma::mbar_wait(empty + s, ((item / kStages) & 1) ^ 1);
      if (warp == 0) AST_STAMP(0, item, 2);
      float4* a_hi = reinterpret_cast<float4*>(a_stage + s * kStageFloats);
      float4* a_lo = a_hi + kAFloats / 4;
#pragma unroll
      for (int i = 0; i < kStage; ++i)
        if (i < sp.n) {
          float4 h, l;
          umma::split_tf32(v[i], h, l);
          a_hi[sp.slot0 + i * sp.slot_step] = v[i];   // raw: the tensor core truncates to TF32 itself (measured)
          a_lo[sp.slot0 + i * sp.slot_step] = l;
        }
      if (warp == 0) AST_STAMP(0, item, 3);
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(full + s);
      if (warp == 0) AST_STAMP(0, item, 4);
      // this group's next block is two blocks further
      advance();
      if (tile < total) advance();
      if (tile < total) {
        sp = plan_block(p, b, oct, t0, j, tg, stages_complete);
        issue_loads(sp);
      }
      if (warp == 0) AST_STAMP(0, item, 1);
    }
    }
  } else if (warp == kTmaWarp) {
    // ================================================================= TMA (one thread)
    if (kTma && lane == 0) {
      for (int i = 0; i < kOctaves; ++i) umma::prefetch_tensormap(&p.maps[i]);
      unsigned stages_complete = 0;
      int item = 0;
      TileFetcher<kQueue> fetch{seq, p.queue, 0, false, 4 * kEpiGroups, 0, false};
      for (int k = 0, tile; (tile = fetch.peek(k)) < total; ++k) {
        int b, oct, t0;
        decode_tile(p, tile, b, oct, t0);
        if (tile_dead(p, b, t0)) continue;
        const int hop = kHop >> oct;
        const int n_blocks = blocks_per_tile(oct);
        const int cols = block_cols(oct), rt = block_rt(oct);
        const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
        const int len = (int)((len0 + (1LL << oct) - 1) >> oct);
        // one thread asks L2 for the signal span of the tile after next, so its box finds it there
        // (Queue variant: no look-ahead - every tile drawn early is a tile another CTA cannot take, and the CTAs should
        // end together: looking two tiles ahead 0.2657 ms per feature step, one 0.2634, none 0.2611.  The one draw in
        // flight stays.)
        const int tile2 = kQueue ? total : fetch.peek(k + 2);
        if (tile2 < total && !(p.debug & 8)) {
          int b2, oct2, t2;
          decode_tile(p, tile2, b2, oct2, t2);
          const long long len2 = ((p.lengths ? p.lengths[b2] : p.max_samples) + (1LL << oct2) - 1) >> oct2;
          const float* x2 = oct2 == 0 ? p.wave + (long long)b2 * p.wave_stride
                                      : p.ws + (long long)b2 * p.ws_clip_stride + p.oct_off[oct2];
          const int hop2 = kHop >> oct2;
          long long lo = (long long)t2 * hop2 - kCqtNfft / 2, hi = lo + (long long)(kM - 1) * hop2 + kCqtNfft;
          if (lo < 0) lo = 0;
          if (hi > len2) hi = len2;
          const uintptr_t a0 = (reinterpret_cast<uintptr_t>(x2 + lo) + 15) & ~(uintptr_t)15;
          const uintptr_t a1 = reinterpret_cast<uintptr_t>(x2 + hi) & ~(uintptr_t)15;
          // (a span the decimator has not produced yet would only prefetch stale lines of the workspace: harmless,
          // the box itself is issued after the counters below say the data is there)
          if (a1 > a0) umma::prefetch_l2_bulk(reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0));
        }
        for (int j = 0; j < n_blocks; ++j, ++item) {
          const int s = item % kNStages;
          const int first = t0 * hop - kCqtNfft / 2 + 32 * j;
          // the decimator tiles that produce the block's samples (octave >= 1), unless their whole stage is known done
          if (oct > 0 && p.flags && !((stages_complete >> (oct - 1)) & 1)) {
            int f;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.stage_done + oct - 1) : "memory");
            if (f >= p.stage_tiles[oct - 1]) {
              __threadfence();
              stages_complete |= 1u << (oct - 1);
            } else {
              const int last = first + (rt - 1) * hop + 4 * cols;
              const int lo = first > 0 ? first : 0, hi = last < len ? last : len;
              if (hi > lo) {
                const int k_lo = lo / p.dec_tile_outputs, k_hi = (hi - 1) / p.dec_tile_outputs;
                const int* dep = p.flags + ((long long)(oct - 1) * p.batch + b) * p.flag_tiles0;
                unsigned long long t_first = 0;
                for (uint32_t spin = 0;; ++spin) {
                  int ok = 1;
                  for (int k = k_lo; k <= k_hi; ++k) {
                    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(dep + k) : "memory");
                    ok &= f >= 4 ? 1 : 0;
                  }
                  if (ok) break;
                  __nanosleep(200);
                  if ((spin & 1023) == 1023) {
                    const unsigned long long now = umma::global_ns();
                    if (t_first == 0) t_first = now;
                    if (now - t_first > umma::kPollTimeoutNs) __trap();
                  }
                }
                __threadfence();   // acquire: the decimator's stores are visible to the loads issued from here on
              }
            }
          }
          umma::mbar_wait(empty + s, ((item / kNStages) & 1) ^ 1);   // the MMAs that last read this stage are done
          umma::mbar_arrive_expect_tx(hi_full + s, (uint32_t)(cols * rt * 16));
          // octaves 4-6: chunk column e of the block = samples first + 4 e + R hop, R < rt: a box of (4 samples, rt rows,
          // 1 clip) at tensor element (col0 + 4 e, row0), first = row0 hop + col0 - rows 16 bytes apart in shared memory
          // are exactly the K-major, no-swizzle operand layout, so the box lands where the MMA reads it
          const int q = 32 * j - kCqtNfft / 2;                         // first - t0 hop, in [-128, 128)
          const int dr = q >= 0 ? q / hop : -((-q + hop - 1) / hop);   // floor(q / hop)
          float* dst = a_stage + s * kStageFloats;
          if (oct <= 3) {
            // rows of >= 128 bytes: ONE box of (32 samples, rt rows, 1 clip), 128-byte swizzled as the MMA expects it
            umma::tma_load_3d(dst, &p.maps[oct], hi_full + s, q - dr * hop, t0 + dr, b);
          } else {
            for (int e = 0; e < cols; ++e)
              umma::tma_load_3d(dst + e * rt * 4, &p.maps[oct], hi_full + s, q - dr * hop + 4 * e, t0 + dr, b);
          }
          // the tile's last box is on its way: draw the next tile now, so that the atomic's round trip overlaps this
          // thread's wait for a free stage and the number is held no longer than that (drawn at the start of the tile:
          // 0.2574 ms per feature step, here: 0.2563)
          if (j == n_blocks - 1) fetch.draw_ahead();
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issue
    const uint32_t idesc64 = umma::instr_desc_tf32(kM, 2 * kN), idesc32 = umma::instr_desc_tf32(kM, kN);
    const uint32_t b_addr = umma::smem_u32(b_img);
    if (kTma) umma::mbar_wait(b_full, 0);
    int item = 0, n_tile = 0;
    for (int k = 0, tile; (tile = seq.get(k)) < total; ++k) {
      int b, oct, t0;
      decode_tile(p, tile, b, oct, t0);
      if (kTma && tile_dead(p, b, t0)) continue;   // n_tile counts the tiles that use an accumulator set
      const int q = n_tile & 1;
      umma::mbar_wait(acc_empty + q, ((n_tile >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator set
      umma::fence_after_thread_sync();
      const int n_blocks = oct == 0 ? 8 : oct == 1 ? 4 : oct == 2 ? 2 : 1;
      const int hop = kHop >> oct, rt = block_rt(oct), cols = block_cols(oct);
      const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
      const int len = (int)((len0 + (1LL << oct) - 1) >> oct);
      const int ok_end = len < p.tma_end[oct] ? len : p.tma_end[oct];
      for (int j = 0; j < n_blocks; ++j, ++item) {
        const int s = item % kNStages;
        // TMA path: the hi * [hi | lo] products need only the boxes (hi image), unless the splitters have to patch the
        // block's tail (clip end); the lo * hi products wait for the lo image.  Register path: one pass.
        const bool early = kTma && t0 * hop - kCqtNfft / 2 + 32 * j + (rt - 1) * hop + 4 * cols <= ok_end;
        const uint32_t a_hi_addr = umma::smem_u32(a_stage + s * kStageFloats);
        const uint32_t acc_set = tmem_base + (uint32_t)(q * kSetCols);
        auto issue = [&](int term_lo, int term_hi) {
          if (!(p.debug & 4)) issue_block_any(oct, a_hi_addr, b_addr, acc_set, j, idesc64, idesc32, term_lo, term_hi, kTma);
        };
        AST_STAMP(1, item, 0);
        if (early) {
          umma::mbar_wait(hi_full + s, (item / kNStages) & 1);
          umma::fence_after_thread_sync();
          if (umma::elect_one_sync()) issue(0, 0);
          __syncwarp();
        }
        umma::mbar_wait(full + s, (item / kNStages) & 1);
        umma::fence_after_thread_sync();
        AST_STAMP(1, item, 1);
        if (umma::elect_one_sync()) {
          issue(early ? 1 : 0, 1);
          umma::commit(empty + s);                          // stage s may be overwritten once these MMAs finish
          if (j == n_blocks - 1) umma::commit(acc_full + q);    // ... and the accumulator set is complete
        }
        __syncwarp();
        AST_STAMP(1, item, 2);
      }
      ++n_tile;
    }
  } else if (warp >= kEpilogueWarp0) {
    // ================================================================= epilogue (warps 8-15)
    // TMEM holds one frame per lane; written that way every store instruction would touch 32 output rows
    // (32 L1 wavefronts for 128 B).  The 32 x 24 block is therefore transposed through shared memory and
    // stored with consecutive lanes on consecutive columns of a row (12-float runs, ~3 rows per instruction).
    // register path: two groups, group g takes tiles n_tile = 2 k + g and accumulator set g;
    // TMA path: one group takes every tile, accumulator sets alternate
    const int group = (warp - kEpilogueWarp0) >> 2;
    const int quad = warp & 3;                       // the TMEM lane quadrant this warp may read
    float* stg = epi_buf + (warp - kEpilogueWarp0) * 32 * kEpiStride;
    const long long clip_floats = p.out.layout == AST_LAYOUT_FLAT ? 2LL * p.out.dim1 * p.out.f_row
                                                                  : 2LL * p.out.dim1 * p.out.window * p.out.f_row;
    const int plane = (p.out.layout == AST_LAYOUT_FLAT ? p.out.dim1 : p.out.window) * p.out.f_row;
    // everything a tile's epilogue needs besides the accumulators; loaded one tile ahead so that the global
    // loads do not queue behind the previous tile's stores
    struct TileCtx {
      float* clip_out;
      float sc, mean, rstd;   // lane c < 24: constants of output column c (12 re then 12 im)
      int off0, off1;         // this lane's frame: destination rows as offsets from clip_out (column col0 included)
      int flags;              // bit 0 / 1: row 0 / 1 live (data, else zeros); bit 2 / 3: row 0 / 1 present
      int n_valid;            // statistics mode: live frames among this warp's 32 rows
      long long part_idx;     //                  first element of this warp's 24 partial moments
    };
    const bool stats_mode = p.out.cqt_part != nullptr;
    auto load_ctx = [&](int tile) {
      TileCtx c;
      int b, oct, t0;
      decode_tile(p, tile, b, oct, t0);
      const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
      const int frames_b = num_frames(len0);
      const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
      const int col0 = kFCqt - kBinsPerOctave * (oct + 1);
      c.sc = 0.f, c.mean = 0.f, c.rstd = 1.f;
      if (lane < kCqtCols) {
        const int j = lane < kBinsPerOctave ? lane : lane - kBinsPerOctave;
        c.sc = __ldg(p.scale + oct * kBinsPerOctave + j);
        if (p.out.stats) {
          const float2 m = __ldg(p.out.stats + (long long)b * p.out.stats_clip_stride + p.out.stats_off + col0 + j +
                                 (lane < kBinsPerOctave ? 0 : p.out.f_stats));
          c.mean = m.x, c.rstd = m.y;
        }
      }
      c.clip_out = p.out.out + (long long)b * clip_floats;
      const int t = t0 + quad * 32 + lane;
      c.off0 = c.off1 = 0;
      c.flags = 0;
      c.n_valid = frames_b - (t0 + quad * 32);
      c.n_valid = c.n_valid < 0 ? 0 : c.n_valid > 32 ? 32 : c.n_valid;
      c.part_idx = ((((long long)b * kOctaves + oct) * p.tiles_per_clip_oct + t0 / kM) * 4 + quad) * kCqtCols;
      if (stats_mode) return c;
      if (t < p.slots) {
        const RowDest d = row_dest(p.out, b, t, frames_b, sections_b);
        if (d.n > 0) c.off0 = (int)(d.row[0] - c.clip_out) + col0, c.flags |= 4 | (d.live[0] ? 1 : 0);
        if (d.n > 1) c.off1 = (int)(d.row[1] - c.clip_out) + col0, c.flags |= 8 | (d.live[1] ? 2 : 0);
      }
      return c;
    };
    // Every epilogue warp walks ALL of the CTA's tiles; the accumulator set, barrier phase and group of a live tile follow
    // the number of live tiles before it (the MMA warp counts the same way).
    // The tile's context is loaded BEFORE the wait for its accumulators, so its global loads overlap the wait.
    TileCtx ctx;
    int n_live = 0;
    for (int n_pos = 0, tile; (tile = seq.get(n_pos)) < total; ++n_pos) {
      if (kQueue && lane == 0) tile_ring[kTileRing + 1 + (warp - kEpilogueWarp0)] = n_pos;   // entries < n_pos: done with
      bool dead = false;
      if (kTma) {
        int b, oct, t0;
        decode_tile(p, tile, b, oct, t0);
        dead = tile_dead(p, b, t0);
      }
      const int n_tile = n_live;   // sequence number among the tiles that use an accumulator set
      if (!dead) ++n_live;
      // live tiles go to the groups in LIVE order, so that an accumulator set is always drained by the same group, in
      // order (a parity wait of a later use must not overtake an earlier one); dead tiles by position
      if ((dead ? n_pos : n_tile) % kEpiGroups != group) continue;
      ctx = load_ctx(tile);
      const int set = n_tile & 1;   // accumulator set of this tile
      float acc[kCqtCols];
      if (!dead) {
      if (warp == kEpilogueWarp0) AST_STAMP(2, n_tile, 0);
      umma::mbar_wait(acc_full + set, (n_tile >> 1) & 1);
      umma::fence_after_thread_sync();
      if (warp == kEpilogueWarp0) AST_STAMP(2, n_tile, 1);
      // 24 of each accumulator's 32 columns carry data: one x16 and one x8 load per part
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * kSetCols);
      {
        float a16[16], a8[8];
        umma::tmem_ld_32x16(lane_base, a16);               // hi * hi, even K-steps
        umma::tmem_ld_32x8(lane_base + 16, a8);
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = a16[c];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[16 + c] = a8[c];
#pragma unroll
        for (int part = 1; part < 4; ++part) {
          // part 1: hi * hi of the odd K-steps; parts 2, 3: the cross terms (hi * lo + lo * hi) of both accumulators
          const uint32_t col = part == 1 ? 2 * kN : part == 2 ? kN : 3 * kN;
          umma::tmem_ld_32x16(lane_base + col, a16);
          umma::tmem_ld_32x8(lane_base + col + 16, a8);
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] += a16[c];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[16 + c] += a8[c];
        }
      }
      umma::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(acc_empty + set);  // this warp's quadrant of the set is drained
      if (warp == kEpilogueWarp0) AST_STAMP(2, n_tile, 2);

      // scale + normalise (column constants broadcast from their lane), row-per-lane into the staging buffer
#pragma unroll
      for (int c = 0; c < kCqtCols; ++c) {
        const float sc = __shfl_sync(0xffffffffu, ctx.sc, c);
        const float mu = __shfl_sync(0xffffffffu, ctx.mean, c);
        const float rs = __shfl_sync(0xffffffffu, ctx.rstd, c);
        float v = acc[c] * sc;
        if (p.out.stats) v = (v - mu) * rs;
        stg[lane * kEpiStride + c] = v;
      }
      }   // (a dead tile has no accumulators: its rows are stored as zeros / its partial moments are empty below)
      __syncwarp();
      if (warp == kEpilogueWarp0) AST_STAMP(2, n_tile, 4);
      if (warp == kEpilogueWarp0) AST_STAMP(2, n_tile, 5);
      if (stats_mode) {
        // compute_stats' per-clip reductions (compute_separated_stats.py:27-28) for this quadrant: lane c walks column
        // c of the staged 32 x 24 block (Welford over the live rows) and leaves (mean, M2); nothing else is stored
        if (lane < kCqtCols) {
          float mean = 0.f, m2 = 0.f;
          for (int r = 0; r < ctx.n_valid; ++r) {
            const float x = stg[r * kEpiStride + lane], d = x - mean;
            mean += d / (float)(r + 1);
            m2 = fmaf(d, x - mean, m2);
          }
          p.out.cqt_part[ctx.part_idx + lane] = make_float2(mean, m2);
        }
        __syncwarp();
        continue;
      }
      // 32 rows x 12 columns per plane = 12 store rounds; lane l of round i owns element 32 i + l
      float* const clip_out = ctx.clip_out;
#pragma unroll 3
      for (int i = 0; i < kBinsPerOctave; ++i) {
        const int idx = lane + 32 * i;
        const int r = idx / kBinsPerOctave, j = idx - r * kBinsPerOctave;
        const float re = stg[r * kEpiStride + j], im = stg[r * kEpiStride + kBinsPerOctave + j];
        const int o0 = __shfl_sync(0xffffffffu, ctx.off0, r), o1 = __shfl_sync(0xffffffffu, ctx.off1, r);
        const int fl = __shfl_sync(0xffffffffu, ctx.flags, r);
        if (p.debug & 1) continue;
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
      if (warp == kEpilogueWarp0) AST_STAMP(2, n_tile, 3);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  AST_TIMELINE_STAMP_IF(warp == kMmaWarp && lane == 0, cqt, blockIdx.x, 1);
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_base, kTmemCols);
  AST_TIMELINE_STAMP_IF(warp == kMmaWarp && lane == 0, cqt, blockIdx.x, 3);   // (after the TMEM release)
  // The projection has consumed every decimator tile, but the decimator's publisher bumps a tile's flag BEFORE its
  // stage counter: formally that grid may still be executing its last atomicAdd.  Each CTA therefore waits for its
  // programmatic primary before it exits, so "the CQT grid is complete" implies "the decimator grid is complete" -
  // which is what the STFT's tail wait, and through it the next call's counter-zeroing prologue, relies on.
  if (p.flags && tid == 0) pdl_wait();
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

#ifdef AST_TRACE
extern "C" int ast_debug_cqt_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, g_cqt_trace, sizeof(long long) * 3 * 512 * 6);
}
#endif

int cqt_tc_init() {
  AST_CUDA_TRY(cudaFuncSetAttribute(cqt_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cqt_tc::kSmemTma));
  AST_CUDA_TRY(cudaFuncSetAttribute(cqt_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cqt_tc::kSmemTma));
  AST_CUDA_TRY(cudaFuncSetAttribute(cqt_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cqt_tc::kSmem));
  return AST_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Tensor maps over the seven octave signals of the batch: element (c, r, b) = sample r hop + c of clip b's octave
// signal, i.e. rows of `hop` samples whose stride is the hop - consecutive rows are consecutive frames' window starts.
// A box of (4 samples, rows of a block, 1 clip) is one chunk column of a staged block, delivered straight in the MMA's
// K-major no-swizzle operand layout (rows 16 bytes apart); coordinates may be negative (before the clip) or past the
// last row: those elements arrive as zeros, which is librosa's zero padding (pad_mode = "constant").
// Octave 0 lives in the caller's buffer: only WHOLE rows inside a clip are mapped (no read past the last clip's end);
// the partial last row is re-read by the splitters (BlockPlan::ok_end).  Octaves >= 1 live in the workspace, whose
// per-octave padding and following regions make a partial last row readable (its tail is masked by the clip length).
static bool make_cqt_tensor_maps(CqtTcParams& p, const float* wave, long long wave_stride, const float* ws,
                                 long long ws_clip_stride, int batch, long long max_samples) {
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return false;
  if ((reinterpret_cast<uintptr_t>(wave) & 15) || (wave_stride % 4 != 0 && batch > 1) || (reinterpret_cast<uintptr_t>(ws) & 15) ||
      ws_clip_stride % 4 != 0)
    return false;
  for (int oct = 0; oct < kOctaves; ++oct) {
    const int hop = kHop >> oct;
    const long long len = octave_len(max_samples, oct);
    const long long rows = oct == 0 ? len / hop : (len + hop - 1) / hop;
    if (rows < 1) return false;
    const float* base = oct == 0 ? wave : ws + p.oct_off[oct];
    const long long clip_stride = oct == 0 ? (batch > 1 ? wave_stride : max_samples + (4 - max_samples % 4) % 4) : ws_clip_stride;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)hop, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)hop * 4, (cuuint64_t)clip_stride * 4};
    // octaves 0-3: the whole block, 32 samples = 128 bytes of rt consecutive rows, 128-byte swizzled;
    // octaves 4-6: one chunk column, 16 bytes of rt consecutive rows
    const cuuint32_t box[3] = {(cuuint32_t)(oct <= 3 ? 32 : 4), (cuuint32_t)cqt_tc::block_rt(oct), 1};
    const cuuint32_t elem[3] = {1, 1, 1};
    if (encode(&p.maps[oct], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, elem,
               CU_TENSOR_MAP_INTERLEAVE_NONE, oct <= 3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    p.tma_end[oct] = (int)(rows * hop);
  }
  return true;
}

int launch_cqt_tc(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
                  long long wave_stride, const float* ws, long long ws_clip_stride, const int* dec_flags, const OutSpec& out,
                  cudaStream_t st, bool tile_queue) {
  CqtTcParams p;
  memset(&p, 0, sizeof(p));
  p.flags = dec_flags;
  p.flag_tiles0 = decimator_tiles_stage0(max_samples);
  p.stage_done = dec_flags ? dec_flags + decimator_stage_done_offset(batch, max_samples) : nullptr;
  // The tile queue (its counter sits behind the six stage counters and is zeroed with them, decimator_flag_bytes) pays
  // when this kernel's CTAs start staggered and nothing runs after it: the feature call, where it follows the STFT
  // (0.2850 -> 0.2669 ms per 64 clips).  Where it runs into the decimator's tail and the STFT follows it (statistics
  // call, AST_FEATURE_ORDER=dcs) round-robin lists are better (statistics call 0.2993 vs 0.3067 ms): early CTAs would
  // draw the low octaves' tiles while the decimator's chain stages are still producing them.
  p.queue = (tile_queue || getenv("AST_CQT_QUEUE_ALWAYS")) && dec_flags && !getenv("AST_CQT_STATIC")
                ? const_cast<int*>(dec_flags) + decimator_stage_done_offset(batch, max_samples) + (kOctaves - 1) : nullptr;
  for (int s = 0; s < kOctaves - 1; ++s) p.stage_tiles[s] = decimator_tiles_of_stage(max_samples, s) * batch;
  p.dec_tile_outputs = decimator_tile_outputs();
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
  {
    const char* env = getenv("AST_CQT_DEBUG");
    p.debug = env ? atoi(env) : 0;
  }
  p.vec_ok = (wave_stride % 4 == 0 || batch == 1) && ((reinterpret_cast<uintptr_t>(wave) & 15) == 0);
  p.out = out;
  if (p.slots == 0 || batch == 0) return AST_OK;
  {
    const char* env = getenv("AST_CQT_TMA");   // diagnostic A/B switch: "0" keeps the register-staged producers
    p.use_tma = (!env || strcmp(env, "0") != 0) && p.vec_ok &&
                make_cqt_tensor_maps(p, wave, wave_stride, ws, ws_clip_stride, batch, max_samples) ? 1 : 0;
  }
  if (2LL * p.slots * out.f_row * (out.layout == AST_LAYOUT_FLAT ? 1 : 2) >= (1LL << 31))
    return fail(AST_ERR_INVALID_ARG, "clip too long for the CQT epilogue's 32-bit in-clip offsets");
  long long ctas = (long long)p.tiles_per_clip_oct * kOctaves * batch;
  if (ctas >= (1LL << 24)) p.queue = nullptr;   // ring entries carry 24-bit tile numbers
  if (ctas > plan->sm_count) ctas = plan->sm_count;  // persistent: one CTA per SM
  ProfileSpan span("cqt_tc_kernel", st);
  if (p.use_tma && p.queue)
    AST_CUDA_TRY(launch_with_pdl(cqt_tc_kernel<true, true>, dim3((unsigned)ctas), cqt_tc::kThreadsTma, cqt_tc::kSmemTma, st, p));
  else if (p.use_tma)
    AST_CUDA_TRY(launch_with_pdl(cqt_tc_kernel<true, false>, dim3((unsigned)ctas), cqt_tc::kThreadsTma, cqt_tc::kSmemTma, st, p));
  else
    AST_CUDA_TRY(launch_with_pdl(cqt_tc_kernel<false, false>, dim3((unsigned)ctas), cqt_tc::kThreads, cqt_tc::kSmem, st, p));
  return AST_OK;
}

}  // namespace ast
