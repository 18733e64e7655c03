// cqt.cu - K3 (FMA-pipe version): the constant-Q projection of librosa.cqt as the reference calls it
// (utilityFunctions.py:52), evaluated in the time domain.
//
// librosa computes, per octave i, resp_i = fft_basis_i @ rfft(frames_i) with rectangular 256-sample
// frames of the i-times-decimated signal at hop 256 / 2^i (zero padded, centred) and a sparsified
// 12 x 129 basis.  fft_basis_i = sqrt(2^i) * fft_basis_0, so with K = fft_basis_0 @ DFT_256 (a dense
// real 256 x 24 matrix, 12 real + 12 imaginary columns, built on the host in plan.cu)
//
//     resp_i[t, :] = sqrt(2^i) * frames_i[t, :] @ K,      V[k, t] = resp / sqrt(length_k)
//
// which is one small dense contraction per frame (the one dense contraction of the path).  This
// version runs it on the FMA pipes: one thread per frame, 24 accumulators, K broadcast from shared
// memory.  The epilogue applies the per-bin scale, (x - mean) * rstd, and scatters to the flat or
// section layout (columns 513..596 of the feature rows).
#include "common.cuh"

namespace ast {

constexpr int kCqtThreads = 128;

struct CqtParams {
  const float* wave;       // octave 0 signal
  long long wave_stride;
  const float* ws;         // octave buffers 1..6
  long long ws_clip_stride;
  long long oct_off[kOctaves];  // offset of octave buffer i inside a clip's workspace (i >= 1)
  const int32_t* lengths;
  long long max_samples;
  int slots;
  int overlap;
  const float* kmat;       // [256][24]
  const float* scale;      // [7][12]
  bool vec_ok;             // octave-0 rows are 16-byte aligned
  OutSpec out;
};

__device__ __forceinline__ float4 load4_zero_ext(const float* __restrict__ x, long long i, long long len, bool vec_ok) {
  // x[i .. i+3] with zeros outside [0, len)  (librosa.stft pad_mode="constant")
  if (i >= 0 && i + 3 < len && vec_ok) return __ldg(reinterpret_cast<const float4*>(x + i));
  float4 v;
  v.x = (i >= 0 && i < len) ? __ldg(x + i) : 0.f;
  v.y = (i + 1 >= 0 && i + 1 < len) ? __ldg(x + i + 1) : 0.f;
  v.z = (i + 2 >= 0 && i + 2 < len) ? __ldg(x + i + 2) : 0.f;
  v.w = (i + 3 >= 0 && i + 3 < len) ? __ldg(x + i + 3) : 0.f;
  return v;
}

__global__ void __launch_bounds__(kCqtThreads) cqt_kernel(const CqtParams p) {
  __shared__ __align__(16) float ks[kCqtNfft * kCqtCols];  // 24 KB
  for (int i = threadIdx.x; i < kCqtNfft * kCqtCols / 4; i += kCqtThreads)
    reinterpret_cast<float4*>(ks)[i] = __ldg(reinterpret_cast<const float4*>(p.kmat) + i);
  __syncthreads();

  const int b = blockIdx.y;
  const int t = blockIdx.x * kCqtThreads + threadIdx.x;
  if (t >= p.slots) return;
  const long long len0 = p.lengths ? p.lengths[b] : p.max_samples;
  const int frames_b = num_frames(len0);
  const int sections_b = p.out.layout == AST_LAYOUT_SECTIONS ? num_sections(frames_b, p.out.window, p.overlap) : 0;
  const RowDest d = row_dest(p.out, b, t, frames_b, sections_b);
  const bool live = t < frames_b;
  const float2* st0 = nullptr;
  if (p.out.stats) st0 = p.out.stats + (long long)b * p.out.stats_clip_stride + p.out.stats_off;

  {
    const int oct = blockIdx.z;  // one octave per CTA: 7x more CTAs to fill the machine
    float acc[kCqtCols];
#pragma unroll
    for (int c = 0; c < kCqtCols; ++c) acc[c] = 0.f;
    if (live) {
      const float* __restrict__ x = oct == 0 ? p.wave + (long long)b * p.wave_stride
                                             : p.ws + (long long)b * p.ws_clip_stride + p.oct_off[oct];
      const long long len = (len0 + (1LL << oct) - 1) >> oct;
      const bool vec_ok = oct == 0 ? p.vec_ok : true;
      const long long start = (long long)t * (kHop >> oct) - kCqtNfft / 2;
      // frames entirely inside the zero padding contribute nothing; skip chunks outside [0, len)
#pragma unroll 2
      for (int n4 = 0; n4 < kCqtNfft / 4; ++n4) {
        const long long i = start + 4 * n4;
        if (i + 3 < 0 || i >= len) continue;
        const float4 xv = load4_zero_ext(x, i, len, vec_ok);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4* krow = reinterpret_cast<const float4*>(ks + (4 * n4 + e) * kCqtCols);
#pragma unroll
          for (int q = 0; q < kCqtCols / 4; ++q) {
            const float4 kv = krow[q];
            acc[4 * q + 0] = fmaf(xs[e], kv.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(xs[e], kv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xs[e], kv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(xs[e], kv.w, acc[4 * q + 3]);
          }
        }
      }
    }
    // bins of octave `oct` are columns 84 - 12 (oct + 1) .. 84 - 12 oct of the CQT block
    const int col0 = kFCqt - kBinsPerOctave * (oct + 1);
#pragma unroll
    for (int j = 0; j < kBinsPerOctave; ++j) {
      const float s = __ldg(p.scale + oct * kBinsPerOctave + j);
      float re = acc[j] * s, im = acc[kBinsPerOctave + j] * s;
      if (st0) {
        const float2 m0 = __ldg(st0 + col0 + j), m1 = __ldg(st0 + p.out.f_stats + col0 + j);
        re = (re - m0.x) * m0.y;
        im = (im - m1.x) * m1.y;
      }
      for (int r = 0; r < d.n; ++r) {
        d.row[r][col0 + j] = d.live[r] ? re : 0.f;
        d.row[r][d.plane + col0 + j] = d.live[r] ? im : 0.f;
      }
    }
  }
}

static int g_use_tc_cqt = 1;
void set_tc_cqt(int on) { g_use_tc_cqt = on; }
bool use_tc_cqt() { return g_use_tc_cqt != 0; }

int launch_cqt(const ast_plan* plan, const float* wave, const int32_t* lengths, int batch, long long max_samples,
               long long wave_stride, const float* ws, long long ws_clip_stride, const int* dec_flags, const OutSpec& out,
               cudaStream_t st, bool tile_queue) {
  if (g_use_tc_cqt || out.cqt_part)   // (the statistics epilogue exists in the tensor-core kernel only)
    return launch_cqt_tc(plan, wave, lengths, batch, max_samples, wave_stride, ws, ws_clip_stride, dec_flags, out, st, tile_queue);
  CqtParams p;
  p.wave = wave;
  p.wave_stride = wave_stride;
  p.ws = ws;
  p.ws_clip_stride = ws_clip_stride;
  for (int i = 0; i < kOctaves; ++i) p.oct_off[i] = i == 0 ? 0 : octave_offset(max_samples, i);
  p.lengths = lengths;
  p.max_samples = max_samples;
  p.slots = frame_slots(out.layout, out.dim1, out.window, out.step);
  p.overlap = out.window - out.step;
  p.kmat = plan->d_cqt_kernel;
  p.scale = plan->d_cqt_scale;
  p.vec_ok = (wave_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(wave) & 15) == 0);
  p.out = out;
  if (p.slots == 0 || batch == 0) return AST_OK;
  dim3 grid((unsigned)((p.slots + kCqtThreads - 1) / kCqtThreads), (unsigned)batch, kOctaves);
  ProfileSpan span("cqt_kernel", st);
  cqt_kernel<<<grid, kCqtThreads, 0, st>>>(p);
  AST_LAUNCH_CHECK("cqt_kernel");
  return AST_OK;
}

}  // namespace ast
