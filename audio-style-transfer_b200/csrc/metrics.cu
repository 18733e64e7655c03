// metrics.cu - the spectrogram-domain evaluation metric on the far side of the iSTFT:
// mse_spectrogram (evaluation_reconstruction.py:105-118, evaluation_style_transfer.py:111-119):
//
//     spec = |librosa.stft(audio, n_fft=1024, hop_length=256)|   (center=True, ZERO padding, periodic Hann)
//     mse  = mean((spec_orig[:, :T] - spec_gen[:, :T]) ** 2),  T = min(T_orig, T_gen)
//
// The two STFTs reuse K1 (stft.cu) with pad_zero = 1; the magnitudes are never stored: a reduction kernel reads the
// real / imaginary planes of both, accumulates the squared differences in double and writes per-CTA partials that
// a single thread folds in a fixed order (deterministic, no atomics).
// instrumentation_similarity (evaluation_style_transfer.py:111-119) follows further down.
#include "common.cuh"

namespace ast {

constexpr int kMseThreads = 256;
constexpr int kMseCtas = 592;  // 4 per SM on a 148-SM part; the partial buffer is sized for this

__global__ void __launch_bounds__(kMseThreads) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                  long long plane_a, long long plane_b, long long n,
                                                                  double* __restrict__ partial) {
  __shared__ double red[kMseThreads / 32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * kMseThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kMseThreads) {
    const float ma = hypotf(__ldg(a + i), __ldg(a + plane_a + i));   // np.abs of a complex64
    const float mb = hypotf(__ldg(b + i), __ldg(b + plane_b + i));
    const float d = ma - mb;
    acc += (double)(d * d);                                           // the reference squares in float32
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kMseThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void mse_final_kernel(const double* __restrict__ partial, int n_partial, double count, double* __restrict__ out) {
  double s = 0.0;
  for (int i = 0; i < n_partial; ++i) s += partial[i];
  out[0] = count > 0 ? s / count : 0.0;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- instrumentation_similarity (evaluation_style_transfer.py:111-119) ----------------------------------------
//
//     S      = |librosa.stft(audio)|          (defaults: n_fft = 2048, hop = 512, center, ZERO padding, periodic Hann)
//     energy = S.sum(axis=1)                  (1025 values, one per frequency bin)
//     corr   = pearsonr(energy1, energy2)[0]  (0.0 when it is NaN)
//
// The 2048-point real transform is a 1024-point complex transform of the even / odd samples (five radix-4 Stockham
// passes in shared memory, 256 threads) followed by the even/odd split; the magnitudes are summed over the frames a
// CTA owns in double and written as one partial row per CTA; one CTA folds the rows in a fixed order and computes the
// correlation in double (deterministic, no atomics).
constexpr int kSimFft = 2048;
constexpr int kSimHop = 512;
constexpr int kSimBins = kSimFft / 2 + 1;   // 1025
constexpr int kSimHalf = kSimFft / 2;       // length of the complex transform
constexpr int kSimThreads = 256;
constexpr int kSimMaxCtas = 296;            // 2 per SM; rows of the partial buffer per signal
constexpr int kSimRow = 1032;               // doubles per partial row (1025 rounded up)


__device__ __forceinline__ void sim_radix4_pass(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw,
                                                int j, int ns) {
  const int k = j & (ns - 1);
  const int m = k * (512 / ns);               // exp(-2 pi i r k / (4 ns)) = tw[r m]
  float2 v0 = in[j], v1 = in[j + 256], v2 = in[j + 512], v3 = in[j + 768];
  v1 = cmul(v1, tw[m]);
  v2 = cmul(v2, tw[2 * m]);
  v3 = cmul(v3, tw[3 * m]);
  const float2 a = make_float2(v0.x + v2.x, v0.y + v2.y), b = make_float2(v0.x - v2.x, v0.y - v2.y);
  const float2 c = make_float2(v1.x + v3.x, v1.y + v3.y), d = make_float2(v1.x - v3.x, v1.y - v3.y);
  const int j0 = ((j - k) << 2) + k;
  out[j0] = make_float2(a.x + c.x, a.y + c.y);
  out[j0 + ns] = make_float2(b.x + d.y, b.y - d.x);       // b - i d
  out[j0 + 2 * ns] = make_float2(a.x - c.x, a.y - c.y);
  out[j0 + 3 * ns] = make_float2(b.x - d.y, b.y + d.x);   // b + i d
}

__global__ void __launch_bounds__(kSimThreads) stft2048_energy_kernel(const float* __restrict__ a, long long n_a, int t_a,
                                                                      const float* __restrict__ b, long long n_b, int t_b,
                                                                      double* __restrict__ partial) {
  __shared__ float2 tw[kSimFft];            // exp(-2 pi i m / 2048)
  __shared__ float2 buf0[kSimHalf], buf1[kSimHalf];
  const int j = threadIdx.x;
  const float* x = blockIdx.y ? b : a;
  const long long n = blockIdx.y ? n_b : n_a;
  const int frames = blockIdx.y ? t_b : t_a;
  for (int m = j; m < kSimFft; m += kSimThreads) {
    float sn, cs;
    sincospif((float)m * (1.0f / 1024.0f), &sn, &cs);
    tw[m] = make_float2(cs, -sn);
  }
  __syncthreads();
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, acc_nyq = 0.0;
  for (int t = blockIdx.x; t < frames; t += gridDim.x) {
    const long long s0 = (long long)t * kSimHop - kSimHalf;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = j + 256 * r;
      const long long s = s0 + 2 * c;
      const float xe = (s >= 0 && s < n) ? __ldg(x + s) : 0.0f;
      const float xo = (s + 1 >= 0 && s + 1 < n) ? __ldg(x + s + 1) : 0.0f;
      buf0[c] = make_float2(xe * (0.5f - 0.5f * tw[2 * c].x), xo * (0.5f - 0.5f * tw[2 * c + 1].x));
    }
    __syncthreads();
    sim_radix4_pass(buf0, buf1, tw, j, 1);
    __syncthreads();
    sim_radix4_pass(buf1, buf0, tw, j, 4);
    __syncthreads();
    sim_radix4_pass(buf0, buf1, tw, j, 16);
    __syncthreads();
    sim_radix4_pass(buf1, buf0, tw, j, 64);
    __syncthreads();
    sim_radix4_pass(buf0, buf1, tw, j, 256);
    __syncthreads();
    // X[k] = E[k] + W^k O[k],  E = (Z[k] + conj Z[N-k]) / 2,  O = (Z[k] - conj Z[N-k]) / 2i
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = j + 256 * r;
      const float2 z = buf1[k], zc = buf1[(kSimHalf - k) & (kSimHalf - 1)];
      const float2 e = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
      const float2 o = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
      const float2 wo = cmul(tw[k], o);
      acc[r] += (double)hypotf(e.x + wo.x, e.y + wo.y);
    }
    if (j == 0) acc_nyq += (double)fabsf(buf1[0].x - buf1[0].y);
    __syncthreads();
  }
  double* row = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kSimRow;
#pragma unroll
  for (int r = 0; r < 4; ++r) row[j + 256 * r] = acc[r];
  if (j == 0) row[kSimHalf] = acc_nyq;
}

__device__ double sim_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < 32; ++w) s += red[w];   // same order on every thread
  return s;
}

// One CTA of 1024 threads: thread k owns bin k (thread 0 also bin 1024).
__global__ void __launch_bounds__(1024) sim_pearson_kernel(const double* __restrict__ partial, int rows, double* __restrict__ out) {
  __shared__ double red[32];
  const int k = threadIdx.x;
  double ea = 0.0, eb = 0.0, ea_n = 0.0, eb_n = 0.0;
  for (int r = 0; r < rows; ++r) {
    ea += partial[(size_t)r * kSimRow + k];
    eb += partial[(size_t)(rows + r) * kSimRow + k];
    if (k == 0) {
      ea_n += partial[(size_t)r * kSimRow + kSimHalf];
      eb_n += partial[(size_t)(rows + r) * kSimRow + kSimHalf];
    }
  }
  // energy = S.sum(axis=1) is a float32 array in the reference
  ea = (double)(float)ea, eb = (double)(float)eb, ea_n = (double)(float)ea_n, eb_n = (double)(float)eb_n;
  const double mean_a = sim_block_sum(ea + ea_n, red) / kSimBins;
  const double mean_b = sim_block_sum(eb + eb_n, red) / kSimBins;
  const double da = ea - mean_a, db = eb - mean_b;
  const double da_n = k == 0 ? ea_n - mean_a : 0.0, db_n = k == 0 ? eb_n - mean_b : 0.0;
  const double sab = sim_block_sum(da * db + da_n * db_n, red);
  const double saa = sim_block_sum(da * da + da_n * da_n, red);
  const double sbb = sim_block_sum(db * db + db_n * db_n, red);
  if (k == 0) {
    double r = sab / (sqrt(saa) * sqrt(sbb));
    if (r != r) r = 0.0;                      // `corr if not np.isnan(corr) else 0.0`
    out[0] = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
  }
}

static int sim_frames(int64_t n) { return 1 + (int)(n / kSimHop); }
static int sim_ctas(int64_t n_a, int64_t n_b) {
  const int t = sim_frames(n_a > n_b ? n_a : n_b);
  return t < kSimMaxCtas ? t : kSimMaxCtas;
}

}  // namespace ast

using namespace ast;

extern "C" {

size_t ast_mse_workspace_bytes(const ast_plan* plan, int64_t n_a, int64_t n_b) {
  (void)plan;
  if (n_a < 0 || n_b < 0) return 0;
  const size_t ta = (size_t)num_frames(n_a), tb = (size_t)num_frames(n_b);
  return align256(sizeof(float) * 2 * ta * kFStft) + align256(sizeof(float) * 2 * tb * kFStft) + align256(sizeof(double) * kMseCtas);
}

int ast_mse_spectrogram(const ast_plan* plan, const float* a, int64_t n_a, const float* b, int64_t n_b, void* workspace,
                        size_t workspace_bytes, double* result, void* stream) {
  if (!plan || !a || !b || !result) return fail(AST_ERR_INVALID_ARG, "ast_mse_spectrogram: null pointer");
  if (n_a < 1 || n_b < 1) return fail(AST_ERR_INVALID_ARG, "ast_mse_spectrogram: empty signal");
  if (n_a >= (1LL << 30) || n_b >= (1LL << 30)) return fail(AST_ERR_INVALID_ARG, "signals longer than 2^30 samples are not supported");
  const size_t need = ast_mse_workspace_bytes(plan, n_a, n_b);
  if (!workspace || workspace_bytes < need)
    return fail(AST_ERR_WORKSPACE, "workspace of %zu bytes is too small, need %zu (ast_mse_workspace_bytes)", workspace_bytes, need);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(AST_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int ta = num_frames(n_a), tb = num_frames(n_b);
  char* w = static_cast<char*>(workspace);
  float* sa = reinterpret_cast<float*>(w);
  w += align256(sizeof(float) * 2 * (size_t)ta * kFStft);
  float* sb = reinterpret_cast<float*>(w);
  w += align256(sizeof(float) * 2 * (size_t)tb * kFStft);
  double* partial = reinterpret_cast<double*>(w);
  OutSpec oa{};
  oa.out = sa, oa.layout = AST_LAYOUT_FLAT, oa.dim1 = ta, oa.f_row = kFStft, oa.f_off = 0;
  oa.window = plan->cfg.window_size, oa.step = plan->cfg.window_size - plan->cfg.overlap_frames;
  oa.stats = nullptr, oa.stats_clip_stride = 0, oa.f_stats = 0, oa.stats_off = 0;
  OutSpec ob = oa;
  ob.out = sb, ob.dim1 = tb;
  int rc = launch_stft(plan, a, nullptr, 1, n_a, n_a, oa, st, /*pad_zero=*/1);
  if (rc != AST_OK) return rc;
  rc = launch_stft(plan, b, nullptr, 1, n_b, n_b, ob, st, /*pad_zero=*/1);
  if (rc != AST_OK) return rc;
  const int t = ta < tb ? ta : tb;
  const long long n = (long long)t * kFStft;  // rows are contiguous: the first t rows of each plane
  long long ctas = (n + kMseThreads - 1) / kMseThreads;
  if (ctas > kMseCtas) ctas = kMseCtas;
  {
    ProfileSpan span("mse_partial_kernel", st);
    mse_partial_kernel<<<(unsigned)ctas, kMseThreads, 0, st>>>(sa, sb, (long long)ta * kFStft, (long long)tb * kFStft, n, partial);
    AST_LAUNCH_CHECK("mse_partial_kernel");
  }
  mse_final_kernel<<<1, 1, 0, st>>>(partial, (int)ctas, (double)n, result);
  AST_LAUNCH_CHECK("mse_final_kernel");
  return AST_OK;
}

size_t ast_instrumentation_similarity_workspace_bytes(const ast_plan* plan, int64_t n_a, int64_t n_b) {
  (void)plan;
  if (n_a < 1 || n_b < 1) return 0;
  return align256(sizeof(double) * 2 * (size_t)sim_ctas(n_a, n_b) * kSimRow);
}

int ast_instrumentation_similarity(const ast_plan* plan, const float* a, int64_t n_a, const float* b, int64_t n_b, void* workspace,
                                   size_t workspace_bytes, double* result, void* stream) {
  if (!plan || !a || !b || !result) return fail(AST_ERR_INVALID_ARG, "ast_instrumentation_similarity: null pointer");
  if (n_a < 1 || n_b < 1) return fail(AST_ERR_INVALID_ARG, "ast_instrumentation_similarity: empty signal");
  if (n_a >= (1LL << 30) || n_b >= (1LL << 30)) return fail(AST_ERR_INVALID_ARG, "signals longer than 2^30 samples are not supported");
  const size_t need = ast_instrumentation_similarity_workspace_bytes(plan, n_a, n_b);
  if (!workspace || workspace_bytes < need)
    return fail(AST_ERR_WORKSPACE, "workspace of %zu bytes is too small, need %zu (ast_instrumentation_similarity_workspace_bytes)",
                workspace_bytes, need);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(AST_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = static_cast<double*>(workspace);
  const int ctas = sim_ctas(n_a, n_b);
  {
    ProfileSpan span("stft2048_energy_kernel", st);
    stft2048_energy_kernel<<<dim3((unsigned)ctas, 2), kSimThreads, 0, st>>>(a, (long long)n_a, sim_frames(n_a), b, (long long)n_b,
                                                                          sim_frames(n_b), partial);
    AST_LAUNCH_CHECK("stft2048_energy_kernel");
  }
  sim_pearson_kernel<<<1, 1024, 0, st>>>(partial, ctas, result);
  AST_LAUNCH_CHECK("sim_pearson_kernel");
  return AST_OK;
}

}  // extern "C"
