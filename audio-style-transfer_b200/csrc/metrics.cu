// metrics.cu - the spectrogram-domain evaluation metric on the far side of the iSTFT:
// mse_spectrogram (evaluation_reconstruction.py:105-118, evaluation_style_transfer.py:111-119):
//
//     spec = |librosa.stft(audio, n_fft=1024, hop_length=256)|   (center=True, ZERO padding, periodic Hann)
//     mse  = mean((spec_orig[:, :T] - spec_gen[:, :T]) ** 2),  T = min(T_orig, T_gen)
//
// The two STFTs reuse K1 (stft.cu) with pad_zero = 1; the magnitudes are never stored: a reduction kernel reads the
// real / imaginary planes of both, accumulates the squared differences in double and writes per-CTA partials that
// a single thread folds in a fixed order (deterministic, no atomics).
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

}  // extern "C"
