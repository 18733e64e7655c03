// plan.cu - plan lifetime, error reporting and host-side derivation of every constant table.
//
// The constant-Q tables follow librosa >= 0.10 as called by utilityFunctions.py:52
// (librosa.cqt(y, sr=22050, n_bins=84, hop_length=256), everything else default):
// filters.wavelet / wavelet_lengths / _relative_bandwidth, __vqt_filter_fft, util.sparsify_rows.
// The sparsified 12 x 129 FFT basis is folded with the 256-point real DFT into one real
// 256 x 24 time-domain matrix (identical for every octave up to sqrt(2^i)), which is what the
// projection kernel contracts against.  The 2:1 decimator is the soxr-HQ-like Kaiser design
// frozen in DESIGN.md (the checker re-derives both independently: tests/test_cabi_host.py::test_plan_constants_match_oracle).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace ast {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

void host_hann(double* w, int n) {
  for (int i = 0; i < n; ++i) w[i] = 0.5 - 0.5 * std::cos(2.0 * M_PI * i / n);  // torch.hann_window(periodic)
}

static double bessel_i0(double x) {
  double term = 1.0, total = 1.0;
  const double y = x * x / 4.0;
  for (int k = 1; k < 64; ++k) {
    term *= y / ((double)k * k);
    total += term;
  }
  return total;
}

// soxr "HQ" recipe restated: 20-bit precision, pass-band end 1 - 0.05 / TO_3dB(rej) of the new
// Nyquist, stop-band at the new Nyquist, (bits + 1) * 6.02 dB rejection, Kaiser-windowed sinc in
// the form of lsx_make_lpf (rho = 0.5), odd length, unit DC gain.
void host_decimator_taps(double* taps) {
  const double bits = 20.0, db2 = 20.0 * std::log10(2.0);
  const double rej = bits * db2;
  const double to_3db = (1.6e-6 * rej - 7.5e-4) * rej + 0.646;
  const double passband_end = 1.0 - 0.05 / to_3db;
  const double att = (bits + 1.0) * db2;
  const double fp = passband_end / 2.0, fs = 0.5;
  const double tr_bw = 0.5 * (fs - fp);
  const double fc = fs - tr_bw;
  const double beta = 0.1102 * (att - 8.7);
  int n = (int)std::ceil((att - 7.95) / (2.285 * M_PI * (fs - fp)) + 1.0);
  if (n % 2 == 0) ++n;
  // the tap count is a compile-time constant of the kernels; the design must reproduce it
  if (n != kDecTaps) {
    set_error("decimator design produced %d taps, kernels expect %d", n, kDecTaps);
    n = kDecTaps;
  }
  const int m = n - 1;
  const double rho = 0.5, i0b = bessel_i0(beta);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    const double z = i - 0.5 * m;
    double h = (z == 0.0) ? fc : std::sin(fc * M_PI * z) / (M_PI * z);
    const double y = z / (0.5 * m + rho);
    h *= bessel_i0(beta * std::sqrt(1.0 - y * y)) / i0b;
    taps[i] = h;
    sum += h;
  }
  for (int i = 0; i < n; ++i) taps[i] /= sum;
}

static void cqt_frequencies(double* freqs) {
  const double fmin = 32.70319566257483;  // librosa.note_to_hz("C1")
  for (int k = 0; k < kFCqt; ++k) freqs[k] = fmin * std::pow(2.0, (double)k / kBinsPerOctave);
}

static void relative_bandwidth(const double* freqs, double* alpha) {
  std::vector<double> logf(kFCqt), bpo(kFCqt);
  for (int k = 0; k < kFCqt; ++k) logf[k] = std::log2(freqs[k]);
  bpo[0] = 1.0 / (logf[1] - logf[0]);
  bpo[kFCqt - 1] = 1.0 / (logf[kFCqt - 1] - logf[kFCqt - 2]);
  for (int k = 1; k < kFCqt - 1; ++k) bpo[k] = 2.0 / (logf[k + 1] - logf[k - 1]);
  for (int k = 0; k < kFCqt; ++k) {
    const double p = std::pow(2.0, 2.0 / bpo[k]);
    alpha[k] = (p - 1.0) / (p + 1.0);
  }
}

void host_cqt_lengths(double* lengths) {
  double freqs[kFCqt], alpha[kFCqt];
  cqt_frequencies(freqs);
  relative_bandwidth(freqs, alpha);
  for (int k = 0; k < kFCqt; ++k) lengths[k] = (1.0 / alpha[k]) * AST_SAMPLE_RATE / freqs[k];
}

// top-octave (bins 72..83 at sr = 22050) wavelets -> sparsified FFT basis -> time-domain kernel
void host_cqt_kernel(double* k_re, double* k_im) {
  typedef std::complex<double> cd;
  double freqs[kFCqt], alpha[kFCqt];
  cqt_frequencies(freqs);
  relative_bandwidth(freqs, alpha);
  const int nfft = kCqtNfft, nb = nfft / 2 + 1;
  const double sr = AST_SAMPLE_RATE;
  std::vector<cd> fft_basis((size_t)kBinsPerOctave * nb);
  for (int j = 0; j < kBinsPerOctave; ++j) {
    const int k = kFCqt - kBinsPerOctave + j;
    const double ilen = (1.0 / alpha[k]) * sr / freqs[k];
    const int n0 = (int)std::floor(-ilen / 2.0), n1 = (int)std::floor(ilen / 2.0);  // arange(-ilen // 2, ilen // 2)
    const int len = n1 - n0;
    std::vector<cd> sig(len);
    double l1 = 0.0;
    for (int i = 0; i < len; ++i) {
      const double ph = (double)(n0 + i) * 2.0 * M_PI * freqs[k] / sr;
      const double w = 0.5 - 0.5 * std::cos(2.0 * M_PI * i / len);
      sig[i] = cd(std::cos(ph), std::sin(ph)) * w;
      l1 += std::abs(sig[i]);
    }
    std::vector<cd> padded(nfft, cd(0, 0));
    const int lpad = (nfft - len) / 2;
    for (int i = 0; i < len; ++i) padded[lpad + i] = sig[i] / l1 * (ilen / nfft);
    for (int f = 0; f < nb; ++f) {
      cd acc(0, 0);
      for (int n = 0; n < nfft; ++n) {
        const double a = -2.0 * M_PI * (double)((f * n) % nfft) / nfft;
        acc += padded[n] * cd(std::cos(a), std::sin(a));
      }
      fft_basis[(size_t)j * nb + f] = acc;
    }
    // util.sparsify_rows(quantile = 0.01)
    std::vector<double> mags(nb), sorted(nb);
    double norm = 0.0;
    for (int f = 0; f < nb; ++f) {
      mags[f] = std::abs(fft_basis[(size_t)j * nb + f]);
      norm += mags[f];
    }
    sorted = mags;
    std::sort(sorted.begin(), sorted.end());
    double cum = 0.0;
    int thr_idx = 0;
    for (int f = 0; f < nb; ++f) {
      cum += sorted[f] / norm;
      if (!(cum < 0.01)) {
        thr_idx = f;
        break;
      }
    }
    for (int f = 0; f < nb; ++f)
      if (!(mags[f] >= sorted[thr_idx])) fft_basis[(size_t)j * nb + f] = cd(0, 0);
  }
  // K[j][n] = sum_f basis[j][f] exp(-2 pi i f n / nfft)
  for (int j = 0; j < kBinsPerOctave; ++j)
    for (int n = 0; n < nfft; ++n) {
      cd acc(0, 0);
      for (int f = 0; f < nb; ++f) {
        const double a = -2.0 * M_PI * (double)((f * n) % nfft) / nfft;
        acc += fft_basis[(size_t)j * nb + f] * cd(std::cos(a), std::sin(a));
      }
      k_re[j * nfft + n] = acc.real();
      k_im[j * nfft + n] = acc.imag();
    }
}

long long octave_len(long long n_samples, int octave) { return (n_samples + (1LL << octave) - 1) >> octave; }

static long long pad_len(long long n) { return (n + 8 + 7) & ~7LL; }  // slack + multiple of 8 floats (32 B)

long long octave_offset(long long max_samples, int octave) {
  long long off = 0;
  for (int i = 1; i < octave; ++i) off += pad_len(octave_len(max_samples, i));
  return off;
}

long long cqt_ws_clip_stride(long long max_samples) { return octave_offset(max_samples, kOctaves); }

}  // namespace ast

using namespace ast;

extern "C" {

const char* ast_last_error(void) { return g_error; }
const char* ast_version(void) { return "audio-style-transfer_b200 0.1.0 (sm_100a)"; }

int ast_default_config(ast_config* cfg) {
  if (!cfg) return fail(AST_ERR_INVALID_ARG, "cfg is null");
  cfg->sample_rate = AST_SAMPLE_RATE;
  cfg->n_fft = AST_N_FFT;
  cfg->hop = AST_HOP;
  cfg->n_bins = AST_F_CQT;
  cfg->window_size = 287;
  cfg->overlap_frames = 96;
  cfg->device = 0;
  return AST_OK;
}

int32_t ast_num_frames(int64_t n_samples) { return n_samples < 0 ? 0 : num_frames(n_samples); }
int32_t ast_num_sections(int32_t n_frames, int32_t window_size, int32_t overlap_frames) {
  return num_sections(n_frames, window_size, overlap_frames);
}
int64_t ast_istft_length(int32_t n_frames) { return n_frames > 0 ? (int64_t)kHop * (n_frames - 1) : 0; }

int ast_host_decimator_taps(double* taps, int32_t capacity, int32_t* n_taps) {
  if (!taps || capacity < kDecTaps) return fail(AST_ERR_INVALID_ARG, "need room for %d taps", kDecTaps);
  g_error[0] = 0;
  host_decimator_taps(taps);
  if (n_taps) *n_taps = kDecTaps;
  return g_error[0] ? AST_ERR_INVALID_ARG : AST_OK;
}

int ast_host_cqt_kernel(double* out) {
  if (!out) return fail(AST_ERR_INVALID_ARG, "out is null");
  std::vector<double> re(kBinsPerOctave * kCqtNfft), im(kBinsPerOctave * kCqtNfft);
  host_cqt_kernel(re.data(), im.data());
  for (int i = 0; i < kBinsPerOctave * kCqtNfft; ++i) {
    out[2 * i] = re[i];
    out[2 * i + 1] = im[i];
  }
  return AST_OK;
}

int ast_host_cqt_lengths(double* lengths) {
  if (!lengths) return fail(AST_ERR_INVALID_ARG, "lengths is null");
  host_cqt_lengths(lengths);
  return AST_OK;
}

int ast_plan_create(const ast_config* cfg, ast_plan** out) {
  if (!cfg || !out) return fail(AST_ERR_INVALID_ARG, "cfg / plan is null");
  *out = nullptr;
  if (cfg->sample_rate != AST_SAMPLE_RATE || cfg->n_fft != AST_N_FFT || cfg->hop != AST_HOP || cfg->n_bins != AST_F_CQT)
    return fail(AST_ERR_INVALID_ARG,
                "unsupported geometry sr=%d n_fft=%d hop=%d n_bins=%d (the reference only ever uses 22050/1024/256/84)",
                cfg->sample_rate, cfg->n_fft, cfg->hop, cfg->n_bins);
  if (cfg->window_size < 2 || cfg->overlap_frames < 0 || 2 * cfg->overlap_frames > cfg->window_size)
    return fail(AST_ERR_INVALID_ARG, "need window >= 2 and 0 <= 2 * overlap <= window (got %d / %d)", cfg->window_size,
                cfg->overlap_frames);
  int n_dev = 0;
  AST_CUDA_TRY(cudaGetDeviceCount(&n_dev));
  if (cfg->device < 0 || cfg->device >= n_dev) return fail(AST_ERR_INVALID_ARG, "no CUDA device %d", cfg->device);
  AST_CUDA_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  AST_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10)
    return fail(AST_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major,
                prop.minor);

  ast_plan* p = new ast_plan();
  std::memset(p, 0, sizeof(*p));
  p->cfg = *cfg;
  p->sm_count = prop.multiProcessorCount;

  std::vector<double> hann(kNfft);
  host_hann(hann.data(), kNfft);
  std::vector<float2> tw1(kTw1Size), tw2(kTw2Size);
  fill_twiddle_tables(tw1.data(), tw2.data());
  std::vector<float> w(kNfft), w_inv(kNfft), w_sq(kNfft);
  for (int m = 0; m < kNfft; ++m) {
    w[m] = (float)hann[m];
    w_inv[m] = (float)(hann[m] / kNfft);
    w_sq[m] = (float)(hann[m] * hann[m]);
  }
  std::vector<double> k_re(kBinsPerOctave * kCqtNfft), k_im(kBinsPerOctave * kCqtNfft), lengths(kFCqt);
  host_cqt_kernel(k_re.data(), k_im.data());
  host_cqt_lengths(lengths.data());
  std::vector<float> kmat((size_t)kCqtNfft * kCqtCols);
  for (int n = 0; n < kCqtNfft; ++n)
    for (int j = 0; j < kBinsPerOctave; ++j) {
      kmat[(size_t)n * kCqtCols + j] = (float)k_re[j * kCqtNfft + n];
      kmat[(size_t)n * kCqtCols + kBinsPerOctave + j] = (float)k_im[j * kCqtNfft + n];
    }
  std::vector<float> scale(kOctaves * kBinsPerOctave);
  for (int i = 0; i < kOctaves; ++i)
    for (int j = 0; j < kBinsPerOctave; ++j) {
      const int k = kFCqt - kBinsPerOctave * (i + 1) + j;
      scale[i * kBinsPerOctave + j] = (float)(std::sqrt(std::pow(2.0, i)) / std::sqrt(lengths[k]));
    }
  std::vector<double> taps(kDecTaps);
  g_error[0] = 0;
  host_decimator_taps(taps.data());
  if (g_error[0]) {
    delete p;
    return AST_ERR_INVALID_ARG;
  }
  std::vector<float> taps_f(kDecTaps);
  for (int i = 0; i < kDecTaps; ++i) taps_f[i] = (float)(taps[i] * std::sqrt(2.0));  // resample(scale=True): / sqrt(0.5)

#define AST_ALLOC_COPY(dst, src, bytes)                                        \
  do {                                                                         \
    cudaError_t e = cudaMalloc((void**)&(dst), (bytes));                       \
    if (e == cudaSuccess) e = cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice); \
    if (e != cudaSuccess) {                                                    \
      ast_plan_destroy(p);                                                     \
      return fail(AST_ERR_CUDA, "plan upload failed: %s", cudaGetErrorString(e)); \
    }                                                                          \
  } while (0)
  AST_ALLOC_COPY(p->d_tw1, tw1.data(), sizeof(float2) * kTw1Size);
  AST_ALLOC_COPY(p->d_tw2, tw2.data(), sizeof(float2) * kTw2Size);
  {
    std::vector<float2> tw32(kTw32Size);
    fill_tw32(tw32.data());
    AST_ALLOC_COPY(p->d_tw32, tw32.data(), sizeof(float2) * kTw32Size);
  }
  AST_ALLOC_COPY(p->d_hann, w.data(), sizeof(float) * kNfft);
  AST_ALLOC_COPY(p->d_hann_inv_n, w_inv.data(), sizeof(float) * kNfft);
  AST_ALLOC_COPY(p->d_hann_sq, w_sq.data(), sizeof(float) * kNfft);
  AST_ALLOC_COPY(p->d_cqt_kernel, kmat.data(), sizeof(float) * kmat.size());
  AST_ALLOC_COPY(p->d_cqt_scale, scale.data(), sizeof(float) * scale.size());
  {
    std::vector<double> kd((size_t)kCqtNfft * kCqtCols);
    for (int n = 0; n < kCqtNfft; ++n)
      for (int j = 0; j < kBinsPerOctave; ++j) {
        kd[(size_t)n * kCqtCols + j] = k_re[j * kCqtNfft + n];
        kd[(size_t)n * kCqtCols + kBinsPerOctave + j] = k_im[j * kCqtNfft + n];
      }
    std::vector<float> images(cqt_tc_image_floats());
    host_cqt_tc_images(kd.data(), images.data());
    AST_ALLOC_COPY(p->d_cqt_tc_images, images.data(), sizeof(float) * images.size());
  }
  {
    std::vector<double> taps_scaled(kDecTaps);
    for (int i = 0; i < kDecTaps; ++i) taps_scaled[i] = taps[i] * std::sqrt(2.0);
    std::vector<float> strip_hi(decimator_strip_floats()), strip_lo(decimator_strip_floats());
    host_decimator_strip(taps_scaled.data(), strip_hi.data(), strip_lo.data());
    AST_ALLOC_COPY(p->d_dec_strip_hi, strip_hi.data(), sizeof(float) * strip_hi.size());
    AST_ALLOC_COPY(p->d_dec_strip_lo, strip_lo.data(), sizeof(float) * strip_lo.size());
    std::vector<uint16_t> h_hi(decimator_strip_h_bytes() / 2), h_lo(decimator_strip_h_bytes() / 2);
    host_decimator_strip_h(taps_scaled.data(), h_hi.data(), h_lo.data());
    AST_ALLOC_COPY(p->d_dec_strip_h_hi, h_hi.data(), h_hi.size() * 2);
    AST_ALLOC_COPY(p->d_dec_strip_h_lo, h_lo.data(), h_lo.size() * 2);
  }
#undef AST_ALLOC_COPY
  int rc = upload_decimator_taps(taps_f.data());
  if (rc == AST_OK) rc = stft_init();
  if (rc == AST_OK) rc = istft_init();
  if (rc == AST_OK) rc = decimate_init();
  if (rc == AST_OK) rc = cqt_tc_init();
  if (const char* env = std::getenv("AST_OVERLAP")) set_overlap_streams(std::strcmp(env, "0") != 0);
  if (const char* env = std::getenv("AST_FEATURE_ORDER")) set_stft_second(std::strcmp(env, "dcs") != 0);
  if (const char* env = std::getenv("AST_CQT"))  // diagnostic A/B switch: "fma" selects the FMA-pipe projection
    set_tc_cqt(std::strcmp(env, "fma") != 0);
  if (const char* env = std::getenv("AST_DECIMATOR")) {  // diagnostic A/B switch: "fma" selects the FMA-pipe kernel, "tf32" the TF32-split tensor kernel
    set_tc_decimator(std::strcmp(env, "fma") != 0);
    set_decimator_half(std::strcmp(env, "tf32") != 0);
  }
  if (rc != AST_OK) {
    ast_plan_destroy(p);
    return rc;
  }
  *out = p;
  return AST_OK;
}

int ast_plan_destroy(ast_plan* p) {
  if (!p) return AST_OK;
  cudaFree(p->d_tw1);
  cudaFree(p->d_tw2);
  cudaFree(p->d_tw32);
  cudaFree(p->d_hann);
  cudaFree(p->d_hann_inv_n);
  cudaFree(p->d_hann_sq);
  cudaFree(p->d_cqt_kernel);
  cudaFree(p->d_cqt_scale);
  cudaFree(p->d_cqt_tc_images);
  cudaFree(p->d_dec_strip_hi);
  cudaFree(p->d_dec_strip_lo);
  cudaFree(p->d_dec_strip_h_hi);
  cudaFree(p->d_dec_strip_h_lo);
  delete p;
  return AST_OK;
}

}  // extern "C"
