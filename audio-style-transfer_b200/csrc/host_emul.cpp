// host_emul.cpp - compiles the kernels' host/device arithmetic (fft_core.h) with g++ and runs it
// thread-by-thread on the CPU, so index / twiddle algebra is checked without a GPU.
// Test-only: built by tests/test_fft_core_host.py, never linked into libast_frontend.so.
#include <cmath>
#include <vector>

#include "fft_core.h"

using namespace ast;

namespace {
std::vector<float2> make_twiddles() {
  std::vector<float2> tw(kFftN);
  for (int m = 0; m < kFftN; ++m) {
    const double a = -2.0 * M_PI * m / kFftN;
    tw[m] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  return tw;
}
struct EmitPair {
  float* a;
  float* b;
  void operator()(int k, float are, float aim, float bre, float bim) {
    a[2 * k] = are, a[2 * k + 1] = aim, b[2 * k] = bre, b[2 * k + 1] = bim;
  }
};
struct EmitComplex {
  float* out;
  int* hits;
  void operator()(int n, float2 z) {
    out[2 * n] = z.x, out[2 * n + 1] = z.y;
    hits[n]++;
  }
};
void run_stages12(const float2* z, std::vector<float2>& buf2) {
  static const std::vector<float2> tw = make_twiddles();
  std::vector<float2> buf1(kBuf1Size);
  for (int tid = 0; tid < kFftThreads; ++tid) {
    float2 v[16];
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = z[64 * n1 + tid];
    fft1024_stage1(v, tid, tw.data(), buf1.data());
  }
  for (int tid = 0; tid < kFftThreads; ++tid) fft1024_stage2(tid, tw.data(), buf1.data(), buf2.data());
}
}  // namespace

extern "C" {
// complex 1024-point forward FFT; returns the number of output indices not written exactly once
int emul_fft1024(const float* in, float* out) {
  std::vector<float2> z(kFftN), buf2(kBuf2Size);
  for (int n = 0; n < kFftN; ++n) z[n] = make_float2(in[2 * n], in[2 * n + 1]);
  run_stages12(z.data(), buf2);
  std::vector<int> hits(kFftN, 0);
  EmitComplex emit{out, hits.data()};
  for (int j = 0; j < kFftThreads; ++j) fft1024_stage3_complex(j, buf2.data(), emit);
  int bad = 0;
  for (int n = 0; n < kFftN; ++n) bad += hits[n] != 1;
  return bad;
}
// two real 1024-sample frames -> two 513-bin half spectra (interleaved re/im)
void emul_rfft_pair(const float* fa, const float* fb, float* xa, float* xb) {
  std::vector<float2> z(kFftN), buf2(kBuf2Size);
  for (int n = 0; n < kFftN; ++n) z[n] = make_float2(fa[n], fb[n]);
  run_stages12(z.data(), buf2);
  for (int k = 0; k <= 512; ++k) xa[2 * k] = xa[2 * k + 1] = xb[2 * k] = xb[2 * k + 1] = NAN;
  EmitPair emit{xa, xb};
  for (int j = 0; j < kFftThreads; ++j) fft1024_stage3_real_pair(j, buf2.data(), emit);
}
// two 513-bin half spectra -> two real 1024-sample frames (irfft with 1/N scaling)
void emul_irfft_pair(const float* xa, const float* xb, float* fa, float* fb) {
  std::vector<float2> z(kFftN), buf2(kBuf2Size);
  for (int m = 0; m < kFftN; ++m) {
    const int kk = m <= 512 ? m : kFftN - m;
    z[m] = pack_conj_hermitian_pair(m, make_float2(xa[2 * kk], xa[2 * kk + 1]), make_float2(xb[2 * kk], xb[2 * kk + 1]));
  }
  run_stages12(z.data(), buf2);
  std::vector<float> out(2 * kFftN);
  std::vector<int> hits(kFftN, 0);
  EmitComplex emit{out.data(), hits.data()};
  for (int j = 0; j < kFftThreads; ++j) fft1024_stage3_complex(j, buf2.data(), emit);
  for (int n = 0; n < kFftN; ++n) {
    fa[n] = out[2 * n] / kFftN;
    fb[n] = -out[2 * n + 1] / kFftN;
  }
}
}
