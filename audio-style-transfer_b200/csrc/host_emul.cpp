// host_emul.cpp - compiles the kernels' host/device arithmetic (fft_core.h) with g++ and runs it
// thread-by-thread on the CPU, so index / twiddle algebra is checked without a GPU.
// Test-only: built by tests/test_fft_core_host.py, never linked into libast_frontend.so.
#include <cmath>
#include <vector>

#include "fft_core.h"

using namespace ast;

namespace {
struct Tables {
  std::vector<float2> t1, t2;
  Tables() : t1(kTw1Size), t2(kTw2Size) { fill_twiddle_tables(t1.data(), t2.data()); }
};
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
struct EmitCols {
  float* out;
  int* hits;
  template <int S>
  void col(int q, const float2 (&z)[4]) {
    for (int a = 0; a < 4; ++a) {
      out[2 * (q + 256 * a)] = z[a].x, out[2 * (q + 256 * a) + 1] = z[a].y;
      hits[q + 256 * a] += 1 + 0 * S;
    }
  }
};
void run_stages12(const float2* z, std::vector<float2>& buf2) {
  static const Tables tb;
  std::vector<float2> buf1(kBuf1Size);
  for (int tid = 0; tid < kFftThreads; ++tid) {
    float2 v[16];
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = z[64 * n1 + tid];
    fft1024_stage1(v, tid, tb.t1.data(), buf1.data());
  }
  for (int tid = 0; tid < kFftThreads; ++tid) fft1024_stage2(tid, tb.t2.data(), buf1.data(), buf2.data());
}
}  // namespace

namespace {
template <int N1>
void pack_all(const float* xa, const float* xb, float2* z) {
  for (int tid = 0; tid < kFftThreads; ++tid) {
    const int kk = ihalf_bin<N1>(tid);
    z[64 * N1 + tid] = ihalf_pack<N1>(tid, make_float2(xa[2 * kk], xa[2 * kk + 1]), make_float2(xb[2 * kk], xb[2 * kk + 1]));
  }
  if constexpr (N1 < 15) pack_all<N1 + 1>(xa, xb, z);
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
  for (int j = 0; j < kFftThreads; ++j) fft1024_stage3_real_pair<true>(j, buf2.data(), emit);
}
// same complex FFT through the column-wise stage 3 used by the iSTFT kernel
int emul_fft1024_columns(const float* in, float* out) {
  std::vector<float2> z(kFftN), buf2(kBuf2Size);
  for (int n = 0; n < kFftN; ++n) z[n] = make_float2(in[2 * n], in[2 * n + 1]);
  run_stages12(z.data(), buf2);
  std::vector<int> hits(kFftN, 0);
  EmitCols emit{out, hits.data()};
  for (int j = 0; j < kFftThreads; ++j) fft1024_stage3_columns(j, buf2.data(), emit);
  int bad = 0;
  for (int n = 0; n < kFftN; ++n) bad += hits[n] != 1;
  return bad;
}
// compile-time-resolved packing (what the iSTFT kernel uses) must equal the generic one; returns mismatches
int emul_pack_check(const float* xa, const float* xb) {
  std::vector<float2> z1(kFftN), z2(kFftN);
  for (int m = 0; m < kFftN; ++m) {
    const int kk = m <= 512 ? m : kFftN - m;
    z1[m] = pack_conj_hermitian_pair(m, make_float2(xa[2 * kk], xa[2 * kk + 1]), make_float2(xb[2 * kk], xb[2 * kk + 1]));
  }
  pack_all<0>(xa, xb, z2.data());
  int bad = 0;
  for (int m = 0; m < kFftN; ++m) bad += !(z1[m].x == z2[m].x && z1[m].y == z2[m].y);
  return bad;
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
// the warp-per-transform 32 x 32 formulation (fft1024w_pass1 / pass2), lane by lane
void emul_fft1024_warp(const float* in, float* out) {
  std::vector<float2> tw(kTw32Size), tile(kTileSize);
  fill_tw32(tw.data());
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[32];
    for (int n1 = 0; n1 < 32; ++n1) v[n1] = make_float2(in[2 * (32 * n1 + lane)], in[2 * (32 * n1 + lane) + 1]);
    fft1024w_pass1(v, lane, tw.data(), tile.data());
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[32];
    fft1024w_pass2(v, lane, tile.data());
    for (int k2 = 0; k2 < 32; ++k2) out[2 * (lane + 32 * k2)] = v[k2].x, out[2 * (lane + 32 * k2) + 1] = v[k2].y;
  }
}
void emul_fft32(const float* in, float* out) {
  float2 v[32];
  for (int n = 0; n < 32; ++n) v[n] = make_float2(in[2 * n], in[2 * n + 1]);
  fft32(v);
  for (int n = 0; n < 32; ++n) out[2 * n] = v[n].x, out[2 * n + 1] = v[n].y;
}
}
