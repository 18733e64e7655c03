// fft_core.h - 1024-point complex FFT split over 64 threads x 16 points (16 x 16 x 4), written
// as host/device functions so the exact arithmetic the kernels run can also be compiled with
// g++ and checked on a CPU box (tests/test_fft_core_host.py drives csrc/host_emul.cpp).
//
// One 1024-point complex FFT carries TWO real frames (frame A in the real lane, frame B in the
// imaginary lane); the Hermitian separation happens in registers in stage 3, with no twiddle,
// so the imaginary parts of bins 0 and 512 come out as exact zeros (torch.stft does the same,
// and dataloader.normalize divides those two columns by 0 + 1e-8, dataloader.py:13).
//
// Index algebra (N = 1024 = 16 * 16 * 4):
//   n = 64 n1 + 4 n2 + n3          k = k1 + 16 k2 + 256 k3
//   stage 1  thread tid = 4 n2 + n3 : 16-point DFT over n1, twiddle W_256^(n2 k1)  -> buf1[k1][tid]
//   stage 2  thread tid = k1 + 16 n3: 16-point DFT over n2, twiddle W_1024^(n3 (k1 + 16 k2)) -> buf2[n3][k1 + 16 k2]
//   stage 3  thread j: 4-point DFTs over n3 for q in {j, 256 - j, 128 - j, 128 + j}  (thread 0: {0, 128, 64, 192})
//            so that bins k and N - k always live in the same thread.
// Shared-memory strides: buf1 rows are 65 float2 apart (conflict-free 64-bit reads in stage 2).
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define AST_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define AST_HD inline
struct float2 {
  float x, y;
};
static inline float2 make_float2(float a, float b) {
  float2 r;
  r.x = a;
  r.y = b;
  return r;
}
#endif

namespace ast {

constexpr int kFftN = 1024;
constexpr int kFftThreads = 64;   // threads cooperating on one 1024-point FFT
constexpr int kBuf1Stride = 65;   // float2 per k1 row
constexpr int kBuf1Size = 16 * kBuf1Stride;
constexpr int kBuf2Size = 1024;

// On the device the complex helpers are the packed FP32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: one issue
// slot per complex add, two per complex multiply; the swaps and sign flips of mul_neg_i and of the multiply fold into
// the instructions' operand modifiers).  Same roundings as the scalar forms: (a.x b.x - round(a.y b.y)) fused once.
#if defined(__CUDA_ARCH__)
AST_HD float2 cmul(float2 a, float2 b) {
  const float2 q = __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x));
  return __ffma2_rn(make_float2(a.x, a.x), b, make_float2(-q.x, q.y));
}
AST_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
AST_HD float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#else
AST_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
AST_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AST_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
// multiply by -i  (forward W_4^1)
AST_HD float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

// forward 4-point DFT, natural order in and out
AST_HD void fft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s02 = cadd(a, c), d02 = csub(a, c);
  const float2 s13 = cadd(b, d), d13 = mul_neg_i(csub(b, d));
  a = cadd(s02, s13);
  b = cadd(d02, d13);
  c = csub(s02, s13);
  d = csub(d02, d13);
}

// forward 16-point DFT in registers, natural order in and out (4 x 4 Cooley-Tukey)
AST_HD void fft16(float2 (&v)[16]) {
  // W_16^m = exp(-2 pi i m / 16)
  const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f, r = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // now v[4 k1 + n2] = A[k1][n2]; twiddle by W_16^(n2 k1)
  v[5] = cmul(v[5], make_float2(c1, -s1));    // k1=1,n2=1 : m=1
  v[6] = cmul(v[6], make_float2(r, -r));      // k1=1,n2=2 : m=2
  v[7] = cmul(v[7], make_float2(s1, -c1));    // k1=1,n2=3 : m=3
  v[9] = cmul(v[9], make_float2(r, -r));      // k1=2,n2=1 : m=2
  v[10] = mul_neg_i(v[10]);                   // k1=2,n2=2 : m=4
  v[11] = cmul(v[11], make_float2(-r, -r));   // k1=2,n2=3 : m=6
  v[13] = cmul(v[13], make_float2(s1, -c1));  // k1=3,n2=1 : m=3
  v[14] = cmul(v[14], make_float2(-r, -r));   // k1=3,n2=2 : m=6
  v[15] = cmul(v[15], make_float2(-c1, s1));  // k1=3,n2=3 : m=9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // now v[4 k1 + k2] = X[k1 + 4 k2]  -> transpose to natural order
  float2 t[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) t[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = t[i];
}

// Twiddle tables (built on the host in double precision):
//   t1[k1 * 16 + n2]  = exp(-2 pi i n2 k1 / 256)                 (stage 1; a warp reads one k1 row -> no bank conflicts)
//   t2[k2 * 64 + tid] = exp(-2 pi i n3 (k1 + 16 k2) / 1024), tid = k1 + 16 n3   (stage 2; lane-contiguous)
constexpr int kTw1Size = 16 * 16;
constexpr int kTw2Size = 16 * 64;

inline void fill_twiddle_tables(float2* t1, float2* t2) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 16; ++n2) {
      const double a = -two_pi * (double)(n2 * k1) / 256.0;
      t1[k1 * 16 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k2 = 0; k2 < 16; ++k2)
    for (int tid = 0; tid < 64; ++tid) {
      const int k1 = tid & 15, n3 = tid >> 4;
      const double a = -two_pi * (double)(n3 * (k1 + 16 * k2)) / 1024.0;
      t2[k2 * 64 + tid] = make_float2((float)cos(a), (float)sin(a));
    }
}

// stage 1: thread tid holds v[n1] = z[64 n1 + tid]
AST_HD void fft1024_stage1(float2 (&v)[16], int tid, const float2* t1, float2* buf1) {
  fft16(v);
  const int n2 = tid >> 2;
  buf1[tid] = v[0];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) buf1[k1 * kBuf1Stride + tid] = cmul(v[k1], t1[k1 * 16 + n2]);
}

// stage 2: thread tid = k1 + 16 n3 gathers over n2, transforms, twiddles, scatters to buf2[n3][q]
AST_HD void fft1024_stage2(int tid, const float2* t2, const float2* buf1, float2* buf2) {
  const int k1 = tid & 15, n3 = tid >> 4;
  float2 v[16];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf1[k1 * kBuf1Stride + 4 * n2 + n3];
  fft16(v);
  // unconditional multiply (t2 = 1 exactly where n3 == 0) keeps the warp convergent
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) buf2[n3 * 256 + k1 + 16 * k2] = cmul(v[k2], t2[k2 * 64 + tid]);
}

// stage 3 helper: the four outputs Z[q + 256 a], a = 0..3
AST_HD void fft1024_stage3_column(int q, const float2* buf2, float2 (&z)[4]) {
  z[0] = buf2[q];
  z[1] = buf2[256 + q];
  z[2] = buf2[512 + q];
  z[3] = buf2[768 + q];
  fft4(z[0], z[1], z[2], z[3]);
}

// Hermitian separation of one conjugate pair: zk = Z[k], zp = Z[(N - k) mod N], k <= 512.
//   frame A bin k = (zk + conj(zp)) / 2          frame B bin k = (zk - conj(zp)) / (2 i)
// With kHalve == false the four values are emitted WITHOUT the factor 1/2 (the caller folds it
// into its own scaling, e.g. the normalisation multiply).
template <bool kHalve, class Emit>
AST_HD void separate_pair(int k, float2 zk, float2 zp, Emit& emit) {
  if (kHalve)
    emit(k, 0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y), 0.5f * (zk.y + zp.y), 0.5f * (zp.x - zk.x));
  else
    emit(k, zk.x + zp.x, zk.y - zp.y, zk.y + zp.y, zp.x - zk.x);
}

// bin emitted by the i-th emit() call of fft1024_stage3_real_pair in thread j (i < 8, thread 0: i < 9); lets a caller
// fetch per-bin constants ahead of the transform's last stage
AST_HD int fft1024_stage3_bin(int j, int i) {
  const int a[9] = {j, 256 + j, 512 - j, 256 - j, 128 - j, 384 - j, 384 + j, 128 + j, 0};
  const int z[9] = {0, 256, 512, 128, 384, 64, 320, 448, 192};
  return j != 0 ? a[i] : z[i];
}

// stage 3 of the forward transform of two real frames: emits all 513 bins of both frames.
// emit(k, a_re, a_im, b_re, b_im).  Thread j emits 8 bins (thread 0: 9).
template <bool kHalve, class Emit>
AST_HD void fft1024_stage3_real_pair(int j, const float2* buf2, Emit& emit) {
  float2 za[4], zb[4];
  if (j != 0) {
    // columns q = j and 256 - j hold bins k = j + 256 a  <->  N - k = (256 - j) + 256 (3 - a)
    fft1024_stage3_column(j, buf2, za);
    fft1024_stage3_column(256 - j, buf2, zb);
    separate_pair<kHalve>(j, za[0], zb[3], emit);
    separate_pair<kHalve>(256 + j, za[1], zb[2], emit);
    separate_pair<kHalve>(512 - j, zb[1], za[2], emit);
    separate_pair<kHalve>(256 - j, zb[0], za[3], emit);
    // columns q = 128 - j and 128 + j: k = 128 - j + 256 a  <->  N - k = (128 + j) + 256 (3 - a)
    fft1024_stage3_column(128 - j, buf2, za);
    fft1024_stage3_column(128 + j, buf2, zb);
    separate_pair<kHalve>(128 - j, za[0], zb[3], emit);
    separate_pair<kHalve>(384 - j, za[1], zb[2], emit);
    separate_pair<kHalve>(384 + j, zb[1], za[2], emit);
    separate_pair<kHalve>(128 + j, zb[0], za[3], emit);
  } else {
    fft1024_stage3_column(0, buf2, za);  // Z[0], Z[256], Z[512], Z[768]
    separate_pair<kHalve>(0, za[0], za[0], emit);
    separate_pair<kHalve>(256, za[1], za[3], emit);
    separate_pair<kHalve>(512, za[2], za[2], emit);
    fft1024_stage3_column(128, buf2, za);  // Z[128], Z[384], Z[640], Z[896]
    separate_pair<kHalve>(128, za[0], za[3], emit);
    separate_pair<kHalve>(384, za[1], za[2], emit);
    fft1024_stage3_column(64, buf2, za);   // Z[64], Z[320], Z[576], Z[832]
    fft1024_stage3_column(192, buf2, zb);  // Z[192], Z[448], Z[704], Z[960]
    separate_pair<kHalve>(64, za[0], zb[3], emit);
    separate_pair<kHalve>(320, za[1], zb[2], emit);
    separate_pair<kHalve>(448, zb[1], za[2], emit);
    separate_pair<kHalve>(192, zb[0], za[3], emit);
  }
}

// stage 3 of a plain complex transform: emit(n, Z[n]) for the 16 outputs of thread j
template <class Emit>
AST_HD void fft1024_stage3_complex(int j, const float2* buf2, Emit& emit) {
  int qs[4];
  if (j != 0) {
    qs[0] = j, qs[1] = 256 - j, qs[2] = 128 - j, qs[3] = 128 + j;
  } else {
    qs[0] = 0, qs[1] = 128, qs[2] = 64, qs[3] = 192;
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float2 z[4];
    fft1024_stage3_column(qs[s], buf2, z);
#pragma unroll
    for (int a = 0; a < 4; ++a) emit(qs[s] + 256 * a, z[a]);
  }
}

// stage 3 of a plain complex transform, column-wise: emit.template col<s>(q, z) with z[a] = Z[q + 256 a],
// s = 0..3 the (compile-time) slot of column q in thread j's set, so per-column state can live in registers.
template <class Emit>
AST_HD void fft1024_stage3_columns(int j, const float2* buf2, Emit& emit) {
  float2 z[4];
  const bool first = j == 0;
  const int q0 = first ? 0 : j, q1 = first ? 128 : 256 - j, q2 = first ? 64 : 128 - j, q3 = first ? 192 : 128 + j;
  fft1024_stage3_column(q0, buf2, z);
  emit.template col<0>(q0, z);
  fft1024_stage3_column(q1, buf2, z);
  emit.template col<1>(q1, z);
  fft1024_stage3_column(q2, buf2, z);
  emit.template col<2>(q2, z);
  fft1024_stage3_column(q3, buf2, z);
  emit.template col<3>(q3, z);
}

// Packs two Hermitian half-spectra (513 bins each, imag of bins 0 / 512 ignored as torch.istft
// does) into conj(Z)[m], Z = XA + i XB extended to 1024 bins, so that
//   forward_fft(conj Z)[n] = N * conj(a[n] + i b[n])   ->  a[n] = Re / N,  b[n] = -Im / N.
// xa = XA[kk], xb = XB[kk] with kk = m <= 512 ? m : 1024 - m.
AST_HD float2 pack_conj_hermitian_pair(int m, float2 xa, float2 xb) {
  if (m == 0 || m == 512) return make_float2(xa.x, -xb.x);
  if (m < 512) return make_float2(xa.x - xb.y, -(xa.y + xb.x));
  return make_float2(xa.x + xb.y, -(xb.x - xa.y));
}

// The same packing for stage-1 register n1 of thread tid (m = 64 n1 + tid), with the case analysis
// resolved at compile time: bin index kk = 64 n1 + tid (n1 < 8) or 64 (16 - n1) - tid (n1 >= 8), and
// only (n1 in {0, 8}, tid == 0) are the self-conjugate bins 0 / 512.
template <int N1>
AST_HD int ihalf_bin(int tid) {
  return N1 < 8 ? 64 * N1 + tid : 64 * (16 - N1) - tid;
}
template <int N1>
AST_HD float2 ihalf_pack(int tid, float2 xa, float2 xb) {
  if ((N1 == 0 || N1 == 8) && tid == 0) {
    xa.y = 0.f;
    xb.y = 0.f;
  }
  if (N1 < 8) return make_float2(xa.x - xb.y, -(xa.y + xb.x));
  return make_float2(xa.x + xb.y, -(xb.x - xa.y));
}

}  // namespace ast

// ------------------------------------------------------------------------------------------------------------
// 1024 = 32 x 32: one WARP per transform, 32 points per lane in registers, ONE exchange through a warp-private
// shared-memory tile, no block-level barrier (stft.cu, istft.cu).
//
//   n = 32 n1 + n2 (n1: register, n2: lane)          k = k1 + 32 k2
//   X[k1 + 32 k2] = sum_n2 W_32^(n2 k2) [ W_1024^(n2 k1) sum_n1 x[32 n1 + n2] W_32^(n1 k1) ]
//   pass 1  lane n2: 32-point DFT over n1 -> A[k1], times W_1024^(n2 k1) -> tile[k1][n2]
//   pass 2  lane k1: reads tile[k1][n2] over n2, 32-point DFT over n2 -> register k2 holds X[k1 + 32 k2]
// Bins k and 1024 - k live in lanes k1 and (32 - k1) % 32, registers k2 and 31 - k2 (lane 0: 32 - k2): the Hermitian
// separation of two real frames carried by one complex transform is one warp shuffle per bin.
namespace ast {

constexpr int kTileStride = 33;                  // float2 per tile row: conflict-free transposed reads
constexpr int kTileSize = 32 * kTileStride;      // one warp's exchange tile (8448 B)
constexpr int kTw32Size = 32 * 32;               // tw32[k1 * 32 + n2] = exp(-2 pi i n2 k1 / 1024)

inline void fill_tw32(float2* tw) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int k1 = 0; k1 < 32; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = -two_pi * (double)(n2 * k1) / 1024.0;
      tw[k1 * 32 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
}

// forward 32-point DFT in registers, natural order in and out: one radix-2 decimation-in-frequency step, then two
// 16-point transforms (X[2m] from x[n] + x[n+16], X[2m+1] from (x[n] - x[n+16]) W_32^n)
AST_HD void fft32(float2 (&v)[32]) {
  // cos / sin of pi n / 16, n = 1..7
  const float c1 = 0.98078528040323044913f, s1 = 0.19509032201612826785f;
  const float c2 = 0.92387953251128673848f, s2 = 0.38268343236508978178f;
  const float c3 = 0.83146961230254523708f, s3 = 0.55557023301960222474f;
  const float r = 0.70710678118654752440f;
  float2 a[16], b[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    a[n] = cadd(v[n], v[n + 16]);
    b[n] = csub(v[n], v[n + 16]);
  }
  b[1] = cmul(b[1], make_float2(c1, -s1));
  b[2] = cmul(b[2], make_float2(c2, -s2));
  b[3] = cmul(b[3], make_float2(c3, -s3));
  b[4] = cmul(b[4], make_float2(r, -r));
  b[5] = cmul(b[5], make_float2(s3, -c3));
  b[6] = cmul(b[6], make_float2(s2, -c2));
  b[7] = cmul(b[7], make_float2(s1, -c1));
  b[8] = mul_neg_i(b[8]);
  b[9] = cmul(b[9], make_float2(-s1, -c1));
  b[10] = cmul(b[10], make_float2(-s2, -c2));
  b[11] = cmul(b[11], make_float2(-s3, -c3));
  b[12] = cmul(b[12], make_float2(-r, -r));
  b[13] = cmul(b[13], make_float2(-c3, -s3));
  b[14] = cmul(b[14], make_float2(-c2, -s2));
  b[15] = cmul(b[15], make_float2(-c1, -s1));
  fft16(a);
  fft16(b);
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    v[2 * m] = a[m];
    v[2 * m + 1] = b[m];
  }
}

// pass 1 of lane n2: v[n1] = z[32 n1 + n2] in, tile[k1][n2] out
AST_HD void fft1024w_pass1(float2 (&v)[32], int lane, const float2* tw32, float2* tile) {
  fft32(v);
  tile[lane] = v[0];
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) tile[k1 * kTileStride + lane] = cmul(v[k1], tw32[k1 * 32 + lane]);
}

// pass 2 of lane k1: v[k2] = Z[k1 + 32 k2] out (the caller synchronises the warp between the passes)
AST_HD void fft1024w_pass2(float2 (&v)[32], int lane, const float2* tile) {
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) v[n2] = tile[lane * kTileStride + n2];
  fft32(v);
}

}  // namespace ast
