// Microbenchmark: tcgen05.mma kind::tf32 M128 issue cost vs N and number of rotating accumulators,
// and whether raw FP32 operands are truncated (RZ) by the tensor core.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../audio-style-transfer_b200/csrc/umma.cuh"
using namespace ast;

struct Res { long long cyc; float v0; };

__global__ void __launch_bounds__(128, 2) bench(int N, int R, int iters, int tmem_cols, float aval, Res* out, int order) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* a = reinterpret_cast<float*>(smem_raw);            // 128 rows x 8 (K) : 2 chunks x 128 rows x 16 B = 4 KB (use 8 KB)
  float* b = a + 4096;                                      // up to 256 rows x 8 : 8 KB
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096; i += 128) { a[i] = aval; b[i] = 1.0f; }
  if (warp == 0) umma::tmem_alloc(slot, tmem_cols);
  if (tid == 0) umma::mbar_init(mbar, 1);
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tb = *slot;
  const uint32_t idesc = umma::instr_desc_tf32(128, N);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (umma::elect_one_sync()) {
      const uint64_t da = umma::smem_desc(umma::smem_u32(a), 128 * 16, 128);
      const uint64_t db = umma::smem_desc(umma::smem_u32(b), 256 * 16, 128);
      // init accumulators
      for (int r = 0; r < R; ++r) umma::mma_tf32(tb + r * N, da, db, idesc, 0u);
      umma::commit(mbar);
      umma::mbar_wait(mbar, 0);
      t0 = clock64();
      if (order == 0) {
        for (int i = 0; i < iters; ++i) umma::mma_tf32(tb + (i % R) * N, da, db, idesc, 1u);
      } else {
        // pairs of dependent MMAs back to back (like the decimator's cross terms), rotating over R
        for (int i = 0; i < iters; i += 2) {
          umma::mma_tf32(tb + ((i / 2) % R) * N, da, db, idesc, 1u);
          umma::mma_tf32(tb + ((i / 2) % R) * N, da, db, idesc, 1u);
        }
      }
      umma::commit(mbar);
      umma::mbar_wait(mbar, 1);
      t1 = clock64();
    }
    __syncwarp();
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  float v[16];
  umma::tmem_ld_32x16(tb + ((uint32_t)(warp * 32) << 16), v);
  if (tid == 0 && blockIdx.x == 0) { out->cyc = t1 - t0; out->v0 = v[0]; }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, tmem_cols);
}

int main() {
  Res* d; cudaMalloc(&d, sizeof(Res));
  const int iters = 512;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  printf("truncation test: A = 1 + 2^-11 + 2^-12, B = 1, K = 8, one MMA init + %d accum on R=1\n", iters);
  for (int grid : {1, 148, 296}) {
    for (int N : {32, 64, 128, 256}) {
      for (int R : {1, 2, 4, 8, 16}) {
        const int cols = grid == 296 ? 256 : 512;
        if (R * N > cols) continue;
        for (int order = 0; order < 2; ++order) {
          Res h;
          bench<<<grid, 128, 40000>>>(N, R, iters, cols, 1.0f + 0.00048828125f + 0.000244140625f, d, order);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
          printf("grid %3d N %3d R %2d order %d : %7.1f cyc/MMA   acc0 = %.6f\n", grid, N, R, order, (double)h.cyc / iters, h.v0);
        }
      }
    }
  }
  return 0;
}
