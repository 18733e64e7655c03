// Microbenchmark 2: tcgen05.mma kind::tf32 M128 K8, fully unrolled issue (compile-time accumulator rotation),
// SS mode with aligned / row-shifted A descriptors, and TS mode (A operand in TMEM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../audio-style-transfer_b200/csrc/umma.cuh"
using namespace ast;

struct Res { long long cyc; float v0; };

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// MODE 0: SS aligned, 1: SS with A start shifted by (i % 7) * 16 B, 2: TS (A in TMEM columns 448..455)
template <int N, int R, int MODE>
__global__ void __launch_bounds__(128, 2) bench(int outer, int tmem_cols, float aval, Res* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* a = reinterpret_cast<float*>(smem_raw);            // 2 chunks x 160 rows x 16 B
  float* b = a + 4096;                                      // 2 chunks x 256 rows x 16 B = 8 KB
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
  constexpr uint32_t idesc = umma::instr_desc_tf32(128, N);
  const uint32_t a_tmem = tb + (uint32_t)(tmem_cols - 8);
  if (MODE == 2) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = aval;
    tmem_st_32x8(a_tmem + ((uint32_t)(warp * 32) << 16), v);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
  }
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (umma::elect_one_sync()) {
      const uint64_t da = umma::smem_desc(umma::smem_u32(a), 160 * 16, 128);
      const uint64_t db = umma::smem_desc(umma::smem_u32(b), 256 * 16, 128);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (MODE == 2) mma_tf32_ts(tb + r * N, a_tmem, db, idesc, 0u);
        else umma::mma_tf32(tb + r * N, da, db, idesc, 0u);
      }
      umma::commit(mbar);
      umma::mbar_wait(mbar, 0);
      t0 = clock64();
      for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int i = 0; i < 56; ++i) {
          if (MODE == 2) mma_tf32_ts(tb + (i % R) * N, a_tmem, db, idesc, 1u);
          else if (MODE == 1) umma::mma_tf32(tb + (i % R) * N, da + (uint64_t)(i % 7), db, idesc, 1u);
          else umma::mma_tf32(tb + (i % R) * N, da, db, idesc, 1u);
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

template <int N, int R, int MODE>
void run(int grid, Res* d) {
  const int outer = 8;
  const int cols = grid > 148 ? 256 : 512;
  if (R * N + 8 > cols) return;
  cudaFuncSetAttribute(bench<N, R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  bench<N, R, MODE><<<grid, 128, 40000>>>(outer, cols, 1.0f, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d R %d MODE %d error %s\n", N, R, MODE, cudaGetErrorString(e)); exit(1); }
  Res h;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("grid %3d N %3d R %2d mode %d : %7.1f cyc/MMA   acc0 = %.1f (expect %.1f)\n", grid, N, R, MODE,
         (double)h.cyc / (outer * 56), h.v0, 8.0 * (1 + outer * ((56 + R - 1) / R)));
}

template <int N, int R>
void run_modes(Res* d) {
  for (int grid : {148, 296}) {
    run<N, R, 0>(grid, d);
    run<N, R, 1>(grid, d);
    run<N, R, 2>(grid, d);
  }
}

int main() {
  Res* d; cudaMalloc(&d, sizeof(Res));
  run_modes<8, 1>(d);  run_modes<8, 4>(d);
  run_modes<16, 1>(d); run_modes<16, 4>(d);
  run_modes<32, 1>(d); run_modes<32, 2>(d); run_modes<32, 4>(d); run_modes<32, 7>(d);
  run_modes<64, 1>(d); run_modes<64, 2>(d); run_modes<64, 3>(d); run_modes<64, 7>(d);
  run_modes<128, 1>(d); run_modes<128, 3>(d);
  run_modes<256, 1>(d);
  return 0;
}
