// umma.cuh - minimal inline-PTX layer over the Blackwell (sm_100a) tensor-core path: TMEM allocation,
// shared-memory matrix descriptors, tcgen05.mma (kind::tf32), tcgen05.commit -> mbarrier, tcgen05.ld.
// Bit layouts follow the CUTLASS sm100 headers (cute/arch/mma_sm100_desc.hpp): SmemDescriptor and
// InstrDescriptor.  Only what the front-end kernels need: cta_group::1, K-major operands, no swizzle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ast {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE ("interleave") canonical layout in 16-byte units: ((8, n), 2) : ((1, SBO), LBO):
// 8 consecutive rows are 16 B apart, 8-row groups SBO apart, the two 16-byte K chunks of one MMA LBO apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);  // version = 1 (sm_100)
}

// K-major, SWIZZLE_128B canonical layout ((8, n), (4, 2)) : ((128 B, SBO), (4 B, 16 B)): rows of 128 bytes (four K-steps
// of 8 TF32 values) whose 16-byte chunks are XOR-swizzled with the row index mod 8, 8-row groups SBO = 1024 B apart -
// what a TMA box with CU_TENSOR_MAP_SWIZZLE_128B writes into a 1024-byte aligned buffer.  The K-step is selected by
// + 32 B on the start address, a shift by whole rows by + 128 B per row: the swizzle is a function of the address bits,
// so the descriptor's base offset stays 0 (measured).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::tf32, FP32 accumulate, both operands K-major, no negate / sparsity
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// round-toward-zero to TF32 (10 explicit mantissa bits), what the tensor core keeps of an FP32 operand
__host__ __device__ __forceinline__ uint32_t tf32_trunc_bits(uint32_t bits) { return bits & 0xFFFFE000u; }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(n_cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(n_cols) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads smem through it)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMA: the calling thread arrives on the barrier and announces `bytes` of asynchronous writes that will complete it
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA tiled load of one 3-D box (coordinates innermost first; out-of-range elements arrive as zeros) into shared
// memory; the barrier receives complete_tx for the whole box
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tensor_map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tensor_map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const void* tensor_map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tensor_map) : "memory");
}

// wall-clock nanoseconds (for the bounds of the cross-CTA polling loops: a count of polls would also trip under a
// sanitizer or a debugger, where everything runs orders of magnitude slower)
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kPollTimeoutNs = 20ull * 1000 * 1000 * 1000;  // 20 s: nothing here legitimately waits that long

// Bounded wait: a broken pipeline traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 24)) __trap();
  }
}

// 16-byte asynchronous global -> shared copy (LDGSTS) and its completion wait
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// fire-and-forget L2 prefetch of a contiguous global range (16-byte aligned address, size a multiple of 16)
__device__ __forceinline__ void prefetch_l2_bulk(const void* gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

// one lane of a fully converged warp
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all prior MMAs of this thread arrive (once) on the mbarrier when they have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread t of the warp receives row (lane base + t), columns c .. c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: issue several, then tmem_wait_ld() once (one TMEM round trip instead of one per load)
__device__ __forceinline__ void tmem_ld_32x16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 8 columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// split 4 floats into their TF32 truncation (hi) and the exact residual (lo)
__device__ __forceinline__ void split_tf32(const float4& v, float4& h, float4& l) {
  h.x = __uint_as_float(tf32_trunc_bits(__float_as_uint(v.x)));
  h.y = __uint_as_float(tf32_trunc_bits(__float_as_uint(v.y)));
  h.z = __uint_as_float(tf32_trunc_bits(__float_as_uint(v.z)));
  h.w = __uint_as_float(tf32_trunc_bits(__float_as_uint(v.w)));
  l.x = v.x - h.x, l.y = v.y - h.y, l.z = v.z - h.z, l.w = v.w - h.w;
}

// x[s .. s+3] with zeros outside [0, len).  A chunk is almost always entirely inside or entirely outside the
// signal; the straddling case (at most two chunks per staged block) is kept out of line so that unrolled staging
// loops stay small - the inlined scalar path made the producers of the tensor-core kernels instruction-bound.
static __device__ __noinline__ float4 load4_partial(const float* __restrict__ x, int s, int len) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s >= 0 && s < len) v.x = __ldg(x + s);
  if (s + 1 >= 0 && s + 1 < len) v.y = __ldg(x + s + 1);
  if (s + 2 >= 0 && s + 2 < len) v.z = __ldg(x + s + 2);
  if (s + 3 >= 0 && s + 3 < len) v.w = __ldg(x + s + 3);
  return v;
}
__device__ __forceinline__ float4 load4_zero_ext(const float* __restrict__ x, int s, int len, bool vec_ok) {
  if (s >= 0 && s + 3 < len && vec_ok) return __ldg(reinterpret_cast<const float4*>(x + s));
  if (s + 3 < 0 || s >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
  return load4_partial(x, s, len);
}

// the same through L2 only (ld.global.cg): for data another kernel / CTA may still be producing, where neither the
// non-coherent path nor a stale L1 line is acceptable
__device__ __forceinline__ float4 ld_cg_f4(const float* ptr) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cg_f(const float* ptr) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(ptr) : "memory");
  return v;
}
static __device__ __noinline__ float4 load4_partial_cg(const float* x, int s, int len) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s >= 0 && s < len) v.x = ld_cg_f(x + s);
  if (s + 1 >= 0 && s + 1 < len) v.y = ld_cg_f(x + s + 1);
  if (s + 2 >= 0 && s + 2 < len) v.z = ld_cg_f(x + s + 2);
  if (s + 3 >= 0 && s + 3 < len) v.w = ld_cg_f(x + s + 3);
  return v;
}
__device__ __forceinline__ float4 load4_zero_ext_cg(const float* x, int s, int len) {
  if (s >= 0 && s + 3 < len) return ld_cg_f4(x + s);
  if (s + 3 < 0 || s >= len) return make_float4(0.f, 0.f, 0.f, 0.f);
  return load4_partial_cg(x, s, len);
}

}  // namespace umma
}  // namespace ast
