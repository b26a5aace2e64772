// Thin PTX wrappers for the sm_100a tensor-core path (tcgen05.mma / TMEM / mbarrier) used by the
// tensor-core recurrence (lstm_tc.cu).  gemm_tc.cu keeps its own private copies.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TC5_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC5_WAIT_DONE;\n\t"
      "bra TC5_WAIT_LOOP;\n\t"
      "TC5_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy shared-memory writes -> visible to the tensor core's (async proxy) operand reads
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is an M x K bf16 matrix, row m in TMEM lane m, elements
// (2c, 2c+1) packed in 32-bit column c (even k in the low half)
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                             uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16 with bf16 A/B (format 1),
// fp32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc512(uint32_t slot_smem_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot_smem_addr)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_addr) : "memory");
}
// the calling warp writes 4 consecutive 32-bit columns of its 32 TMEM lanes (lane field of taddr
// = 32 * (warp % 4)); thread l supplies lane l's values
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a),
               "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// raw bits: convert with __uint_as_float only AFTER tmem_wait_ld()
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// bulk copy global -> this CTA's shared memory on the TMA engine; completion = transaction bytes
// on `bar`.  dst, src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// orders generic-proxy accesses (st.global by any CTA, made visible by a release/acquire chain)
// against this thread's later async-proxy accesses (bulk copies), and vice versa
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
// the same, restricted to the global state space (does not wait on shared-memory async traffic)
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
// instruction descriptor for kind::f16 with fp16 A/B (format 0), fp32 accumulate, K-major operands
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// x = a + 2^-11 * b with a, b fp16: a = fp16(x), b = fp16(2^11 (x - a)); |x - a - 2^-11 b| <= 2^-24 |x|
// (for |x| < 65504; tiny x: absolute error <= 2^-36)
__device__ __forceinline__ void split2h(float x, uint32_t& a, uint32_t& b) {
  const __half h1 = __float2half_rn(x);
  const float r = x - __half2float(h1);
  const __half h2 = __float2half_rn(r * 2048.f);
  a = (uint32_t)__half_as_ushort(h1);
  b = (uint32_t)__half_as_ushort(h2);
}
__device__ __forceinline__ void split2hx8(const float (&x)[8], uint4& o1, uint4& o2) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2h(x[i], a[i], b[i]);
  o1 = make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), a[4] | (a[5] << 16), a[6] | (a[7] << 16));
  o2 = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4] | (b[5] << 16), b[6] | (b[7] << 16));
}

// x = a + b + c with a, b, c bf16 (24 significant bits: exact to fp32 rounding)
__device__ __forceinline__ void split3(float x, uint32_t& a, uint32_t& b, uint32_t& c) {
  const __nv_bfloat16 ha = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(ha);
  const __nv_bfloat16 hb = __float2bfloat16_rn(r);
  r -= __bfloat162float(hb);
  const __nv_bfloat16 hc = __float2bfloat16_rn(r);
  a = (uint32_t)__bfloat16_as_ushort(ha);
  b = (uint32_t)__bfloat16_as_ushort(hb);
  c = (uint32_t)__bfloat16_as_ushort(hc);
}
// 8 floats -> three 16-byte chunks of 8 bf16 (element i in bits [16*(i%2), +16) of word i/2)
__device__ __forceinline__ void split3x8(const float (&x)[8], uint4& o1, uint4& o2, uint4& o3) {
  uint32_t a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split3(x[i], a[i], b[i], c[i]);
  o1 = make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), a[4] | (a[5] << 16), a[6] | (a[7] << 16));
  o2 = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4] | (b[5] << 16), b[6] | (b[7] << 16));
  o3 = make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), c[6] | (c[7] << 16));
}

// ---- cp.async (LDGSTS): global -> shared without staging registers; completion is per thread ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- cross-CTA flags in global memory (release / acquire at gpu scope) ----
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// polling load: coherent at gpu scope but WITHOUT acquire semantics -- an acquire load makes ptxas
// emit CCTL.IVALL (invalidate the SM's whole L1) on every iteration of a spin loop, which stalls
// every other warp's memory instructions; poll relaxed, then fence once (acquire_fence_gpu)
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void acquire_fence_gpu() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace tc5
