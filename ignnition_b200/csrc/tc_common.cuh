// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (sm_100a only).
//
// Operand images are [rows][32 fp32] blocks in the canonical K-major SWIZZLE_128B layout
// (8-row groups of 1024 bytes, 16-byte chunks XOR-swizzled by row & 7); every fp32 operand is
// stored twice, as hi = rna_tf32(v) and lo = rna_tf32(v - hi), and every product is issued as
// A_hi B_hi + A_lo B_hi + A_hi B_lo (3xTF32) so results keep fp32 accuracy.
#pragma once

#include "common.cuh"

namespace ign_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, 16-byte units
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32: D fp32, A/B tf32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[128, n] (+)= A[128, 32] B[n, 32]^T as 3xTF32 over the 4 K = 8 steps of one 32-float chunk
__device__ __forceinline__ void umma_chunk_3x(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                              uint32_t b_lo, int n, bool accumulate_first) {
  const uint32_t idesc = umma_idesc(n);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t ko = kk * 32;                      // 8 tf32 = 32 bytes along K inside the swizzle row
    umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(b_hi + ko), idesc, (accumulate_first || kk > 0) ? 1u : 0u);
    umma_tf32(tmem_d, umma_desc(a_lo + ko), umma_desc(b_hi + ko), idesc, 1u);
    umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(b_lo + ko), idesc, 1u);
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one try_wait (may suspend for a hardware-chosen time); true once the phase with this parity completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (expect_tx armed first)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {     // one warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {        // one warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// four adjacent floats added to global memory with one reduction (16-byte aligned address)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// byte offset of float (row r, k) inside a [rows][32] K-major SWIZZLE_128B image
__host__ __device__ __forceinline__ int sw128_off(int r, int k) {
  return r * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4;
}
// store 4 consecutive floats (16-byte chunk c4 of row r) as hi / lo into two swizzled images
__device__ __forceinline__ void store_split(unsigned char* img_hi, unsigned char* img_lo, int r, int c4, float4 v) {
  float4 hi, lo;
  tf32_split(v.x, hi.x, lo.x);
  tf32_split(v.y, hi.y, lo.y);
  tf32_split(v.z, hi.z, lo.z);
  tf32_split(v.w, hi.w, lo.w);
  const int off = r * 128 + ((c4 ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(img_hi + off) = hi;
  *reinterpret_cast<float4*>(img_lo + off) = lo;
}

// 1 / (1 + 2^(-x log2 e)) on the SFU approximations (ex2 2 ulp, rcp 1 ulp): 4 instructions, saturates
// cleanly (ex2 -> inf gives rcp -> 0)
__device__ __forceinline__ float fast_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
// tanh for |x| < 1/4 as an odd polynomial (relative error 1e-8): 2 / (1 + e) - 1 has an ABSOLUTE error of
// ~3 ulp of 1.0, which is a large relative error on a small candidate state and accumulates over the
// steps of a walk (measured: path states 3.7e-5 -> fp32-kernel level with this branch)
__device__ __forceinline__ float tanh_small(float x) {
  const float x2 = x * x;
  return x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 0.021869488f, -0.053968254f), 0.13333334f), -0.33333334f), 1.0f);
}
__device__ __forceinline__ float fast_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fabsf(x) < 0.25f ? tanh_small(x) : fmaf(2.0f, r, -1.0f);
}
// exp(x) - 1 for x <= 0 on one SFU operation: the ABSOLUTE error (2^-22 of a value in (0, 1]) is what the
// next GEMM sees, so the relative accuracy of expm1 near 0 is not needed (2.4e-7 x scale x alpha = 4e-7,
// below the fp32 accumulation error of the dot products that consume it; measured: the model's parity
// figure is the same 6.1e-6 with a Taylor branch near zero, the readout 0.5 ms slower)
__device__ __forceinline__ float fast_expm1_neg(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return e - 1.0f;
}
// activation of a tensor-core epilogue.  Call it with a COMPILE-TIME act wherever possible: with a run-time
// act every element drags the whole if-chain (and the libm slow paths of act_fwd) through the epilogue --
// measured 2.4x on the fused readout (profiles/r1_mlp_head_phases.md)
__device__ __forceinline__ float act_epi(int act, float x) {
  if (act == IGN_ACT_SELU) return x > 0.0f ? IGN_SELU_SCALE * x : (IGN_SELU_SCALE * IGN_SELU_ALPHA) * fast_expm1_neg(x);
  if (act == IGN_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == IGN_ACT_LINEAR) return x;
  if (act == IGN_ACT_ELU) return x > 0.0f ? x : fast_expm1_neg(x);
  if (act == IGN_ACT_SIGMOID) return fast_sigmoid(x);
  if (act == IGN_ACT_TANH) return fast_tanh(x);
  return act_fwd(act, x);
}

// One GRU element on 5 SFU operations instead of 6: the reciprocals of the update gate and of the
// candidate's tanh share one rcp,  z = b / (a b),  tanh = 2 a / (a b) - 1  with a = 1 + e^-pz,
// b = 1 + e^-2ph.  The exponents are clamped (NaN-propagating min) so that a b stays finite:
// sigmoid(x) below 2^-40 and 1 - |tanh| below 2^-39 are returned as those bounds.
__device__ __forceinline__ float fast_gru_gate(float pz, float pr, float pxh, float phh, float h) {
  float tz, ez, er, r, th, eh, inv;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(tz) : "f"(pz * -1.4426950408889634f), "f"(40.0f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ez) : "f"(tz));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(er) : "f"(pr * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + er));
  const float ph = fmaf(r, phh, pxh);
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(th) : "f"(ph * -2.8853900817779268f), "f"(40.0f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eh) : "f"(th));
  const float a = 1.0f + ez, b = 1.0f + eh;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(a * b));
  const float z = b * inv;
  const float hh = fabsf(ph) < 0.25f ? tanh_small(ph) : fmaf(2.0f * a, inv, -1.0f);
  return fmaf(z, h - hh, hh);
}

}  // namespace ign_tc
