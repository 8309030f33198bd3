// Shared helpers for the ignnition_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ignnition_b200.h"

#define IGN_NUM_SMS 148   // B200: 2 dies x 74 SMs

// thread-local error text, read back through ign_last_error()
void ign_set_error(const char* fmt, ...);

#define IGN_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ign_set_error(__VA_ARGS__);           \
      return (code);                        \
    }                                       \
  } while (0)

// true the first time this call site runs on the CURRENT DEVICE in this thread: function attributes
// (cudaFuncSetAttribute) are per device, so a process that drives a second GPU has to set them there too
#define IGN_ONCE_PER_DEVICE()                                                    \
  ([] {                                                                          \
    static thread_local unsigned long long seen_ = 0;                            \
    int d_ = 0;                                                                  \
    cudaGetDevice(&d_);                                                          \
    const unsigned long long bit_ = 1ull << (d_ & 63);                           \
    if (seen_ & bit_) return false;                                              \
    seen_ |= bit_;                                                               \
    return true;                                                                 \
  }())

// process-wide count of kernels launched by this library (ign_launch_count)
void ign_count_launch();

// after a launch: report asynchronous-launch errors without synchronising
#define IGN_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    ign_count_launch();                                                                 \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ign_set_error("IGNNITION: %s launch failed: %s", name, cudaGetErrorString(e__));  \
      return (int)e__;                                                                  \
    }                                                                                   \
  } while (0)

#define IGN_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ign_set_error("IGNNITION: %s failed: %s", #call, cudaGetErrorString(e__));         \
      return (int)e__;                                                                   \
    }                                                                                    \
  } while (0)

static inline cudaStream_t ign_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ign_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t ign_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
static inline size_t ign_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

#define IGN_SELU_ALPHA 1.6732632423543772848170429916717f
#define IGN_SELU_SCALE 1.0507009873554804934193349852946f

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case IGN_ACT_RELU: return fmaxf(x, 0.0f);
    case IGN_ACT_SELU: return x > 0.0f ? IGN_SELU_SCALE * x : IGN_SELU_SCALE * IGN_SELU_ALPHA * expm1f(x);
    case IGN_ACT_SIGMOID: return sigmoid_f(x);
    case IGN_ACT_TANH: return tanhf(x);
    case IGN_ACT_ELU: return x > 0.0f ? x : expm1f(x);
    case IGN_ACT_SOFTPLUS: return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.0f);
    case IGN_ACT_LEAKY_RELU: return x > 0.0f ? x : 0.2f * x;
    default: return x;
  }
}
// derivative of the activation w.r.t. its pre-activation input
__device__ __forceinline__ float act_bwd(int act, float pre) {
  switch (act) {
    case IGN_ACT_RELU: return pre > 0.0f ? 1.0f : 0.0f;
    case IGN_ACT_SELU: return pre > 0.0f ? IGN_SELU_SCALE : IGN_SELU_SCALE * IGN_SELU_ALPHA * expf(pre);
    case IGN_ACT_SIGMOID: { float s = sigmoid_f(pre); return s * (1.0f - s); }
    case IGN_ACT_TANH: { float t = tanhf(pre); return 1.0f - t * t; }
    case IGN_ACT_ELU: return pre > 0.0f ? 1.0f : expf(pre);
    case IGN_ACT_SOFTPLUS: return sigmoid_f(pre);
    case IGN_ACT_LEAKY_RELU: return pre > 0.0f ? 1.0f : 0.2f;
    default: return 1.0f;
  }
}

// the same derivative from the activation's OUTPUT y = act(pre) (IGN_ACT_FROM_OUTPUT)
__device__ __forceinline__ float act_bwd_from_output(int act, float y) {
  switch (act) {
    case IGN_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
    case IGN_ACT_SELU: return y > 0.0f ? IGN_SELU_SCALE : y + IGN_SELU_SCALE * IGN_SELU_ALPHA;
    case IGN_ACT_SIGMOID: return y * (1.0f - y);
    case IGN_ACT_TANH: return 1.0f - y * y;
    case IGN_ACT_ELU: return y > 0.0f ? 1.0f : y + 1.0f;
    case IGN_ACT_SOFTPLUS: return 1.0f - expf(-y);
    case IGN_ACT_LEAKY_RELU: return y > 0.0f ? 1.0f : 0.2f;
    default: return 1.0f;
  }
}

// cp.async 16-byte global->shared copy (LDGSTS), used to prefetch gathered rows
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 3xTF32 operand split: hi = round-to-nearest TF32 of v, lo = round-to-nearest TF32 of (v - hi).
// |v - (hi + lo)| <= 2^-24 |v|, i.e. the pair carries v to fp32 precision; the tensor core reads
// both exactly (their low 13 mantissa bits are zero).
// round to nearest, ties away from zero, onto the 10-bit tf32 mantissa: what cvt.rna.tf32.f32 computes
// for every finite input, in two integer instructions (ptxas expands the cvt into four, with an
// inf / nan guard the split does not need: inf stays inf, and a nan reaches the product through
// lo = v - hi)
__device__ __forceinline__ float tf32_rna(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) {
  hi = tf32_rna(v);
  lo = tf32_rna(v - hi);
}
