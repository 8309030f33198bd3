// Shared pieces of the fused GRU kernels (forward and backward): tile geometry, weight staging
// and the register-tiled gate GEMM.
//
// Keras GRUCell v2 (reset_after=True, gates z|r|h, bias[2,3u]) as the reference instantiates it
// (code/utils/auxilary_classes.py:740-750, called at :764 and :785-790):
//   mx = x K + b0 ; mh = h R + b1
//   z = sigmoid(mx_z + mh_z) ; r = sigmoid(mx_r + mh_r) ; hh = tanh(mx_h + r * mh_h)
//   h' = z h + (1 - z) hh
//
// One CTA of 256 threads owns a tile of R destinations.  K, R and the biases stay in shared
// memory for the CTA's whole life (persistent grid); x and h tiles are staged in shared memory
// with a +4 float pad; each thread owns TR rows x TU units and keeps, per (row, unit), the four
// pre-activations  a_z, a_r (x and h parts merged), a_xh, a_hh  in registers, so the gate math
// needs no exchange between threads.  fp32 FMA on the CUDA cores: the weight GEMMs here have
// K = 32..64 and the result must match the fp32 reference to 1e-5 (see DESIGN.md for the
// tensor-core plan).
#pragma once

#include "common.cuh"

// gru.cu and gru_bwd.cu are compiled twice: as they are (tiles of 128 / 64 destinations per CTA, for graphs that
// fill the machine) and with -DIGN_GRU_SMALL_TILE (16 destinations per CTA: a batch of 3 samples is 546 paths = 5 big
// tiles on 148 SMs; small tiles spread the same rows over 35 CTAs and the kernel's latency drops with the rows per
// thread).  The second compilation lives in its own namespace and exports its launchers with a _small_tile suffix.
#ifdef IGN_GRU_SMALL_TILE
#define ign_gru ign_gru_small_tile
#define IGN_GRU_FN(name) name##_small_tile
#else
#define IGN_GRU_FN(name) name
#endif

// rows below which the small tiles win: up to four of them per SM
inline bool ign_gru_use_small_tile(int64_t rows) {
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return rows <= (int64_t)16 * 4 * sms;
}

namespace ign_gru {

constexpr int THREADS = 256;

template <int U>
struct Geo;                       // TU units x TR rows per thread, R rows per CTA tile
#ifdef IGN_GRU_SMALL_TILE
template <>
struct Geo<16> { static constexpr int TU = 1, TR = 1; };
template <>
struct Geo<32> { static constexpr int TU = 2, TR = 1; };
template <>
struct Geo<64> { static constexpr int TU = 4, TR = 1; };
#else
template <>
struct Geo<16> { static constexpr int TU = 1, TR = 8; };
template <>
struct Geo<32> { static constexpr int TU = 2, TR = 8; };
template <>
struct Geo<64> { static constexpr int TU = 4, TR = 4; };
#endif

template <int U>
struct Tile {
  static constexpr int TU = Geo<U>::TU;
  static constexpr int TR = Geo<U>::TR;
  static constexpr int NUG = U / TU;                // unit groups
  static constexpr int NRG = THREADS / NUG;         // row groups
  static constexpr int R = NRG * TR;                // rows per tile
};

// shared-memory carve-up (floats): K [FI][3U] | Rk [U][3U] | b [2][3U]
template <int FI, int U>
struct WeightSmem {
  static constexpr int K_OFF = 0;
  static constexpr int R_OFF = FI * 3 * U;
  static constexpr int B_OFF = R_OFF + U * 3 * U;
  static constexpr int FLOATS = B_OFF + 6 * U;
};

template <int FI, int U>
__device__ __forceinline__ void load_weights(float* sw, const float* __restrict__ kernel,
                                             const float* __restrict__ rkernel, const float* __restrict__ bias) {
  using W = WeightSmem<FI, U>;
  for (int i = threadIdx.x * 4; i < FI * 3 * U; i += THREADS * 4) st_f4(sw + W::K_OFF + i, ldg_f4(kernel + i));
  for (int i = threadIdx.x * 4; i < U * 3 * U; i += THREADS * 4) st_f4(sw + W::R_OFF + i, ldg_f4(rkernel + i));
  for (int i = threadIdx.x; i < 6 * U; i += THREADS) sw[W::B_OFF + i] = bias[i];
}

template <int TU>
struct VecLoad;
template <>
struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float* o) { o[0] = p[0]; }
};
template <>
struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = v.x; o[1] = v.y;
  }
};
template <>
struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};

// Pre-activations for this thread's TR x TU block.  X: [R][FI+4], H: [R][U+4] in shared memory.
// The thread's rows are rg + i*NRG (strided, so the two row groups of a warp read rows that are
// FI+4 floats apart: different banks).  Output arrays are [TR][TU]; az / ar hold
// (x part + h part + both biases).
template <int FI, int U>
__device__ __forceinline__ void gate_gemm(const float* __restrict__ sw, const float* __restrict__ X,
                                          const float* __restrict__ H, int rg, int u0,
                                          float (&az)[Geo<U>::TR][Geo<U>::TU], float (&ar)[Geo<U>::TR][Geo<U>::TU],
                                          float (&axh)[Geo<U>::TR][Geo<U>::TU],
                                          float (&ahh)[Geo<U>::TR][Geo<U>::TU]) {
  using W = WeightSmem<FI, U>;
  constexpr int TU = Geo<U>::TU, TR = Geo<U>::TR;
  constexpr int XS = FI + 4, HS = U + 4, NRG = Tile<U>::NRG;
  const float* b0 = sw + W::B_OFF;
  const float* b1 = b0 + 3 * U;
#pragma unroll
  for (int j = 0; j < TU; ++j) {
    const float bz = b0[u0 + j] + b1[u0 + j];
    const float br = b0[U + u0 + j] + b1[U + u0 + j];
    const float bx = b0[2 * U + u0 + j];
    const float bh = b1[2 * U + u0 + j];
#pragma unroll
    for (int i = 0; i < TR; ++i) { az[i][j] = bz; ar[i][j] = br; axh[i][j] = bx; ahh[i][j] = bh; }
  }
  const float* Kz = sw + W::K_OFF + u0;
#pragma unroll 2
  for (int k = 0; k < FI; k += 4) {
    float4 xv[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (rg + i * NRG) * XS + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float wz[TU], wr[TU], wh[TU];
      VecLoad<TU>::ld(Kz + (k + kk) * 3 * U, wz);
      VecLoad<TU>::ld(Kz + (k + kk) * 3 * U + U, wr);
      VecLoad<TU>::ld(Kz + (k + kk) * 3 * U + 2 * U, wh);
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const float x = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
#pragma unroll
        for (int j = 0; j < TU; ++j) {
          az[i][j] = fmaf(x, wz[j], az[i][j]);
          ar[i][j] = fmaf(x, wr[j], ar[i][j]);
          axh[i][j] = fmaf(x, wh[j], axh[i][j]);
        }
      }
    }
  }
  const float* Rz = sw + W::R_OFF + u0;
#pragma unroll 2
  for (int k = 0; k < U; k += 4) {
    float4 hv[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) hv[i] = *reinterpret_cast<const float4*>(H + (rg + i * NRG) * HS + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float wz[TU], wr[TU], wh[TU];
      VecLoad<TU>::ld(Rz + (k + kk) * 3 * U, wz);
      VecLoad<TU>::ld(Rz + (k + kk) * 3 * U + U, wr);
      VecLoad<TU>::ld(Rz + (k + kk) * 3 * U + 2 * U, wh);
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const float h = kk == 0 ? hv[i].x : kk == 1 ? hv[i].y : kk == 2 ? hv[i].z : hv[i].w;
#pragma unroll
        for (int j = 0; j < TU; ++j) {
          az[i][j] = fmaf(h, wz[j], az[i][j]);
          ar[i][j] = fmaf(h, wr[j], ar[i][j]);
          ahh[i][j] = fmaf(h, wh[j], ahh[i][j]);
        }
      }
    }
  }
}

// h' from the pre-activations and the old state
__device__ __forceinline__ float gru_out(float az, float ar, float axh, float ahh, float hold) {
  const float z = sigmoid_f(az);
  const float r = sigmoid_f(ar);
  const float hh = tanhf(fmaf(r, ahh, axh));
  return fmaf(z, hold - hh, hh);                  // z*h + (1-z)*hh
}

}  // namespace ign_gru
