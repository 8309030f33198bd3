// Fused readout MLP on the tcgen05 tensor cores (sm_100a):
//     out[m] = act2( act1(x W1 + b1) W2 + b2 ) . w3 + b3
// i.e. the predict stack of RouteNet / Q-size (32 -> 256 selu -> 256 selu -> 1 linear,
// examples/Routenet/model_description.json:119-141; built by Feed_forward_model.construct_tf_model,
// code/utils/auxilary_classes.py:918-975, run at code/utils/generate_model.py:623-624).
// The two hidden activations never leave the SM: per 128-row tile
//   1. D1[128, N1] = x W1 as 3xTF32 tcgen05.mma (TMEM accumulator #1);
//   2. for every 32-column chunk c of D1: tcgen05.ld -> + b1 -> act1 -> hi/lo split -> the swizzled
//      A image of chunk c in shared memory, W2 chunk c streamed in with cp.async, 12 tcgen05.mma
//      accumulate into D2[128, N2] (TMEM accumulator #2); two stages, so the MMAs of chunk c run
//      while chunk c+1 is being produced;
//   3. D2 -> + b2 -> act2 -> dot with w3 -> out (four partial sums per row, combined in a fixed order
//      through shared memory: deterministic).
// HBM traffic is 4 K1 + 4 bytes per row instead of 4 (K1 + 2 N1 + 2 N2 + 1).  Inference only (the
// train step keeps the per-layer kernels, which save the pre-activations).

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int NPART = 4;                      // warps per TMEM lane group
constexpr int TC_THREADS = 128 * NPART;
constexpr int ROWS = 128;
constexpr int A_IMG = ROWS * 128;

__device__ __forceinline__ float fast_expm1_neg(float x) {       // x <= 0
  const float p = x * (1.0f + x * (0.5f + x * (0.16666667f + x * (0.041666668f + x * 0.0083333338f))));
  return x > -0.125f ? p : __expf(x) - 1.0f;
}
__device__ __forceinline__ float act_epi(int act, float x) {
  if (act == IGN_ACT_SELU) return x > 0.0f ? IGN_SELU_SCALE * x : (IGN_SELU_SCALE * IGN_SELU_ALPHA) * fast_expm1_neg(x);
  if (act == IGN_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == IGN_ACT_LINEAR) return x;
  if (act == IGN_ACT_ELU) return x > 0.0f ? x : fast_expm1_neg(x);
  return act_fwd(act, x);
}

// W[K,N] -> chunk images [N rows][32] hi then lo (same format as dense_tc_prep)
__global__ void mlp_prep_kernel(const float* __restrict__ w, int K, int N, float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i % N;
  float hi, lo;
  tf32_split(w[i], hi, lo);
  char* base = reinterpret_cast<char*>(img) + (size_t)(k / 32) * (2 * N * 128);
  *reinterpret_cast<float*>(base + sw128_off(n, k % 32)) = hi;
  *reinterpret_cast<float*>(base + N * 128 + sw128_off(n, k % 32)) = lo;
}

__global__ void __launch_bounds__(TC_THREADS, 1) mlp_head_tc_kernel(
    const float* __restrict__ x, int64_t M, int K1, const float* __restrict__ w1img, const float* __restrict__ b1,
    int N1, int act1, const float* __restrict__ w2img, const float* __restrict__ b2, int N2, int act2,
    const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nmax = N1 > N2 ? N1 : N2;
  const int stage_bytes = 2 * A_IMG + 2 * nmax * 128;
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_b[2];                       // weight chunk landed (TMA bulk copy, complete_tx)
  __shared__ uint64_t bar_d1, bar_d2;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_b1[256];
  __shared__ __align__(16) float s_b2[256];
  __shared__ __align__(16) float s_w3[256];
  __shared__ float s_part[NPART][ROWS];               // per-row partial dot products of the head

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_stage[0], 1); mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_b[0], 1); mbar_init(&bar_b[1], 1);
    mbar_init(&bar_d1, 1); mbar_init(&bar_d2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid < N1) s_b1[tid] = b1 ? b1[tid] : 0.0f;
  if (tid < N2) { s_b2[tid] = b2 ? b2[tid] : 0.0f; s_w3[tid] = w3[tid]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d1 = tmem_base_s, d2 = tmem_base_s + 256;
  const int nc1 = K1 / 32, nc2 = N1 / 32;
  const int64_t ntiles = (M + ROWS - 1) / ROWS;
  uint32_t use[2] = {0, 0}, tiles_done = 0, ctr = 0;
  const int q = warp & 3, part = warp >> 2;          // part in [0, NPART)
  const float head_b = b3 ? __ldg(b3) : 0.0f;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * ROWS;
    // ---- layer 1: D1 = x W1
    for (int c = 0; c < nc1; ++c, ++ctr) {
      const int s = ctr & 1;
      unsigned char* st = smem + s * stage_bytes;
      if (use[s] > 0) mbar_wait(&bar_stage[s], (use[s] - 1) & 1);
      if (tid == 0) {                                   // weight chunk: one TMA bulk copy, no LSU traffic
        mbar_expect_tx(&bar_b[s], 2 * N1 * 128);
        bulk_g2s(st + 2 * A_IMG, reinterpret_cast<const char*>(w1img) + (size_t)c * (2 * N1 * 128), 2 * N1 * 128,
                 &bar_b[s]);
      }
#pragma unroll
      for (int j = 0; j < 1024 / TC_THREADS; ++j) {
        const int idx = tid + j * TC_THREADS;
        const int r = idx >> 3, c4 = idx & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < M) v = ldg_f4(x + (m0 + r) * K1 + c * 32 + c4 * 4);
        store_split(st, st + A_IMG, r, c4, v);
      }
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(&bar_b[s], use[s] & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(st), b_hi = a_hi + 2 * A_IMG;
        umma_chunk_3x(d1, a_hi, a_hi + A_IMG, b_hi, b_hi + N1 * 128, N1, c > 0);
        umma_commit(&bar_stage[s]);
        if (c == nc1 - 1) umma_commit(&bar_d1);
      }
      use[s] += 1;
    }
    mbar_wait(&bar_d1, tiles_done & 1);
    tc_fence_after();
    // ---- layer 2: chunk c of act1(D1 + b1) is produced into shared memory and multiplied at once
    for (int c = 0; c < nc2; ++c, ++ctr) {
      const int s = ctr & 1;
      unsigned char* st = smem + s * stage_bytes;
      if (use[s] > 0) mbar_wait(&bar_stage[s], (use[s] - 1) & 1);
      if (tid == 0) {
        mbar_expect_tx(&bar_b[s], 2 * N2 * 128);
        bulk_g2s(st + 2 * A_IMG, reinterpret_cast<const char*>(w2img) + (size_t)c * (2 * N2 * 128), 2 * N2 * 128,
                 &bar_b[s]);
      }
      {
        const int col = c * 32 + part * 8;                // this thread: row q*32+lane, 8 of the chunk's 32 columns
        uint32_t r[8];
        tmem_ld8_nowait(d1 + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
        tmem_ld_wait();
        const int row = q * 32 + lane;
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(s_b1 + col + i);
          float4 v;
          v.x = act_epi(act1, __uint_as_float(r[i]) + bv.x);
          v.y = act_epi(act1, __uint_as_float(r[i + 1]) + bv.y);
          v.z = act_epi(act1, __uint_as_float(r[i + 2]) + bv.z);
          v.w = act_epi(act1, __uint_as_float(r[i + 3]) + bv.w);
          store_split(st, st + A_IMG, row, part * 2 + i / 4, v);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(&bar_b[s], use[s] & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(st), b_hi = a_hi + 2 * A_IMG;
        umma_chunk_3x(d2, a_hi, a_hi + A_IMG, b_hi, b_hi + N2 * 128, N2, c > 0);
        umma_commit(&bar_stage[s]);
        if (c == nc2 - 1) umma_commit(&bar_d2);
      }
      use[s] += 1;
    }
    // ---- head: out = act2(D2 + b2) . w3 + b3
    mbar_wait(&bar_d2, tiles_done & 1);
    tc_fence_after();
    {
      const int64_t row = m0 + q * 32 + lane;
      const int quarter = N2 / NPART;                  // N2 % 32 == 0: a multiple of 8
      float acc = 0.0f;
      for (int cb = 0; cb < quarter; cb += 8) {
        const int col = part * quarter + cb;
        uint32_t r[8];
        tmem_ld8_nowait(d2 + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(s_b2 + col + i);
          const float4 hw = *reinterpret_cast<const float4*>(s_w3 + col + i);
          acc = fmaf(act_epi(act2, __uint_as_float(r[i]) + bv.x), hw.x, acc);
          acc = fmaf(act_epi(act2, __uint_as_float(r[i + 1]) + bv.y), hw.y, acc);
          acc = fmaf(act_epi(act2, __uint_as_float(r[i + 2]) + bv.z), hw.z, acc);
          acc = fmaf(act_epi(act2, __uint_as_float(r[i + 3]) + bv.w), hw.w, acc);
        }
      }
      s_part[part][q * 32 + lane] = acc;
    }
    tiles_done += 1;
    tc_fence_before();
    __syncthreads();                  // D1 / D2 drained before the next tile's MMAs; partials visible
    if (tid < ROWS && m0 + tid < M) { // fixed summation order: deterministic
      float v = s_part[0][tid];
#pragma unroll
      for (int p = 1; p < NPART; ++p) v += s_part[p][tid];
      out[m0 + tid] = v + head_b;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

}  // namespace

bool ign_mlp_head_tc_supported(int k1, int n1, int n2) {
  return k1 % 32 == 0 && k1 >= 32 && n1 % 32 == 0 && n1 >= 32 && n1 <= 256 && n2 % 32 == 0 && n2 >= 32 && n2 <= 256;
}
size_t ign_mlp_head_tc_ws(int k1, int n1, int n2) {
  return (size_t)(k1 / 32) * 2 * n1 * 128 + (size_t)(n1 / 32) * 2 * n2 * 128;
}

int ign_mlp_head_tc_launch(const float* x, int64_t m, int k1, const float* w1, const float* b1, int n1, int act1,
                           const float* w2, const float* b2, int n2, int act2, const float* w3, const float* b3,
                           float* out, void* ws, cudaStream_t st) {
  float* img1 = reinterpret_cast<float*>(ws);
  float* img2 = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)(k1 / 32) * 2 * n1 * 128);
  mlp_prep_kernel<<<(unsigned)ign_cdiv((int64_t)k1 * n1, 256), 256, 0, st>>>(w1, k1, n1, img1);
  IGN_CHECK_LAUNCH("mlp_prep");
  mlp_prep_kernel<<<(unsigned)ign_cdiv((int64_t)n1 * n2, 256), 256, 0, st>>>(w2, n1, n2, img2);
  IGN_CHECK_LAUNCH("mlp_prep");
  const int nmax = n1 > n2 ? n1 : n2;
  const size_t smem = 1024 + 2 * (size_t)(2 * A_IMG + 2 * nmax * 128);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    IGN_CUDA(cudaFuncSetAttribute(mlp_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(m, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  mlp_head_tc_kernel<<<grid, TC_THREADS, smem, st>>>(x, m, k1, img1, b1, n1, act1, img2, b2, n2, act2, w3, b3, out);
  IGN_CHECK_LAUNCH("mlp_head_tc");
  return IGN_OK;
}
