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

#include <stdlib.h>

#include "tc_common.cuh"

using namespace ign_tc;

// -DIGN_MLP_PROFILE: per-phase clock64() sums of thread 0, printed after every launch.
#ifdef IGN_MLP_PROFILE
#include <stdio.h>
#include <string.h>
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

namespace {
PROF(__device__ unsigned long long mlp_prof[8];)

constexpr int NPART = 4;                      // warps per TMEM lane group
constexpr int EPI_GROUP = 256;                // threads of one epilogue group (8 warps)
constexpr int MMA_WARP = 16, TMA_WARP = 17;
constexpr int MLP_THREADS = 2 * EPI_GROUP + 64;
constexpr int ROWS = 128;
constexpr int A_IMG = ROWS * 128;

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

// Roles (warp-specialised, no CTA-wide barrier inside the tile loop):
//   warps 0-7 / 8-15  two epilogue groups; the chunk jobs of a tile (nc1 of layer 1, then nc2 of layer 2)
//                     alternate between them, job ctr -> group / stage ctr & 1: produce the A image of the
//                     job (x rows, or act1(D1 chunk + b1)) and arrive on full[stage]
//   warp 16           MMA issuer: waits full[stage] + the weight chunk, issues 12 UMMAs, commits to
//                     stage[stage] (and to d1 / d2 after a layer's last chunk)
//   warp 17           TMA producer: as soon as a stage is free, the next weight chunk for it
// so the activation of chunk c+1 runs under the UMMAs of chunk c, and the UMMAs are issued in chunk order.
// ACT1 / ACT2 >= 0: the activation is a compile-time constant (the selu / relu stacks of the examples);
// -1: taken from the arguments
template <int ACT1, int ACT2>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_head_tc_kernel(
    const float* __restrict__ x, int64_t M, int K1, const float* __restrict__ w1img, const float* __restrict__ b1,
    int N1, int act1_, const float* __restrict__ w2img, const float* __restrict__ b2, int N2, int act2_,
    const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nmax = N1 > N2 ? N1 : N2;
  const int stage_bytes = 2 * A_IMG + 2 * nmax * 128;
  __shared__ uint64_t bar_stage[2];                   // UMMAs that read the stage are done (tcgen05.commit)
  __shared__ uint64_t bar_full[2];                    // A image of the job is in place (256 arrivals)
  __shared__ uint64_t bar_b[2];                       // weight chunk landed (TMA bulk copy, complete_tx)
  __shared__ uint64_t bar_d1, bar_d2;                 // a layer's accumulator is complete
  __shared__ uint64_t bar_drained;                    // every epilogue thread has read D2 (512 arrivals)
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_b1[256];
  __shared__ __align__(16) float s_b2[256];
  __shared__ __align__(16) float s_w3[256];
  __shared__ float s_part[NPART][ROWS];               // per-row partial dot products of the head

  const int act1 = ACT1 >= 0 ? ACT1 : act1_, act2 = ACT2 >= 0 ? ACT2 : act2_;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_stage[0], 1); mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_full[0], EPI_GROUP); mbar_init(&bar_full[1], EPI_GROUP);
    mbar_init(&bar_b[0], 1); mbar_init(&bar_b[1], 1);
    mbar_init(&bar_d1, 1); mbar_init(&bar_d2, 1);
    mbar_init(&bar_drained, 2 * EPI_GROUP);
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
  uint32_t use[2] = {0, 0}, tiles_done = 0, ctr = 0;   // every role walks the same job sequence

  if (warp == MMA_WARP) {
    // ================================ MMA issuer ================================
    // job order: layer 1 of the first tile; then per tile its layer-2 chunks followed by layer 1 of the
    // NEXT tile (D1 is free once the last chunk has been activated), so that the head of a tile runs
    // under the next tile's first UMMAs
    auto mma_job = [&](bool l2, int c, bool first_of_layer, bool last_of_layer) {
      const int s = ctr & 1;
      if (lane == 0) {
        unsigned char* st = smem + s * stage_bytes;
        mbar_wait(&bar_full[s], use[s] & 1);
        mbar_wait(&bar_b[s], use[s] & 1);
        if (l2 && c == 0 && tiles_done > 0) mbar_wait(&bar_drained, (tiles_done - 1) & 1);   // D2 of the last tile read
        tc_fence_after();
        const uint32_t a_hi = smem_u32(st), b_hi = a_hi + 2 * A_IMG;
        const int n = l2 ? N2 : N1;
        umma_chunk_3x(l2 ? d2 : d1, a_hi, a_hi + A_IMG, b_hi, b_hi + n * 128, n, !first_of_layer);
        umma_commit(&bar_stage[s]);
        if (last_of_layer) umma_commit(l2 ? &bar_d2 : &bar_d1);
      }
      __syncwarp();
      use[s] += 1;
      ++ctr;
    };
    if ((int64_t)blockIdx.x < ntiles)
      for (int c = 0; c < nc1; ++c) mma_job(false, c, c == 0, c == nc1 - 1);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tiles_done) {
      for (int c = 0; c < nc2; ++c) mma_job(true, c, c == 0, c == nc2 - 1);
      if (tile + gridDim.x < ntiles)
        for (int c = 0; c < nc1; ++c) mma_job(false, c, c == 0, c == nc1 - 1);
    }
  } else if (warp == TMA_WARP) {
    // ================================ weight producer ================================
    auto tma_job = [&](bool l2, int c) {
      const int s = ctr & 1;
      if (lane == 0) {
        unsigned char* st = smem + s * stage_bytes;
        if (use[s] > 0) mbar_wait(&bar_stage[s], (use[s] - 1) & 1);
        const uint32_t bytes = 2 * (l2 ? N2 : N1) * 128;
        const char* src = (l2 ? reinterpret_cast<const char*>(w2img) : reinterpret_cast<const char*>(w1img)) +
                          (size_t)c * bytes;
        mbar_expect_tx(&bar_b[s], bytes);
        bulk_g2s(st + 2 * A_IMG, src, bytes, &bar_b[s]);
      }
      __syncwarp();
      use[s] += 1;
      ++ctr;
    };
    if ((int64_t)blockIdx.x < ntiles)
      for (int c = 0; c < nc1; ++c) tma_job(false, c);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = 0; c < nc2; ++c) tma_job(true, c);
      if (tile + gridDim.x < ntiles)
        for (int c = 0; c < nc1; ++c) tma_job(false, c);
    }
  } else {
    // ================================ epilogue groups ================================
    const int g = warp >> 3, gw = warp & 7, gtid = tid & (EPI_GROUP - 1);
    const int q = warp & 3;                               // TMEM lane group of this warp
    const int row = q * 32 + lane;
    const int half = gw >> 2;                             // which 16 of a chunk's 32 columns
    const int part = warp >> 2;                           // head: which quarter of the N2 columns
    const float head_b = b3 ? __ldg(b3) : 0.0f;
    PROF(long long p_wait = 0, p_epi = 0, p_head = 0, p_d2 = 0, c0, c1;)
    // one chunk job: the A image of layer-1 chunk c of the tile starting at row m_l1, or of layer-2 chunk c
    // of the current tile; only the group whose turn it is (ctr & 1) does anything
    auto epi_job = [&](bool l2, int c, int64_t m_l1) {
      const int s = ctr & 1;
      ++ctr;
      if (s != g) return;
      unsigned char* st = smem + s * stage_bytes;
      PROF(c0 = clock64();)
      if (use[s] > 0) mbar_wait(&bar_stage[s], (use[s] - 1) & 1);
      if (!l2) {
        // 128 rows x 32 floats of x, 8 lanes per row
#pragma unroll
        for (int i = 0; i < 1024 / EPI_GROUP; ++i) {
          const int idx = gtid + i * EPI_GROUP;
          const int r = idx >> 3, c4 = idx & 7;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m_l1 + r < M) v = ldg_f4(x + (m_l1 + r) * K1 + c * 32 + c4 * 4);
          store_split(st, st + A_IMG, r, c4, v);
        }
        PROF(c1 = clock64(); p_epi += c1 - c0;)
      } else {
        // chunk c of act1(D1 + b1), this thread: one row x 16 columns
        if (c < 2) {                                      // this group's first layer-2 job of the tile
          mbar_wait(&bar_d1, tiles_done & 1);
          tc_fence_after();
        }
        PROF(c1 = clock64(); p_wait += c1 - c0;)
        const int col = c * 32 + half * 16;
        uint32_t r[16];
        tmem_ld16_nowait(d1 + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(s_b1 + col + i);
          float4 v;
          v.x = act_epi(act1, __uint_as_float(r[i]) + bv.x);
          v.y = act_epi(act1, __uint_as_float(r[i + 1]) + bv.y);
          v.z = act_epi(act1, __uint_as_float(r[i + 2]) + bv.z);
          v.w = act_epi(act1, __uint_as_float(r[i + 3]) + bv.w);
          store_split(st, st + A_IMG, row, half * 4 + i / 4, v);
        }
        PROF(c0 = clock64(); p_epi += c0 - c1;)
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_full[s]);
      use[s] += 1;
    };
    if ((int64_t)blockIdx.x < ntiles)
      for (int c = 0; c < nc1; ++c) epi_job(false, c, (int64_t)blockIdx.x * ROWS);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tiles_done) {
      const int64_t m0 = tile * ROWS;
      for (int c = 0; c < nc2; ++c) epi_job(true, c, 0);
      if (tile + gridDim.x < ntiles)
        for (int c = 0; c < nc1; ++c) epi_job(false, c, (tile + gridDim.x) * ROWS);
      // ---- head: out = act2(D2 + b2) . w3 + b3
      PROF(c0 = clock64();)
      mbar_wait(&bar_d2, tiles_done & 1);
      tc_fence_after();
      PROF(c1 = clock64(); p_d2 += c1 - c0;)
      {
        const int quarter = N2 / NPART;                  // N2 % 32 == 0: a multiple of 8
        float acc = 0.0f, acc2 = 0.0f;
        int cb = 0;
        for (; cb + 32 <= quarter; cb += 32) {           // 32 columns per round trip to TMEM
          const int col = part * quarter + cb;
          uint32_t r0[16], r1[16];
          tmem_ld16_nowait(d2 + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r0);
          tmem_ld16_nowait(d2 + ((uint32_t)(q * 32) << 16) + (uint32_t)(col + 16), r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(s_b2 + col + i);
            const float4 hw = *reinterpret_cast<const float4*>(s_w3 + col + i);
            const float4 bw = *reinterpret_cast<const float4*>(s_b2 + col + 16 + i);
            const float4 hx = *reinterpret_cast<const float4*>(s_w3 + col + 16 + i);
            acc = fmaf(act_epi(act2, __uint_as_float(r0[i]) + bv.x), hw.x, acc);
            acc2 = fmaf(act_epi(act2, __uint_as_float(r1[i]) + bw.x), hx.x, acc2);
            acc = fmaf(act_epi(act2, __uint_as_float(r0[i + 1]) + bv.y), hw.y, acc);
            acc2 = fmaf(act_epi(act2, __uint_as_float(r1[i + 1]) + bw.y), hx.y, acc2);
            acc = fmaf(act_epi(act2, __uint_as_float(r0[i + 2]) + bv.z), hw.z, acc);
            acc2 = fmaf(act_epi(act2, __uint_as_float(r1[i + 2]) + bw.z), hx.z, acc2);
            acc = fmaf(act_epi(act2, __uint_as_float(r0[i + 3]) + bv.w), hw.w, acc);
            acc2 = fmaf(act_epi(act2, __uint_as_float(r1[i + 3]) + bw.w), hx.w, acc2);
          }
        }
        for (; cb < quarter; cb += 8) {
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
        tc_fence_before();
        mbar_arrive(&bar_drained);                       // D2 may be overwritten by the next tile
        s_part[part][row] = acc + acc2;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(2 * EPI_GROUP) : "memory");   // partials visible (epilogue warps only)
      if (tid < ROWS && m0 + tid < M) {                  // fixed summation order: deterministic
        float v = s_part[0][tid];
#pragma unroll
        for (int p = 1; p < NPART; ++p) v += s_part[p][tid];
        out[m0 + tid] = v + head_b;
      }
      PROF(c0 = clock64(); p_head += c0 - c1;)
    }
    PROF(if (gtid == 32) {
      const long long v[4] = {p_wait, p_epi, p_d2, p_head};
      for (int i = 0; i < 4; ++i) atomicAdd(&mlp_prof[i], (unsigned long long)v[i]);
    })
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
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(m, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  auto launch = [&](auto kernel) -> int {
    IGN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, MLP_THREADS, smem, st>>>(x, m, k1, img1, b1, n1, act1, img2, b2, n2, act2, w3, b3, out);
    return IGN_OK;
  };
  int rc;
  if (act1 == IGN_ACT_SELU && act2 == IGN_ACT_SELU) rc = launch(mlp_head_tc_kernel<IGN_ACT_SELU, IGN_ACT_SELU>);
  else if (act1 == IGN_ACT_RELU && act2 == IGN_ACT_RELU) rc = launch(mlp_head_tc_kernel<IGN_ACT_RELU, IGN_ACT_RELU>);
  else rc = launch(mlp_head_tc_kernel<-1, -1>);
  if (rc != IGN_OK) return rc;
  IGN_CHECK_LAUNCH("mlp_head_tc");
  PROF({
    unsigned long long h[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, mlp_prof, sizeof(h));
    const double nt = (double)tiles * 2, nc = (double)tiles * (n1 / 32 + k1 / 32);
    fprintf(stderr, "mlp_head (warp 1 of each group) per job: wait stage/d1 %.0f produce %.0f | per tile: wait d2 %.0f head %.0f\n",
            h[0] / nc, h[1] / nc, h[2] / nt, h[3] / nt);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(mlp_prof, h, sizeof(h));
  })
  return IGN_OK;
}
