// One GRU step for every destination on the tcgen05 tensor cores (sm_100a), units = width in {32, 64}.
//
// perform_unsorted_update of the reference (code/utils/auxilary_classes.py:752-765): x = aggregated
// messages, h = old state.  Same math as gru_cell_kernel (gru.cu) with the two gate GEMMs as 3xTF32
// tcgen05.mma into a TMEM accumulator  D[128 rows, 4U cols] = [ z | r | xh | hh ]:
//     x chunks : D[:, 0:3U]   (+)= x . K            (N = 3U)
//     h chunks : D[:, 0:2U]    += h . R[:, z|r]     (N = 2U)
//                D[:, 3U:4U]  (+)= h . R[:, h]      (N = U)
// K and R do not fit in shared memory next to the operand tiles at U = 64, so their split / swizzled
// images are prepared once per call into a caller workspace (gru_cell_tc_prep) and streamed per
// 32-float K chunk by TMA bulk copies; two shared-memory stages and two TMEM accumulators, with the
// work split by role (loaders / MMA issuer / TMA producer / gate epilogue) so that the loads + MMAs of
// tile i+1 run under the epilogue of tile i.  History (10 M rows, U = 64): 5.31 ms in lock-step with
// thread 0 issuing in line -> 3.57 ms; a third weight slot made it slower (4.11 ms: less L1 left for the
// row-per-lane state loads / stores of the epilogue, which is what bounds it now).

#include <stdlib.h>

#include "tc_common.cuh"

using namespace ign_tc;

// -DIGN_CELL_PROFILE: per-phase clock64() sums of thread 0, printed after every launch.
#ifdef IGN_CELL_PROFILE
#include <stdio.h>
#include <string.h>
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

namespace {
PROF(__device__ unsigned long long cell_prof[8];)

constexpr int GROUP = 256;                    // threads of the loader group and of the epilogue group
constexpr int MMA_WARP = 16, TMA_WARP = 17;
constexpr int CELL_THREADS = 2 * GROUP + 64;
constexpr int ROWS = 128;
constexpr int A_IMG = ROWS * 128;

// image of chunk c (0..NC-1: x chunks from K, NC..2NC-1: h chunks from R): [3U rows][32] hi then lo
__global__ void gru_cell_tc_prep_kernel(const float* __restrict__ kernel, const float* __restrict__ rkernel, int U,
                                        float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = U * 3 * U;
  if (i >= 2 * per) return;
  const int which = i / per, j = i % per;
  const int k = j / (3 * U), n = j % (3 * U);
  const float v = which ? rkernel[j] : kernel[j];
  float hi, lo;
  tf32_split(v, hi, lo);
  const int nc = U / 32;
  const int c = which * nc + k / 32, kk = k % 32;
  char* base = reinterpret_cast<char*>(img) + (size_t)c * (2 * 3 * U * 128);
  *reinterpret_cast<float*>(base + sw128_off(n, kk)) = hi;
  *reinterpret_cast<float*>(base + 3 * U * 128 + sw128_off(n, kk)) = lo;
}

// Roles (warp-specialised, mbarriers only inside the tile loop):
//   warps 0-7    loaders: operand chunk (128 rows x 32 floats of x or h) -> hi / lo images of the stage,
//                arrive on full[stage]
//   warps 8-15   epilogue: gates + new state of tile i from accumulator i & 1 while the loaders and the
//                tensor core work on tile i + 1
//   warp 16      MMA issuer (one lane): waits full[stage] + the weight chunk, issues, commits to
//                stage[stage] and, after a tile's last chunk, to acc[ab]
//   warp 17      TMA producer (one lane): the weight chunk of a stage as soon as the stage is free
template <int U>
__global__ void __launch_bounds__(CELL_THREADS, 1) gru_cell_tc_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ h, int64_t n,
                                                                      const float* __restrict__ wimg,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ out) {
  constexpr int NC = U / 32;                 // K chunks per operand
  constexpr int B_IMG = 3 * U * 128;         // bytes of one weight image (hi or lo) of a chunk
  constexpr int STAGE = 2 * A_IMG + 2 * B_IMG;
  constexpr int DCOLS = 4 * U;               // accumulator columns
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_stage[2];          // the UMMAs that read the stage are done
  __shared__ uint64_t bar_full[2];           // operand images of the stage are in place (256 arrivals)
  __shared__ uint64_t bar_b[2];              // weight chunk landed (complete_tx)
  __shared__ uint64_t bar_acc[2];            // accumulator complete
  __shared__ uint64_t bar_drained[2];        // accumulator read by every epilogue thread (256 arrivals)
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];          // merged gate biases [bz | br | bxh | bhh]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_stage[i], 1); mbar_init(&bar_full[i], GROUP); mbar_init(&bar_b[i], 1);
      mbar_init(&bar_acc[i], 1); mbar_init(&bar_drained[i], GROUP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 2 * DCOLS);
  if (tid < U) {
    s_gb[tid] = bias[tid] + bias[3 * U + tid];
    s_gb[U + tid] = bias[U + tid] + bias[4 * U + tid];
    s_gb[2 * U + tid] = bias[2 * U + tid];
    s_gb[3 * U + tid] = bias[5 * U + tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t ntiles = (n + ROWS - 1) / ROWS;
  uint32_t stage_uses[2] = {0, 0}, acc_uses[2] = {0, 0}, ctr = 0;   // every role walks the same chunk sequence

  if (warp == MMA_WARP) {
    int ab = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ab ^= 1) {
      const uint32_t d = tmem_base + ab * DCOLS;
      for (int c = 0; c < 2 * NC; ++c, ++ctr) {
        const int s = ctr & 1;
        if (lane == 0) {
          unsigned char* st = smem + s * STAGE;
          if (c == 0 && acc_uses[ab] > 0) mbar_wait(&bar_drained[ab], (acc_uses[ab] - 1) & 1);
          mbar_wait(&bar_full[s], stage_uses[s] & 1);
          mbar_wait(&bar_b[s], stage_uses[s] & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + A_IMG, b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + B_IMG;
          if (c < NC) {
            umma_chunk_3x(d, a_hi, a_lo, b_hi, b_lo, 3 * U, c > 0);
          } else {
            umma_chunk_3x(d, a_hi, a_lo, b_hi, b_lo, 2 * U, true);
            umma_chunk_3x(d + 3 * U, a_hi, a_lo, b_hi + 2 * U * 128, b_lo + 2 * U * 128, U, c > NC);
          }
          umma_commit(&bar_stage[s]);
          if (c == 2 * NC - 1) umma_commit(&bar_acc[ab]);
        }
        __syncwarp();
        stage_uses[s] += 1;
      }
      acc_uses[ab] += 1;
    }
  } else if (warp == TMA_WARP) {
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = 0; c < 2 * NC; ++c, ++ctr) {
        const int s = ctr & 1;
        if (lane == 0) {
          if (stage_uses[s] > 0) mbar_wait(&bar_stage[s], (stage_uses[s] - 1) & 1);
          mbar_expect_tx(&bar_b[s], 2 * B_IMG);
          bulk_g2s(smem + s * STAGE + 2 * A_IMG, reinterpret_cast<const char*>(wimg) + (size_t)c * (2 * B_IMG),
                   2 * B_IMG, &bar_b[s]);
        }
        __syncwarp();
        stage_uses[s] += 1;
      }
    }
  } else if (warp < GROUP / 32) {
    // ---- loaders: every chunk is loaded into registers two chunks ahead of its use (32 KB per SM in
    // flight; with one chunk in flight the kernel was bound by bytes in flight at 1.5 TB/s)
    PROF(long long p_stage = 0, p_load = 0, c0, c1;)
    constexpr int PER = 1024 / GROUP;                      // float4 per thread per chunk
    float4 v[2][PER];                                      // chunk parity -> register set (2 NC is even)
    auto load_chunk = [&](float4 (&dst)[PER], int64_t tile, int c) {
      const int64_t m0 = tile * ROWS;
      const float* base = (c < NC) ? x : h;
      const int koff = (c < NC ? c : c - NC) * 32;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int idx = tid + j * GROUP;
        const int r = idx >> 3, c4 = idx & 7;
        dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile < ntiles && m0 + r < n) dst[j] = ldg_f4(base + (m0 + r) * U + koff + c4 * 4);
      }
    };
    load_chunk(v[0], blockIdx.x, 0);
    load_chunk(v[1], blockIdx.x, 1);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
      for (int c = 0; c < 2 * NC; ++c, ++ctr) {
        const int s = ctr & 1;
        unsigned char* st = smem + s * STAGE;
        PROF(c0 = clock64();)
        if (stage_uses[s] > 0) mbar_wait(&bar_stage[s], (stage_uses[s] - 1) & 1);
        PROF(c1 = clock64(); p_stage += c1 - c0;)
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int idx = tid + j * GROUP;
          store_split(st, st + A_IMG, idx >> 3, idx & 7, v[c & 1][j]);
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(&bar_full[s]);
        if (c + 2 < 2 * NC) load_chunk(v[c & 1], tile, c + 2);        // two chunks ahead
        else load_chunk(v[c & 1], tile + gridDim.x, c + 2 - 2 * NC);
        PROF(c0 = clock64(); p_load += c0 - c1;)
        stage_uses[s] += 1;
      }
    }
    PROF(if (tid == 32) { atomicAdd(&cell_prof[0], (unsigned long long)p_stage); atomicAdd(&cell_prof[1], (unsigned long long)p_load); })
  } else {
    // ---- epilogue: gates + new state, thread = one row x U / 2 units, 16 at a time
    PROF(long long p_acc = 0, p_gate = 0, c0, c1;)
    const int ew = warp - GROUP / 32;
    const int q = ew & 3, half = ew >> 2;
    int ab = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ab ^= 1) {
      const int64_t row = tile * ROWS + q * 32 + lane;
      const uint32_t tb = tmem_base + ab * DCOLS + ((uint32_t)(q * 32) << 16);
      float4 hold_all[U / 8];                              // old state of this thread's U / 2 units
#pragma unroll
      for (int j = 0; j < U / 8; ++j)
        hold_all[j] = row < n ? ldg_f4(h + row * U + half * (U / 2) + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      PROF(c0 = clock64();)
      mbar_wait(&bar_acc[ab], acc_uses[ab] & 1);
      tc_fence_after();
      PROF(c1 = clock64(); p_acc += c1 - c0;)
#pragma unroll
      for (int u0 = half * (U / 2); u0 < (half + 1) * (U / 2); u0 += 16) {
        uint32_t az[16], ar[16], axh[16], ahh[16];
        tmem_ld16_nowait(tb + u0, az);
        tmem_ld16_nowait(tb + U + u0, ar);
        tmem_ld16_nowait(tb + 2 * U + u0, axh);
        tmem_ld16_nowait(tb + 3 * U + u0, ahh);
        const float4* ho = hold_all + (u0 - half * (U / 2)) / 4;
        tmem_ld_wait();
        if (u0 + 16 >= (half + 1) * (U / 2)) {           // last read of this accumulator by this thread
          tc_fence_before();
          mbar_arrive(&bar_drained[ab]);
        }
        if (row < n) {
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 vz = *reinterpret_cast<const float4*>(s_gb + u0 + j4);
            const float4 vr = *reinterpret_cast<const float4*>(s_gb + U + u0 + j4);
            const float4 vx = *reinterpret_cast<const float4*>(s_gb + 2 * U + u0 + j4);
            const float4 vh = *reinterpret_cast<const float4*>(s_gb + 3 * U + u0 + j4);
            const float bz[4] = {vz.x, vz.y, vz.z, vz.w}, br[4] = {vr.x, vr.y, vr.z, vr.w};
            const float bxh[4] = {vx.x, vx.y, vx.z, vx.w}, bhh[4] = {vh.x, vh.y, vh.z, vh.w};
            const float hold[4] = {ho[j4 / 4].x, ho[j4 / 4].y, ho[j4 / 4].z, ho[j4 / 4].w};
            float hn[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 + jj;
              hn[jj] = fast_gru_gate(__uint_as_float(az[j]) + bz[jj], __uint_as_float(ar[j]) + br[jj],
                                     __uint_as_float(axh[j]) + bxh[jj], __uint_as_float(ahh[j]) + bhh[jj], hold[jj]);
            }
            st_f4(out + row * U + u0 + j4, make_float4(hn[0], hn[1], hn[2], hn[3]));
          }
        }
      }
      acc_uses[ab] += 1;
      PROF(c0 = clock64(); p_gate += c0 - c1;)
    }
    PROF(if (tid == GROUP + 32) { atomicAdd(&cell_prof[2], (unsigned long long)p_acc); atomicAdd(&cell_prof[3], (unsigned long long)p_gate); })
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * DCOLS);
}

}  // namespace

bool ign_gru_cell_tc_supported(int f_in, int units) { return f_in == units && (units == 32 || units == 64); }
size_t ign_gru_cell_tc_ws(int units) { return (size_t)2 * (units / 32) * 2 * 3 * units * 128; }

// split / swizzled weight images of one cell into the caller's workspace (also used by agg_gru_tc.cu)
int ign_gru_cell_tc_prep(const float* kernel, const float* rkernel, int units, void* ws, cudaStream_t st) {
  gru_cell_tc_prep_kernel<<<(unsigned)ign_cdiv(2 * units * 3 * units, 256), 256, 0, st>>>(
      kernel, rkernel, units, reinterpret_cast<float*>(ws));
  IGN_CHECK_LAUNCH("gru_cell_tc_prep");
  return IGN_OK;
}

int ign_gru_cell_tc_launch(const float* x, const float* h, int64_t n, int units, const float* kernel,
                           const float* rkernel, const float* bias, float* out, void* ws, cudaStream_t st) {
  float* img = reinterpret_cast<float*>(ws);
  int prc = ign_gru_cell_tc_prep(kernel, rkernel, units, ws, st);
  if (prc) return prc;
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(n, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  if (units == 64) {
    const size_t smem = 1024 + 2 * (size_t)(2 * A_IMG + 2 * 3 * 64 * 128);
    if (IGN_ONCE_PER_DEVICE()) {
      IGN_CUDA(cudaFuncSetAttribute(gru_cell_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    gru_cell_tc_kernel<64><<<grid, CELL_THREADS, smem, st>>>(x, h, n, img, bias, out);
  } else {
    const size_t smem = 1024 + 2 * (size_t)(2 * A_IMG + 2 * 3 * 32 * 128);
    if (IGN_ONCE_PER_DEVICE()) {
      IGN_CUDA(cudaFuncSetAttribute(gru_cell_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    gru_cell_tc_kernel<32><<<grid, CELL_THREADS, smem, st>>>(x, h, n, img, bias, out);
  }
  IGN_CHECK_LAUNCH("gru_cell_tc");
  PROF({
    unsigned long long hp[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(hp, cell_prof, sizeof(hp));
    const double nt = (double)tiles, nc = nt * 2 * (units / 32);
    fprintf(stderr, "gru_cell_tc<%d> loader warp 1 per chunk: wait stage %.0f split+store %.0f | epilogue warp 1 per tile: "
                    "wait acc %.0f gates %.0f\n", units, hp[0] / nc, hp[1] / nc, hp[2] / nt, hp[3] / nt);
    memset(hp, 0, sizeof(hp));
    cudaMemcpyToSymbol(cell_prof, hp, sizeof(hp));
  })
  return IGN_OK;
}
