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
// 32-float K chunk with cp.async; two shared-memory stages and two TMEM accumulators let the
// loads + MMAs of tile i+1 run under the epilogue (gates, new state) of tile i.

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int NPART = 4;                      // warps per TMEM lane group
constexpr int TC_THREADS = 128 * NPART;       // 16 warps: the gate epilogue is latency-bound with fewer
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

template <int U>
__global__ void __launch_bounds__(TC_THREADS, 1) gru_cell_tc_kernel(const float* __restrict__ x,
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
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_acc[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];          // merged gate biases [bz | br | bxh | bhh]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_stage[0], 1); mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_acc[0], 1); mbar_init(&bar_acc[1], 1);
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
  uint32_t stage_uses[2] = {0, 0}, acc_uses[2] = {0, 0};
  uint32_t chunk_ctr = 0;                                 // global chunk counter -> stage = ctr & 1

  // loads + MMAs of one tile into accumulator `ab`
  auto produce = [&](int64_t tile, int ab) {
    const int64_t m0 = tile * ROWS;
    const uint32_t d = tmem_base + ab * DCOLS;
#pragma unroll 1
    for (int c = 0; c < 2 * NC; ++c) {
      const int s = chunk_ctr & 1;
      unsigned char* st = smem + s * STAGE;
      if (stage_uses[s] > 0) mbar_wait(&bar_stage[s], (stage_uses[s] - 1) & 1);
      {   // weight chunk image (hi + lo): straight copy
        const char* src = reinterpret_cast<const char*>(wimg) + (size_t)c * (2 * B_IMG);
        unsigned char* dst = st + 2 * A_IMG;
        for (int i = tid * 16; i < 2 * B_IMG; i += TC_THREADS * 16) cp_async16(dst + i, src + i);
        cp_async_commit();
      }
      {   // operand chunk: 128 rows x 32 floats of x (c < NC) or h
        const float* base = (c < NC) ? x : h;
        const int koff = (c < NC ? c : c - NC) * 32;
#pragma unroll
        for (int j = 0; j < 1024 / TC_THREADS; ++j) {
          const int idx = tid + j * TC_THREADS;
          const int r = idx >> 3, c4 = idx & 7;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r < n) v = ldg_f4(base + (m0 + r) * U + koff + c4 * 4);
          store_split(st, st + A_IMG, r, c4, v);
        }
      }
      cp_async_wait<0>();
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
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
      stage_uses[s] += 1;
      chunk_ctr += 1;
    }
    acc_uses[ab] += 1;
  };

  // gates + new state of one tile from accumulator `ab`
  auto consume = [&](int64_t tile, int ab) {
    const int64_t m0 = tile * ROWS;
    mbar_wait(&bar_acc[ab], (acc_uses[ab] - 1) & 1);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;
    const int64_t row = m0 + q * 32 + lane;
    const uint32_t tb = tmem_base + ab * DCOLS + ((uint32_t)(q * 32) << 16);
    constexpr int UPT = U / NPART;                       // units per thread: 16 (U = 64) or 8 (U = 32)
    const int u0 = part * UPT;
    uint32_t az[UPT], ar[UPT], axh[UPT], ahh[UPT];
    if constexpr (UPT == 16) {
      tmem_ld16_nowait(tb + u0, az);
      tmem_ld16_nowait(tb + U + u0, ar);
      tmem_ld16_nowait(tb + 2 * U + u0, axh);
      tmem_ld16_nowait(tb + 3 * U + u0, ahh);
    } else {
      tmem_ld8_nowait(tb + u0, az);
      tmem_ld8_nowait(tb + U + u0, ar);
      tmem_ld8_nowait(tb + 2 * U + u0, axh);
      tmem_ld8_nowait(tb + 3 * U + u0, ahh);
    }
    tmem_ld_wait();
    if (row < n) {
#pragma unroll
      for (int j4 = 0; j4 < UPT; j4 += 4) {
        const float4 vz = *reinterpret_cast<const float4*>(s_gb + u0 + j4);
        const float4 vr = *reinterpret_cast<const float4*>(s_gb + U + u0 + j4);
        const float4 vx = *reinterpret_cast<const float4*>(s_gb + 2 * U + u0 + j4);
        const float4 vh = *reinterpret_cast<const float4*>(s_gb + 3 * U + u0 + j4);
        const float4 ho = ldg_f4(h + row * U + u0 + j4);
        const float bz[4] = {vz.x, vz.y, vz.z, vz.w}, br[4] = {vr.x, vr.y, vr.z, vr.w};
        const float bxh[4] = {vx.x, vx.y, vx.z, vx.w}, bhh[4] = {vh.x, vh.y, vh.z, vh.w};
        const float hold[4] = {ho.x, ho.y, ho.z, ho.w};
        float hn[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = j4 + jj;
          const float z = fast_sigmoid(__uint_as_float(az[j]) + bz[jj]);
          const float r = fast_sigmoid(__uint_as_float(ar[j]) + br[jj]);
          const float hh = fast_tanh(fmaf(r, __uint_as_float(ahh[j]) + bhh[jj], __uint_as_float(axh[j]) + bxh[jj]));
          hn[jj] = fmaf(z, hold[jj] - hh, hh);
        }
        st_f4(out + row * U + u0 + j4, make_float4(hn[0], hn[1], hn[2], hn[3]));
      }
    }
    tc_fence_before();
  };

  // software pipeline: MMAs of tile i+1 are in flight while the gates of tile i are computed
  int64_t tile = blockIdx.x;
  int ab = 0;
  if (tile < ntiles) produce(tile, ab);
  while (tile < ntiles) {
    const int64_t next = tile + gridDim.x;
    if (next < ntiles) produce(next, ab ^ 1);
    consume(tile, ab);
    __syncthreads();                 // accumulator `ab` fully read before it is produced again
    tile = next;
    ab ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * DCOLS);
}

}  // namespace

bool ign_gru_cell_tc_supported(int f_in, int units) { return f_in == units && (units == 32 || units == 64); }
size_t ign_gru_cell_tc_ws(int units) { return (size_t)2 * (units / 32) * 2 * 3 * units * 128; }

int ign_gru_cell_tc_launch(const float* x, const float* h, int64_t n, int units, const float* kernel,
                           const float* rkernel, const float* bias, float* out, void* ws, cudaStream_t st) {
  float* img = reinterpret_cast<float*>(ws);
  gru_cell_tc_prep_kernel<<<(unsigned)ign_cdiv(2 * units * 3 * units, 256), 256, 0, st>>>(kernel, rkernel, units, img);
  IGN_CHECK_LAUNCH("gru_cell_tc_prep");
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(n, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  static thread_local bool configured[2] = {false, false};
  if (units == 64) {
    const size_t smem = 1024 + 2 * (size_t)(2 * A_IMG + 2 * 3 * 64 * 128);
    if (!configured[1]) {
      IGN_CUDA(cudaFuncSetAttribute(gru_cell_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[1] = true;
    }
    gru_cell_tc_kernel<64><<<grid, TC_THREADS, smem, st>>>(x, h, n, img, bias, out);
  } else {
    const size_t smem = 1024 + 2 * (size_t)(2 * A_IMG + 2 * 3 * 32 * 128);
    if (!configured[0]) {
      IGN_CUDA(cudaFuncSetAttribute(gru_cell_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[0] = true;
    }
    gru_cell_tc_kernel<32><<<grid, TC_THREADS, smem, st>>>(x, h, n, img, bias, out);
  }
  IGN_CHECK_LAUNCH("gru_cell_tc");
  return IGN_OK;
}
