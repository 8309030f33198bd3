// Step-synchronous ordered / interleave aggregation on the tcgen05 tensor cores (sm_100a), 32-wide.
//
// Same result as ign_gru_seq (reference code/utils/auxilary_classes.py:767-796, 421-440), organised
// the other way round: instead of one tile walking its sequences for len steps (a latency chain per
// tile), step t of ALL destinations that still have a step t is one streaming launch.  Destinations
// are sorted by descending length, so the rows alive at step t are the prefix [0, nt[t]) of the
// sorted order; the running state lives in a sorted-order buffer hs[num_dst, 32] between launches
// (HBM round trip of 256 B per row-step: cheap on B200) and the final state is scattered to out[d]
// by the launch that executes a destination's last step.
//
// Per 128-row tile: gather x_t rows by the step-major table steps_T, read h rows, split both into the
// swizzled hi / lo operand images, 36 tcgen05.mma (3xTF32; K, R images resident in shared memory),
// gate math from the TMEM accumulator.  Tiles are independent, so the kernel is a plain software
// pipeline: row loads of tile i+2 and MMAs of tile i+1 are in flight while tile i's gates are computed
// (two operand stages, two TMEM accumulators, 16 warps).

#include <type_traits>

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int NPART = 4;
constexpr int TC_THREADS = 128 * NPART;
constexpr int ROWS = 128;
constexpr int U = 32;
constexpr int UPT = U / NPART;            // 8 units (two 16-byte chunks) per thread
constexpr int IMG = ROWS * 128;
constexpr int BIMG = 96 * 128;
constexpr int STAGE = 4 * IMG;            // Ax_hi | Ax_lo | Ah_hi | Ah_lo

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};
__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}

struct RowIdx {     // what a thread needs to know about its row of a tile
  int d;            // destination (-1: row not alive in this launch)
  int lo, len;      // first step, number of steps
  int entry;        // step-table entry of this step
};
struct RowData {
  float4 x[UPT / 4];
  float4 h[UPT / 4];
};

__global__ void __launch_bounds__(TC_THREADS, 1) gru_step_tc_kernel(
    int t, const int* __restrict__ nt, const int* __restrict__ off, int64_t num_dst, const int4* __restrict__ meta,
    const int* __restrict__ steps_T, SrcPtrs srcs, const float* __restrict__ h0, float* __restrict__ hs,
    float* __restrict__ out, float* __restrict__ h_seq, const float* __restrict__ kernel,
    const float* __restrict__ rkernel, const float* __restrict__ bias) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* bx_hi = smem;
  unsigned char* bx_lo = bx_hi + BIMG;
  unsigned char* bh_hi = bx_lo + BIMG;
  unsigned char* bh_lo = bh_hi + BIMG;
  unsigned char* stages = bh_lo + BIMG;
  __shared__ uint64_t bar_acc[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, part = warp >> 2;
  const int row = q * 32 + lane;
  const int u0 = part * UPT;

  if (tid == 0) {
    mbar_init(&bar_acc[0], 1);
    mbar_init(&bar_acc[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < U * 3 * U; i += TC_THREADS) {
    const int k = i / (3 * U), n = i % (3 * U);
    const int o = sw128_off(n, k);
    float hi, lo;
    tf32_split(__ldg(kernel + i), hi, lo);
    *reinterpret_cast<float*>(bx_hi + o) = hi;
    *reinterpret_cast<float*>(bx_lo + o) = lo;
    tf32_split(__ldg(rkernel + i), hi, lo);
    *reinterpret_cast<float*>(bh_hi + o) = hi;
    *reinterpret_cast<float*>(bh_lo + o) = lo;
  }
  if (tid < U) {
    s_gb[tid] = bias[tid] + bias[3 * U + tid];
    s_gb[U + tid] = bias[U + tid] + bias[4 * U + tid];
    s_gb[2 * U + tid] = bias[2 * U + tid];
    s_gb[3 * U + tid] = bias[5 * U + tid];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // rows of this launch: step 0 covers every destination (those with no step just copy their state)
  const int64_t n_rows = (t == 0) ? num_dst : (int64_t)__ldg(nt + t);
  const int64_t n_alive = (int64_t)__ldg(nt + t);
  const int* entries = steps_T + __ldg(off + t);
  const int64_t ntiles = (n_rows + ROWS - 1) / ROWS;
  const int G = gridDim.x;

  auto load_idx = [&](int64_t tile, RowIdx& ri) {
    const int64_t i = tile * ROWS + row;
    ri.d = -1; ri.lo = 0; ri.len = 0; ri.entry = IGN_STEP_ZERO;
    if (tile < ntiles && i < n_rows) {
      const int4 m = __ldg(meta + i);
      ri.d = m.x; ri.lo = m.y; ri.len = m.z;
      if (i < n_alive) ri.entry = __ldg(entries + i);
    }
  };
  // ---- pipeline state
  RowIdx idx_next;            // indices of the tile whose rows are loaded next
  RowIdx idx_rows;            // indices of the tile whose rows are in `rows`
  RowData rows;               // prefetched rows (tile to be produced next)
  RowIdx idx_acc0, idx_acc1;  // indices of the tiles sitting in the two accumulators
  float hkeep0[UPT], hkeep1[UPT];   // their old state (this thread's units), exact fp32
  uint32_t acc_uses0 = 0, acc_uses1 = 0;

  auto fetch_rows = [&](int64_t tile, const RowIdx& ri) {
    const int64_t i = tile * ROWS + row;
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j) { rows.x[j] = make_float4(0.f, 0.f, 0.f, 0.f); rows.h[j] = rows.x[j]; }
    if (ri.d >= 0) {
      const float* hp = (t == 0) ? h0 + (int64_t)ri.d * U + u0 : hs + i * U + u0;
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j) rows.h[j] = ldg_f4(hp + 4 * j);
      if (ri.entry >= 0) {
        const float* xp = pick_src(srcs, ri.entry >> IGN_STEP_SRC_SHIFT) + (int64_t)(ri.entry & IGN_STEP_ROW_MASK) * U + u0;
#pragma unroll
        for (int j = 0; j < UPT / 4; ++j) rows.x[j] = ldg_f4(xp + 4 * j);
      }
    }
  };

  // store the prefetched rows as operand images of stage `ab` and launch the tile's MMAs
  auto produce = [&](auto abc) {
    constexpr int ab = decltype(abc)::value;
    float (&hkeep)[UPT] = *(ab == 0 ? &hkeep0 : &hkeep1);
    unsigned char* st = stages + ab * STAGE;
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j) {
      store_split(st, st + IMG, row, part * (UPT / 4) + j, rows.x[j]);
      store_split(st + 2 * IMG, st + 3 * IMG, row, part * (UPT / 4) + j, rows.h[j]);
      hkeep[4 * j] = rows.h[j].x; hkeep[4 * j + 1] = rows.h[j].y;
      hkeep[4 * j + 2] = rows.h[j].z; hkeep[4 * j + 3] = rows.h[j].w;
    }
    if (ab == 0) idx_acc0 = idx_rows; else idx_acc1 = idx_rows;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ax_hi = smem_u32(st), ax_lo = ax_hi + IMG, ah_hi = ax_lo + IMG, ah_lo = ah_hi + IMG;
      const uint32_t d = tmem_base + ab * 128;
      umma_chunk_3x(d, ax_hi, ax_lo, smem_u32(bx_hi), smem_u32(bx_lo), 96, false);           // z | r | xh  (x part)
      umma_chunk_3x(d, ah_hi, ah_lo, smem_u32(bh_hi), smem_u32(bh_lo), 64, true);            // z | r      (h part)
      umma_chunk_3x(d + 96, ah_hi, ah_lo, smem_u32(bh_hi) + 64 * 128, smem_u32(bh_lo) + 64 * 128, 32, false);   // hh
      umma_commit(&bar_acc[ab]);
    }
    if (ab == 0) acc_uses0 += 1; else acc_uses1 += 1;
  };

  auto consume = [&](int64_t tile, auto abc) {
    constexpr int ab = decltype(abc)::value;
    const float (&hkeep)[UPT] = *(ab == 0 ? &hkeep0 : &hkeep1);
    mbar_wait(&bar_acc[ab], ((ab == 0 ? acc_uses0 : acc_uses1) - 1) & 1);
    tc_fence_after();
    const RowIdx ri = ab == 0 ? idx_acc0 : idx_acc1;
    const uint32_t tb = tmem_base + ab * 128 + ((uint32_t)(q * 32) << 16) + u0;
    uint32_t az[UPT], ar[UPT], axh[UPT], ahh[UPT];
    tmem_ld8_nowait(tb, az);
    tmem_ld8_nowait(tb + 32, ar);
    tmem_ld8_nowait(tb + 64, axh);
    tmem_ld8_nowait(tb + 96, ahh);
    tmem_ld_wait();
    tc_fence_before();
    if (ri.d < 0) return;
    const int64_t i = tile * ROWS + row;
    float hn[UPT];
    const bool alive = ri.len > t;
#pragma unroll
    for (int j4 = 0; j4 < UPT; j4 += 4) {
      const float4 vz = *reinterpret_cast<const float4*>(s_gb + u0 + j4);
      const float4 vr = *reinterpret_cast<const float4*>(s_gb + U + u0 + j4);
      const float4 vx = *reinterpret_cast<const float4*>(s_gb + 2 * U + u0 + j4);
      const float4 vh = *reinterpret_cast<const float4*>(s_gb + 3 * U + u0 + j4);
      const float bz[4] = {vz.x, vz.y, vz.z, vz.w}, br[4] = {vr.x, vr.y, vr.z, vr.w};
      const float bxh[4] = {vx.x, vx.y, vx.z, vx.w}, bhh[4] = {vh.x, vh.y, vh.z, vh.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j4 + jj;
        const float hold = hkeep[j];
        const float z = fast_sigmoid(__uint_as_float(az[j]) + bz[jj]);
        const float r = fast_sigmoid(__uint_as_float(ar[j]) + br[jj]);
        const float hh = fast_tanh(fmaf(r, __uint_as_float(ahh[j]) + bhh[jj], __uint_as_float(axh[j]) + bxh[jj]));
        hn[j] = alive ? fmaf(z, hold - hh, hh) : hold;
      }
    }
    const bool last = ri.len <= t + 1;              // this launch produces the destination's final state
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j) {
      const float4 v = make_float4(hn[4 * j], hn[4 * j + 1], hn[4 * j + 2], hn[4 * j + 3]);
      if (!last) st_f4(hs + i * U + u0 + 4 * j, v);
      else st_f4(out + (int64_t)ri.d * U + u0 + 4 * j, v);
      if (h_seq && alive) st_f4(h_seq + (int64_t)(ri.lo + t) * U + u0 + 4 * j, v);
    }
  };

  // ---- software pipeline over this CTA's tiles: blockIdx.x, +G, +2G, ...
  int64_t tile = blockIdx.x;
  if (tile < ntiles) {
    load_idx(tile, idx_rows);
    fetch_rows(tile, idx_rows);
    load_idx(tile + G, idx_next);
    produce(std::integral_constant<int, 0>{});         // tile 0 -> accumulator 0
    idx_rows = idx_next;
    fetch_rows(tile + G, idx_rows);                    // rows of tile 1 in flight
    load_idx(tile + 2 * (int64_t)G, idx_next);
  }
  auto body = [&](auto abc) {                          // accumulator index is a compile-time constant
    constexpr int ab = decltype(abc)::value;
    const int64_t next = tile + G;
    if (next < ntiles) {
      produce(std::integral_constant<int, ab ^ 1>{});  // tile i+1: images + MMAs (its rows were prefetched)
      idx_rows = idx_next;
      fetch_rows(next + G, idx_rows);                  // rows of tile i+2 in flight during the gate math
      load_idx(next + 2 * (int64_t)G, idx_next);
    }
    consume(tile, abc);
    __syncthreads();                                   // accumulator / stage `ab` free again
    tile = next;
  };
  while (tile < ntiles) {
    body(std::integral_constant<int, 0>{});
    if (tile >= ntiles) break;
    body(std::integral_constant<int, 1>{});
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

int ign_gru_step_tc_launch(int t, const int* nt, const int* off, int64_t num_dst, int64_t rows_bound, const int* meta,
                           const int* steps_T, int n_src, const float* const* srcs, const float* h0, float* hs,
                           float* out, float* h_seq, const float* kernel, const float* rkernel, const float* bias,
                           cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  const size_t smem = 1024 + 4 * (size_t)BIMG + 2 * (size_t)STAGE;
  static thread_local bool configured = false;
  if (!configured) {
    IGN_CUDA(cudaFuncSetAttribute(gru_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(rows_bound > 0 ? rows_bound : 1, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  gru_step_tc_kernel<<<grid, TC_THREADS, smem, st>>>(t, nt, off, num_dst, reinterpret_cast<const int4*>(meta), steps_T,
                                                      sp, h0, hs, out, h_seq, kernel, rkernel, bias);
  IGN_CHECK_LAUNCH("gru_step_tc");
  return IGN_OK;
}
