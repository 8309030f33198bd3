// Ordered / interleave aggregation + GRU update on the tcgen05 tensor cores (sm_100a), 32-wide.
//
// Same contract as gru_seq_kernel (gru.cu): for every destination d walk its step list,
//   h <- GRU(x = message(step), h)        (reference code/utils/auxilary_classes.py:767-796, 421-440)
// but the two gate GEMMs of every step run as 3xTF32 tcgen05.mma with the accumulator in TMEM:
//
//   D[128 rows, 128 cols] = [ z | r | xh | hh ] pre-activations of 128 destinations
//     x chunk : D[:, 0:96]   = x_t  . K[:, z|r|h]          (N = 96)
//     h chunk : D[:, 0:64]  += h    . R[:, z|r]            (N = 64)
//               D[:, 96:128] = h    . R[:, h]              (N = 32)
//   each product as  A_hi B_hi + A_lo B_hi + A_hi B_lo  (hi = top 19 bits, lo = remainder), so the
//   result keeps fp32 accuracy (parity bar 1e-5).
//
// One CTA (256 threads, persistent) owns TWO tiles of 128 destinations at a time: while the tensor
// core computes step t of one tile, all 8 warps run the epilogue of the other (tcgen05.ld of the
// gates, sigmoid / tanh, new state, re-split of h and of the next gathered message into the
// swizzled shared-memory operand images).  K and R are split and laid out once per CTA.  The
// running state h stays in registers in full fp32 (thread = one destination x 16 units).

#include <stdlib.h>

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int NSPLIT = 4;                 // warps per TMEM lane group: each owns 32 / NSPLIT units of a row
constexpr int EPI_THREADS = 128 * NSPLIT;  // 16 epilogue warps
constexpr int TC_THREADS = EPI_THREADS + 32;   // + one warp that only issues the tcgen05.mma stream
constexpr int UPT = 32 / NSPLIT;          // units per thread
constexpr int ROWS = 128;                 // destinations per tile = UMMA M
constexpr int U = 32;                     // units = message width
constexpr int IMG = ROWS * 128;           // bytes of one [128 x 32] fp32 operand image
constexpr int BIMG = 96 * 128;            // bytes of one [96 x 32] weight image
constexpr int SLOT_BYTES = 4 * IMG;       // Ax_hi | Ax_lo | Ah_hi | Ah_lo

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};
__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}

struct Slot {
  int d;        // destination of this thread's row (-1: none)
  int lo;       // first step
  int len;      // number of steps
  int entry;    // step-table entry of the next message to gather (one step of lookahead)
  float h[UPT]; // running state, this thread's units
};

template <bool FAST>
__global__ void __launch_bounds__(TC_THREADS, 1) gru_seq_tc_kernel(
    const int* __restrict__ steps_rowptr, const int* __restrict__ steps, const int* __restrict__ order, SrcPtrs srcs,
    const float* __restrict__ h0, int64_t num_dst, const float* __restrict__ kernel, const float* __restrict__ rkernel,
    const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ h_seq,
    const int4* __restrict__ meta) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* bx_hi = smem;
  unsigned char* bx_lo = bx_hi + BIMG;
  unsigned char* bh_hi = bx_lo + BIMG;
  unsigned char* bh_lo = bh_hi + BIMG;
  unsigned char* slots = bh_lo + BIMG;                  // 4 * 12288 = 49152 = 48 * 1024: still 1024-aligned
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[6 * U];
  __shared__ __align__(16) float s_gb[4 * U];
  __shared__ int s_maxlen[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool mma_warp = warp == EPI_THREADS / 32;          // warp 16: MMA issuer, no rows
  const int q = warp & 3, split = (warp >> 2) & (NSPLIT - 1);
  const int row = q * 32 + lane;                         // TMEM lane == row of the tile
  const int u0 = split * UPT;                            // this thread's units [u0, u0 + UPT)

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // weight images: Bx[n][k] = K[k][n], Bh[n][k] = R[k][n], n in [0,96), hi / lo, swizzled
  for (int i = tid; i < U * 3 * U; i += TC_THREADS) {
    const int k = i / (3 * U), n = i % (3 * U);
    const int off = n * 128 + ((((k >> 2) ^ (n & 7)) & 7) << 4) + (k & 3) * 4;
    float hi, lo;
    tf32_split(__ldg(kernel + i), hi, lo);
    *reinterpret_cast<float*>(bx_hi + off) = hi;
    *reinterpret_cast<float*>(bx_lo + off) = lo;
    tf32_split(__ldg(rkernel + i), hi, lo);
    *reinterpret_cast<float*>(bh_hi + off) = hi;
    *reinterpret_cast<float*>(bh_lo + off) = lo;
  }
  for (int i = tid; i < 6 * U; i += TC_THREADS) s_bias[i] = bias[i];
  fence_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // merged gate biases in shared memory: [bz | br | bxh | bhh], read as broadcast float4
  __syncthreads();
  if (tid < U) {
    const float b0z = s_bias[tid], b0r = s_bias[U + tid], b0h = s_bias[2 * U + tid];
    const float b1z = s_bias[3 * U + tid], b1r = s_bias[4 * U + tid], b1h = s_bias[5 * U + tid];
    s_gb[tid] = b0z + b1z; s_gb[U + tid] = b0r + b1r; s_gb[2 * U + tid] = b0h; s_gb[3 * U + tid] = b1h;
  }
  __syncthreads();

  const int64_t ntiles = (num_dst + ROWS - 1) / ROWS;
  const int64_t npairs = (ntiles + 1) / 2;
  uint32_t uses[2] = {0, 0};

  // gather this thread's 64 bytes of the message of step t (its entry was fetched one step earlier,
  // so the row load does not wait on an index load), then fetch the entry of step t + 1
  auto load_x = [&](Slot& s, int t, float4 (&x)[UPT / 4]) {
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j) x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int entry = s.entry;
    s.entry = (t + 1 < s.len) ? __ldg(steps + s.lo + t + 1) : IGN_STEP_ZERO;
    if (t < s.len && entry >= 0) {
      const float* p = pick_src(srcs, entry >> IGN_STEP_SRC_SHIFT) + (int64_t)(entry & IGN_STEP_ROW_MASK) * U + u0;
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j) x[j] = ldg_f4(p + 4 * j);
    }
  };
  auto store_x = [&](int slot, const float4 (&x)[UPT / 4]) {
    unsigned char* b = slots + slot * SLOT_BYTES;
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j) store_split(b, b + IMG, row, split * (UPT / 4) + j, x[j]);
  };
  auto store_h = [&](int slot, const Slot& s) {
    unsigned char* b = slots + slot * SLOT_BYTES + 2 * IMG;
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j)
      store_split(b, b + IMG, row, split * (UPT / 4) + j,
                  make_float4(s.h[4 * j], s.h[4 * j + 1], s.h[4 * j + 2], s.h[4 * j + 3]));
  };
  auto issue_mma = [&](int slot) {           // one thread: 36 UMMAs of one step, then commit
    const uint32_t ax_hi = smem_u32(slots + slot * SLOT_BYTES), ax_lo = ax_hi + IMG;
    const uint32_t ah_hi = ax_lo + IMG, ah_lo = ah_hi + IMG;
    const uint32_t bxh_ = smem_u32(bx_hi), bxl_ = smem_u32(bx_lo), bhh_ = smem_u32(bh_hi), bhl_ = smem_u32(bh_lo);
    const uint32_t d = tmem_base + slot * 128;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t ko = kk * 32;
      umma_tf32(d, umma_desc(ax_hi + ko), umma_desc(bxh_ + ko), umma_idesc(96), kk > 0 ? 1u : 0u);
      umma_tf32(d, umma_desc(ax_lo + ko), umma_desc(bxh_ + ko), umma_idesc(96), 1u);
      umma_tf32(d, umma_desc(ax_hi + ko), umma_desc(bxl_ + ko), umma_idesc(96), 1u);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t ko = kk * 32;
      umma_tf32(d, umma_desc(ah_hi + ko), umma_desc(bhh_ + ko), umma_idesc(64), 1u);
      umma_tf32(d, umma_desc(ah_lo + ko), umma_desc(bhh_ + ko), umma_idesc(64), 1u);
      umma_tf32(d, umma_desc(ah_hi + ko), umma_desc(bhl_ + ko), umma_idesc(64), 1u);
      umma_tf32(d + 96, umma_desc(ah_hi + ko), umma_desc(bhh_ + 64 * 128 + ko), umma_idesc(32), kk > 0 ? 1u : 0u);
      umma_tf32(d + 96, umma_desc(ah_lo + ko), umma_desc(bhh_ + 64 * 128 + ko), umma_idesc(32), 1u);
      umma_tf32(d + 96, umma_desc(ah_hi + ko), umma_desc(bhl_ + 64 * 128 + ko), umma_idesc(32), 1u);
    }
    umma_commit(&bar[slot]);
  };

  // walk plan of this thread's rows in the NEXT pair of tiles (prefetched one pair ahead)
  int4 nmeta[2];
  auto fetch_meta = [&](int64_t pr) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int64_t didx = (pr * 2 + s) * ROWS + row;
      nmeta[s] = make_int4(-1, 0, 0, IGN_STEP_ZERO);
      if (!mma_warp && pr < npairs && didx < num_dst) {
        if (meta) {
          nmeta[s] = __ldg(meta + didx);
        } else {
          const int d = order ? __ldg(order + didx) : (int)didx;
          const int lo = __ldg(steps_rowptr + d), len = __ldg(steps_rowptr + d + 1) - lo;
          nmeta[s] = make_int4(d, lo, len, len > 0 ? __ldg(steps + lo) : IGN_STEP_ZERO);
        }
      }
    }
  };
  fetch_meta(blockIdx.x);

  for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    Slot sl[2];
    if (tid < 2) s_maxlen[tid] = 0;
    __syncthreads();
    // ---- tile set-up: meta, h0 -> registers + image, x_0 -> image
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      sl[s].d = nmeta[s].x; sl[s].lo = nmeta[s].y; sl[s].len = nmeta[s].z; sl[s].entry = nmeta[s].w;
#pragma unroll
      for (int j = 0; j < UPT; ++j) sl[s].h[j] = 0.f;
      if (sl[s].d >= 0) {
#pragma unroll
        for (int j = 0; j < UPT / 4; ++j) {
          const float4 v = ldg_f4(h0 + (int64_t)sl[s].d * U + u0 + 4 * j);
          sl[s].h[4 * j] = v.x; sl[s].h[4 * j + 1] = v.y; sl[s].h[4 * j + 2] = v.z; sl[s].h[4 * j + 3] = v.w;
        }
      }
      if (split == 0 && sl[s].len > 0) atomicMax(&s_maxlen[s], sl[s].len);
      if (!mma_warp) {
        float4 x[UPT / 4];
        load_x(sl[s], 0, x);
        store_x(s, x);
        store_h(s, sl[s]);
      }
    }
    fetch_meta(pair + gridDim.x);                        // lands while this pair is being walked
    fence_async_smem();
    __syncthreads();
    const int maxlen0 = s_maxlen[0], maxlen1 = s_maxlen[1];
    const int maxlen = max(maxlen0, maxlen1);
    if (mma_warp) {
      // ---- MMA issuer warp: one elected lane feeds the tensor core; the epilogue warps never wait on it
      if (lane == 0) {
        if (maxlen0 > 0) issue_mma(0);
        if (maxlen1 > 0) issue_mma(1);
      }
      for (int t = 0; t < maxlen; ++t) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int ml = s == 0 ? maxlen0 : maxlen1;
          if (t + 1 >= ml) continue;
          // operands of step t + 1 of slot s are in shared memory once all epilogue threads arrived
          asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "r"(TC_THREADS) : "memory");
          if (lane == 0) issue_mma(s);
        }
      }
      continue;                                          // next pair (joins the set-up barriers)
    }
    if (maxlen0 > 0) uses[0] += 1;
    if (maxlen1 > 0) uses[1] += 1;

    for (int t = 0; t < maxlen; ++t) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int ml = s == 0 ? maxlen0 : maxlen1;
        if (t >= ml) continue;                           // uniform over the CTA
        Slot& S = sl[s];
        float4 xn[UPT / 4];
        load_x(S, t + 1, xn);                            // next message: in flight while we wait for the MMA
        mbar_wait(&bar[s], (uses[s] - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tbase = tmem_base + s * 128 + ((uint32_t)(q * 32) << 16) + u0;
        uint32_t az[UPT], ar[UPT], axh[UPT], ahh[UPT];
        tmem_ld8_nowait(tbase, az);
        tmem_ld8_nowait(tbase + 32, ar);
        tmem_ld8_nowait(tbase + 64, axh);
        tmem_ld8_nowait(tbase + 96, ahh);
        tmem_ld_wait();
        if (t < S.len) {
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
              const float pz = __uint_as_float(az[j]) + bz[jj], pr = __uint_as_float(ar[j]) + br[jj];
              const float z = FAST ? fast_sigmoid(pz) : sigmoid_f(pz);
              const float r = FAST ? fast_sigmoid(pr) : sigmoid_f(pr);
              const float ph = fmaf(r, __uint_as_float(ahh[j]) + bhh[jj], __uint_as_float(axh[j]) + bxh[jj]);
              const float hh = FAST ? fast_tanh(ph) : tanhf(ph);
              S.h[j] = fmaf(z, S.h[j] - hh, hh);
            }
          }
          if (h_seq) {
            float* p = h_seq + (int64_t)(S.lo + t) * U + u0;
#pragma unroll
            for (int j = 0; j < UPT / 4; ++j)
              st_f4(p + 4 * j, make_float4(S.h[4 * j], S.h[4 * j + 1], S.h[4 * j + 2], S.h[4 * j + 3]));
          }
        }
        if (t + 1 < ml) {                                // operands of the next step
          store_h(s, S);
          store_x(s, xn);
          fence_async_smem();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (t + 1 < ml) {                                // hand the slot to the MMA warp, do not wait
          asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "r"(TC_THREADS) : "memory");
          uses[s] += 1;
        }
      }
    }
    // ---- results
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (sl[s].d >= 0) {
        float* p = out + (int64_t)sl[s].d * U + u0;
#pragma unroll
        for (int j = 0; j < UPT / 4; ++j)
          st_f4(p + 4 * j, make_float4(sl[s].h[4 * j], sl[s].h[4 * j + 1], sl[s].h[4 * j + 2], sl[s].h[4 * j + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

}  // namespace

int ign_gru_seq_tc_launch(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                          const float* const* srcs, const float* h0, int64_t num_dst, const float* kernel,
                          const float* rkernel, const float* bias, float* out, float* h_seq, const int* meta,
                          cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  const size_t smem = 1024 + 4 * (size_t)BIMG + 2 * (size_t)SLOT_BYTES;
  static thread_local bool configured = false;
  static const bool fast = getenv("IGN_GRU_TC_EXACT_MATH") == nullptr;   // default: ex2.approx-based sigmoid / tanh
  if (!configured) {
    IGN_CUDA(cudaFuncSetAttribute(gru_seq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IGN_CUDA(cudaFuncSetAttribute(gru_seq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t npairs = (ign_cdiv(num_dst, ROWS) + 1) / 2;
  const int grid = (int)(npairs < sms ? npairs : sms);
  if (fast)
    gru_seq_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(steps_rowptr, steps, order, sp, h0, num_dst, kernel,
                                                             rkernel, bias, out, h_seq,
                                                             reinterpret_cast<const int4*>(meta));
  else
    gru_seq_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(steps_rowptr, steps, order, sp, h0, num_dst, kernel,
                                                              rkernel, bias, out, h_seq,
                                                              reinterpret_cast<const int4*>(meta));
  IGN_CHECK_LAUNCH("gru_seq_tc");
  return IGN_OK;
}
