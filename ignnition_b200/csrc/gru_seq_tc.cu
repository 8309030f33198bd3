// Ordered / interleave aggregation + GRU update on the tcgen05 tensor cores (sm_100a), 32-wide.
//
// Same contract as gru_seq_kernel (gru.cu): for every destination d walk its step list,
//   h <- GRU(x = message(step), h)        (reference code/utils/auxilary_classes.py:767-796, 421-440)
// but the two gate GEMMs of every step run as 3xTF32 tcgen05.mma with the accumulator in TMEM:
//
//   D[128 rows, 128 cols] = [ z | r | xh | hh ] pre-activations of 128 destinations
//     D  = x_t . [ Kz | Kr | Kh | 0  ]         (N = 128, 12 UMMAs)
//     D += h   . [ Rz | Rr | 0  | Rh ]         (N = 128, 12 UMMAs; a UMMA of N <= 128 costs the same
//                                               71 cycles whatever N is, profiles/r1_umma_tf32_rate.md)
//   each product as  A_hi B_hi + A_lo B_hi + A_hi B_lo  (hi = rna_tf32(v), lo = rna_tf32(v - hi)), so
//   the result keeps fp32 accuracy (parity bar 1e-5).
//
// One persistent CTA per SM (512 threads) runs TWO independent walkers of 8 warps each.  A walker
// owns one tile of 128 destinations at a time: it waits for the tensor core (mbarrier armed by
// tcgen05.commit), reads the gates from TMEM, applies sigmoid / tanh and the state update (h stays in
// registers in full fp32; thread = one destination x 16 units), re-splits h and the next gathered
// message into the swizzled operand images, meets on its own named barrier and lets one thread issue
// the 24 UMMAs of the next step.  The walkers never synchronise with each other: while one sets up
// its next tile or runs its epilogue, the tensor core works for the other.  K and R are split and
// laid out once per CTA.

#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "tc_common.cuh"

using namespace ign_tc;

// -DIGN_WALK_PROFILE: per-phase clock64() sums of one warp per walker, printed after every launch
// (profiles/r1_walk_phases.md).  Compiles to nothing otherwise.
#ifdef IGN_WALK_PROFILE
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

namespace {
PROF(__device__ unsigned long long prof_cyc[12];)

constexpr int WALKERS = 2;
constexpr int WALKER_THREADS = 256;                       // 8 warps: 4 TMEM lane groups x 2 unit halves
constexpr int TC_THREADS = WALKERS * WALKER_THREADS;
constexpr int UPT = 16;                                   // units per thread
constexpr int ROWS = 128;                                 // destinations per tile = UMMA M
constexpr int U = 32;                                     // units = message width
constexpr int IMG = ROWS * 128;                           // bytes of one [128 x 32] fp32 operand image
constexpr int SLOT_BYTES = 4 * IMG;                       // Ax_hi | Ax_lo | Ah_hi | Ah_lo

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};
__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}

template <bool FAST>
__global__ void __launch_bounds__(TC_THREADS, 1) gru_seq_tc_kernel(
    const int* __restrict__ steps_rowptr, const int* __restrict__ steps, const int* __restrict__ order, SrcPtrs srcs,
    const float* __restrict__ h0, int64_t num_dst, const float* __restrict__ kernel, const float* __restrict__ rkernel,
    const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ h_seq,
    const int4* __restrict__ meta PROF(, int dbg)) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* bx_hi = smem;                          // [128 n][32 k] images of [Kz|Kr|Kh|0] and [Rz|Rr|0|Rh]
  unsigned char* bx_lo = bx_hi + IMG;
  unsigned char* bh_hi = bx_lo + IMG;
  unsigned char* bh_lo = bh_hi + IMG;
  unsigned char* slots = bh_lo + IMG;
  __shared__ uint64_t bar_acc[WALKERS];                 // tensor core -> walker: gates are in TMEM
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];
  __shared__ int s_maxlen[WALKERS][2];                  // per tile parity: reset one tile ahead of its use
  __shared__ __align__(16) int4 s_meta[WALKERS][3][ROWS];  // ring of walk plans: this tile, the next, the one after

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int g = 0; g < WALKERS; ++g) mbar_init(&bar_acc[g], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256u);
  // weight images, n = output column of D, k = input unit
  for (int i = tid; i < 128 * U; i += TC_THREADS) {
    const int n = i >> 5, k = i & 31;
    const int off = sw128_off(n, k);
    const float kx = n < 96 ? __ldg(kernel + k * 96 + n) : 0.f;
    const float kh = n < 64 ? __ldg(rkernel + k * 96 + n) : n < 96 ? 0.f : __ldg(rkernel + k * 96 + n - 32);
    float hi, lo;
    tf32_split(kx, hi, lo);
    *reinterpret_cast<float*>(bx_hi + off) = hi;
    *reinterpret_cast<float*>(bx_lo + off) = lo;
    tf32_split(kh, hi, lo);
    *reinterpret_cast<float*>(bh_hi + off) = hi;
    *reinterpret_cast<float*>(bh_lo + off) = lo;
  }
  // merged gate biases [bz | br | bxh | bhh] (Keras bias[2, 3u]: input row, recurrent row)
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
  const int64_t ntiles = (num_dst + ROWS - 1) / ROWS;

  {
    // ================================ walkers ================================
    const int g = warp >> 3;                               // walker of this thread
    const int gw = warp & 7, gtid = tid & (WALKER_THREADS - 1);
    // owner mapping (epilogue): thread = one destination (TMEM lane) x 16 units
    const int q = gw & 3, split = gw >> 2;
    const int row = q * 32 + lane;
    const int u0 = split * UPT;
    // loader mapping (global <-> shared): 8 lanes cover the 128 bytes of one row, a warp instruction
    // touches 4 whole rows (4 L1 wavefronts instead of 32 with one row per lane)
    const int lc = lane & 7;                               // 16-byte chunk of the row
    const int lr0 = gw * 16 + (lane >> 3);                 // rows lr0 + 4 i, i = 0..3
    unsigned char* slot = slots + g * SLOT_BYTES;
    const uint32_t tbase = tmem_base + g * 128 + ((uint32_t)(q * 32) << 16) + u0;
    const int64_t tile_stride = (int64_t)WALKERS * gridDim.x;
    uint32_t acc_phase = 0;

    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(WALKER_THREADS) : "memory"); };
    // walk plan of this thread's row in a later tile -> ring slot, asynchronously (no registers held
    // across the walk); rows past the end get an empty plan
    auto plan_to_smem = [&](int64_t tile, int ring) {
      if (split != 0) return;
      int4* dst = &s_meta[g][ring][row];
      const int64_t didx = tile * ROWS + row;
      if (tile < ntiles && didx < num_dst) {
        if (meta) {
          cp_async16(dst, meta + didx);
        } else {
          const int d = order ? __ldg(order + didx) : (int)didx;
          const int lo = __ldg(steps_rowptr + d), len = __ldg(steps_rowptr + d + 1) - lo;
          *dst = make_int4(d, lo, len, len > 0 ? __ldg(steps + lo) : IGN_STEP_ZERO);
        }
      } else {
        *dst = make_int4(-1, 0, 0, IGN_STEP_ZERO);
      }
      cp_async_commit();
    };
    int l_lo[4], l_len[4], l_ent[4];               // loader rows: destination, first step, steps, next entry
    int s_len, s_lo;                                       // owner row: steps, first step
    float h[UPT];
    // gather this thread's 16 bytes of the messages of step t of its 4 loader rows (the entries were
    // fetched one step earlier, so the row loads do not wait on an index load), then fetch the entries
    // of step t + 1
    auto load_x = [&](int t, float4 (&x)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int entry = l_ent[i];
        if (t < l_len[i] && entry >= 0)
          x[i] = ldg_f4(pick_src(srcs, entry >> IGN_STEP_SRC_SHIFT) + (int64_t)(entry & IGN_STEP_ROW_MASK) * U + lc * 4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        l_ent[i] = (t + 1 < l_len[i]) ? __ldg(steps + l_lo[i] + t + 1) : IGN_STEP_ZERO;
    };
    auto store_x = [&](const float4 (&x)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) store_split(slot, slot + IMG, lr0 + 4 * i, lc, x[i]);
    };
    auto store_h = [&]() {
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j)
        store_split(slot + 2 * IMG, slot + 3 * IMG, row, split * (UPT / 4) + j,
                    make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
    };
    const uint32_t bxh_ = smem_u32(bx_hi), bxl_ = smem_u32(bx_lo), bhh_ = smem_u32(bh_hi), bhl_ = smem_u32(bh_lo);
    auto publish = [&]() {                                 // operands written: visible to the async proxy, walker-wide
      fence_async_smem();
      tc_fence_before();
      group_sync();
    };
    auto issue = [&]() {                                   // one thread: the 24 UMMAs of one step, then commit
      if (gtid == 0) {
        const uint32_t ax_hi = smem_u32(slot), ax_lo = ax_hi + IMG, ah_hi = ax_lo + IMG, ah_lo = ah_hi + IMG;
        tc_fence_after();
        umma_chunk_3x(tmem_base + g * 128, ax_hi, ax_lo, bxh_, bxl_, 128, false);
        umma_chunk_3x(tmem_base + g * 128, ah_hi, ah_lo, bhh_, bhl_, 128, true);
        umma_commit(&bar_acc[g]);
      }
      __syncwarp();
    };

    PROF(long long p_load = 0, p_setup = 0, p_wait = 0, p_epi = 0, p_store = 0, p_pub = 0, p_out = 0, p_tiles = 0,
         p_steps = 0, p_t0 = clock64(), c0, c1;)
    // Loads of the NEXT tile's rows (plan from shared memory): h0 chunks into registers, first messages
    // into x, entries of step 1; its longest list into the other parity's slot.  Issued before the wait
    // of the current tile's last step, so the set-up of the next tile never sits on a memory round trip.
    auto next_loads = [&](int ring, int np, float4 (&hv)[4], float4 (&x)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int4 m = s_meta[g][ring][lr0 + 4 * i];
        l_lo[i] = m.y; l_len[i] = m.z; l_ent[i] = m.w;
        hv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m.x >= 0) hv[i] = ldg_f4(h0 + (int64_t)m.x * U + lc * 4);
      }
      load_x(0, x);
      const int wmax = __reduce_max_sync(0xffffffffu, max(max(l_len[0], l_len[1]), max(l_len[2], l_len[3])));
      if (lane == 0 && wmax > 0) atomicMax(&s_maxlen[g][np], wmax);
    };
    // h0 as hi + lo with lo = v - hi unrounded (the tensor core truncates it: 2^-21 |v| for this one step
    // instead of 2^-22), so that the owners read hi + lo == h0 back exactly
    auto store_h0 = [&](const float4 (&hv)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = lr0 + 4 * i, off = r * 128 + ((lc ^ (r & 7)) << 4);
        float4 hi;
        hi.x = tf32_rna(hv[i].x); hi.y = tf32_rna(hv[i].y); hi.z = tf32_rna(hv[i].z); hi.w = tf32_rna(hv[i].w);
        *reinterpret_cast<float4*>(slot + 2 * IMG + off) = hi;
        *reinterpret_cast<float4*>(slot + 3 * IMG + off) =
            make_float4(hv[i].x - hi.x, hv[i].y - hi.y, hv[i].z - hi.z, hv[i].w - hi.w);
      }
    };

    const int64_t tile0 = (int64_t)blockIdx.x * WALKERS + g;
    int par = 0, ring = 0;                                 // parity / ring slot of the current tile
    {
      // prologue: plans of the first two tiles, the first tile's operands
      plan_to_smem(tile0, 0);
      plan_to_smem(tile0 + tile_stride, 1);
      if (gtid == 0) s_maxlen[g][0] = s_maxlen[g][1] = 0;
      cp_async_wait<0>();
      group_sync();
      float4 hv[4], x[4];
      next_loads(0, 0, hv, x);
      store_h0(hv);
      store_x(x);
    }
    for (int64_t tile = tile0; tile < ntiles PROF(&& !(dbg == 1 && g == 1)); tile += tile_stride) {
      // ---- the tile's operand images are in place (prologue / tail of the previous tile)
      PROF(c0 = clock64(); ++p_tiles;)
      const int ring1 = ring == 2 ? 0 : ring + 1, ring2 = ring1 == 2 ? 0 : ring1 + 1;
      {
        const int4 m = s_meta[g][ring][row];               // owner row
        s_len = m.z; s_lo = m.y;
      }
      if (gtid == 0) s_maxlen[g][par ^ 1] = 0;             // last read one tile ago
      cp_async_wait<0>();                                  // my part of the next tile's plan has landed
      PROF(c1 = clock64(); p_load += c1 - c0;)
      publish();
      const int maxlen = s_maxlen[g][par];
      if (maxlen > 0) issue();
      // the plan after the next one (its slot was last read before the barrier above); the next tile's
      // state rows into L2
      plan_to_smem(tile + 2 * tile_stride, ring2);
      if (split == 0) {
        const int nd = s_meta[g][ring1][row].x;
        if (nd >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(h0 + (int64_t)nd * U));
      }
      // owners: the running state in full fp32, read back from the state images (rows without a
      // destination hold zeros)
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j) {
        const int off = row * 128 + (((split * (UPT / 4) + j) ^ (row & 7)) << 4);
        const float4 hi = *reinterpret_cast<const float4*>(slot + 2 * IMG + off);
        const float4 lo = *reinterpret_cast<const float4*>(slot + 3 * IMG + off);
        h[4 * j] = hi.x + lo.x; h[4 * j + 1] = hi.y + lo.y; h[4 * j + 2] = hi.z + lo.z; h[4 * j + 3] = hi.w + lo.w;
      }
      PROF(c0 = clock64(); p_setup += c0 - c1;)

      float4 xn[4];                                        // next messages; after the last step: x_0 of the next tile
      if (maxlen == 0) {                                   // nothing to walk: the state images are idle already
        float4 hv[4];
        next_loads(ring1, par ^ 1, hv, xn);
        group_sync();                                      // every owner has read its h0 back
        store_h0(hv);
      }
      for (int t = 0; t < maxlen; ++t) {
        const bool last = t + 1 == maxlen;
        float4 hv[4];
        PROF(c0 = clock64(); ++p_steps;)
        if (!last) load_x(t + 1, xn);                      // in flight while we wait for the MMA
        else next_loads(ring1, par ^ 1, hv, xn);
        mbar_wait(&bar_acc[g], acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        PROF(c1 = clock64(); p_wait += c1 - c0;)
        if (last) {                                        // the last MMA is done: the state images are idle
          if (maxlen == 1) group_sync();                   // ... once every owner has read its h0 back
          store_h0(hv);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {             // 8 units at a time keeps the live registers down
          uint32_t az[8], ar[8], axh[8], ahh[8];
          tmem_ld8_nowait(tbase + half * 8, az);
          tmem_ld8_nowait(tbase + 32 + half * 8, ar);
          tmem_ld8_nowait(tbase + 64 + half * 8, axh);
          tmem_ld8_nowait(tbase + 96 + half * 8, ahh);
          tmem_ld_wait();
          if (t < s_len) {
#pragma unroll
            for (int j4 = 0; j4 < 8; j4 += 4) {
              const int ub = u0 + half * 8 + j4;
              const float4 vz = *reinterpret_cast<const float4*>(s_gb + ub);
              const float4 vr = *reinterpret_cast<const float4*>(s_gb + U + ub);
              const float4 vx = *reinterpret_cast<const float4*>(s_gb + 2 * U + ub);
              const float4 vh = *reinterpret_cast<const float4*>(s_gb + 3 * U + ub);
              const float bz[4] = {vz.x, vz.y, vz.z, vz.w}, br[4] = {vr.x, vr.y, vr.z, vr.w};
              const float bxh[4] = {vx.x, vx.y, vx.z, vx.w}, bhh[4] = {vh.x, vh.y, vh.z, vh.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j4 + jj;
                const float pz = __uint_as_float(az[j]) + bz[jj], pr = __uint_as_float(ar[j]) + br[jj];
                float& hv_ = h[half * 8 + j];
                if (FAST) {
                  hv_ = fast_gru_gate(pz, pr, __uint_as_float(axh[j]) + bxh[jj], __uint_as_float(ahh[j]) + bhh[jj], hv_);
                } else {
                  const float z = sigmoid_f(pz), r = sigmoid_f(pr);
                  const float hh = tanhf(fmaf(r, __uint_as_float(ahh[j]) + bhh[jj], __uint_as_float(axh[j]) + bxh[jj]));
                  hv_ = fmaf(z, hv_ - hh, hh);
                }
              }
            }
          }
        }
        if (h_seq && t < s_len) {
          float* p = h_seq + (int64_t)(s_lo + t) * U + u0;
#pragma unroll
          for (int j = 0; j < UPT / 4; ++j)
            st_f4(p + 4 * j, make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
        }
        PROF(c0 = clock64(); p_epi += c0 - c1;)
        if (!last) {                                       // operands of the next step, then do not wait
          store_h();
          store_x(xn);
          PROF(c1 = clock64(); p_store += c1 - c0;)
          publish();
          issue();
          PROF(c0 = clock64(); p_pub += c0 - c1;)
        }
      }
      // ---- results: owners park the final state in the (idle) message image, loaders write whole rows,
      // then drop the next tile's first messages into the same image
      PROF(c0 = clock64();)
      {
        unsigned char* stage = slot;
#pragma unroll
        for (int j = 0; j < UPT / 4; ++j) {
          const int c4 = split * (UPT / 4) + j;
          *reinterpret_cast<float4*>(stage + row * 128 + ((c4 ^ (row & 7)) << 4)) =
              make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        }
        group_sync();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = lr0 + 4 * i;
          const int d = s_meta[g][ring][r].x;              // this tile's plan stays until the next tile's barrier
          if (d >= 0)
            st_f4(out + (int64_t)d * U + lc * 4,
                  *reinterpret_cast<const float4*>(stage + r * 128 + ((lc ^ (r & 7)) << 4)));
        }
        store_x(xn);
        par ^= 1;
        ring = ring1;
      }
      PROF(c1 = clock64(); p_out += c1 - c0;)
    }
    PROF(if (gtid == 32) {
      const long long v[10] = {p_load, p_setup, p_wait, p_epi, p_store, p_pub, p_out, p_tiles, p_steps, clock64() - p_t0};
      for (int i = 0; i < 10; ++i) atomicAdd(&prof_cyc[i], (unsigned long long)v[i]);
      atomicAdd(&prof_cyc[10], 1ull);
    })
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256u);
}

}  // namespace

int ign_gru_seq_tc_launch(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                          const float* const* srcs, const float* h0, int64_t num_dst, const float* kernel,
                          const float* rkernel, const float* bias, float* out, float* h_seq, const int* meta,
                          cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  const size_t smem = 1024 + 4 * (size_t)IMG + WALKERS * (size_t)SLOT_BYTES;
  static const bool fast = getenv("IGN_GRU_TC_EXACT_MATH") == nullptr;   // default: ex2.approx-based sigmoid / tanh
  if (IGN_ONCE_PER_DEVICE()) {
    IGN_CUDA(cudaFuncSetAttribute(gru_seq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IGN_CUDA(cudaFuncSetAttribute(gru_seq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t nctas = ign_cdiv(ign_cdiv(num_dst, ROWS), WALKERS);
  const int grid = (int)(nctas < sms ? nctas : sms);
  if (fast)
    gru_seq_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(steps_rowptr, steps, order, sp, h0, num_dst, kernel,
                                                             rkernel, bias, out, h_seq,
                                                             reinterpret_cast<const int4*>(meta) PROF(, getenv("IGN_DBG") ? atoi(getenv("IGN_DBG")) : 0));
  else
    gru_seq_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(steps_rowptr, steps, order, sp, h0, num_dst, kernel,
                                                              rkernel, bias, out, h_seq,
                                                              reinterpret_cast<const int4*>(meta) PROF(, getenv("IGN_DBG") ? atoi(getenv("IGN_DBG")) : 0));
  IGN_CHECK_LAUNCH("gru_seq_tc");
  PROF({
    unsigned long long h[12];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, prof_cyc, sizeof(h));
    const double w = (double)h[10], nt = (double)h[7], ns = (double)h[8];
    fprintf(stderr, "gru_seq_tc walkers %.0f tiles/walker %.1f steps/walker %.1f cycles/walker %.0f | per tile: head %.0f "
                    "publish+issue+readback %.0f results+x0 %.0f | per step: wait %.0f gates %.0f split+store %.0f publish+issue %.0f\n",
            w, nt / w, ns / w, h[9] / w, h[0] / nt, h[1] / nt, h[6] / nt, h[2] / ns, h[3] / ns, h[4] / ns, h[5] / ns);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(prof_cyc, h, sizeof(h));
  })
  return IGN_OK;
}
