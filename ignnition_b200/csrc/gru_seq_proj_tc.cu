// Ordered / interleave aggregation + GRU update with the INPUT PROJECTION HOISTED out of the walk (sm_100a,
// 32-wide states).  Same contract as ign_gru_seq (reference code/utils/auxilary_classes.py:767-796, 421-440):
// for every destination d walk its step list,  h <- GRUCell(x = message(step), h).
//
// The Keras GRUCell (reset_after) splits into an input half and a recurrent half:
//     z = sigmoid(x Kz + bxz + h Rz + bhz),  r = sigmoid(x Kr + bxr + h Rr + bhr),
//     c = tanh(x Kh + bxh + r (h Rh + bhh)),  h' = z h + (1 - z) c
// A message is the state row of a SOURCE entity, and a source row is walked over many times per update
// (RouteNet GEANT2 x 4096: 303 k link rows, 6.14 M path-link steps), so
//   1. project_kernel computes  xp[row] = x K + [bxz + bhz | bxr + bhr | bxh]  ONCE per source row (fp32 FMA:
//      exact fp32 products, 35 us for 303 k rows) into a 96-wide table that stays in L2 (116 MB), and
//   2. the walk gathers xp rows (cp.async, 8 lanes per 128-byte segment) instead of x rows and runs only the
//      recurrent GEMM  D[128, 96] = h [Rz | Rr | Rh]  per step: 12 tcgen05.mma (3xTF32, K = 32) instead of 24,
//      no hi / lo split and no operand image for x at all.
// The h operand lives in TENSOR MEMORY (tcgen05.mma TS form: the owner thread of a row writes its hi / lo
// values with tcgen05.st, no swizzled shared-memory stores, no proxy fence), so a walker needs only the 50 KB
// staging tile of its gathered xp rows in shared memory and 160 TMEM columns (D 96 | A_hi 32 | A_lo 32):
// THREE independent 4-warp walkers per SM (thread = one destination, all 32 units) instead of two 8-warp ones.
// A walker's chain per step: MMA -> TMEM load + staged xp -> gates (5 SFU ops per element) -> tcgen05.st ->
// barrier -> MMA; the other two walkers fill the tensor pipe and the SFUs meanwhile.  At a tile boundary the
// next tile's plan, h0 rows and second entries are already in registers (loaded under the last step), the
// first MMA of the next tile needs no xp, and its xp gather flies under that MMA.

#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

using namespace ign_tc;

// -DIGN_PROJ_PROFILE: per-phase clock64() sums of warp 0 of every walker, printed after every launch
// (profiles/r2_gru_seq_proj.md).  Compiles to nothing otherwise.
#ifdef IGN_PROJ_PROFILE
#include <stdio.h>
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

namespace {
PROF(__device__ unsigned long long pj_prof[12];)

#ifndef IGN_PROJ_WALKERS
#define IGN_PROJ_WALKERS 3
#endif
constexpr int WALKERS = IGN_PROJ_WALKERS;      // (-DIGN_PROJ_WALKERS=1|2: phase-profile probes, profiles/r2_gru_seq_proj.md)
constexpr int WTHREADS = 128;                              // 4 warps = the 4 TMEM lane quarters
constexpr int THREADS = WALKERS * WTHREADS;
constexpr int ROWS = 128;                                  // destinations per tile = UMMA M
constexpr int U = 32;
constexpr int XP = 3 * U;                                  // projected message width
constexpr int XROW = XP * 4 + 16;                          // staged row: 384 bytes + 16 of padding (bank spread)
constexpr int STAGE = ROWS * XROW;
constexpr int BIMG = XP * 128;                             // [96 n][32 k] weight image
constexpr int TCOLS = 160;                                 // TMEM columns of a walker: D | A_hi | A_lo
constexpr int ENT_IDLE = -2;                               // loader code: the row has no step t

struct Tables {
  const float* p[IGN_MAX_SOURCES];
};
__device__ __forceinline__ const float* pick_table(const Tables& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// fast_gru_gate (tc_common.cuh) with the exponent scalings folded into the weights: tz = -log2e pz, tr = -log2e pr,
// xc_s / phh_s = the candidate's input and recurrent halves times -2 log2e, so the three multiplications in front of
// the ex2 are gone.  y = -2 log2e ph; tanh for |ph| < 1/4 (|y| < 0.7213) is tanh_small's odd polynomial written in y.
__device__ __forceinline__ float gru_gate_scaled(float tz, float tr, float xc_s, float phh_s, float h) {
  float ez, er, r, th, eh, inv;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(tz) : "f"(tz), "f"(40.0f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ez) : "f"(tz));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(er) : "f"(tr));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + er));
  const float y = fmaf(r, phh_s, xc_s);
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(th) : "f"(y), "f"(40.0f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eh) : "f"(th));
  const float a = 1.0f + ez, b = 1.0f + eh;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(a * b));
  const float z = b * inv;
  const float y2 = y * y;
  const float small = y * fmaf(y2, fmaf(y2, fmaf(y2, fmaf(y2, -1.577603293e-06f, 3.241205935e-05f), -6.666779407e-04f),
                                        1.387602744e-02f), -3.465735903e-01f);
  const float hh = fabsf(y) < 0.72134752f ? small : fmaf(a + a, inv, -1.0f);
  return fmaf(z, h - hh, hh);
}
// mbarrier wait that sleeps in hardware between polls (the spin of mbar_wait takes issue slots from the two other
// walkers' warps on the same scheduler)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)     // suspend-time hint, ns; an arrival wakes the warp at once
        : "memory");
  }
}
// explicit shared-space 16-byte load (through the generic row pointer the compiler emitted generic LD)
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}
constexpr float SCALE_ZR = -1.4426950408889634f;           // -log2 e
constexpr float SCALE_C = -2.8853900817779268f;            // -2 log2 e

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- projection
// xp[row, 0:96] = x[row, 0:32] K[32, 96] + fold(bias).  CTA = 128 rows x 3 gates (thread = one row, one gate);
// x tile and output tile staged in shared memory so that every global access is a full line.
constexpr int PJ_THREADS = 384;
constexpr int PJ_XS = U + 1;                               // padded x row (floats)
constexpr int PJ_OS = XP + 4;                              // padded output row (floats): 16-byte stores spread over the banks
__global__ void __launch_bounds__(PJ_THREADS, 2) project_kernel(const float* __restrict__ x, int64_t n,
                                                                const float* __restrict__ kernel,
                                                                const float* __restrict__ bias, float* __restrict__ xp,
                                                                float scale_zr, float scale_c) {
  extern __shared__ __align__(16) unsigned char pj_smem[];
  float* s_w = reinterpret_cast<float*>(pj_smem);                  // [32 k][96 n]
  float* s_x = s_w + U * XP;                                       // [128][33]
  float* s_o = s_x + ROWS * PJ_XS;                                 // [128][100] (offset 7296 floats: 16-byte aligned)
  __shared__ float s_b[XP];
  const int tid = threadIdx.x;
  for (int i = tid; i < U * XP; i += PJ_THREADS) s_w[i] = __ldg(kernel + i);
  if (tid < XP) s_b[tid] = bias[tid] + (tid < 2 * U ? bias[XP + tid] : 0.0f);
  const int gate = tid >> 7, r = tid & 127;
  const float gscale = gate < 2 ? scale_zr : scale_c;      // exponent scaling of the walk's gates (1 = none)
  constexpr int NPF = (ROWS * (U / 4) + PJ_THREADS - 1) / PJ_THREADS;     // 16-byte pieces of an x tile per thread
  float4 pf[NPF];
  // the x tile of the NEXT iteration travels in registers while this one is multiplied (two CTAs per SM cover the rest)
  auto fetch = [&](int64_t m0) {
#pragma unroll
    for (int j = 0; j < NPF; ++j) {
      const int i = tid + j * PJ_THREADS;
      const int rr = i >> 3, c4 = i & 7;
      pf[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < ROWS * (U / 4) && m0 + rr < n) pf[j] = ldg_f4(x + (m0 + rr) * U + c4 * 4);
    }
  };
  const int64_t stride = (int64_t)gridDim.x * ROWS;
  int64_t m0 = (int64_t)blockIdx.x * ROWS;
  if (m0 < n) fetch(m0);
  for (; m0 < n; m0 += stride) {
    __syncthreads();                                               // s_w / s_b ready; previous tile fully written out
#pragma unroll
    for (int j = 0; j < NPF; ++j) {
      const int i = tid + j * PJ_THREADS;
      if (i < ROWS * (U / 4)) {
        float* d = s_x + (i >> 3) * PJ_XS + (i & 7) * 4;
        d[0] = pf[j].x; d[1] = pf[j].y; d[2] = pf[j].z; d[3] = pf[j].w;
      }
    }
    __syncthreads();
    if (m0 + stride < n) fetch(m0 + stride);
    float acc[U];
#pragma unroll
    for (int j = 0; j < U; ++j) acc[j] = s_b[gate * U + j];
#pragma unroll 4
    for (int k = 0; k < U; ++k) {
      const float xv = s_x[r * PJ_XS + k];
      const float4* w4 = reinterpret_cast<const float4*>(s_w + k * XP + gate * U);
#pragma unroll
      for (int j4 = 0; j4 < U / 4; ++j4) {
        const float4 w = w4[j4];
        acc[4 * j4] = fmaf(xv, w.x, acc[4 * j4]);
        acc[4 * j4 + 1] = fmaf(xv, w.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(xv, w.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(xv, w.w, acc[4 * j4 + 3]);
      }
    }
#pragma unroll
    for (int j4 = 0; j4 < U / 4; ++j4)
      *reinterpret_cast<float4*>(s_o + r * PJ_OS + gate * U + 4 * j4) =
          make_float4(gscale * acc[4 * j4], gscale * acc[4 * j4 + 1], gscale * acc[4 * j4 + 2], gscale * acc[4 * j4 + 3]);
    __syncthreads();
    // the tile's rows are adjacent in global memory: one contiguous block written 16 bytes per thread
    const int64_t left = n - m0;
    const int n_f4 = (int)(left < ROWS ? left : ROWS) * (XP / 4);
    float4* dst = reinterpret_cast<float4*>(xp + m0 * XP);
    for (int i = tid; i < n_f4; i += PJ_THREADS) {
      const int rr = i / (XP / 4), c4 = i - rr * (XP / 4);
      dst[i] = *reinterpret_cast<const float4*>(s_o + rr * PJ_OS + 4 * c4);
    }
  }
}

// ---------------------------------------------------------------------------------------------- walk
// UPT = units per thread: 32 -> four warps per walker (thread = one destination), 16 -> eight warps per walker (two
// threads per destination, the second warp of a TMEM lane quarter takes units 16..31)
template <bool FAST, int UPT>
__global__ void __launch_bounds__(WALKERS * 128 * (U / UPT), 1) gru_seq_proj_kernel(
    const int* __restrict__ steps, Tables xp, const float* __restrict__ h0, int64_t num_dst,
    const float* __restrict__ rkernel, const float* __restrict__ bias, float* __restrict__ out,
    float* __restrict__ h_seq, const int4* __restrict__ meta, int opt) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int WT = 128 * (U / UPT);                      // threads of a walker
  constexpr int NTHREADS = WALKERS * WT;
  const uint32_t hint_ns = (uint32_t)opt >> 8;             // IGN_PROJ_OPT: suspend hint ns << 8
  unsigned char* b_hi = smem;                              // [96 n][32 k] images of [Rz | Rr | Rh], hi and lo
  unsigned char* b_lo = b_hi + BIMG;
  unsigned char* stages = b_lo + BIMG;
  __shared__ uint64_t bar_acc[WALKERS];                    // tensor core -> walker: D is complete
  __shared__ uint64_t bar_xp[WALKERS];                     // TMA -> walker: the xp rows of the step have landed
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bhh[U];                 // recurrent bias of the candidate (multiplied by r)
  __shared__ __align__(16) float s_xb[XP];                 // xp row of a ZERO message: the folded biases
  __shared__ int s_max[WALKERS][2][4];                     // longest list of a tile, per warp, per tile parity

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int g = 0; g < WALKERS; ++g) {
      mbar_init(&bar_acc[g], 1);
      mbar_init(&bar_xp[g], WT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512u);
  for (int i = tid; i < XP * U; i += NTHREADS) {
    const int n = i >> 5, k = i & 31;
    float hi, lo;
    tf32_split((FAST ? (n < 2 * U ? SCALE_ZR : SCALE_C) : 1.0f) * __ldg(rkernel + k * XP + n), hi, lo);
    const int off = sw128_off(n, k);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  if (tid < U) s_bhh[tid] = (FAST ? SCALE_C : 1.0f) * bias[XP + 2 * U + tid];
  if (tid < XP)
    s_xb[tid] = (FAST ? (tid < 2 * U ? SCALE_ZR : SCALE_C) : 1.0f) * (bias[tid] + (tid < 2 * U ? bias[XP + tid] : 0.0f));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t ntiles = (num_dst + ROWS - 1) / ROWS;

  const int g = warp / (WT / 32);                          // walker
  const int wq = warp % (WT / 32);
  const int q = wq & 3;                                    // TMEM lane quarter of this warp
  const int split = wq >> 2;                               // which half of the units (UPT = 16)
  const int u0 = split * UPT;
  const int gtid = tid % WT;
  const int row = q * 32 + lane;                           // this thread's destination of the tile
  unsigned char* stage = stages + g * STAGE;
  const uint32_t tD = tmem_base + g * TCOLS;               // MMA operand addresses (lane 0)
  const uint32_t tl = tD + ((uint32_t)(q * 32) << 16);     // this warp's lanes, for tcgen05.ld / st
  const uint32_t bhi_s = smem_u32(b_hi), blo_s = smem_u32(b_lo);
  const int64_t tile_stride = (int64_t)WALKERS * gridDim.x;
  uint32_t acc_phase = 0;

  auto wsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(WT) : "memory"); };
  auto load_plan = [&](int64_t tile) -> int4 {
    const int64_t i = tile * ROWS + row;
    if (tile < ntiles && i < num_dst) return __ldg(meta + i);
    return make_int4(-1, 0, 0, IGN_STEP_ZERO);
  };
  // gather the xp rows of one step into the staging tile: lane (sub, lc) copies 16-byte piece lc of the z, r and c
  // segments of rows 4 i + sub of this warp's quarter with cp.async, and the copies of every thread complete on the
  // walker's mbarrier (cp.async.mbarrier.arrive.noinc), so the consumers need no barrier of their own.  `e` = this
  // thread's own entry for that step (ENT_IDLE: no such step; IGN_STEP_ZERO: a zero message -- nothing is copied for
  // either, the owner of a zero-message row reads the folded biases s_xb instead of its staging row).  (One TMA bulk copy per row was measured slower: the per-lane UBLKCP loop costs 12
  // instructions per row, 20 % of all instructions issued -- profiles/r2_gru_seq_proj.md.)
  const int sub = lane >> 3, lc = lane & 7;
  const uint32_t stage_s = smem_u32(stage);
  unsigned char* my_row = stage + row * XROW;
  auto gather_xp = [&](int e) {
    constexpr int NIT = 8 / (WT / 128);                      // the warps of a lane quarter share its 32 rows
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int rl = 4 * (i + NIT * split) + sub;
      const int ev = __shfl_sync(0xffffffffu, e, rl);
      const uint32_t dst = stage_s + (uint32_t)((q * 32 + rl) * XROW + lc * 16);
      if (ev >= 0) {
        const float* src = pick_table(xp, ev >> IGN_STEP_SRC_SHIFT) + (int64_t)(ev & IGN_STEP_ROW_MASK) * XP + lc * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 128), "l"(src + U) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 256), "l"(src + 2 * U) : "memory");
      }
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_xp[g])) : "memory");
  };
  uint32_t xp_phase = 0;
  float h[UPT];
  // the state as the A operand of the next MMA: hi / lo columns of this thread's TMEM lane
  auto store_a = [&]() {
#pragma unroll
    for (int hb = 0; hb < UPT / 16; ++hb) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a, b;
        tf32_split(h[16 * hb + k], a, b);
        hi[k] = __float_as_uint(a);
        lo[k] = __float_as_uint(b);
      }
      tmem_st16(tl + XP + u0 + 16 * hb, hi);
      tmem_st16(tl + XP + U + u0 + 16 * hb, lo);
    }
    tmem_st_wait();
    tc_fence_before();
  };
  auto issue = [&]() {                                     // one thread: the 12 UMMAs of one step, then commit
    if (gtid == 0) {
      constexpr uint32_t idesc = umma_idesc(XP);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        umma_tf32_ts(tD, tD + XP + 8 * kk, umma_desc(bhi_s + 32 * kk), idesc, kk > 0 ? 1u : 0u);
        umma_tf32_ts(tD, tD + XP + U + 8 * kk, umma_desc(bhi_s + 32 * kk), idesc, 1u);
        umma_tf32_ts(tD, tD + XP + 8 * kk, umma_desc(blo_s + 32 * kk), idesc, 1u);
      }
      umma_commit(&bar_acc[g]);
    }
    __syncwarp();
  };

  PROF(long long p_mma = 0, p_xp = 0, p_gate = 0, p_sta = 0, p_iss = 0, p_gat = 0, p_tile = 0, p_steps = 0, p_tiles = 0,
       p_t0 = clock64(), c0, c1;)
  // ---- prologue: plan and state of the first tile, its first MMA and xp gather
  int64_t tile = (int64_t)blockIdx.x * WALKERS + g;
  int4 plan = load_plan(tile);
  int par = 0;
  {
    const int wmax = __reduce_max_sync(0xffffffffu, plan.z);
    if (lane == 0) s_max[g][0][q] = wmax;
#pragma unroll
    for (int j = 0; j < UPT / 4; ++j) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (plan.x >= 0) v = ldg_f4(h0 + (int64_t)plan.x * U + u0 + 4 * j);
      h[4 * j] = v.x; h[4 * j + 1] = v.y; h[4 * j + 2] = v.z; h[4 * j + 3] = v.w;
    }
  }
  int e1 = plan.z > 1 ? __ldg(steps + plan.y + 1) : ENT_IDLE;        // entry of step 1
  int e_cur = plan.z > 0 ? plan.w : ENT_IDLE;                        // entry of the step whose xp is being consumed
  if (tile < ntiles) {
    store_a();
    wsync();
  }
  int maxlen = max(max(s_max[g][0][0], s_max[g][0][1]), max(s_max[g][0][2], s_max[g][0][3]));
  if (tile < ntiles && maxlen > 0) {
    issue();
    gather_xp(plan.z > 0 ? plan.w : ENT_IDLE);
  }

  for (; tile < ntiles; tile += tile_stride) {
    const int s_d = plan.x, s_lo = plan.y, s_len = plan.z;
    int4 plan_n = load_plan(tile + tile_stride);           // arrives long before the last step
    float4 h0n[UPT / 4];
    int e1n = ENT_IDLE;
    // loads for the NEXT tile, issued under the wait of this tile's last step
    auto next_loads = [&]() {
      const int wmax = __reduce_max_sync(0xffffffffu, plan_n.z);
      if (lane == 0) s_max[g][par ^ 1][q] = wmax;
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j) {
        h0n[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (plan_n.x >= 0) h0n[j] = ldg_f4(h0 + (int64_t)plan_n.x * U + u0 + 4 * j);
      }
      e1n = plan_n.z > 1 ? __ldg(steps + plan_n.y + 1) : ENT_IDLE;
    };
    if (maxlen == 0) next_loads();
    for (int t = 0; t < maxlen; ++t) {
      const bool last = t + 1 == maxlen;
      const int e2 = (t + 2 < s_len) ? __ldg(steps + s_lo + t + 2) : ENT_IDLE;
      if (last) next_loads();
      PROF(c0 = clock64(); ++p_steps;)
      mbar_wait_sleep(&bar_acc[g], acc_phase, hint_ns);
      acc_phase ^= 1;
      PROF(c1 = clock64(); p_mma += c1 - c0;)
      mbar_wait_sleep(&bar_xp[g], xp_phase, hint_ns);            // the xp rows of step t have landed
      xp_phase ^= 1;
      tc_fence_after();
      PROF(c0 = clock64(); p_xp += c0 - c1;)
      {
        const bool act = t < s_len;                        // rows of one warp may differ in length: the TMEM loads are
        // .sync.aligned and run for every lane, only the math is guarded
        const uint32_t xr = e_cur == IGN_STEP_ZERO ? smem_u32(s_xb) : smem_u32(my_row);
        // [z | r | c] pre-activations of 8 units; double-buffered where the registers allow it (four-warp walkers)
        constexpr int NBUF = UPT == 32 ? 2 : 1;
        uint32_t acc[NBUF][24];
        const uint32_t tu = tl + u0;
        if (NBUF == 2) {
          tmem_ld8_nowait(tu, *reinterpret_cast<uint32_t(*)[8]>(&acc[0][0]));
          tmem_ld8_nowait(tu + U, *reinterpret_cast<uint32_t(*)[8]>(&acc[0][8]));
          tmem_ld8_nowait(tu + 2 * U, *reinterpret_cast<uint32_t(*)[8]>(&acc[0][16]));
        }
#pragma unroll
        for (int c = 0; c < UPT / 8; ++c) {                // 8 units at a time; the loads of chunk c + 1 fly under the math
          if (NBUF == 1) {
            tmem_ld8_nowait(tu + 8 * c, *reinterpret_cast<uint32_t(*)[8]>(&acc[0][0]));
            tmem_ld8_nowait(tu + U + 8 * c, *reinterpret_cast<uint32_t(*)[8]>(&acc[0][8]));
            tmem_ld8_nowait(tu + 2 * U + 8 * c, *reinterpret_cast<uint32_t(*)[8]>(&acc[0][16]));
          }
          tmem_ld_wait();
          if (NBUF == 2 && c + 1 < UPT / 8) {
            tmem_ld8_nowait(tu + 8 * (c + 1), *reinterpret_cast<uint32_t(*)[8]>(&acc[(c + 1) & 1][0]));
            tmem_ld8_nowait(tu + U + 8 * (c + 1), *reinterpret_cast<uint32_t(*)[8]>(&acc[(c + 1) & 1][8]));
            tmem_ld8_nowait(tu + 2 * U + 8 * (c + 1), *reinterpret_cast<uint32_t(*)[8]>(&acc[(c + 1) & 1][16]));
          }
          if (act) {
            const uint32_t* a = acc[NBUF == 2 ? (c & 1) : 0];
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              const float4 vz = lds_f4(xr + (u0 + 8 * c + 4 * j4) * 4);
              const float4 vr = lds_f4(xr + 128 + (u0 + 8 * c + 4 * j4) * 4);
              const float4 vc = lds_f4(xr + 256 + (u0 + 8 * c + 4 * j4) * 4);
              const float4 vb = *reinterpret_cast<const float4*>(s_bhh + u0 + 8 * c + 4 * j4);
              const float xz[4] = {vz.x, vz.y, vz.z, vz.w}, xg[4] = {vr.x, vr.y, vr.z, vr.w};
              const float xc[4] = {vc.x, vc.y, vc.z, vc.w}, bh[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * j4 + jj;
                const float pz = __uint_as_float(a[j]) + xz[jj], pr = __uint_as_float(a[8 + j]) + xg[jj];
                const float phh = __uint_as_float(a[16 + j]) + bh[jj];
                float& hv = h[8 * c + j];
                if (FAST) {
                  hv = gru_gate_scaled(pz, pr, xc[jj], phh, hv);
                } else {
                  const float z = sigmoid_f(pz), r = sigmoid_f(pr);
                  const float hh = tanhf(fmaf(r, phh, xc[jj]));
                  hv = fmaf(z, hv - hh, hh);
                }
              }
            }
          }
        }
        if (h_seq && act) {
          float* p = h_seq + (int64_t)(s_lo + t) * U + u0;
#pragma unroll
          for (int j = 0; j < UPT / 4; ++j) st_f4(p + 4 * j, make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
        }
      }
      PROF(c1 = clock64(); p_gate += c1 - c0;)
      if (!last) {
        store_a();
        PROF(c0 = clock64(); p_sta += c0 - c1;)
        wsync();                                           // staging tile and D are free, every A row is written
        issue();
        PROF(c1 = clock64(); p_iss += c1 - c0;)
        gather_xp(e1);
        PROF(c0 = clock64(); p_gat += c0 - c1;)
        e_cur = e1;
        e1 = e2;
      }
    }
    PROF(c0 = clock64(); ++p_tiles;)
    // ---- tile boundary: results out, the next tile's state in, its first MMA and xp gather
    if (s_d >= 0) {
      float* p = out + (int64_t)s_d * U + u0;
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j) st_f4(p + 4 * j, make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
    }
    if (tile + tile_stride < ntiles) {
#pragma unroll
      for (int j = 0; j < UPT / 4; ++j) {
        h[4 * j] = h0n[j].x; h[4 * j + 1] = h0n[j].y; h[4 * j + 2] = h0n[j].z; h[4 * j + 3] = h0n[j].w;
      }
      store_a();
      wsync();                                             // every A row is written; s_max of the next tile is visible
      par ^= 1;
      maxlen = max(max(s_max[g][par][0], s_max[g][par][1]), max(s_max[g][par][2], s_max[g][par][3]));
      plan = plan_n;
      e1 = e1n;
      e_cur = plan.z > 0 ? plan.w : ENT_IDLE;
      if (maxlen > 0) {
        issue();
        gather_xp(plan.z > 0 ? plan.w : ENT_IDLE);
      }
    }
    PROF(c1 = clock64(); p_tile += c1 - c0;)
  }
  PROF(if (gtid == 0) {
    const long long v[10] = {p_mma, p_xp, p_gate, p_sta, p_iss, p_gat, p_tile, p_tiles, p_steps, clock64() - p_t0};
    for (int i = 0; i < 10; ++i) atomicAdd(&pj_prof[i], (unsigned long long)v[i]);
    atomicAdd(&pj_prof[10], 1ull);
  })
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

}  // namespace

extern "C" size_t ign_gru_seq_proj_ws_bytes(int n_src, const int64_t* src_rows, int f_in, int units) {
  if (f_in != U || units != U || n_src < 1 || n_src > IGN_MAX_SOURCES || !src_rows) return 0;
  size_t total = 0;
  for (int k = 0; k < n_src; ++k) total += ((size_t)(src_rows[k] > 0 ? src_rows[k] : 0) * XP * 4 + 255) & ~(size_t)255;
  return total + 256;
}

extern "C" int ign_gru_seq_proj(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* meta, int n_src,
                                const float* const* srcs, const int64_t* src_rows, int f_in, const float* h0,
                                int64_t num_dst, int units, const float* kernel, const float* recurrent_kernel,
                                const float* bias, float* out, float* h_seq, void* ws, size_t ws_bytes, void* stream) {
  (void)steps_rowptr;
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: gru_seq_proj: negative size");
  IGN_REQUIRE(n_src >= 1 && n_src <= IGN_MAX_SOURCES && srcs && src_rows, IGN_ERR_INVALID,
              "IGNNITION: gru_seq_proj: between 1 and %d sources", IGN_MAX_SOURCES);
  IGN_REQUIRE(f_in == U && units == U, IGN_ERR_UNSUPPORTED,
              "IGNNITION: gru_seq_proj: built for message width == units == 32 (got %d, %d)", f_in, units);
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(steps && meta && h0 && out && kernel && recurrent_kernel && bias, IGN_ERR_INVALID,
              "IGNNITION: gru_seq_proj: null pointer (the walk plan of ign_seq_meta is required)");
  const size_t need = ign_gru_seq_proj_ws_bytes(n_src, src_rows, f_in, units);
  IGN_REQUIRE(ws && ws_bytes >= need, IGN_ERR_WORKSPACE, "IGNNITION: gru_seq_proj: workspace too small (%zu < %zu)",
              ws_bytes, need);
  cudaStream_t st = ign_stream(stream);
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const bool fast = getenv("IGN_GRU_TC_EXACT_MATH") == nullptr;   // default: ex2.approx-based sigmoid / tanh
  // 1. projected tables, one per source
  Tables tb;
  unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  const size_t pj_smem = (size_t)(U * XP + ROWS * PJ_XS + ROWS * PJ_OS) * 4;
  IGN_CUDA(cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pj_smem));
  for (int k = 0; k < IGN_MAX_SOURCES; ++k) {
    tb.p[k] = nullptr;
    if (k >= n_src) continue;
    IGN_REQUIRE(srcs[k] || src_rows[k] == 0, IGN_ERR_INVALID, "IGNNITION: gru_seq_proj: null source state");
    tb.p[k] = reinterpret_cast<const float*>(w);
    if (src_rows[k] > 0) {
      const int64_t tiles = ign_cdiv(src_rows[k], ROWS);
      const int grid = (int)(tiles < 2 * sms ? tiles : 2 * sms);
      project_kernel<<<grid, PJ_THREADS, pj_smem, st>>>(srcs[k], src_rows[k], kernel, bias, reinterpret_cast<float*>(w),
                                                        fast ? SCALE_ZR : 1.0f, fast ? SCALE_C : 1.0f);
      IGN_CHECK_LAUNCH("gru_seq_proj/project");
    }
    w += ((size_t)src_rows[k] * XP * 4 + 255) & ~(size_t)255;
  }
  // 2. the walk
  const size_t smem = 1024 + 2 * (size_t)BIMG + WALKERS * (size_t)STAGE;
  static const int upt = getenv("IGN_PROJ_UPT") ? atoi(getenv("IGN_PROJ_UPT")) : 32;
  auto launch = [&](auto kern, int threads) -> int {
    IGN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int opt_ = getenv("IGN_PROJ_OPT") ? atoi(getenv("IGN_PROJ_OPT")) : (500 << 8);
    const int64_t nctas_ = ign_cdiv(ign_cdiv(num_dst, ROWS), WALKERS);
    const int grid_ = (int)(nctas_ < sms ? nctas_ : sms);
    kern<<<grid_, threads, smem, st>>>(steps, tb, h0, num_dst, recurrent_kernel, bias, out, h_seq,
                                       reinterpret_cast<const int4*>(meta), opt_);
    return IGN_OK;
  };
  int rc;
  if (upt == 16) rc = fast ? launch(gru_seq_proj_kernel<true, 16>, 2 * THREADS) : launch(gru_seq_proj_kernel<false, 16>, 2 * THREADS);
  else rc = fast ? launch(gru_seq_proj_kernel<true, 32>, THREADS) : launch(gru_seq_proj_kernel<false, 32>, THREADS);
  if (rc) return rc;
  IGN_CHECK_LAUNCH("gru_seq_proj");
  PROF({
    unsigned long long hh[12];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(hh, pj_prof, sizeof(hh));
    const double w = (double)hh[10], nt = (double)hh[7], ns = (double)hh[8];
    fprintf(stderr, "gru_seq_proj walkers %.0f tiles/walker %.1f steps/walker %.1f cycles/walker %.0f | per step: wait MMA %.0f "
                    "wait xp + barrier %.0f gates %.0f store A %.0f barrier + issue %.0f gather issue %.0f | per tile: boundary %.0f\n",
            w, nt / w, ns / w, hh[9] / w, hh[0] / ns, hh[1] / ns, hh[2] / ns, hh[3] / ns, hh[4] / ns, hh[5] / ns, hh[6] / nt);
    memset(hh, 0, sizeof(hh));
    cudaMemcpyToSymbol(pj_prof, hh, sizeof(hh));
  })
  return IGN_OK;
}
