// Backward pass of the ordered / interleave aggregation on the tcgen05 tensor cores (sm_100a), 32-wide:
// tf.gradients through keras.layers.RNN(GRUCell) over every destination's message sequence
// (reference code/utils/generate_model.py:791 through auxilary_classes.py:767-796), organised like the
// step-synchronous forward (gru_step_tc.cu): BPTT step t of ALL destinations that have a step t is one
// streaming launch, t = max_len-1 .. 0.  Destinations are sorted by descending length, so the rows alive
// at step t are the prefix [0, nt[t]) of the sorted order; the running dL/dh lives in a sorted-order
// buffer dhs[num_dst, 32] between launches.
//
// Per 128-row tile (3xTF32 everywhere, fp32 accumulation in TMEM):
//   GEMM1  P[128, 128] = [x_t | h_{t-1}] . [K ; R]       gate pre-activations z | r | xh | hh (as the forward)
//   gates  -> gate gradients G = [d_az | d_ar | d_axh | d_ahh] and the direct term dh z, in registers;
//          G goes to global memory (the weight-gradient kernel below reads it) and, split into hi / lo, into
//          the A images of GEMM2 -- which are the x / h operand images of GEMM1, dead by then
//   GEMM2  [dx | dh][128, 64] = G[128, 128] . W2[128, 64] with W2 = [K^T ; R^T] arranged to G's columns; the
//          128 G columns are walked as 4 chunks of (8 units x 4 gates), each chunk issued as soon as the
//          epilogue warps that own those units have stored it
//   out    dx -> d_steps row, dh_{t-1} = dh z + dh part -> dhs (or dh0[d] at t = 0)
// Roles: 4 producer warps (gather x_t rows by the step-major table, h_{t-1} rows from h_seq / h0: 8 lanes per
// 128-byte row; next tile's rows are in registers while the current tile computes), 1 MMA-issuer warp,
// 8 epilogue warps; mbarriers only.
//
// The weight gradients dK = sum x^T GX, dR = sum h^T GH, db = colsum(G) are a tall-skinny product over all
// row-steps: gru_dw_tc_kernel runs it with MN-major operands and TMEM-resident accumulators (see dw_tc.cu
// for the layout), gathering [x_t | h_{t-1}] again and reading G.

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int ROWS = 128;
constexpr int U = 32;
constexpr int IMG = ROWS * 128;           // bytes of one [128 x 32] fp32 image
constexpr int W2IMG = 64 * 128;           // bytes of one [64 x 32] image
constexpr int PROD_WARPS = 4;
constexpr int EPI_WARPS = 8;
constexpr int EPI_WARP0 = PROD_WARPS;                       // warps 4..11: TMEM lane groups 0..3 twice
constexpr int MMA1_WARP = EPI_WARP0 + EPI_WARPS;
constexpr int MMA2_WARP = MMA1_WARP + 1;
constexpr int PF_WARP0 = MMA2_WARP + 1;                       // two warps that run ahead and pull rows into L2
constexpr int PF_WARPS = 2;
constexpr int PF_AHEAD = 1;                                 // tiles
constexpr int BW_THREADS = 32 * (PF_WARP0 + PF_WARPS);      // 16 warps

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};
// h operand: lo = v - hi kept exactly (the tensor core truncates it to TF32 on read), so that the epilogue can
// rebuild the old state bit for bit as hi + lo from the operand images
__device__ __forceinline__ void store_split_exact(unsigned char* img_hi, unsigned char* img_lo, int r, int c4, float4 v) {
  float4 hi, lo;
  hi.x = tf32_rna(v.x); lo.x = v.x - hi.x;
  hi.y = tf32_rna(v.y); lo.y = v.y - hi.y;
  hi.z = tf32_rna(v.z); lo.z = v.z - hi.z;
  hi.w = tf32_rna(v.w); lo.w = v.w - hi.w;
  const int off = r * 128 + ((c4 ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(img_hi + off) = hi;
  *reinterpret_cast<float4*>(img_lo + off) = lo;
}
__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// D[128, n] (+)= A B^T with A[128 lanes, 8 columns] in TENSOR MEMORY (tf32 bit patterns, one K element per column)
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMA bulk store shared -> global (linear bytes) and its bookkeeping
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#ifdef IGN_BWD_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_DECL unsigned long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pc_ = clock64();
#define PROF(k) { const long long n_ = clock64(); pt_[k] += (unsigned long long)(n_ - pc_); pc_ = n_; }
#define PROF_FLUSH(base, n) if (lane == 0) { for (int k_ = 0; k_ < n; ++k_) atomicAdd(&g_prof[base + k_], pt_[k_]); }
#else
#define PROF_DECL
#define PROF(k)
#define PROF_FLUSH(base, n)
#endif

__global__ void __launch_bounds__(BW_THREADS, 1) gru_step_bwd_tc_kernel(
    int t, const int* __restrict__ nt, const int* __restrict__ off, const int4* __restrict__ meta,
    const int* __restrict__ steps_T, SrcPtrs srcs, const float* __restrict__ h0, const float* __restrict__ h_seq,
    const float* __restrict__ d_out, float* __restrict__ dhs, float* __restrict__ dh0, float* __restrict__ d_steps,
    float* __restrict__ g_out, int64_t g_rows_bound, const float* __restrict__ kernel,
    const float* __restrict__ rkernel, const float* __restrict__ bias, int dbg) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // forward gate weights, 128 rows each: Bx = [K_z | K_r | K_h | 0], Bh = [R_z | R_r | 0 | R_h]
  unsigned char* bx_hi = smem;
  unsigned char* bx_lo = bx_hi + IMG;
  unsigned char* bh_hi = bx_lo + IMG;
  unsigned char* bh_lo = bh_hi + IMG;
  // W2 chunk q (units 8q .. 8q+7, column k = gate * 8 + jj): [hi 64 rows | lo 64 rows], rows 0-31 -> dx, 32-63 -> dh
  unsigned char* w2 = bh_lo + IMG;
  unsigned char* stage = w2 + 8 * W2IMG;                     // X_hi | X_lo | H_hi | H_lo of the tile in GEMM1
  unsigned char* gstage = stage + 4 * IMG;                   // per epilogue warp: [32 rows][128 B] of G on its way out
  __shared__ uint64_t bar_full, bar_acc1[2], bar_a1free[2], bar_gfull[2], bar_gdone[2], bar_acc2, bar_d2free, bar_hfree;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];                // merged gate biases [bz | br | bxh | bhh]
  __shared__ volatile int s_progress;                        // tiles whose GEMM1 has been issued (throttles the prefetch)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    s_progress = 0;
    mbar_init(&bar_full, PROD_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_acc1[s], 1);
      mbar_init(&bar_a1free[s], EPI_WARPS);
    }
    mbar_init(&bar_gfull[0], EPI_WARPS / 2);
    mbar_init(&bar_gfull[1], EPI_WARPS / 2);
    mbar_init(&bar_gdone[0], 1);
    mbar_init(&bar_gdone[1], 1);
    mbar_init(&bar_acc2, 1);
    mbar_init(&bar_d2free, EPI_WARPS);
    mbar_init(&bar_hfree, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA1_WARP) tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < 4 * IMG / 16; i += BW_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < U * 3 * U; i += BW_THREADS) {
    const int k = i / (3 * U), n = i % (3 * U);
    float hi, lo;
    tf32_split(__ldg(kernel + i), hi, lo);                   // K column n -> Bx row n (n < 96)
    *reinterpret_cast<float*>(bx_hi + sw128_off(n, k)) = hi;
    *reinterpret_cast<float*>(bx_lo + sw128_off(n, k)) = lo;
    tf32_split(__ldg(rkernel + i), hi, lo);                  // R column n -> Bh row n (z, r) or n + 32 (h)
    const int nh = n < 2 * U ? n : n + U;
    *reinterpret_cast<float*>(bh_hi + sw128_off(nh, k)) = hi;
    *reinterpret_cast<float*>(bh_lo + sw128_off(nh, k)) = lo;
  }
  for (int i = tid; i < 4 * 64 * 32; i += BW_THREADS) {
    const int q = i >> 11, n = (i >> 5) & 63, k = i & 31;
    const int gate = k >> 3, u = 8 * q + (k & 7);
    float v = 0.0f;
    if (n < U) {                                             // dx_n = sum GX[c] K[n][c]
      if (gate < 3) v = __ldg(kernel + n * 3 * U + gate * U + u);
    } else {                                                 // dh_m = sum GH[c] R[m][c]
      const int mrow = n - U;
      if (gate < 2) v = __ldg(rkernel + mrow * 3 * U + gate * U + u);
      else if (gate == 3) v = __ldg(rkernel + mrow * 3 * U + 2 * U + u);
    }
    float hi, lo;
    tf32_split(v, hi, lo);
    *reinterpret_cast<float*>(w2 + q * 2 * W2IMG + sw128_off(n, k)) = hi;
    *reinterpret_cast<float*>(w2 + q * 2 * W2IMG + W2IMG + sw128_off(n, k)) = lo;
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

  const int64_t n_alive = (int64_t)__ldg(nt + t);
  const int* entries = steps_T + __ldg(off + t);
  const int64_t ntiles = (n_alive + ROWS - 1) / ROWS;
  const int G = gridDim.x;
  const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + G - 1) / G : 0;

  if (warp < PROD_WARPS) {
    // ================================ producers ================================
    const int ptid = tid;                                    // 0..127
    const int c4 = ptid & 7;
    PROF_DECL
    // (meta, step entry) of this thread's 8 rows, fetched one tile ahead so that the row loads of a tile can be
    // issued at once (the gathers are two dependent loads deep)
    int4 m_nx[8];
    int ent_nx[8];
    auto fetch_index = [&](int64_t j) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {                          // rows (ptid >> 3) + 16 k
        const int64_t i = (blockIdx.x + j * G) * ROWS + (ptid >> 3) + 16 * k;
        const bool ok = j < my_tiles && i < n_alive;
        m_nx[k] = ok ? __ldg(meta + i) : make_int4(-1, 0, 0, 0);
        ent_nx[k] = ok ? __ldg(entries + i) : IGN_STEP_ZERO;
      }
    };
    fetch_index(0);
    for (int64_t j = 0; j < my_tiles; ++j) {
      int4 m[8];
      int ent[8];
      bool inb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        m[k] = m_nx[k];
        ent[k] = ent_nx[k];
        inb[k] = m[k].x >= 0;
        if (!inb[k]) m[k] = make_int4(0, 0, 1, 0);           // any readable row; masked below
      }
      float4 xv[8], hv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = inb[k] ? ent[k] : IGN_STEP_ZERO;
        const float* hp = (t == 0) ? h0 + (int64_t)m[k].x * U : h_seq + (int64_t)(m[k].y + t - 1) * U;
        const float* xp = e >= 0 ? pick_src(srcs, e >> IGN_STEP_SRC_SHIFT) + (int64_t)(e & IGN_STEP_ROW_MASK) * U : srcs.p[0];
        if (dbg & 32) {                                      // timing experiment only: every row from the same 128 KB
          hp = h0 + (int64_t)(((ptid >> 3) + 16 * k) * U);
          xp = srcs.p[0] + (int64_t)(((ptid >> 3) + 16 * k) * U);
        }
        hv[k] = ldg_f4(hp + c4 * 4);
        xv[k] = ldg_f4(xp + c4 * 4);
        if (e < 0) xv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!inb[k]) hv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      fetch_index(j + 1);
      // GEMM1 of the previous tile has read the stage
#ifdef IGN_BWD_PROFILE
      if (__float_as_uint(xv[0].x + hv[0].x + xv[7].x + hv[7].x) == 0x7fc12345u) __nanosleep(1);   // loads landed
#endif
      PROF(0)
      if (j > 0) {
        mbar_wait(&bar_acc1[(j - 1) & 1], (uint32_t)((j - 1) >> 1) & 1);
        mbar_wait(&bar_hfree, (uint32_t)(j - 1) & 1);        // ... and the epilogue has taken h_{t-1} from the h images
      }
      PROF(1)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = (ptid >> 3) + 16 * k;
        store_split(stage, stage + IMG, r, c4, xv[k]);
        store_split_exact(stage + 2 * IMG, stage + 3 * IMG, r, c4, hv[k]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full);
      PROF(2)
    }
    if (warp == 0) { PROF_FLUSH(0, 3) }
  } else if (warp == MMA1_WARP) {
    // ================================ issuer of GEMM1 ================================
    const uint32_t ax_hi = smem_u32(stage), ax_lo = ax_hi + IMG, ah_hi = ax_lo + IMG, ah_lo = ah_hi + IMG;
    for (int64_t j = 0; j < my_tiles; ++j) {
      const int buf = (int)(j & 1);
      mbar_wait(&bar_full, (uint32_t)j & 1);
      if (j >= 2) mbar_wait(&bar_a1free[buf], (uint32_t)((j >> 1) - 1) & 1);   // epilogue of tile j - 2 has read it
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d1 = tmem_base + buf * 128;
        umma_chunk_3x(d1, ax_hi, ax_lo, smem_u32(bx_hi), smem_u32(bx_lo), 128, false);    // z | r | xh | 0
        umma_chunk_3x(d1, ah_hi, ah_lo, smem_u32(bh_hi), smem_u32(bh_lo), 128, true);     // z | r | 0  | hh
        umma_commit(&bar_acc1[buf]);
        s_progress = (int)(j + 1);
      }
      __syncwarp();
    }
  } else if (warp >= PF_WARP0) {
    // ================================ L2 prefetch ================================
    // The gathers are two dependent loads deep (meta / step entry -> row); with one tile of rows in flight per SM
    // the kernel ran at the latency of that chain (1.6 TB/s).  These warps resolve the chain PF_AHEAD tiles early
    // and leave the rows in L2, holding no registers across the wait.
    const int pl = tid - PF_WARP0 * 32;                      // 0..63: rows pl, pl + 64
    for (int64_t j = 0; j < my_tiles; ++j) {
      if (dbg & 8) break;
      while (j - s_progress > PF_AHEAD) __nanosleep(200);
      const int64_t tile = blockIdx.x + j * G;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int64_t i = tile * ROWS + pl + 64 * k;
        if (i >= n_alive) continue;
        const int4 m = __ldg(meta + i);
        const int e = __ldg(entries + i);
        if (e >= 0) prefetch_l2(pick_src(srcs, e >> IGN_STEP_SRC_SHIFT) + (int64_t)(e & IGN_STEP_ROW_MASK) * U);
        prefetch_l2((t == 0) ? h0 + (int64_t)m.x * U : h_seq + (int64_t)(m.y + t - 1) * U);
        prefetch_l2((m.z == t + 1) ? d_out + (int64_t)m.x * U : dhs + i * U);
      }
    }
  } else if (warp == MMA2_WARP) {
    // ================================ issuer of GEMM2 ================================
    // A = G chunk in tensor memory (slot of the half that wrote it: hi 32 columns | lo 32 columns), B = W2 chunk images
    const uint32_t d2 = tmem_base + 256;
    constexpr uint32_t idesc = umma_idesc(64);
    for (int64_t j = 0; j < my_tiles; ++j) {
#pragma unroll 1
      for (int n = 0; n < 4; ++n) {                          // fills (half, c) in the order (0,0) (1,0) (0,1) (1,1)
        const int hf = n & 1, c = n >> 1, q = hf * 2 + c;
        mbar_wait(&bar_gfull[hf], (uint32_t)c);
        if (n == 0 && j > 0) mbar_wait(&bar_d2free, (uint32_t)(j - 1) & 1);   // previous [dx | dh] has been read
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = tmem_base + 320 + hf * 64, a_lo = a_hi + 32;
          const uint32_t w_hi = smem_u32(w2 + q * 2 * W2IMG), w_lo = w_hi + W2IMG;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_tf32_ts(d2, a_hi + 8 * kk, umma_desc(w_hi + 32 * kk), idesc, (n > 0 || kk > 0) ? 1u : 0u);
            umma_tf32_ts(d2, a_lo + 8 * kk, umma_desc(w_hi + 32 * kk), idesc, 1u);
            umma_tf32_ts(d2, a_hi + 8 * kk, umma_desc(w_lo + 32 * kk), idesc, 1u);
          }
          umma_commit(&bar_gdone[hf]);
          if (n == 3) umma_commit(&bar_acc2);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue ================================
    const int e = warp - EPI_WARP0;                          // 0..7
    const int q = warp & 3;                                  // TMEM lane group this warp may access
    const int half = e >> 2;                                 // units [16 half, 16 half + 16)
    const int row = q * 32 + lane;
    unsigned char* my_stage = gstage + e * 4096;
    const int64_t warp_row0 = q * 32;                        // first tile row of this warp
    PROF_DECL
    // (dL/dh, h_{t-1}) of this thread's 16 units, prefetched one tile ahead
    float4 pre_d[4];
    int4 mr_next = make_int4(-1, 0, 0, 0);
    auto fetch_meta = [&](int64_t j) {
      const int64_t i = (blockIdx.x + j * G) * ROWS + row;
      mr_next = (j < my_tiles && i < n_alive) ? __ldg(meta + i) : make_int4(-1, 0, 0, 0);
    };
    auto fetch_rows = [&](int64_t j) {
      if (mr_next.x < 0 || (dbg & 2)) return;
      const int64_t i = (blockIdx.x + j * G) * ROWS + row;
      const float* dsrc = ((mr_next.z == t + 1) ? d_out + (int64_t)mr_next.x * U : dhs + i * U) + half * 16;
#pragma unroll
      for (int k = 0; k < 4; ++k) pre_d[k] = ldg_f4(dsrc + 4 * k);
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) pre_d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    fetch_meta(0);
    fetch_rows(0);
    for (int64_t j = 0; j < my_tiles; ++j) {
      const int buf = (int)(j & 1);
      const int64_t i = (blockIdx.x + j * G) * ROWS + row;
      const int4 mr = mr_next;
      const bool alive = mr.x >= 0;
      float dh[16], hold[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dh[4 * k] = pre_d[k].x; dh[4 * k + 1] = pre_d[k].y; dh[4 * k + 2] = pre_d[k].z; dh[4 * k + 3] = pre_d[k].w;
      }
      fetch_meta(j + 1);
      float direct[16];
      PROF(0)
      mbar_wait(&bar_acc1[buf], (uint32_t)(j >> 1) & 1);     // gate pre-activations of the tile are in TMEM
      tc_fence_after();
      // h_{t-1} of this thread's 16 units = hi + lo of the h operand images (GEMM1 has read them; the producers wait
      // for bar_hfree before the next tile overwrites them)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = row * 128 + (((half * 4 + k) ^ (row & 7)) << 4);
        const float4 a = *reinterpret_cast<const float4*>(stage + 2 * IMG + o);
        const float4 b = *reinterpret_cast<const float4*>(stage + 3 * IMG + o);
        hold[4 * k] = a.x + b.x; hold[4 * k + 1] = a.y + b.y; hold[4 * k + 2] = a.z + b.z; hold[4 * k + 3] = a.w + b.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_hfree);
      PROF(1)
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int u0 = half * 16 + 8 * c;
        uint32_t az[8], ar[8], axh[8], ahh[8];
        tmem_ld8_nowait(tb + buf * 128 + u0, az);
        tmem_ld8_nowait(tb + buf * 128 + 32 + u0, ar);
        tmem_ld8_nowait(tb + buf * 128 + 64 + u0, axh);
        tmem_ld8_nowait(tb + buf * 128 + 96 + u0, ahh);
        tmem_ld_wait();
        if (c == 1) {                                        // acc1[buf] is free for the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_a1free[buf]);
        }
        float gz[8], gr[8], gx[8], gh[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float z = fast_sigmoid(__uint_as_float(az[u]) + s_gb[u0 + u]);
          const float r = fast_sigmoid(__uint_as_float(ar[u]) + s_gb[U + u0 + u]);
          const float phh = __uint_as_float(ahh[u]) + s_gb[3 * U + u0 + u];
          const float hh = fast_tanh(fmaf(r, phh, __uint_as_float(axh[u]) + s_gb[2 * U + u0 + u]));
          const float dv = alive ? dh[8 * c + u] : 0.0f;
          const float t1 = dv * (1.0f - z);
          gx[u] = t1 * (1.0f - hh * hh);
          gh[u] = gx[u] * r;
          gr[u] = gx[u] * phh * r * (1.0f - r);
          gz[u] = dv * (hold[8 * c + u] - hh) * z * (1.0f - z);
          direct[8 * c + u] = dv * z;
        }
        // G chunk (k = gate * 8 + unit) -> tensor memory as the A operand of GEMM2 (hi | lo), slot of this half; the
        // slot's previous chunk (c == 0 of this tile) must have been consumed
        PROF(2)
        if (c == 1) mbar_wait(&bar_gdone[half], 0);
        PROF(3)
        {
          const float* gsrc[4] = {gz, gr, gx, gh};
          uint32_t hi[16], lo[16];
          const uint32_t ta = tb + 320 + half * 64;
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {                   // gates (0, 1) then (2, 3)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float v = gsrc[2 * hb + (k >> 3)][k & 7];
              const float vh_ = tf32_rna(v);
              hi[k] = __float_as_uint(vh_);
              lo[k] = __float_as_uint(tf32_rna(v - vh_));
            }
            tmem_st16(ta + 16 * hb, hi);
            tmem_st16(ta + 32 + 16 * hb, lo);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_gfull[half]);
        }
        PROF(4)
        // the same chunk to global memory for the weight-gradient kernel: chunk-major G[q][row][32] with the
        // 16-byte pieces of a row XOR-ed by (row & 7), staged per warp and written as one 4 KB TMA bulk store
        if (!(dbg & 1)) {
          if (lane == 0) bulk_wait_read();                   // the previous bulk store has read the staging rows
          __syncwarp();
          unsigned char* srow = my_stage + lane * 128;
          const int sw = lane & 7;
          *reinterpret_cast<float4*>(srow + ((0 ^ sw) << 4)) = make_float4(gz[0], gz[1], gz[2], gz[3]);
          *reinterpret_cast<float4*>(srow + ((1 ^ sw) << 4)) = make_float4(gz[4], gz[5], gz[6], gz[7]);
          *reinterpret_cast<float4*>(srow + ((2 ^ sw) << 4)) = make_float4(gr[0], gr[1], gr[2], gr[3]);
          *reinterpret_cast<float4*>(srow + ((3 ^ sw) << 4)) = make_float4(gr[4], gr[5], gr[6], gr[7]);
          *reinterpret_cast<float4*>(srow + ((4 ^ sw) << 4)) = make_float4(gx[0], gx[1], gx[2], gx[3]);
          *reinterpret_cast<float4*>(srow + ((5 ^ sw) << 4)) = make_float4(gx[4], gx[5], gx[6], gx[7]);
          *reinterpret_cast<float4*>(srow + ((6 ^ sw) << 4)) = make_float4(gh[0], gh[1], gh[2], gh[3]);
          *reinterpret_cast<float4*>(srow + ((7 ^ sw) << 4)) = make_float4(gh[4], gh[5], gh[6], gh[7]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int qc = half * 2 + c;
            const int64_t r0 = (blockIdx.x + j * G) * ROWS + warp_row0;
            bulk_s2g(g_out + ((int64_t)qc * g_rows_bound + r0) * U, my_stage, 4096);
          }
        }
        PROF(5)
      }
      fetch_rows(j + 1);                                     // lands while GEMM2 finishes
      PROF(0)
      mbar_wait(&bar_acc2, (uint32_t)j & 1);                 // [dx | dh] of the tile is in TMEM
      tc_fence_after();
      PROF(6)
      uint32_t vx[16], vh[16];
      tmem_ld16_nowait(tb + 256 + half * 16, vx);
      tmem_ld16_nowait(tb + 256 + U + half * 16, vh);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_d2free);
      if (!(dbg & 4)) {
        // [dx | dh] leave through the staging rows of this lane group's two warps (half 0's buffer: dx rows, half 1's:
        // dh rows; 16-byte pieces XOR-ed by row & 7), then each warp writes one of the two with 8 lanes per 128-byte row
        unsigned char* buf_x = gstage + (e & 3) * 4096;
        unsigned char* buf_h = buf_x + 4 * 4096;
        if (lane == 0) bulk_wait_read();                     // the bulk store of the last G chunk has read the rows
        __syncwarp();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");            // ... in both warps of the pair
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int o = lane * 128 + (((half * 4 + k) ^ (lane & 7)) << 4);
          *reinterpret_cast<float4*>(buf_x + o) = make_float4(__uint_as_float(vx[4 * k]), __uint_as_float(vx[4 * k + 1]),
                                                               __uint_as_float(vx[4 * k + 2]), __uint_as_float(vx[4 * k + 3]));
          *reinterpret_cast<float4*>(buf_h + o) =
              make_float4(direct[4 * k] + __uint_as_float(vh[4 * k]), direct[4 * k + 1] + __uint_as_float(vh[4 * k + 1]),
                          direct[4 * k + 2] + __uint_as_float(vh[4 * k + 2]), direct[4 * k + 3] + __uint_as_float(vh[4 * k + 3]));
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const int64_t dst_row = (half == 0) ? (int64_t)(mr.y + t) : ((t == 0) ? (int64_t)mr.x : i);
        float* dst_base = (half == 0) ? d_steps : ((t == 0) ? dh0 : dhs);
        const unsigned char* buf = half == 0 ? buf_x : buf_h;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = (lane >> 3) + 4 * k, ch = lane & 7;
          const long long drow = __shfl_sync(0xffffffffu, (long long)dst_row, rr);
          const int al = __shfl_sync(0xffffffffu, alive ? 1 : 0, rr);
          const float4 v = *reinterpret_cast<const float4*>(buf + rr * 128 + ((ch ^ (rr & 7)) << 4));
          if (al) st_f4(dst_base + drow * U + ch * 4, v);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");            // rows are free for the next tile's G chunks
      }
      PROF(7)
    }
    if (warp == EPI_WARP0) { PROF_FLUSH(3, 8) }
  }
  if (warp >= EPI_WARP0 && warp < MMA1_WARP && lane == 0) bulk_wait_all();   // G rows are in global memory
  tc_fence_before();
  __syncthreads();
  if (warp == MMA1_WARP) tmem_dealloc(tmem_base, 512);
}

// destinations without any step: the state passes through, so does its gradient
__global__ void gru_bwd_passthrough_kernel(const int* __restrict__ nt, int64_t num_dst, const int4* __restrict__ meta,
                                           const float* __restrict__ d_out, float* __restrict__ dh0) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = (int64_t)__ldg(nt) + (idx >> 3);
  if (i >= num_dst) return;
  const int d = __ldg(meta + i).x;
  const int c4 = (int)(idx & 7);
  st_f4(dh0 + (int64_t)d * U + c4 * 4, ldg_f4(d_out + (int64_t)d * U + c4 * 4));
}

// ---------------------------------------------------------------------------------------------------------
// weight gradients of one step: D[128 lanes = x_hi | h_hi | x_lo | h_lo features][128 = G columns] accumulated
// over the CTA's rows in TMEM (MN-major SWIZZLE_128B_BASE32B operands, see dw_tc.cu)
// ---------------------------------------------------------------------------------------------------------
constexpr int RW = 32;                     // rows per stage = 4 K-steps
constexpr int DIMG = RW * 128;             // bytes of one [32 x 32] image
constexpr int DW_BLOCKS = 12;              // x_hi h_hi x_lo h_lo | G_hi[4] | G_lo[4]
constexpr int DW_STAGE = DW_BLOCKS * DIMG;
constexpr int DW_STAGES = 2;
constexpr int DW_PROD_WARPS = 8;           // two CTAs per SM: one splits and stores while the other one's loads land
constexpr int DW_PROD_THREADS = 32 * DW_PROD_WARPS;
constexpr int DW_THREADS = DW_PROD_THREADS;                 // warp 0 also issues the MMAs; 2 x 8 warps keep 128 registers

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                    // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void store_split_mn(unsigned char* img_hi, unsigned char* img_lo, int r, int c4, float4 v) {
  float4 hi, lo;
  tf32_split(v.x, hi.x, lo.x);
  tf32_split(v.y, hi.y, lo.y);
  tf32_split(v.z, hi.z, lo.z);
  tf32_split(v.w, hi.w, lo.w);
  const int o = r * 128 + ((((c4 >> 1) ^ (r & 3)) << 5) | ((c4 & 1) << 4));
  *reinterpret_cast<float4*>(img_hi + o) = hi;
  *reinterpret_cast<float4*>(img_lo + o) = lo;
}

__global__ void __launch_bounds__(DW_THREADS, 2) gru_dw_tc_kernel(
    int t, const int* __restrict__ nt, const int* __restrict__ off, const int4* __restrict__ meta,
    const int* __restrict__ steps_T, SrcPtrs srcs, const float* __restrict__ h0, const float* __restrict__ h_seq,
    const float* __restrict__ g_in, int64_t g_rows_bound, float* __restrict__ dK, float* __restrict__ dR,
    float* __restrict__ dB, int dbg) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[DW_STAGES], bar_free[DW_STAGES], bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(&bar_full[s], DW_PROD_WARPS);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int64_t n_alive = (int64_t)__ldg(nt + t);
  const int* entries = steps_T + __ldg(off + t);
  const int64_t nchunks = (n_alive + RW - 1) / RW;
  // at least 64 chunks (2048 rows) per CTA: a short step must not pay a flush of the accumulators (12 k reductions)
  // in every CTA
  int64_t per = (nchunks + gridDim.x - 1) / gridDim.x;
  if (per < 64) per = 64;
  const int64_t c0 = (int64_t)blockIdx.x * per;
  const int64_t c1 = c0 + per < nchunks ? c0 + per : nchunks;
  const int64_t n_my = c1 > c0 ? c1 - c0 : 0;

  {
    // per 32-row chunk a thread loads two float4 of [x | h] (rows tid / 16 + 16 p, chunk tid % 16) and four of G
    // (rows tid / 32 + 8 p, chunk tid % 32): its G column chunk never changes, so the bias sums stay in registers
    const int ra = tid >> 4, wa = tid & 15;
    const int rg = tid >> 5, wg = tid & 31;
    float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);
    // the gather is two loads deep (step entry / meta -> row): the raw indices of a chunk are fetched one iteration
    // before its row pointers are formed, so no iteration waits on a load it has just issued
    struct RawIdx { int a, b; };                             // x part: (entry, -); h part: (destination, first step)
    auto fetch_raw = [&](int64_t chunk, int p) -> RawIdx {
      const int64_t i = chunk * RW + ra + 16 * p;
      RawIdx r{IGN_STEP_ZERO, -1};
      if (chunk < c1 && i < n_alive) {
        if (wa < 8) r.a = __ldg(entries + i);
        else { const int4 m = __ldg(meta + i); r.a = m.x; r.b = m.y; }
      }
      return r;
    };
    auto make_ptr = [&](const RawIdx& r) -> const float* {
      if (wa < 8)
        return r.a >= 0 ? pick_src(srcs, r.a >> IGN_STEP_SRC_SHIFT) + (int64_t)(r.a & IGN_STEP_ROW_MASK) * U + wa * 4 : nullptr;
      if (r.b < 0) return nullptr;
      return ((t == 0) ? h0 + (int64_t)r.a * U : h_seq + (int64_t)(r.b + t - 1) * U) + (wa - 8) * 4;
    };
    auto a_ptr = [&](int64_t chunk, int p) -> const float* { return make_ptr(fetch_raw(chunk, p)); };
    auto load = [&](int64_t chunk, const float* const (&ap)[2], float4 (&va)[2], float4 (&vg)[4]) {
#pragma unroll
      for (int p = 0; p < 2; ++p) va[p] = ap[p] ? ldg_f4(ap[p]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int64_t i = chunk * RW + rg + 8 * p;
        // chunk-major G[q][row][32], 16-byte pieces XOR-ed by (row & 7) (as the backward kernel stores it)
        vg[p] = i < n_alive ? ld_stream_f4(g_in + ((int64_t)(wg >> 3) * g_rows_bound + i) * U + (((wg & 7) ^ (int)(i & 7)) << 2))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    // rows of chunks i + 1 and i + 2 are in registers while chunk i is split and stored (the loads see ~3 us)
    float4 ca[2], cg[4], na[2], ng[4], fa[2], fg[4];
    const float* p_next[2] = {nullptr, nullptr};
#pragma unroll
    for (int p = 0; p < 2; ++p) na[p] = fa[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < 4; ++p) ng[p] = fg[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_my > 0) {
      const float* p0[2] = {a_ptr(c0, 0), a_ptr(c0, 1)};
      load(c0, p0, ca, cg);
      if (n_my > 1) {
        const float* p1[2] = {a_ptr(c0 + 1, 0), a_ptr(c0 + 1, 1)};
        load(c0 + 1, p1, na, ng);
      }
      if (n_my > 2) { p_next[0] = a_ptr(c0 + 2, 0); p_next[1] = a_ptr(c0 + 2, 1); }
    }
    RawIdx raw_next[2] = {fetch_raw(c0 + 3, 0), fetch_raw(c0 + 3, 1)};
    for (int64_t i = 0; i < n_my; ++i) {
      const int s = (int)(i % DW_STAGES);
      if (i + 2 < n_my) load(c0 + i + 2, p_next, fa, fg);
      if (i + 3 < n_my) {
        p_next[0] = make_ptr(raw_next[0]);
        p_next[1] = make_ptr(raw_next[1]);
        raw_next[0] = fetch_raw(c0 + i + 4, 0);
        raw_next[1] = fetch_raw(c0 + i + 4, 1);
        if ((wa & 7) == 0) {                                 // the gather is resolved one chunk early: start it now
          if (p_next[0]) prefetch_l2(p_next[0]);
          if (p_next[1]) prefetch_l2(p_next[1]);
        }
        const int64_t gi = (c0 + i + 3) * RW + rg;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if ((wg & 7) == 0 && gi + 8 * p < n_alive)
            prefetch_l2(g_in + ((int64_t)(wg >> 3) * g_rows_bound + gi + 8 * p) * U);
      }
      if (i >= DW_STAGES) mbar_wait(&bar_free[s], (uint32_t)((i / DW_STAGES) - 1) & 1);
      unsigned char* st = smem + (size_t)s * DW_STAGE;
#pragma unroll
      for (int p = 0; p < 2; ++p) {                          // blocks: x_hi 0, h_hi 1, x_lo 2, h_lo 3
        unsigned char* hi = st + (wa >> 3) * DIMG;
        store_split_mn(hi, hi + 2 * DIMG, ra + 16 * p, wa & 7, ca[p]);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        unsigned char* hi = st + (4 + (wg >> 3)) * DIMG;
        store_split_mn(hi, hi + 4 * DIMG, rg + 8 * p, wg & 7, cg[p]);
        colsum.x += cg[p].x; colsum.y += cg[p].y; colsum.z += cg[p].z; colsum.w += cg[p].w;
      }
      if (!(dbg & 32)) fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
      if (warp == 0) {                                       // issue the stage once every warp has stored its part
        mbar_wait(&bar_full[s], (uint32_t)(i / DW_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(st);
          const uint32_t b_hi = sa + 4 * DIMG, b_lo = b_hi + 4 * DIMG;
#pragma unroll
          for (int ks = 0; ks < RW / 8; ++ks) {
            const uint32_t ko = ks * 1024;
            umma_tf32(tmem_base, umma_desc_mn(sa + ko, DIMG), umma_desc_mn(b_hi + ko, DIMG), idesc, (i > 0 || ks > 0) ? 1u : 0u);
            umma_tf32(tmem_base, umma_desc_mn(sa + ko, DIMG), umma_desc_mn(b_lo + ko, DIMG), idesc, 1u);
          }
          umma_commit(&bar_free[s]);
          if (i + 1 == n_my) umma_commit(&bar_done);
        }
        __syncwarp();
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) { ca[p] = na[p]; na[p] = fa[p]; }
#pragma unroll
      for (int p = 0; p < 4; ++p) { cg[p] = ng[p]; ng[p] = fg[p]; }
    }
    if (n_my > 0) {
      // G column c' = q * 32 + gate * 8 + jj holds gate `gate` of unit 8 q + jj (gates: z, r, xh, hh)
      // bias gradients: db[0] takes z, r, xh; db[1] takes z, r and hh (third block)
      const float cs[4] = {colsum.x, colsum.y, colsum.z, colsum.w};
      {
        const int gate = (wg & 7) >> 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int unit = 8 * (wg >> 3) + (wg & 1) * 4 + k;
          if (gate < 3) atomicAdd(dB + gate * U + unit, cs[k]);
          if (gate < 2) atomicAdd(dB + 3 * U + gate * U + unit, cs[k]);
          if (gate == 3) atomicAdd(dB + 3 * U + 2 * U + unit, cs[k]);
        }
      }
      mbar_wait(&bar_done, 0);
      tc_fence_after();
      const int lg = warp & 3;                               // lanes: x_hi | h_hi | x_lo | h_lo
      const bool is_h = lg & 1;
#pragma unroll 1
      for (int qc = warp >> 2; qc < 4; qc += DW_PROD_WARPS / 4)   // columns: chunk q
#pragma unroll 1
      for (int gate = 0; gate < 4; ++gate) {
        float* wrow = (is_h ? dR : dK) + lane * 3 * U + 8 * qc;
        // x rows take z, r, xh (gates 0-2); h rows take z, r (0, 1) and hh (3) -> recurrent column block 2
        if (is_h ? gate == 2 : gate == 3) continue;
        uint32_t v[8];
        tmem_ld8_nowait(tmem_base + ((uint32_t)(lg * 32) << 16) + qc * U + gate * 8, v);
        tmem_ld_wait();
        float* base = wrow + (gate == 3 ? 2 : gate) * U;
        red_add_v4(base, __uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        red_add_v4(base + 4, __uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace

#ifdef IGN_BWD_PROFILE
extern "C" void ign_debug_bwd_prof(unsigned long long* out, int reset) {
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); return; }
  cudaMemcpyFromSymbol(out, g_prof, 16 * sizeof(unsigned long long));
}
// timing experiments (tools/ordered_bwd_bench.py <flags>): switch parts of the kernels off -- RESULTS ARE WRONG with a
// non-zero value, so the switch only exists in the profiling build (tools/build_profile_lib.sh)
static int g_dbg = 0;
extern "C" void ign_debug_bwd(int v) { g_dbg = v; }
#else
static const int g_dbg = 0;
#endif

static inline int64_t g_bound(int64_t num_dst) { return ign_cdiv(num_dst > 0 ? num_dst : 1, ROWS) * ROWS; }

size_t ign_gru_step_bwd_tc_ws(int64_t num_dst) {
  return ign_align((size_t)g_bound(num_dst) * 4 * U * sizeof(float)) +                 // G rows of one step (whole tiles)
         ign_align((size_t)(num_dst > 0 ? num_dst : 1) * U * sizeof(float));           // running dL/dh
}

int ign_gru_step_bwd_tc_launch(int max_steps, const int* nt, const int* off, int64_t num_dst, const int* meta,
                               const int* steps_T, int n_src, const float* const* srcs, const float* h0,
                               const float* h_seq, const float* kernel, const float* rkernel, const float* bias,
                               const float* d_out, float* d_steps, float* dh0, float* dK, float* dR, float* dB,
                               void* ws, cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  float* g_rows = reinterpret_cast<float*>(ws);
  float* dhs = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ign_align((size_t)g_bound(num_dst) * 4 * U * sizeof(float)));
  const size_t smem_a = 1024 + 4 * (size_t)IMG + 8 * (size_t)W2IMG + 4 * (size_t)IMG + EPI_WARPS * 4096;
  const size_t smem_w = 1024 + (size_t)DW_STAGES * DW_STAGE;
  if (IGN_ONCE_PER_DEVICE()) {
    IGN_CUDA(cudaFuncSetAttribute(gru_step_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    IGN_CUDA(cudaFuncSetAttribute(gru_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(num_dst, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  const int4* meta4 = reinterpret_cast<const int4*>(meta);
  gru_bwd_passthrough_kernel<<<(unsigned)ign_cdiv(num_dst * 8, 256), 256, 0, st>>>(nt, num_dst, meta4, d_out, dh0);
  IGN_CHECK_LAUNCH("gru_bwd_passthrough");
  // enough 64-row chunks per CTA that the atomic flush of the accumulators stays small
  int64_t grid_w = ign_cdiv(ign_cdiv(num_dst, RW), 16);
  if (grid_w > 2 * sms) grid_w = 2 * sms;                    // two CTAs per SM
  if (grid_w < 1) grid_w = 1;
  for (int t = max_steps - 1; t >= 0; --t) {
    gru_step_bwd_tc_kernel<<<grid, BW_THREADS, smem_a, st>>>(t, nt, off, meta4, steps_T, sp, h0, h_seq, d_out, dhs, dh0,
                                                              d_steps, g_rows, g_bound(num_dst), kernel, rkernel, bias, g_dbg);
    IGN_CHECK_LAUNCH("gru_step_bwd_tc");
    if (g_dbg & 16) continue;
    gru_dw_tc_kernel<<<(unsigned)grid_w, DW_THREADS, smem_w, st>>>(t, nt, off, meta4, steps_T, sp, h0, h_seq, g_rows,
                                                                    g_bound(num_dst), dK, dR, dB, g_dbg);
    IGN_CHECK_LAUNCH("gru_dw_tc");
  }
  return IGN_OK;
}
