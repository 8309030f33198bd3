// Dense layer on the 5th-generation tensor cores (tcgen05 + TMEM), fp32 in / fp32 out, sm_100a.
//
// y[M,N] = act(x[M,K] W[K,N] + b) for the readout / message / update MLPs of the generated model
// (reference code/utils/auxilary_classes.py:918-975, called at code/utils/generate_model.py:468,
// :600, :624).  The reference computes in fp32 and the parity bar is 1e-5 relative, so the GEMM
// runs as 3xTF32: every fp32 operand is split into hi = rna_tf32(x) and
// lo = rna_tf32(x - hi) (so hi + lo carries x to 2^-24), and  D = A_hi B_hi + A_lo B_hi + A_hi B_lo  accumulates in fp32 in TMEM
// (the dropped lo*lo term is ~2^-22 relative).
//
// Layout: one CTA owns a 128-row tile of x and all N <= 256 output columns; the accumulator is
// 128 TMEM lanes x N columns.  K is walked in chunks of 32 floats (= one 128-byte swizzle row):
//   * W is split and laid out ONCE per call by dense_tc_prep into the exact shared-memory image
//     of each chunk ([N rows][32 k] K-major, SWIZZLE_128B, hi image then lo image), so the main
//     kernel fetches B chunks with plain 16-byte cp.async;
//   * the x chunk is loaded with coalesced float4 loads, split in registers and stored into the
//     swizzled A_hi / A_lo images;
//   * one elected thread issues 12 tcgen05.mma (kind::tf32, M=128, N, K=8) per chunk and commits
//     to the stage's mbarrier; two shared-memory stages let the loads of chunk c+1 overlap the
//     MMAs of chunk c;
//   * epilogue: 8 warps read the accumulator with tcgen05.ld (32 lanes x 16 columns), add the
//     bias, apply the activation and write 64-byte row segments.
// Shapes outside (K % 32 == 0, N % 16 == 0, 16 <= N <= 256) use the fp32 CUDA-core path (dense.cu).

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_M = 128;
constexpr int TC_KC = 32;                       // floats per K chunk = 128 bytes per row
constexpr int A_IMG = TC_M * 128;               // bytes of one A image (hi or lo) of a chunk

// Fused gather + concat of the A operand (the message network's input, generate_model.py:432-475): row r of the
// [M, K] operand is  concat_k src[k][idx[k] ? idx[k][r] : r, :]  and is never written to memory -- the A loaders fetch
// each 32-float chunk straight from the source state array its columns belong to.  Every width is a multiple of 32 (a
// chunk lies inside one source); n == 0: plain x.  A negative index gives a zero row (as ign_gather_concat).
struct GatherA {
  const float* src[IGN_MAX_SOURCES];
  const int* idx[IGN_MAX_SOURCES];
  int width[IGN_MAX_SOURCES];
  int col0[IGN_MAX_SOURCES];
  int n;
};
// address of the 16-byte piece c4 of chunk c of operand row r, or nullptr for a zero row
__device__ __forceinline__ const float* a_piece(const GatherA& ga, const float* x, int K, int64_t r, int c, int c4) {
  if (ga.n == 0) return x + r * K + c * TC_KC + c4 * 4;
  const int col = c * TC_KC;
  int k = 0;
#pragma unroll
  for (int j = 1; j < IGN_MAX_SOURCES; ++j) k += (j < ga.n && col >= ga.col0[j]) ? 1 : 0;
  const float* base = k == 0 ? ga.src[0] : k == 1 ? ga.src[1] : k == 2 ? ga.src[2] : ga.src[3];
  const int* ix = k == 0 ? ga.idx[0] : k == 1 ? ga.idx[1] : k == 2 ? ga.idx[2] : ga.idx[3];
  const int w = k == 0 ? ga.width[0] : k == 1 ? ga.width[1] : k == 2 ? ga.width[2] : ga.width[3];
  const int c0 = k == 0 ? ga.col0[0] : k == 1 ? ga.col0[1] : k == 2 ? ga.col0[2] : ga.col0[3];
  const int64_t row = ix ? (int64_t)__ldg(ix + r) : r;
  if (row < 0) return nullptr;
  return base + row * w + (col - c0) + c4 * 4;
}

// W[K,N] -> per chunk c: [hi image: N rows x 128 B][lo image]  (the shared-memory layout of B)
// transposed: the GEMM uses W^T, i.e. element (k, n) of the [K, N] operand is w[n * K + k] (w stored [N, K])
__global__ void dense_tc_prep_kernel(const float* __restrict__ w, int K, int N, float* __restrict__ img,
                                     bool transposed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i % N;
  const float v = transposed ? w[(size_t)n * K + k] : w[i];
  float hi, lo;
  tf32_split(v, hi, lo);
  const int c = k / TC_KC, kk = k % TC_KC;
  char* base = reinterpret_cast<char*>(img) + (size_t)c * (2 * N * 128);
  *reinterpret_cast<float*>(base + sw128_off(n, kk)) = hi;
  *reinterpret_cast<float*>(base + N * 128 + sw128_off(n, kk)) = lo;
}

// ACT >= 0: compile-time activation (see act_epi); -1: taken from the argument
template <int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1) dense_tc_kernel(const float* __restrict__ x, int64_t M, int K,
                                                                 const float* __restrict__ wimg,
                                                                 const float* __restrict__ bias, int N, int act_,
                                                                 float* __restrict__ y, float* __restrict__ pre,
                                                                 int tmem_cols, const float* __restrict__ head_w,
                                                                 const float* __restrict__ head_b, float* __restrict__ head_out,
                                                                 const GatherA ga) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int act = ACT >= 0 ? ACT : act_;
  // carve-up (1024-byte aligned images): stage s: A_hi | A_lo | B_hi | B_lo
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_img = N * 128;
  const int stage_bytes = 2 * A_IMG + 2 * b_img;
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_acc;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[256];
  __shared__ __align__(16) float s_head[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid < N) {
    s_bias[tid] = bias ? bias[tid] : 0.0f;
    s_head[tid] = head_w ? head_w[tid] : 0.0f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc(N);
  const int nchunks = K / TC_KC;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  uint32_t use[2] = {0, 0};        // how many times each stage has been committed
  uint32_t acc_uses = 0;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m0 = tile * TC_M;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1;
      unsigned char* st = smem + s * stage_bytes;
      // the MMAs that last read this stage must have completed
      if (use[s] > 0) mbar_wait(&bar_stage[s], (use[s] - 1) & 1);
      // B chunk: straight copy of the prepared image (hi + lo)
      {
        const char* src = reinterpret_cast<const char*>(wimg) + (size_t)c * (2 * b_img);
        unsigned char* dst = st + 2 * A_IMG;
        for (int i = tid * 16; i < 2 * b_img; i += TC_THREADS * 16) cp_async16(dst + i, src + i);
        cp_async_commit();
      }
      // A chunk: 128 rows x 32 floats, split into hi / lo, swizzled store
      {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = tid + j * TC_THREADS;        // 0..1023 : (row, 16-byte column)
          const int r = idx >> 3, c4 = idx & 7;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r < M) {
            const float* pz = a_piece(ga, x, K, m0 + r, c, c4);
            if (pz) v = ldg_f4(pz);
          }
          float4 hi, lo;
          tf32_split(v.x, hi.x, lo.x);
          tf32_split(v.y, hi.y, lo.y);
          tf32_split(v.z, hi.z, lo.z);
          tf32_split(v.w, hi.w, lo.w);
          const int off = r * 128 + ((c4 ^ (r & 7)) << 4);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + A_IMG + off) = lo;
        }
      }
      cp_async_wait<0>();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(st), a_lo = a_hi + A_IMG;
        const uint32_t b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + b_img;
#pragma unroll
        for (int kk = 0; kk < TC_KC / 8; ++kk) {
          const uint32_t ko = kk * 32;                 // 8 tf32 = 32 bytes along K inside the swizzle row
          const uint32_t first = (c == 0 && kk == 0) ? 0u : 1u;
          umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(b_hi + ko), idesc, first);
          umma_tf32(tmem_d, umma_desc(a_lo + ko), umma_desc(b_hi + ko), idesc, 1u);
          umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(b_lo + ko), idesc, 1u);
        }
        umma_commit(&bar_stage[s]);
        if (c == nchunks - 1) umma_commit(&bar_acc);
      }
      use[s] += 1;
    }
    // epilogue: accumulator -> registers -> bias / activation -> global
    mbar_wait(&bar_acc, acc_uses & 1);
    acc_uses += 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const int q = warp & 3, half = warp >> 2;
      const int64_t row = m0 + q * 32 + lane;
      const int ncol_half = N / 2;
      float head_acc = 0.0f;
      for (int cb = 0; cb < ncol_half; cb += 32) {       // two 16-column TMEM loads in flight per wait
        const int col = half * ncol_half + cb;
        const bool second = cb + 16 < ncol_half;
        uint32_t ra[16], rb[16];
        tmem_ld16_nowait(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col, ra);
        if (second) tmem_ld16_nowait(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(col + 16), rb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
#pragma unroll
          for (int hblk = 0; hblk < 2; ++hblk) {
            if (hblk == 1 && !second) break;
            const int c0 = col + hblk * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + i);
              float4 v;
              v.x = __uint_as_float(hblk ? rb[i] : ra[i]) + bv.x;
              v.y = __uint_as_float(hblk ? rb[i + 1] : ra[i + 1]) + bv.y;
              v.z = __uint_as_float(hblk ? rb[i + 2] : ra[i + 2]) + bv.z;
              v.w = __uint_as_float(hblk ? rb[i + 3] : ra[i + 3]) + bv.w;
              if (pre) st_f4(pre + row * N + c0 + i, v);
              v.x = act_epi(act, v.x); v.y = act_epi(act, v.y); v.z = act_epi(act, v.z); v.w = act_epi(act, v.w);
              if (y) st_f4(y + row * N + c0 + i, v);
              if (head_out) {
                const float4 hw = *reinterpret_cast<const float4*>(s_head + c0 + i);
                head_acc = fmaf(v.x, hw.x, head_acc); head_acc = fmaf(v.y, hw.y, head_acc);
                head_acc = fmaf(v.z, hw.z, head_acc); head_acc = fmaf(v.w, hw.w, head_acc);
              }
            }
          }
        }
      }
      // fused single-output head: the two column halves of a row add their partial dot products
      // (two commutative float adds onto a zeroed output: deterministic)
      if (head_out && row < M) atomicAdd(head_out + row, head_acc + ((half == 0 && head_b) ? __ldg(head_b) : 0.0f));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                 // accumulator drained before the next tile's MMAs
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)tmem_cols));
  }
}


// ---------------------------------------------------------------------------------------------------------
// Pipelined variant for large M (the train step's readout layers and their dX): the same math, warp-specialised
// so that nothing waits on a CTA-wide barrier inside the tile loop:
//   warps 0-3   A producers: x chunk [128 rows x 32] -> hi / lo images of the stage (next job's rows already in
//               registers), arrive on full[stage]
//   warp  4     TMA producer: the prepared weight chunk (hi + lo, 2 N 128 bytes) as one cp.async.bulk per job
//   warp  5     MMA issuer: 12 tcgen05.mma per job, commit -> stage_free[stage]; last chunk of a tile -> acc_full[buf]
//   warps 8-15  epilogue of the PREVIOUS tile (two TMEM accumulators): tcgen05.ld -> bias -> (pre) -> activation ->
//               32-column chunks through a per-warp staging block so that 8 lanes write each 128-byte row segment
// N % 64 == 0 or N == 32, N <= 256.
constexpr int P_PROD_WARPS = 4, P_TMA_WARP = 4, P_MMA_WARP = 5, P_EPI_WARP0 = 8, P_EPI_WARPS = 8;
constexpr int P_THREADS = 32 * (P_EPI_WARP0 + P_EPI_WARPS);
constexpr int P_STAGING = P_EPI_WARPS * 4096;

template <int ACT>
__global__ void __launch_bounds__(P_THREADS, 1) dense_pipe_tc_kernel(const float* __restrict__ x, int64_t M, int K,
                                                                     const float* __restrict__ wimg,
                                                                     const float* __restrict__ bias, int N, int act_,
                                                                     float* __restrict__ y, float* __restrict__ pre,
                                                                     int nst, int tmem_cols, const GatherA ga) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int act = ACT >= 0 ? ACT : act_;
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_img = N * 128;
  const int stage_bytes = 2 * A_IMG + 2 * b_img;
  unsigned char* staging = smem + (size_t)nst * stage_bytes;
  __shared__ uint64_t bar_full[4], bar_b[4], bar_free[4], bar_acc_full[2], bar_acc_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bar_full[s], P_PROD_WARPS);
      mbar_init(&bar_b[s], 1);
      mbar_init(&bar_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_acc_full[b], 1);
      mbar_init(&bar_acc_free[b], P_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == P_MMA_WARP) tmem_alloc(&tmem_base_s, (uint32_t)tmem_cols);
  if (tid < N) s_bias[tid] = bias ? bias[tid] : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_stride = tmem_cols / 2;
  const int nchunks = K / TC_KC;
  const int64_t ntiles = (M + TC_M - 1) / TC_M;
  const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t njobs = my_tiles * nchunks;

  if (warp < P_PROD_WARPS) {
    // ================================ A producers ================================
    const int r0 = tid >> 3, c4 = tid & 7;                   // rows r0 + 16 k
    float4 cur[8], nxt[8];
    auto load = [&](int64_t job, float4 (&v)[8]) {
      const int64_t m0 = (blockIdx.x + (job / nchunks) * gridDim.x) * TC_M;
      const int c = (int)(job % nchunks);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int64_t r = m0 + r0 + 16 * k;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < M) {
          const float* pz = a_piece(ga, x, K, r, c, c4);
          // gathered state rows are read again by other edges: keep them cacheable; a plain x is streamed
          if (pz) v[k] = ga.n ? ldg_f4(pz) : ld_stream_f4(pz);
        }
      }
    };
    if (njobs > 0) load(0, cur);
    for (int64_t job = 0; job < njobs; ++job) {
      const int s = (int)(job % nst);
      if (job + 1 < njobs) load(job + 1, nxt);
      if (job >= nst) mbar_wait(&bar_free[s], (uint32_t)((job / nst) - 1) & 1);
      unsigned char* st = smem + (size_t)s * stage_bytes;
#pragma unroll
      for (int k = 0; k < 8; ++k) store_split(st, st + A_IMG, r0 + 16 * k, c4, cur[k]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
#pragma unroll
      for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
    }
  } else if (warp == P_TMA_WARP) {
    // ================================ weight chunks ================================
    if (lane == 0) {
      const uint32_t bytes = 2u * (uint32_t)b_img;
      for (int64_t job = 0; job < njobs; ++job) {
        const int s = (int)(job % nst);
        if (job >= nst) mbar_wait(&bar_free[s], (uint32_t)((job / nst) - 1) & 1);
        mbar_expect_tx(&bar_b[s], bytes);
        bulk_g2s(smem + (size_t)s * stage_bytes + 2 * A_IMG,
                 reinterpret_cast<const char*>(wimg) + (size_t)(job % nchunks) * bytes, bytes, &bar_b[s]);
      }
    }
    __syncwarp();
  } else if (warp == P_MMA_WARP) {
    // ================================ MMA issuer ================================
    for (int64_t job = 0; job < njobs; ++job) {
      const int s = (int)(job % nst);
      const int64_t tl = job / nchunks;                      // index of the tile among this CTA's tiles
      const int c = (int)(job % nchunks), buf = (int)(tl & 1);
      const uint32_t ph = (uint32_t)(job / nst) & 1;
      mbar_wait(&bar_full[s], ph);
      mbar_wait(&bar_b[s], ph);
      if (c == 0 && tl >= 2) mbar_wait(&bar_acc_free[buf], (uint32_t)((tl >> 1) - 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), b_hi = a_hi + 2 * A_IMG;
        umma_chunk_3x(tmem_base + buf * acc_stride, a_hi, a_hi + A_IMG, b_hi, b_hi + b_img, N, c > 0);
        umma_commit(&bar_free[s]);
        if (c == nchunks - 1) umma_commit(&bar_acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= P_EPI_WARP0) {
    // ================================ epilogue ================================
    const int e = warp - P_EPI_WARP0, q = warp & 3, half = e >> 2;
    unsigned char* my_stage = staging + e * 4096;
    // N = 32: the four warps of half 0 take the whole row, the others only hand the accumulator back
    const bool solo = N == 32;
    const int ncol_half = solo ? 32 : N / 2;
    // one 32-column chunk of this warp's 32 rows -> global, 8 lanes per 128-byte row segment
    auto emit = [&](const float (&v)[32], float* __restrict__ dst, int64_t m0, int col0) {
      __syncwarp();
      unsigned char* srow = my_stage + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(srow + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = (lane >> 3) + 4 * k, ch = lane & 7;
        const int64_t grow = m0 + q * 32 + rr;
        const float4 val = *reinterpret_cast<const float4*>(my_stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
        if (grow < M) st_f4(dst + grow * N + col0 + ch * 4, val);
      }
    };
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int buf = (int)(tl & 1);
      const int64_t m0 = (blockIdx.x + tl * gridDim.x) * TC_M;
      mbar_wait(&bar_acc_full[buf], (uint32_t)(tl >> 1) & 1);
      tc_fence_after();
      if (solo && half == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_free[buf]);
        continue;
      }
      const uint32_t tb = tmem_base + buf * acc_stride + ((uint32_t)(q * 32) << 16);
      for (int cb = 0; cb < ncol_half; cb += 32) {
        const int col0 = (solo ? 0 : half * ncol_half) + cb;
        uint32_t ra[16], rb[16];
        tmem_ld16_nowait(tb + col0, ra);
        tmem_ld16_nowait(tb + col0 + 16, rb);
        tmem_ld_wait();
        if (cb + 32 >= ncol_half) {                          // the accumulator has been read: free for the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_free[buf]);
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] = __uint_as_float(ra[i]) + s_bias[col0 + i];
          v[16 + i] = __uint_as_float(rb[i]) + s_bias[col0 + 16 + i];
        }
        if (pre) emit(v, pre, m0, col0);
        if (y) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_epi(act, v[i]);
          emit(v, y, m0, col0);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P_MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

}  // namespace

// true when the tensor-core path is built for this shape
bool ign_dense_tc_supported(int k, int n) { return k % TC_KC == 0 && k >= TC_KC && n % 32 == 0 && n >= 32 && n <= 256; }

size_t ign_dense_tc_ws(int k, int n) { return (size_t)(k / TC_KC) * 2 * n * 128; }

static int dense_tc_launch_g(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                             float* y, float* pre_act, void* ws, cudaStream_t st, const float* head_w,
                             const float* head_b, float* head_out, bool w_transposed, const GatherA* gather) {
  GatherA ga;
  memset(&ga, 0, sizeof(ga));
  if (gather) ga = *gather;
  float* img = reinterpret_cast<float*>(ws);
  dense_tc_prep_kernel<<<(unsigned)ign_cdiv((int64_t)k * n, 256), 256, 0, st>>>(w, k, n, img, w_transposed);
  IGN_CHECK_LAUNCH("dense_tc_prep");
  const size_t smem = 1024 + 2 * (size_t)(2 * A_IMG + 2 * n * 128);
  int cols = 32;
  while (cols < n) cols <<= 1;
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(m, TC_M);
  const int grid = (int)(tiles < sms ? tiles : sms);
  if (head_out) IGN_CUDA(cudaMemsetAsync(head_out, 0, (size_t)m * sizeof(float), st));
  if (!head_out && (n % 64 == 0 || n == 32) && m >= 4096) {
    // large M: the warp-specialised pipeline (two accumulators, weight chunks by TMA, staged coalesced stores)
    const size_t stage_bytes = 2 * (size_t)A_IMG + 2 * (size_t)n * 128;
    int nst = (int)((227 * 1024 - 1024 - 2048 - (size_t)P_STAGING) / stage_bytes);
    if (nst > 4) nst = 4;
    if (nst >= 2) {
      const size_t psmem = 1024 + (size_t)nst * stage_bytes + P_STAGING;
      int pcols = 64;
      while (pcols < 2 * n) pcols <<= 1;
      auto plaunch = [&](auto kernel) -> int {
        IGN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        kernel<<<grid, P_THREADS, psmem, st>>>(x, m, k, img, bias, n, act, y, pre_act, nst, pcols, ga);
        return IGN_OK;
      };
      int prc;
      switch (act) {
        case IGN_ACT_LINEAR: prc = plaunch(dense_pipe_tc_kernel<IGN_ACT_LINEAR>); break;
        case IGN_ACT_RELU: prc = plaunch(dense_pipe_tc_kernel<IGN_ACT_RELU>); break;
        case IGN_ACT_SELU: prc = plaunch(dense_pipe_tc_kernel<IGN_ACT_SELU>); break;
        case IGN_ACT_TANH: prc = plaunch(dense_pipe_tc_kernel<IGN_ACT_TANH>); break;
        default: prc = plaunch(dense_pipe_tc_kernel<-1>); break;
      }
      if (prc != IGN_OK) return prc;
      IGN_CHECK_LAUNCH("dense_pipe_tc");
      return IGN_OK;
    }
  }
  auto launch = [&](auto kernel) -> int {
    IGN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, TC_THREADS, smem, st>>>(x, m, k, img, bias, n, act, y, pre_act, cols, head_w, head_b, head_out, ga);
    return IGN_OK;
  };
  int rc;
  switch (act) {
    case IGN_ACT_LINEAR: rc = launch(dense_tc_kernel<IGN_ACT_LINEAR>); break;
    case IGN_ACT_RELU: rc = launch(dense_tc_kernel<IGN_ACT_RELU>); break;
    case IGN_ACT_SELU: rc = launch(dense_tc_kernel<IGN_ACT_SELU>); break;
    case IGN_ACT_SIGMOID: rc = launch(dense_tc_kernel<IGN_ACT_SIGMOID>); break;
    case IGN_ACT_TANH: rc = launch(dense_tc_kernel<IGN_ACT_TANH>); break;
    default: rc = launch(dense_tc_kernel<-1>); break;
  }
  if (rc != IGN_OK) return rc;
  IGN_CHECK_LAUNCH("dense_tc");
  return IGN_OK;
}

int ign_dense_tc_launch(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                        float* y, float* pre_act, void* ws, cudaStream_t st, const float* head_w,
                        const float* head_b, float* head_out, bool w_transposed) {
  return dense_tc_launch_g(x, m, k, w, bias, n, act, y, pre_act, ws, st, head_w, head_b, head_out, w_transposed, nullptr);
}

// ---------------------------------------------------------------------------------------------------------
// First Dense layer of a message network with its input gathered on the fly (see GatherA above).
extern "C" size_t ign_gather_dense_ws_bytes(int n_src, const int32_t* widths, int n) {
  if (n_src < 1 || n_src > IGN_MAX_SOURCES || !widths) return 0;
  int k = 0;
  for (int j = 0; j < n_src; ++j) {
    if (widths[j] <= 0 || widths[j] % TC_KC) return 0;
    k += widths[j];
  }
  return (k <= 256 && ign_dense_tc_supported(k, n)) ? ign_dense_tc_ws(k, n) : 0;
}

bool ign_tensor_cores_enabled();

extern "C" int ign_gather_dense(int n_src, const float* const* srcs, const int32_t* const* idx, const int32_t* widths,
                                int64_t rows, const float* w, const float* bias, int n, int act, float* y, void* ws,
                                size_t ws_bytes, void* stream) {
  IGN_REQUIRE(rows >= 0 && n > 0, IGN_ERR_INVALID, "IGNNITION: gather_dense: bad shape");
  IGN_REQUIRE(act >= IGN_ACT_LINEAR && act <= IGN_ACT_LEAKY_RELU, IGN_ERR_INVALID,
              "IGNNITION: gather_dense: unknown activation %d", act);
  const size_t need = ign_gather_dense_ws_bytes(n_src, widths, n);
  IGN_REQUIRE(need > 0 && rows >= TC_M && ign_tensor_cores_enabled(), IGN_ERR_UNSUPPORTED,
              "IGNNITION: gather_dense: built for 1..%d sources of widths that are multiples of 32 (sum <= 256), "
              "32 <= units <= 256, at least 128 rows, tensor cores on", IGN_MAX_SOURCES);
  IGN_REQUIRE(srcs && idx && w && y, IGN_ERR_INVALID, "IGNNITION: gather_dense: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= need, IGN_ERR_WORKSPACE, "IGNNITION: gather_dense: workspace too small (%zu < %zu)",
              ws_bytes, need);
  GatherA ga;
  memset(&ga, 0, sizeof(ga));
  ga.n = n_src;
  int k = 0;
  for (int j = 0; j < n_src; ++j) {
    IGN_REQUIRE(srcs[j], IGN_ERR_INVALID, "IGNNITION: gather_dense: null source %d", j);
    ga.src[j] = srcs[j];
    ga.idx[j] = idx[j];
    ga.width[j] = widths[j];
    ga.col0[j] = k;
    k += widths[j];
  }
  return dense_tc_launch_g(nullptr, rows, k, w, bias, n, act, y, nullptr, ws, ign_stream(stream), nullptr, nullptr,
                           nullptr, false, &ga);
}
