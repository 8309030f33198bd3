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
// Tiles of 128 rows are independent, so the kernel is a warp-specialised pipeline over two slots
// (operand stage + TMEM accumulator), synchronised only by mbarriers:
//   warps 0-3 / 4-7 : producer group of slot 0 / slot 1 -- gather the x_t rows by the step-major table
//                     (8 lanes per 128-byte row: coalesced), read the h rows, split both into the
//                     swizzled hi / lo operand images of the slot, arrive on full[slot];
//   warp 8          : waits full[slot], issues 24 tcgen05.mma (3xTF32, N = 128: z | r | xh | hh in one
//                     accumulator, K and R images resident in shared memory), commits to acc_full[slot];
//   warps 9-16      : epilogue -- tcgen05.ld of the gates, sigmoid / tanh, new state into a staging tile,
//                     coalesced write to hs / out / h_seq, arrive on slot_free[slot].
// The tensor pipe needs 24 x 71 = 1.7 k cycles per tile (profiles/r1_umma_tf32_rate.md); producers of the
// other slot and the epilogue of the previous tile run underneath.

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int ROWS = 128;
constexpr int U = 32;
constexpr int IMG = ROWS * 128;           // bytes of one [128 x 32] fp32 image
constexpr int STAGE = 4 * IMG;            // Ax_hi | Ax_lo | Ah_hi | Ah_lo
constexpr int PROD_WARPS = 4;             // per slot
constexpr int EPI_WARPS = 8;
constexpr int MMA_WARP = 2 * PROD_WARPS;
constexpr int EPI_WARP0 = MMA_WARP + 1;
constexpr int TC_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);    // 17 warps
constexpr int OUT_STRIDE = U + 4;         // floats: staging tile row stride (bank-conflict-free float4 rows)

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};
__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}
// h operand: lo = v - hi kept exactly (the tensor core truncates it to TF32 on read), so that the
// epilogue can rebuild the old state bit for bit as hi + lo
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

__global__ void __launch_bounds__(TC_THREADS, 1) gru_step_tc_kernel(
    int t, const int* __restrict__ nt, const int* __restrict__ off, int64_t num_dst, const int4* __restrict__ meta,
    const int* __restrict__ steps_T, SrcPtrs srcs, const float* __restrict__ h0, float* __restrict__ hs,
    float* __restrict__ out, float* __restrict__ h_seq, const float* __restrict__ kernel,
    const float* __restrict__ rkernel, const float* __restrict__ bias) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // weight images, 128 rows each: Bx = [K_z | K_r | K_h | 0], Bh = [R_z | R_r | 0 | R_h]  (row n = gate column n)
  unsigned char* bx_hi = smem;
  unsigned char* bx_lo = bx_hi + IMG;
  unsigned char* bh_hi = bx_lo + IMG;
  unsigned char* bh_lo = bh_hi + IMG;
  unsigned char* stages = bh_lo + IMG;                       // 2 slots x STAGE
  __shared__ uint64_t bar_full[2], bar_acc[2], bar_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];                // merged gate biases [bz | br | bxh | bhh]
  __shared__ int4 s_meta[2][ROWS];                           // per slot: (d, lo, len, alive) of every row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], PROD_WARPS);
      mbar_init(&bar_acc[s], 1);
      mbar_init(&bar_free[s], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < 4 * IMG / 16; i += TC_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < U * 3 * U; i += TC_THREADS) {
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
  const int64_t n_alive = (int64_t)__ldg(nt + t);
  const int64_t n_rows = (t == 0) ? num_dst : n_alive;
  const int* entries = steps_T + __ldg(off + t);
  const int* entries_safe = n_alive > 0 ? entries : steps_T;   // always readable at index 0
  const int64_t ntiles = (n_rows + ROWS - 1) / ROWS;
  const int G = gridDim.x;
  // this CTA's tiles: blockIdx.x + j*G, j = 0, 1, ...; tile j lives in slot j & 1
  const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + G - 1) / G : 0;

  if (warp < MMA_WARP) {
    // ================================ producers: group g fills slot g ================================
    const int g = warp / PROD_WARPS;
    const int ptid = tid - g * PROD_WARPS * 32;              // 0..127 inside the group
    unsigned char* st = stages + g * STAGE;
    const int c4 = ptid & 7;                                 // 16-byte chunk of the row: 8 lanes read one row
    for (int64_t j = g, n = 0; j < my_tiles; j += 2, ++n) {
      const int64_t tile = blockIdx.x + j * G;
      // three batches of independent loads (clamped addresses, masked afterwards) so that the eight rows of
      // this thread cost two memory latencies, not sixteen: {meta, step entry} -> {h row, x row}
      int4 m[8];
      int ent[8];
      bool inb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {                          // rows (ptid >> 3) + 16 k
        const int64_t i = tile * ROWS + (ptid >> 3) + 16 * k;
        inb[k] = i < n_rows;
        m[k] = __ldg(meta + (inb[k] ? i : n_rows - 1));
        ent[k] = __ldg(entries_safe + ((i < n_alive) ? i : 0));
      }
      float4 xv[8], hv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int64_t i = tile * ROWS + (ptid >> 3) + 16 * k;
        const bool alive = i < n_alive;
        const int e = alive ? ent[k] : IGN_STEP_ZERO;
        const float* hp = (t == 0) ? h0 + (int64_t)m[k].x * U : hs + (inb[k] ? i : n_rows - 1) * U;
        const float* xp = e >= 0 ? pick_src(srcs, e >> IGN_STEP_SRC_SHIFT) + (int64_t)(e & IGN_STEP_ROW_MASK) * U : srcs.p[0];
        hv[k] = ldg_f4(hp + c4 * 4);
        xv[k] = ldg_f4(xp + c4 * 4);
        if (e < 0) xv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!inb[k]) { hv[k] = make_float4(0.f, 0.f, 0.f, 0.f); m[k] = make_int4(-1, 0, 0, 0); }
        m[k].w = alive ? 1 : 0;
      }
      if (n > 0) mbar_wait(&bar_free[g], (uint32_t)(n - 1) & 1);   // epilogue of this slot's previous tile done
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = (ptid >> 3) + 16 * k;
        store_split(st, st + IMG, r, c4, xv[k]);
        store_split_exact(st + 2 * IMG, st + 3 * IMG, r, c4, hv[k]);   // hi + lo == h exactly (epilogue reads it back)
        if (c4 == 0) s_meta[g][r] = m[k];
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[g]);
    }
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer ================================
    for (int64_t j = 0; j < my_tiles; ++j) {
      const int s = (int)(j & 1);
      mbar_wait(&bar_full[s], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t ax_hi = smem_u32(stages + s * STAGE), ax_lo = ax_hi + IMG, ah_hi = ax_lo + IMG, ah_lo = ah_hi + IMG;
        const uint32_t d = tmem_base + s * 128;
        umma_chunk_3x(d, ax_hi, ax_lo, smem_u32(bx_hi), smem_u32(bx_lo), 128, false);    // z | r | xh | 0
        umma_chunk_3x(d, ah_hi, ah_lo, smem_u32(bh_hi), smem_u32(bh_lo), 128, true);     // z | r | 0  | hh
        umma_commit(&bar_acc[s]);
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue ================================
    const int e = warp - EPI_WARP0;                          // 0..7
    const int q = warp & 3;                                  // TMEM lane group this warp may access
    const int half = e >> 2;
    const int row = q * 32 + lane;
    const int etid = tid - EPI_WARP0 * 32;                   // 0..255
    for (int64_t j = 0; j < my_tiles; ++j) {
      const int s = (int)(j & 1);
      const uint32_t ph = (uint32_t)(j >> 1) & 1;
      const int64_t tile = blockIdx.x + j * G;
      unsigned char* st = stages + s * STAGE;
      float* stg = reinterpret_cast<float*>(st);             // staging tile [128][OUT_STRIDE] over the x images
      mbar_wait(&bar_full[s], ph);                           // s_meta of this tile is visible
      mbar_wait(&bar_acc[s], ph);                            // gates are in TMEM, the A images are free
      tc_fence_after();
      const int4 mr = s_meta[s][row];
      const uint32_t tb = tmem_base + s * 128 + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int ub = 0; ub < 16; ub += 8) {
        const int u0 = half * 16 + ub;
        uint32_t az[8], ar[8], axh[8], ahh[8];
        tmem_ld8_nowait(tb + u0, az);
        tmem_ld8_nowait(tb + 32 + u0, ar);
        tmem_ld8_nowait(tb + 64 + u0, axh);
        tmem_ld8_nowait(tb + 96 + u0, ahh);
        float hold[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {                        // old state = hi + lo of the h operand image
          const int o = row * 128 + ((((u0 >> 2) + c) ^ (row & 7)) << 4);
          const float4 a = *reinterpret_cast<const float4*>(st + 2 * IMG + o);
          const float4 b = *reinterpret_cast<const float4*>(st + 3 * IMG + o);
          hold[4 * c] = a.x + b.x; hold[4 * c + 1] = a.y + b.y; hold[4 * c + 2] = a.z + b.z; hold[4 * c + 3] = a.w + b.w;
        }
        tmem_ld_wait();
        float hn[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; j4 += 4) {
          const float4 vz = *reinterpret_cast<const float4*>(s_gb + u0 + j4);
          const float4 vr = *reinterpret_cast<const float4*>(s_gb + U + u0 + j4);
          const float4 vx = *reinterpret_cast<const float4*>(s_gb + 2 * U + u0 + j4);
          const float4 vh = *reinterpret_cast<const float4*>(s_gb + 3 * U + u0 + j4);
          const float bz[4] = {vz.x, vz.y, vz.z, vz.w}, br[4] = {vr.x, vr.y, vr.z, vr.w};
          const float bxh[4] = {vx.x, vx.y, vx.z, vx.w}, bhh[4] = {vh.x, vh.y, vh.z, vh.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int u = j4 + jj;
            const float z = fast_sigmoid(__uint_as_float(az[u]) + bz[jj]);
            const float r = fast_sigmoid(__uint_as_float(ar[u]) + br[jj]);
            const float hh = fast_tanh(fmaf(r, __uint_as_float(ahh[u]) + bhh[jj], __uint_as_float(axh[u]) + bxh[jj]));
            hn[u] = mr.w ? fmaf(z, hold[u] - hh, hh) : hold[u];
          }
        }
        // the x images are free (MMAs done): new state into the staging tile
        st_f4(stg + row * OUT_STRIDE + u0, make_float4(hn[0], hn[1], hn[2], hn[3]));
        st_f4(stg + row * OUT_STRIDE + u0 + 4, make_float4(hn[4], hn[5], hn[6], hn[7]));
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");      // staging tile complete
#pragma unroll
      for (int k = 0; k < 4; ++k) {                          // coalesced write-out: 8 lanes per row
        const int idx = etid + k * (EPI_WARPS * 32);
        const int r = idx >> 3, c4 = idx & 7;
        const int4 m = s_meta[s][r];
        if (m.x < 0) continue;
        const float4 v = *reinterpret_cast<const float4*>(stg + r * OUT_STRIDE + c4 * 4);
        const int64_t i = tile * ROWS + r;
        if (m.z > t + 1) st_f4(hs + i * U + c4 * 4, v);                   // more steps to come
        else st_f4(out + (int64_t)m.x * U + c4 * 4, v);                    // final state of the destination
        if (h_seq && m.w) st_f4(h_seq + (int64_t)(m.y + t) * U + c4 * 4, v);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_free[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, 256);
}

}  // namespace

int ign_gru_step_tc_launch(int t, const int* nt, const int* off, int64_t num_dst, int64_t rows_bound, const int* meta,
                           const int* steps_T, int n_src, const float* const* srcs, const float* h0, float* hs,
                           float* out, float* h_seq, const float* kernel, const float* rkernel, const float* bias,
                           cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  const size_t smem = 1024 + 4 * (size_t)IMG + 2 * (size_t)STAGE;
  if (IGN_ONCE_PER_DEVICE()) {
    IGN_CUDA(cudaFuncSetAttribute(gru_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
