// Fused gather + segmented aggregation + GRU update + state exchange on sm_100a (one kernel per
// message passing of a sum / mean / max aggregation with a recurrent update):
//
//   agg[d]  = op_{slots of d} src_states[col[slot]]            generate_model.py:432,479-490 + Sum_aggr
//   h'[d]   = GRUCell(agg[d], h[d])                            auxilary_classes.py:254-262, 752-765
//   outs[k][out_row0 + d] = h'[d]  for k < n_out               generate_model.py:602 (state write-back)
//
// The aggregated messages never leave the SM: gather warps reduce the rows of a 128-destination tile in
// registers (one sub-warp of U/4 lanes per destination, single accumulator in slot order = the
// sequential fp32 sum of ign_segment_reduce, bit for bit) and drop them, split hi / lo, straight into
// the swizzled A-operand images of the gate GEMMs (3xTF32 tcgen05.mma, accumulator [z | r | xh | hh]
// in TMEM, same math as gru_cell_tc.cu).  The new states leave through a swizzled staging tile and
// TMA TENSOR stores (cp.async.bulk.tensor.2d, one per output buffer): with n_out > 1 the outputs are
// the caller's own state buffer and every peer GPU's mapped copy (cudaIpc over NVLink), i.e. the
// all-gather of a destination-partitioned graph happens tile by tile from the epilogue of the kernel
// that computes the states, while the gather warps are already reducing the next tile.
//
// Roles of the 512-thread persistent CTA (one per SM), mbarriers between them:
//   warps 0-7    gather: rowptr slice -> edge-balanced row ranges per sub-warp -> flat walk over the tile's
//                slots through a cp.async ring in shared memory (16 rows in flight per sub-warp behind the 4
//                being summed, column indices one chunk ahead) -> x operand images
//   warps 8-11   epilogue: TMEM -> gates -> staging tile -> TMA stores (thread = one destination)
//   warp  12     MMA issuer (one lane)
//   warp  13     TMA producer of the prepared weight chunks (gru_cell_tc prep layout)
//   warps 14-15  h loaders: the old state rows -> h operand images
// HBM-bound by the gather: algorithmic bytes E (4 + 4F) + n (8U + 4)  (DESIGN.md section 4).

#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace ign_tc;

// prepared weight images of gru_cell_tc.cu (same layout: chunk-major, [3U rows][32] hi then lo)
size_t ign_gru_cell_tc_ws(int units);
int ign_gru_cell_tc_prep(const float* kernel, const float* rkernel, int units, void* ws, cudaStream_t st);

namespace {

constexpr int GATHER_WARPS = 8;
constexpr int GATHER_THREADS = GATHER_WARPS * 32;
constexpr int EPI_WARP0 = 8, EPI_THREADS = 128;
constexpr int MMA_WARP = 12, TMA_WARP = 13, HLOAD_WARP0 = 14, HLOAD_THREADS = 64;
constexpr int AGG_THREADS = 512;
constexpr int ROWS = 128;
constexpr int A_IMG = ROWS * 128;                       // one [128 x 32] fp32 image
constexpr int BATCH = 4;                                // gathered rows per cp.async group
// -DIGN_AGG_PROBE_NB=n: profiling build for the gather-only mode (IGN_AGG_DBG=1): n batches per ring, the ring laid
// over the (then unused) operand stages so that it fits (profiles/r2_agg_gru.md)
#ifdef IGN_AGG_PROBE_NB
constexpr int NB = IGN_AGG_PROBE_NB;
constexpr bool PROBE = true;
#else
constexpr int NB = 5;                                   // batches in a sub-warp's ring: NB - 1 in flight
constexpr bool PROBE = false;
#endif
constexpr int BAR_GATHER = 1, BAR_EPI = 2, BAR_HLOAD = 3;   // named barriers (0 = __syncthreads)

struct OutMaps {
  CUtensorMap m[IGN_MAX_PEERS];
  float* p[IGN_MAX_PEERS];                  // the same arrays as plain pointers (1-D bulk stores)
};
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// mbarrier wait of a role that waits for most of a tile's time: back off between polls so that the spin does not
// take issue slots from the gather warps (profiles/r2_agg_gru.md: 56 % of the warp instructions were polls)
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)      // suspend-time hint, ns: sleep in hardware, wake on arrive
        : "memory");
  }
}

// explicit shared-space accesses (through a generic pointer the ring reads compiled to generic LD.E)
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
  return r;
}
__device__ __forceinline__ void sts_f4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t saddr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem) : "memory");
}

template <int OP>
__device__ __forceinline__ void acc_row(float4& a, const float4& v) {
  if (OP == IGN_OP_MAX) {
    a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
  } else {
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
}

// TMA_OUT: the new states leave through the swizzled staging tile and TMA tensor stores (any n_out); otherwise
// (one output, U = 64) the epilogue stores its rows itself and the staging tile's shared memory holds XRAW.
// XRAW: the gather warps park the aggregated rows of a tile as plain fp32 and convert them into the operand images
// at the end of the tile, so that reducing tile t + 1 does not wait for the UMMAs that still read tile t's images.
template <int U, int OP, bool TMA_OUT>
__global__ void __launch_bounds__(AGG_THREADS, 1) agg_gru_tc_kernel(
    const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ src,
    const float* __restrict__ h, int64_t n, const float* __restrict__ wimg, const float* __restrict__ bias,
    const __grid_constant__ OutMaps maps, int n_out, int out_row0, float* __restrict__ agg_out,
    float* __restrict__ out_direct, int dbg) {
  constexpr int NC = U / 32;                 // K chunks per operand
  constexpr int B_IMG = 3 * U * 128;         // one weight image (hi or lo) of a chunk
  constexpr int STAGE_A = 2 * A_IMG;         // operand stage: hi + lo image of one 32-column chunk
  constexpr int DCOLS = 4 * U;               // accumulator columns
  constexpr int G = U / 4;                   // gather lanes per destination
  constexpr int NG = GATHER_THREADS / G;     // destinations reduced at the same time
  constexpr bool XRAW = (U == 32) || !TMA_OUT;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* b_buf = smem + 2 * STAGE_A;             // ONE weight chunk (hi + lo): the tile time is the gather's
  unsigned char* out_stage = b_buf + 2 * B_IMG;          // NC boxes of [128 x 32] fp32, SWIZZLE_128B
  unsigned char* ring = PROBE ? smem : out_stage + NC * A_IMG;   // gathered rows in flight: NG rings of NB x BATCH rows
  // [128][U] fp32 aggregated rows of the tile being reduced: its own region at U = 32, the staging tile's at U = 64
  float* xraw = reinterpret_cast<float*>(PROBE ? smem + NG * (NB * BATCH * U * 4)
                                               : U == 32 ? ring + NG * (NB * BATCH * U * 4) : out_stage);
  constexpr int RING_GROUP = NB * BATCH * U * 4;
  // bar_mma[c]: the UMMAs of chunk c of a tile are done (phase = tile count of the CTA).  One barrier per chunk of
  // the tile, not per stage: a stage is used twice per tile at U = 64 (x chunk, then h chunk) and its two users wait
  // for each other's use only -- with one barrier per stage the gather warps would test a parity two phases back
  // and pass while the x chunk they overwrite is still being read.
  __shared__ uint64_t bar_mma[4];
  __shared__ uint64_t bar_full[2];           // operand images of the stage are in place
  __shared__ uint64_t bar_b;                 // weight chunk landed (complete_tx)
  __shared__ uint64_t bar_acc[2];            // accumulator complete
  __shared__ uint64_t bar_drained[2];        // accumulator read by every epilogue thread
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_gb[4 * U];            // merged gate biases [bz | br | bxh | bhh]
  __shared__ int s_rp[ROWS + 2];                         // rowptr slice of the tile being gathered

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_acc[i], 1); mbar_init(&bar_drained[i], EPI_THREADS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bar_mma[i], 1);
    mbar_init(&bar_b, 1);
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
  // every role walks the same chunk sequence: chunk ctr of the CTA (= chunk ctr % 2NC of its tile ctr / 2NC) uses
  // stage ctr & 1 for the (ctr >> 1)-th time
  uint32_t ctr = 0;
  uint32_t lt = 0;                           // tiles this CTA has finished (phase of the per-chunk barriers)
  // IGN_AGG_DBG=1 (profiling): only the gather warps run, nothing waits for anything (results are garbage)
  const bool gather_only = dbg & 1;
  const bool bulk1d = !(dbg & 2);            // IGN_AGG_DBG=2: tensor stores (cp.async.bulk.tensor) instead of 1-D bulk stores
  const int64_t ntiles_other = gather_only ? 0 : ntiles;

  if (warp == MMA_WARP) {
    int ab = 0;
    uint32_t acc_uses[2] = {0, 0};
    for (int64_t tile = blockIdx.x; tile < ntiles_other; tile += gridDim.x, ab ^= 1) {
      const uint32_t d = tmem_base + ab * DCOLS;
      for (int c = 0; c < 2 * NC; ++c, ++ctr) {
        const int s = ctr & 1;
        const uint32_t use = ctr >> 1;
        if (lane == 0) {
          unsigned char* st = smem + s * STAGE_A;
          if (c == 0 && acc_uses[ab] > 0) mbar_wait(&bar_drained[ab], (acc_uses[ab] - 1) & 1);
          mbar_wait_idle(&bar_full[s], use & 1);
          mbar_wait(&bar_b, ctr & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + A_IMG, b_hi = smem_u32(b_buf), b_lo = b_hi + B_IMG;
          if (c < NC) {
            umma_chunk_3x(d, a_hi, a_lo, b_hi, b_lo, 3 * U, c > 0);
          } else {
            umma_chunk_3x(d, a_hi, a_lo, b_hi, b_lo, 2 * U, true);
            umma_chunk_3x(d + 3 * U, a_hi, a_lo, b_hi + 2 * U * 128, b_lo + 2 * U * 128, U, c > NC);
          }
          umma_commit(&bar_mma[c]);
          if (c == 2 * NC - 1) umma_commit(&bar_acc[ab]);
        }
        __syncwarp();
      }
      acc_uses[ab] += 1;
    }
  } else if (warp == TMA_WARP) {
    for (int64_t tile = blockIdx.x; tile < ntiles_other; tile += gridDim.x, ++lt) {
      for (int c = 0; c < 2 * NC; ++c, ++ctr) {
        if (lane == 0) {
          // the weight buffer is free when the UMMAs of the previous chunk are done
          if (c > 0) mbar_wait_idle(&bar_mma[c - 1], lt & 1);
          else if (lt > 0) mbar_wait_idle(&bar_mma[2 * NC - 1], (lt - 1) & 1);
          mbar_expect_tx(&bar_b, 2 * B_IMG);
          bulk_g2s(b_buf, reinterpret_cast<const char*>(wimg) + (size_t)c * (2 * B_IMG), 2 * B_IMG, &bar_b);
        }
        __syncwarp();
      }
    }
  } else if (warp >= HLOAD_WARP0) {
    // ---- h loaders: chunk c (32 columns of the old state rows) -> hi / lo images of its stage
    const int ht = tid - HLOAD_WARP0 * 32;
    constexpr int PER = 1024 / HLOAD_THREADS;              // float4 per thread per chunk
    for (int64_t tile = blockIdx.x; tile < ntiles_other; tile += gridDim.x, ++lt) {
      const int64_t m0 = tile * ROWS;
      ctr += NC;                                           // the x chunks of the tile
      for (int c = 0; c < NC; ++c, ++ctr) {
        const int s = ctr & 1;
        float4 v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int idx = ht + j * HLOAD_THREADS;
          const int r = idx >> 3, c4 = idx & 7;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r < n) v[j] = ldg_f4(h + (m0 + r) * U + c * 32 + c4 * 4);
        }
        unsigned char* st = smem + s * STAGE_A;
        // previous user of this stage: x chunk c of this tile (U = 64), or the h chunk of the previous tile (U = 32)
        if (NC == 2) mbar_wait_idle(&bar_mma[c], lt & 1);
        else if (lt > 0) mbar_wait_idle(&bar_mma[1], (lt - 1) & 1);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int idx = ht + j * HLOAD_THREADS;
          store_split(st, st + A_IMG, idx >> 3, idx & 7, v[j]);
        }
        fence_async_smem();
        named_sync(BAR_HLOAD, HLOAD_THREADS);
        if (ht == 0) mbar_arrive(&bar_full[s]);
      }
    }
  } else if (warp < GATHER_WARPS) {
    // ---- gather: one sub-warp of G lanes per destination range, rows handed out in edge-balanced ranges.
    // Rows in flight live in shared memory, not in registers: every lane copies its 16 bytes of a gathered row
    // with cp.async into the sub-warp's private ring (NB batches of BATCH rows) and reads back only what it
    // copied itself, so the ring needs no barrier; NB - 1 batches (16 rows) are in flight behind the one being
    // summed, 16 sub-warps (U = 64) x 16 rows x 256 B = 64 KB per SM, which is what the HBM latency asks for.
    const int gl = lane & (G - 1);                         // lane inside the sub-warp
    const int grp = tid / G;                               // sub-warp of the CTA, 0..NG-1
    const int chunk = gl >> 3, c4 = gl & 7;                // which x chunk / 16-byte column of it this lane feeds
    const float init = (OP == IGN_OP_MAX) ? -INFINITY : 0.0f;
    const uint32_t my_ring = smem_u32(ring + grp * RING_GROUP + gl * 16);      // shared-space address
    const int n_slots = __ldg(rowptr + n);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int64_t m0 = tile * ROWS;
      // the x stages of this tile are free once the UMMAs of their previous use are done (long ago): the h chunks of
      // the previous tile at U = 64 (same stages), the x chunk of the previous tile at U = 32
      auto wait_x_stages = [&]() {
        if (lt > 0 && !gather_only) {
          if (NC == 2) { mbar_wait(&bar_mma[2], (lt - 1) & 1); mbar_wait(&bar_mma[3], (lt - 1) & 1); }
          else mbar_wait(&bar_mma[0], (lt - 1) & 1);
        }
      };
      bool stages_free = false;
      for (int i = tid; i <= ROWS; i += GATHER_THREADS) {
        const int64_t r = m0 + i;
        s_rp[i] = __ldg(rowptr + (r < n ? r : n));
      }
      named_sync(BAR_GATHER, GATHER_THREADS);
      const int e0 = s_rp[0], e1 = s_rp[ROWS];
      // rows [b0, b1) of the tile: first row whose slots start at or after the sub-warp's share of the edges
      auto bound = [&](int g) -> int {
        if (g <= 0) return 0;
        if (g >= NG) return ROWS;
        const int target = e0 + (int)(((int64_t)(e1 - e0) * g) / NG);
        int lo = 0, hi = ROWS;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_rp[mid] < target) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      const int b0 = bound(grp), b1 = bound(grp + 1);
      const int eb0 = s_rp[b0], eb1 = s_rp[b1];
      unsigned char* img_hi = smem + ((ctr + chunk) & 1) * STAGE_A;
      unsigned char* img_lo = img_hi + A_IMG;

      int row = b0;
      int row_end = s_rp[b0 + (b0 < ROWS ? 1 : 0)];
      float4 acc = make_float4(init, init, init, init);
      auto flush = [&]() {                                 // destination `row` is complete
        float4 r = acc;
        const int len = row_end - s_rp[row];
        if (OP == IGN_OP_MEAN) {
          const float inv = 1.0f / (float)max(len, 1);
          r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
        }
        if (OP == IGN_OP_MAX && len == 0) r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (XRAW) {
          sts_f4(smem_u32(xraw + row * U + gl * 4), r);
        } else {
          if (!stages_free) { wait_x_stages(); stages_free = true; }
          store_split(img_hi, img_lo, row, c4, r);
        }
        if (agg_out && m0 + row < n) st_f4(agg_out + (m0 + row) * U + gl * 4, r);
        ++row;
        row_end = s_rp[min(row + 1, ROWS)];
        acc = make_float4(init, init, init, init);
      };
      // Batch k of the tile covers the slots [4 k, 4 k + 4) of the whole col array (absolute alignment).  Lane j of the
      // sub-warp holds the four column indices of batch (chunk * G + j): one coalesced 16-byte load per lane covers G
      // batches, is issued a whole chunk (G batches) before its first use, and the batch being issued gets its four
      // indices by full-warp shuffles (both sub-warps of a warp run the same number of iterations, so the warp stays
      // converged).  The loop body is kept SMALL (one copy of every path, nothing unrolled across batches): the first
      // version ran out of the instruction caches (35 KB of loop body: profiles/r2_agg_gru.md).
      const int kb0 = eb0 >> 2, nb = eb1 > eb0 ? ((eb1 + 3) >> 2) - kb0 : 0;
      auto load_chunk = [&](int cb) -> int4 {              // indices of batch cb * G + gl
        const int b = cb * G + gl, k = kb0 + b;
        if (b >= nb) return make_int4(-1, -1, -1, -1);
        if (4 * k + 3 < n_slots) return __ldg(reinterpret_cast<const int4*>(col) + k);
        int4 r;                                            // the last, partial batch of the array
        r.x = 4 * k < n_slots ? __ldg(col + 4 * k) : -1;
        r.y = 4 * k + 1 < n_slots ? __ldg(col + 4 * k + 1) : -1;
        r.z = 4 * k + 2 < n_slots ? __ldg(col + 4 * k + 2) : -1;
        r.w = -1;
        return r;
      };
      {
        // the issue pointer runs NB - 1 batches ahead of the consume pointer; empty groups are committed past the end
        // so that "at most NB - 1 groups pending" always means "the batch being consumed has landed"
        const int total = nb > 0 ? nb + NB - 1 : 0;
        const int wtotal = __reduce_max_sync(0xffffffffu, total);
        int4 ch = load_chunk(0), ch_next = load_chunk(1);
        int islot = 0, cslot = 0;
        const int lane_base = lane & ~(G - 1);
#pragma unroll 1
        for (int b = 0; b < wtotal; ++b) {
          const int jl = lane_base + (b & (G - 1));
          int4 q;
          q.x = __shfl_sync(0xffffffffu, ch.x, jl);
          q.y = __shfl_sync(0xffffffffu, ch.y, jl);
          q.z = __shfl_sync(0xffffffffu, ch.z, jl);
          q.w = __shfl_sync(0xffffffffu, ch.w, jl);
          if ((b & (G - 1)) == G - 1) {
            ch = ch_next;
            ch_next = load_chunk((b >> (G == 16 ? 4 : 3)) + 2);
          }
          if (b < total) {
            {                                              // ---- issue batch b
              const int e = 4 * (kb0 + b);
              const uint32_t dst = my_ring + islot * (BATCH * U * 4);
              if (b < nb) {
                if (e >= eb0 && e + BATCH <= eb1 && (q.x | q.y | q.z | q.w) >= 0) {
                  cp_async16_s(dst, src + (int64_t)q.x * U + gl * 4);
                  cp_async16_s(dst + U * 4, src + (int64_t)q.y * U + gl * 4);
                  cp_async16_s(dst + 2 * U * 4, src + (int64_t)q.z * U + gl * 4);
                  cp_async16_s(dst + 3 * U * 4, src + (int64_t)q.w * U + gl * 4);
                } else {
                  const int c[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                  for (int u = 0; u < BATCH; ++u) {
                    if (e + u >= eb0 && e + u < eb1) {
                      if (c[u] >= 0) cp_async16_s(dst + u * (U * 4), src + (int64_t)c[u] * U + gl * 4);
                      else sts_f4(dst + u * (U * 4), make_float4(init, init, init, init));   // a slot no edge claimed
                    }
                  }
                }
              }
              cp_async_commit();
              islot = islot + 1 == NB ? 0 : islot + 1;
            }
            if (b >= NB - 1) {                             // ---- consume batch b - (NB - 1)
              cp_async_wait<NB - 1>();
              const int e = 4 * (kb0 + b - (NB - 1));
              const uint32_t p = my_ring + cslot * (BATCH * U * 4);
              if (e >= eb0 && e + BATCH <= eb1 && e + BATCH <= row_end) {      // all four belong to the current row
#pragma unroll
                for (int u = 0; u < BATCH; ++u) acc_row<OP>(acc, lds_f4(p + u * (U * 4)));
              } else {
#pragma unroll 1
                for (int u = 0; u < BATCH; ++u) {
                  if (e + u >= eb0 && e + u < eb1) {
                    while (e + u >= row_end) flush();
                    acc_row<OP>(acc, lds_f4(p + u * (U * 4)));
                  }
                }
              }
              cslot = cslot + 1 == NB ? 0 : cslot + 1;
            }
          }
        }
      }
      while (row < b1) flush();                            // the last destination, and ones without slots
      if (XRAW) {
        named_sync(BAR_GATHER, GATHER_THREADS);            // every row of the tile is in xraw
        wait_x_stages();                                   // (done long ago: the UMMAs of the previous tile)
        for (int i = tid; i < ROWS * (U / 4); i += GATHER_THREADS) {
          const int rr = i / (U / 4), cc = i % (U / 4);
          unsigned char* ih = smem + ((ctr + (cc >> 3)) & 1) * STAGE_A;
          store_split(ih, ih + A_IMG, rr, cc & 7, lds_f4(smem_u32(xraw + rr * U + cc * 4)));
        }
      }
      fence_async_smem();
      named_sync(BAR_GATHER, GATHER_THREADS);              // every image row written; s_rp / xraw may be reused
      if (tid == 0 && !gather_only) {
#pragma unroll
        for (int c = 0; c < NC; ++c) mbar_arrive(&bar_full[(ctr + c) & 1]);
      }
      ctr += 2 * NC;
    }
  } else {
    // ---- epilogue: gates + new state; thread = one destination of the tile, 8 units at a time
    const int q = warp - EPI_WARP0;                        // == warp % 4: the TMEM lane group of this warp
    const int et = tid - EPI_WARP0 * 32;
    const int r = q * 32 + lane;                           // row of the tile
    uint32_t acc_uses[2] = {0, 0};
    int ab = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles_other; tile += gridDim.x, ab ^= 1) {
      const int64_t row = tile * ROWS + r;
      const uint32_t tb = tmem_base + ab * DCOLS + ((uint32_t)(q * 32) << 16);
      mbar_wait_idle(&bar_acc[ab], acc_uses[ab] & 1);
      tc_fence_after();
      if (TMA_OUT) {
        if (et == 0) bulk_wait_read();                     // the stores of the previous tile have read the staging tile
        named_sync(BAR_EPI, EPI_THREADS);
      }
#pragma unroll 1
      for (int u0 = 0; u0 < U; u0 += 8) {
        float4 ho[2];
        ho[0] = ho[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n) {
          ho[0] = ldg_f4(h + row * U + u0);
          ho[1] = ldg_f4(h + row * U + u0 + 4);
        }
        uint32_t az[8], ar[8], axh[8], ahh[8];
        tmem_ld8_nowait(tb + u0, az);
        tmem_ld8_nowait(tb + U + u0, ar);
        tmem_ld8_nowait(tb + 2 * U + u0, axh);
        tmem_ld8_nowait(tb + 3 * U + u0, ahh);
        tmem_ld_wait();
        if (u0 + 8 >= U) {                                 // last read of this accumulator by this thread
          tc_fence_before();
          mbar_arrive(&bar_drained[ab]);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; j4 += 4) {
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
          const int cc = u0 + j4;                          // column -> box cc / 32, 16-byte chunk (cc % 32) / 4
          if (TMA_OUT) {
            // tensor stores: two swizzled [128 x 32] boxes; bulk stores: the tile as it lies in global memory
            const int so = bulk1d ? r * (U * 4) + cc * 4
                                  : (cc >> 5) * A_IMG + r * 128 + ((((cc & 31) >> 2) ^ (r & 7)) << 4);
            *reinterpret_cast<float4*>(out_stage + so) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          }
          else if (row < n)
            st_f4(out_direct + (row + out_row0) * U + cc, make_float4(hn[0], hn[1], hn[2], hn[3]));
        }
      }
      acc_uses[ab] += 1;
      if (!TMA_OUT) continue;
      fence_async_smem();
      named_sync(BAR_EPI, EPI_THREADS);
      if (et == 0) {
        const int r0 = out_row0 + (int)(tile * ROWS);
        if (bulk1d) {       // one contiguous store per output array: the tile's rows are adjacent in global memory
          const int64_t left = n - tile * ROWS;
          const uint32_t bytes = (uint32_t)(left < ROWS ? left : ROWS) * (U * 4);
          for (int k = 0; k < n_out; ++k) bulk_s2g(maps.p[k] + (int64_t)r0 * U, out_stage, bytes);
        } else {
          for (int k = 0; k < n_out; ++k) {
#pragma unroll
            for (int bx = 0; bx < NC; ++bx) tma_store_2d(&maps.m[k], out_stage + bx * A_IMG, bx * 32, r0);
          }
        }
        bulk_commit();
      }
    }
    if (TMA_OUT && et == 0) bulk_wait_all();               // every state row has left the SM
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * DCOLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

template <int U, int OP, bool TMA_OUT>
int launch_v(const int* rowptr, const int* col, const float* src, const float* h, int64_t n, const float* wimg,
             const float* bias, const OutMaps& maps, int n_out, int out_row0, float* agg_out, float* out_direct,
             int grid, cudaStream_t st) {
  constexpr size_t smem = PROBE ? 1024 + (size_t)(GATHER_THREADS / (U / 4)) * NB * BATCH * U * 4 + ROWS * U * 4 + 4096
                                : 1024 + 2 * (size_t)(2 * A_IMG) + 2 * (size_t)(3 * U * 128) + (size_t)(U / 32) * A_IMG +
                                      (size_t)(GATHER_THREADS / (U / 4)) * NB * BATCH * U * 4 + (U == 32 ? ROWS * U * 4 : 0);
  IGN_CUDA(cudaFuncSetAttribute(agg_gru_tc_kernel<U, OP, TMA_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  agg_gru_tc_kernel<U, OP, TMA_OUT><<<grid, AGG_THREADS, smem, st>>>(rowptr, col, src, h, n, wimg, bias, maps, n_out,
                                                                     out_row0, agg_out, out_direct,
                                                                     getenv("IGN_AGG_DBG") ? atoi(getenv("IGN_AGG_DBG")) : 0);
  IGN_CHECK_LAUNCH("agg_gru_tc");
  return IGN_OK;
}
// U = 64 with one output: the staging tile's shared memory is worth more as the raw-x buffer (see XRAW)
template <int U, int OP>
int launch(const int* rowptr, const int* col, const float* src, const float* h, int64_t n, const float* wimg,
           const float* bias, const OutMaps& maps, int n_out, int out_row0, float* agg_out, float* out0, int grid,
           cudaStream_t st) {
  static const bool force_tma = getenv("IGN_AGG_TMA_OUT") != nullptr;
  if (U == 64 && n_out == 1 && !force_tma)
    return launch_v<U, OP, false>(rowptr, col, src, h, n, wimg, bias, maps, n_out, out_row0, agg_out, out0, grid, st);
  return launch_v<U, OP, true>(rowptr, col, src, h, n, wimg, bias, maps, n_out, out_row0, agg_out, out0, grid, st);
}

}  // namespace

extern "C" size_t ign_agg_gru_cell_tc_ws_bytes(int f_in, int units) {
  return (f_in == units && (units == 32 || units == 64)) ? ign_gru_cell_tc_ws(units) : 0;
}

extern "C" int ign_agg_gru_cell_tc(int op, const int32_t* rowptr, const int32_t* col, const float* src_states,
                                   int f_in, const float* h_dst, int64_t num_dst, int units, const float* kernel,
                                   const float* recurrent_kernel, const float* bias, int n_out, float* const* outs,
                                   int64_t out_row0, float* agg_out, void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(op == IGN_OP_SUM || op == IGN_OP_MEAN || op == IGN_OP_MAX, IGN_ERR_INVALID,
              "IGNNITION: agg_gru_cell_tc: unknown aggregation %d", op);
  IGN_REQUIRE(num_dst >= 0 && out_row0 >= 0, IGN_ERR_INVALID, "IGNNITION: agg_gru_cell_tc: negative size");
  IGN_REQUIRE(ign_agg_gru_cell_tc_ws_bytes(f_in, units) > 0, IGN_ERR_UNSUPPORTED,
              "IGNNITION: agg_gru_cell_tc: built for message width == units in {32, 64} (got %d, %d)", f_in, units);
  IGN_REQUIRE(n_out >= 1 && n_out <= IGN_MAX_PEERS && outs, IGN_ERR_INVALID,
              "IGNNITION: agg_gru_cell_tc: between 1 and %d output buffers", IGN_MAX_PEERS);
  IGN_REQUIRE(out_row0 + num_dst < ((int64_t)1 << 31), IGN_ERR_UNSUPPORTED, "IGNNITION: agg_gru_cell_tc: int32 rows only");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && src_states && h_dst && kernel && recurrent_kernel && bias, IGN_ERR_INVALID,   // col: null iff no slot
              "IGNNITION: agg_gru_cell_tc: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= ign_gru_cell_tc_ws(units), IGN_ERR_WORKSPACE,
              "IGNNITION: agg_gru_cell_tc: workspace too small (%zu < %zu)", ws_bytes, ign_gru_cell_tc_ws(units));
  EncodeTiledFn enc = encode_tiled();
  IGN_REQUIRE(enc, IGN_ERR_UNSUPPORTED, "IGNNITION: agg_gru_cell_tc: cuTensorMapEncodeTiled is not available");
  OutMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int k = 0; k < n_out; ++k) {
    IGN_REQUIRE(outs[k] && ((uintptr_t)outs[k] & 15) == 0, IGN_ERR_INVALID,
                "IGNNITION: agg_gru_cell_tc: output buffer %d is null or not 16-byte aligned", k);
    // rows past out_row0 + num_dst are clipped by the TMA unit: the last tile cannot touch the next owner's rows
    const cuuint64_t dims[2] = {(cuuint64_t)units, (cuuint64_t)(out_row0 + num_dst)};
    const cuuint64_t strides[1] = {(cuuint64_t)units * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)ROWS};
    const cuuint32_t estr[2] = {1, 1};
    maps.p[k] = outs[k];
    const CUresult r = enc(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, outs[k], dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IGN_REQUIRE(r == CUDA_SUCCESS, IGN_ERR_INVALID, "IGNNITION: agg_gru_cell_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  cudaStream_t st = ign_stream(stream);
  int rc = ign_gru_cell_tc_prep(kernel, recurrent_kernel, units, ws, st);
  if (rc) return rc;
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ign_cdiv(num_dst, ROWS);
  const int grid = (int)(tiles < sms ? tiles : sms);
  const float* wimg = reinterpret_cast<const float*>(ws);
#define IGN_AGG_LAUNCH(UU)                                                                                        \
  switch (op) {                                                                                                   \
    case IGN_OP_SUM:                                                                                              \
      return launch<UU, IGN_OP_SUM>(rowptr, col, src_states, h_dst, num_dst, wimg, bias, maps, n_out,             \
                                    (int)out_row0, agg_out, outs[0], grid, st);                                            \
    case IGN_OP_MEAN:                                                                                             \
      return launch<UU, IGN_OP_MEAN>(rowptr, col, src_states, h_dst, num_dst, wimg, bias, maps, n_out,            \
                                     (int)out_row0, agg_out, outs[0], grid, st);                                           \
    default:                                                                                                      \
      return launch<UU, IGN_OP_MAX>(rowptr, col, src_states, h_dst, num_dst, wimg, bias, maps, n_out,             \
                                    (int)out_row0, agg_out, outs[0], grid, st);                                            \
  }
  if (units == 64) { IGN_AGG_LAUNCH(64) }
  IGN_AGG_LAUNCH(32)
#undef IGN_AGG_LAUNCH
}
