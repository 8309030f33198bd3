// Fused GRU update kernels (forward), sm_100a.
//
//  ign_gru_seq      : ordered / interleave aggregation + recurrent update.  Replaces
//                     keras.layers.RNN(GRUCell)(padded [num_dst,max_len,F], initial_state,
//                     mask=sequence_mask(len)) + gather_nd(outputs,[i,len-1])
//                     (reference code/utils/auxilary_classes.py:767-796) and Interleave_aggr
//                     (:421-440): a persistent CTA walks each destination's CSR slot list, no
//                     padding, the next step's message rows prefetched with cp.async while the
//                     current step's gates are computed.
//  ign_agg_gru_cell : gather + sum aggregation + one GRU step, fused
//                     (code/utils/generate_model.py:432,490; auxilary_classes.py:254-262,752-765).
//  ign_gru_cell     : one GRU step on a dense x (perform_unsorted_update after any aggregation).
//
// Tile geometry and the gate GEMM are in gru.cuh.

#include "gru.cuh"

using namespace ign_gru;

namespace {

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};

__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}

template <int FI, int U, int NX>
struct Smem {
  using T = Tile<U>;
  static constexpr int XS = FI + 4, HS = U + 4;
  static constexpr int W_OFF = 0;
  static constexpr int X_OFF = (WeightSmem<FI, U>::FLOATS + 3) / 4 * 4;
  static constexpr int H_OFF = X_OFF + NX * T::R * XS;
  static constexpr int META_OFF = H_OFF + T::R * HS;          // 3 ints per row
  static constexpr size_t BYTES = (size_t)(META_OFF + 3 * T::R + 4) * 4;
};

// ---------------------------------------------------------------------------------------------
template <int FI, int U>
__global__ void __launch_bounds__(THREADS) gru_seq_kernel(const int* __restrict__ steps_rowptr,
                                                          const int* __restrict__ steps,
                                                          const int* __restrict__ order, SrcPtrs srcs,
                                                          const float* __restrict__ h0, int64_t num_dst,
                                                          const float* __restrict__ kernel,
                                                          const float* __restrict__ rkernel,
                                                          const float* __restrict__ bias, float* __restrict__ out,
                                                          float* __restrict__ h_seq) {
  using S = Smem<FI, U, 2>;
  using T = Tile<U>;
  constexpr int R = T::R, TU = T::TU, TR = T::TR, NUG = T::NUG, NRG = T::NRG, XS = S::XS, HS = S::HS;
  extern __shared__ float4 smem_f4[];
  float* smem = reinterpret_cast<float*>(smem_f4);
  float* sw = smem + S::W_OFF;
  float* Xb = smem + S::X_OFF;
  float* Hs = smem + S::H_OFF;
  int* row_dst = reinterpret_cast<int*>(smem + S::META_OFF);
  int* row_lo = row_dst + R;
  int* row_len = row_lo + R;
  int* s_maxlen = row_len + R;

  load_weights<FI, U>(sw, kernel, rkernel, bias);
  const int tid = threadIdx.x;
  const int ug = tid % NUG, rg = tid / NUG, u0 = ug * TU;
  const int64_t ntiles = (num_dst + R - 1) / R;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();                      // previous tile fully written out before smem is reused
    if (tid == 0) *s_maxlen = 0;
    if (tid < R) {
      const int64_t didx = tile * R + tid;
      int d = -1, lo = 0, len = 0;
      if (didx < num_dst) {
        d = order ? order[didx] : (int)didx;
        lo = steps_rowptr[d];
        len = steps_rowptr[d + 1] - lo;
      }
      row_dst[tid] = d; row_lo[tid] = lo; row_len[tid] = len;
    }
    __syncthreads();
    if (tid < R && row_len[tid] > 0) atomicMax(s_maxlen, row_len[tid]);
    // old state of the tile's destinations
    for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
      const int r = idx / (U / 4), c4 = idx % (U / 4);
      const int d = row_dst[r];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d >= 0) v = ldg_f4(h0 + (int64_t)d * U + c4 * 4);
      st_f4(Hs + r * HS + c4 * 4, v);
    }
    __syncthreads();
    const int maxlen = *s_maxlen;

    auto prefetch = [&](int t, float* X) {
      for (int idx = tid; idx < R * (FI / 4); idx += THREADS) {
        const int r = idx / (FI / 4), c4 = idx % (FI / 4);
        if (row_len[r] > t) {
          const int entry = steps[row_lo[r] + t];
          if (entry >= 0) {
            const float* base = pick_src(srcs, entry >> IGN_STEP_SRC_SHIFT);
            cp_async16(X + r * XS + c4 * 4, base + (int64_t)(entry & IGN_STEP_ROW_MASK) * FI + c4 * 4);
          } else {
            st_f4(X + r * XS + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
      }
      cp_async_commit();
    };

    if (maxlen > 0) prefetch(0, Xb);
    for (int t = 0; t < maxlen; ++t) {
      float* X = Xb + (t & 1) * R * XS;
      if (t + 1 < maxlen) {
        prefetch(t + 1, Xb + ((t + 1) & 1) * R * XS);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      float az[TR][TU], ar[TR][TU], axh[TR][TU], ahh[TR][TU];
      gate_gemm<FI, U>(sw, X, Hs, rg, u0, az, ar, axh, ahh);
      float hn[TR][TU];
      bool act[TR];
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const int r = rg + i * NRG;
        act[i] = row_len[r] > t;
#pragma unroll
        for (int j = 0; j < TU; ++j) {
          const float hold = Hs[r * HS + u0 + j];
          hn[i][j] = act[i] ? gru_out(az[i][j], ar[i][j], axh[i][j], ahh[i][j], hold) : hold;
        }
      }
      __syncthreads();                    // every thread has finished reading Hs
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        if (!act[i]) continue;
        const int r = rg + i * NRG;
#pragma unroll
        for (int j = 0; j < TU; ++j) Hs[r * HS + u0 + j] = hn[i][j];
        if (h_seq) {
          float* dstp = h_seq + (int64_t)(row_lo[r] + t) * U + u0;
#pragma unroll
          for (int j = 0; j < TU; ++j) dstp[j] = hn[i][j];
        }
      }
      __syncthreads();
    }
    for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
      const int r = idx / (U / 4), c4 = idx % (U / 4);
      const int d = row_dst[r];
      if (d >= 0) st_f4(out + (int64_t)d * U + c4 * 4, *reinterpret_cast<const float4*>(Hs + r * HS + c4 * 4));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MODE 0: x = segment sum over CSR (fused aggregation); MODE 1: x given as dense rows
template <int FI, int U, int MODE>
__global__ void __launch_bounds__(THREADS, (U <= 32 && FI <= 32) ? 2 : 1) gru_cell_kernel(const int* __restrict__ rowptr,
                                                           const int* __restrict__ col,
                                                           const float* __restrict__ src,   // states (MODE 0) / x (MODE 1)
                                                           const float* __restrict__ h, int64_t num_dst,
                                                           const float* __restrict__ kernel,
                                                           const float* __restrict__ rkernel,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           float* __restrict__ agg_out) {
  using S = Smem<FI, U, 1>;
  using T = Tile<U>;
  constexpr int R = T::R, TU = T::TU, TR = T::TR, NUG = T::NUG, NRG = T::NRG, XS = S::XS, HS = S::HS;
  constexpr int G = FI / 4;                       // lanes per destination in the aggregation phase
  static_assert((G & (G - 1)) == 0 && G <= 32, "FI/4 must be a power of two <= 32");
  extern __shared__ float4 smem_f4[];
  float* smem = reinterpret_cast<float*>(smem_f4);
  float* sw = smem + S::W_OFF;
  float* X = smem + S::X_OFF;
  float* Hs = smem + S::H_OFF;

  load_weights<FI, U>(sw, kernel, rkernel, bias);
  const int tid = threadIdx.x;
  const int ug = tid % NUG, rg = tid / NUG, u0 = ug * TU;
  const int64_t ntiles = (num_dst + R - 1) / R;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t d0 = tile * R;
    // old state rows -> Hs (async, overlaps the aggregation)
    for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
      const int r = idx / (U / 4), c4 = idx % (U / 4);
      if (d0 + r < num_dst) cp_async16(Hs + r * HS + c4 * 4, h + (d0 + r) * U + c4 * 4);
      else st_f4(Hs + r * HS + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (MODE == 1) {
      for (int idx = tid; idx < R * (FI / 4); idx += THREADS) {
        const int r = idx / (FI / 4), c4 = idx % (FI / 4);
        if (d0 + r < num_dst) cp_async16(X + r * XS + c4 * 4, src + (d0 + r) * FI + c4 * 4);
        else st_f4(X + r * XS + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    cp_async_commit();
    if (MODE == 0) {
      // sum aggregation in slot order, one accumulator: G lanes per destination
      const int lane = tid & 31, gl = lane & (G - 1);
      const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
      for (int r = tid / G; r < R; r += THREADS / G) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t d = d0 + r;
        if (d < num_dst) {
          const int lo = rowptr[d], hi = rowptr[d + 1];
          for (int e = lo; e < hi; e += G) {
            const int mine = (e + gl < hi) ? (col ? col[e + gl] : e + gl) : -1;
            const int cnt = min(G, hi - e);
            for (int j = 0; j < cnt; j += 4) {
              int c[4];
              float4 v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) c[u] = __shfl_sync(gmask, mine, (j + u) & (G - 1), G);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (j + u < cnt) v[u] = ldg_f4(src + (int64_t)c[u] * FI + gl * 4);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (j + u < cnt) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
            }
          }
          if (agg_out) st_f4(agg_out + d * FI + gl * 4, acc);
        }
        st_f4(X + r * XS + gl * 4, acc);
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    float az[TR][TU], ar[TR][TU], axh[TR][TU], ahh[TR][TU];
    gate_gemm<FI, U>(sw, X, Hs, rg, u0, az, ar, axh, ahh);
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      const int r = rg + i * NRG;
      const int64_t d = d0 + r;
      if (d >= num_dst) continue;
      float hn[TU];
#pragma unroll
      for (int j = 0; j < TU; ++j)
        hn[j] = gru_out(az[i][j], ar[i][j], axh[i][j], ahh[i][j], Hs[r * HS + u0 + j]);
      float* o = out + d * U + u0;
#pragma unroll
      for (int j = 0; j < TU; ++j) o[j] = hn[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <typename KernelT>
int persistent_grid(KernelT k, size_t smem, int64_t ntiles, int* grid) {
  static thread_local const void* cached_fn[32];
  static thread_local int cached_occ[32];
  static thread_local int cached_dev[32];
  const void* fn = reinterpret_cast<const void*>(k);
  int occ = 0, cur = 0;
  cudaGetDevice(&cur);                             // function attributes and occupancy are per device
  for (int i = 0; i < 32; ++i) {
    if (cached_fn[i] == fn && cached_dev[i] == cur) { occ = cached_occ[i]; break; }
    if (cached_fn[i] == nullptr) {
      cached_dev[i] = cur;
      IGN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      IGN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, THREADS, smem));
      IGN_REQUIRE(occ >= 1, IGN_ERR_UNSUPPORTED, "IGNNITION: GRU kernel does not fit on an SM (smem %zu B)", smem);
      cached_fn[i] = fn; cached_occ[i] = occ;
      break;
    }
  }
  IGN_REQUIRE(occ >= 1, IGN_ERR_UNSUPPORTED, "IGNNITION: GRU kernel occupancy cache is full");
  int sms = IGN_NUM_SMS;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = (int64_t)sms * occ;
  *grid = (int)(ntiles < cap ? ntiles : cap);
  return IGN_OK;
}

template <int FI, int U>
int launch_gru_seq(const int* steps_rowptr, const int* steps, const int* order, const SrcPtrs& srcs,
                   const float* h0, int64_t num_dst, const float* k, const float* rk, const float* b, float* out,
                   float* h_seq, cudaStream_t st) {
  using S = Smem<FI, U, 2>;
  int grid = 0;
  int rc = persistent_grid(gru_seq_kernel<FI, U>, S::BYTES, ign_cdiv(num_dst, Tile<U>::R), &grid);
  if (rc) return rc;
  gru_seq_kernel<FI, U><<<grid, THREADS, S::BYTES, st>>>(steps_rowptr, steps, order, srcs, h0, num_dst, k, rk, b,
                                                          out, h_seq);
  IGN_CHECK_LAUNCH("gru_seq");
  return IGN_OK;
}

template <int FI, int U, int MODE>
int launch_gru_cell(const int* rowptr, const int* col, const float* src, const float* h, int64_t num_dst,
                    const float* k, const float* rk, const float* b, float* out, float* agg_out, cudaStream_t st) {
  using S = Smem<FI, U, 1>;
  int grid = 0;
  int rc = persistent_grid(gru_cell_kernel<FI, U, MODE>, S::BYTES, ign_cdiv(num_dst, Tile<U>::R), &grid);
  if (rc) return rc;
  gru_cell_kernel<FI, U, MODE><<<grid, THREADS, S::BYTES, st>>>(rowptr, col, src, h, num_dst, k, rk, b, out, agg_out);
  IGN_CHECK_LAUNCH(MODE == 0 ? "agg_gru_cell" : "gru_cell");
  return IGN_OK;
}

#define IGN_GRU_DISPATCH(FI_, U_, CALL)                                   \
  if ((FI_) == 16 && (U_) == 16) { constexpr int FI = 16, U = 16; CALL; } \
  if ((FI_) == 32 && (U_) == 32) { constexpr int FI = 32, U = 32; CALL; } \
  if ((FI_) == 64 && (U_) == 64) { constexpr int FI = 64, U = 64; CALL; } \
  if ((FI_) == 16 && (U_) == 32) { constexpr int FI = 16, U = 32; CALL; } \
  if ((FI_) == 64 && (U_) == 32) { constexpr int FI = 64, U = 32; CALL; } \
  if ((FI_) == 32 && (U_) == 64) { constexpr int FI = 32, U = 64; CALL; } \
  if ((FI_) == 32 && (U_) == 16) { constexpr int FI = 32, U = 16; CALL; }

int check_gru_args(const char* who, int f_in, int units, const float* k, const float* rk, const float* b) {
  IGN_REQUIRE(k && rk && b, IGN_ERR_INVALID, "IGNNITION: %s: null weight pointer", who);
  const bool ok = (f_in == 16 || f_in == 32 || f_in == 64) && (units == 16 || units == 32 || units == 64) &&
                  !(f_in == 16 && units == 64) && !(f_in == 64 && units == 16);
  IGN_REQUIRE(ok, IGN_ERR_UNSUPPORTED,
              "IGNNITION: %s: message width %d with %d GRU units is not built "
              "(supported: widths and units in {16,32,64})", who, f_in, units);
  return IGN_OK;
}

}  // namespace

// fp32 launchers of this compilation's tile geometry (see gru.cuh)
int IGN_GRU_FN(ign_gru_seq_fp32)(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                                 const float* const* srcs, int f_in, const float* h0, int64_t num_dst, int units,
                                 const float* kernel, const float* recurrent_kernel, const float* bias, float* out,
                                 float* h_seq, cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  IGN_GRU_DISPATCH(f_in, units, return (launch_gru_seq<FI, U>(steps_rowptr, steps, order, sp, h0, num_dst, kernel,
                                                               recurrent_kernel, bias, out, h_seq, st)));
  return IGN_ERR_UNSUPPORTED;
}

// mode 0: gather + sum + GRU step (rowptr / col), mode 1: GRU step on a dense x
int IGN_GRU_FN(ign_gru_cell_fp32)(int mode, const int* rowptr, const int* col, const float* src, const float* h,
                                  int64_t num_dst, int f_in, int units, const float* kernel,
                                  const float* recurrent_kernel, const float* bias, float* out, float* agg_out,
                                  cudaStream_t st) {
  if (mode == 0) {
    IGN_GRU_DISPATCH(f_in, units, return (launch_gru_cell<FI, U, 0>(rowptr, col, src, h, num_dst, kernel,
                                                                     recurrent_kernel, bias, out, agg_out, st)));
  } else {
    IGN_GRU_DISPATCH(f_in, units, return (launch_gru_cell<FI, U, 1>(nullptr, nullptr, src, h, num_dst, kernel,
                                                                     recurrent_kernel, bias, out, nullptr, st)));
  }
  return IGN_ERR_UNSUPPORTED;
}

#ifndef IGN_GRU_SMALL_TILE
int ign_gru_seq_fp32_small_tile(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                                const float* const* srcs, int f_in, const float* h0, int64_t num_dst, int units,
                                const float* kernel, const float* recurrent_kernel, const float* bias, float* out,
                                float* h_seq, cudaStream_t st);
int ign_gru_cell_fp32_small_tile(int mode, const int* rowptr, const int* col, const float* src, const float* h,
                                 int64_t num_dst, int f_in, int units, const float* kernel,
                                 const float* recurrent_kernel, const float* bias, float* out, float* agg_out,
                                 cudaStream_t st);

// tensor-core variant (gru_seq_tc.cu), 32-wide messages and states
int ign_gru_seq_tc_launch(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                          const float* const* srcs, const float* h0, int64_t num_dst, const float* kernel,
                          const float* rkernel, const float* bias, float* out, float* h_seq, const int* meta,
                          cudaStream_t st);
bool ign_tensor_cores_enabled();

extern "C" int ign_gru_seq(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* order, int n_src,
                           const float* const* srcs, int f_in, const float* h0, int64_t num_dst, int units,
                           const float* kernel, const float* recurrent_kernel, const float* bias, float* out,
                           float* h_seq, const int32_t* meta, void* stream) {
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: gru_seq: negative size");
  IGN_REQUIRE(n_src >= 1 && n_src <= IGN_MAX_SOURCES && srcs, IGN_ERR_INVALID,
              "IGNNITION: gru_seq: between 1 and %d sources", IGN_MAX_SOURCES);
  int rc = check_gru_args("gru_seq", f_in, units, kernel, recurrent_kernel, bias);
  if (rc) return rc;
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(steps_rowptr && steps && h0 && out, IGN_ERR_INVALID, "IGNNITION: gru_seq: null pointer");
  for (int i = 0; i < n_src; ++i)
    IGN_REQUIRE(srcs[i], IGN_ERR_INVALID, "IGNNITION: gru_seq: null source state");
  cudaStream_t st = ign_stream(stream);
  // (below a few tiles per SM the tensor-core walker's fixed latency -- TMEM, barriers, weight images -- is the whole
  // kernel: 53 us for 546 rows; the 16-row fp32 tiles take over there)
  if (f_in == 32 && units == 32 && ign_tensor_cores_enabled() && !ign_gru_use_small_tile(num_dst))
    return ign_gru_seq_tc_launch(steps_rowptr, steps, order, n_src, srcs, h0, num_dst, kernel, recurrent_kernel,
                                 bias, out, h_seq, meta, st);
  return (ign_gru_use_small_tile(num_dst) ? ign_gru_seq_fp32_small_tile : ign_gru_seq_fp32)(
      steps_rowptr, steps, order, n_src, srcs, f_in, h0, num_dst, units, kernel, recurrent_kernel, bias, out, h_seq, st);
}

extern "C" int ign_agg_gru_cell(const int32_t* rowptr, const int32_t* col, const float* src_states, int f_in,
                                const float* h_dst, int64_t num_dst, int units, const float* kernel,
                                const float* recurrent_kernel, const float* bias, float* out, float* agg_out,
                                void* stream) {
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: agg_gru_cell: negative size");
  int rc = check_gru_args("agg_gru_cell", f_in, units, kernel, recurrent_kernel, bias);
  if (rc) return rc;
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && src_states && h_dst && out, IGN_ERR_INVALID, "IGNNITION: agg_gru_cell: null pointer");
  cudaStream_t st = ign_stream(stream);
  return (ign_gru_use_small_tile(num_dst) ? ign_gru_cell_fp32_small_tile : ign_gru_cell_fp32)(
      0, rowptr, col, src_states, h_dst, num_dst, f_in, units, kernel, recurrent_kernel, bias, out, agg_out, st);
}

// step-synchronous tensor-core variant of the ordered update (gru_step_tc.cu)
int ign_gru_step_tc_launch(int t, const int* nt, const int* off, int64_t num_dst, int64_t rows_bound, const int* meta,
                           const int* steps_T, int n_src, const float* const* srcs, const float* h0, float* hs,
                           float* out, float* h_seq, const float* kernel, const float* rkernel, const float* bias,
                           cudaStream_t st);

extern "C" int ign_gru_seq_step(int t, const int32_t* nt, const int32_t* off, const int32_t* meta,
                                const int32_t* steps_T, int n_src, const float* const* srcs, int f_in, const float* h0,
                                float* hs, int64_t num_dst, int units, const float* kernel,
                                const float* recurrent_kernel, const float* bias, float* out, float* h_seq,
                                void* stream) {
  IGN_REQUIRE(num_dst >= 0 && t >= 0, IGN_ERR_INVALID, "IGNNITION: gru_seq_step: bad argument");
  IGN_REQUIRE(n_src >= 1 && n_src <= IGN_MAX_SOURCES && srcs, IGN_ERR_INVALID, "IGNNITION: gru_seq_step: bad sources");
  IGN_REQUIRE(f_in == 32 && units == 32, IGN_ERR_UNSUPPORTED,
              "IGNNITION: gru_seq_step: built for 32-wide messages and states (got %d, %d)", f_in, units);
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(nt && off && meta && steps_T && h0 && hs && out && kernel && recurrent_kernel && bias, IGN_ERR_INVALID,
              "IGNNITION: gru_seq_step: null pointer");
  return ign_gru_step_tc_launch(t, nt, off, num_dst, num_dst, meta, steps_T, n_src, srcs, h0, hs, out, h_seq, kernel,
                                recurrent_kernel, bias, ign_stream(stream));
}

// tensor-core variant (gru_cell_tc.cu)
bool ign_gru_cell_tc_supported(int f_in, int units);
size_t ign_gru_cell_tc_ws(int units);
int ign_gru_cell_tc_launch(const float* x, const float* h, int64_t n, int units, const float* kernel,
                           const float* rkernel, const float* bias, float* out, void* ws, cudaStream_t st);

extern "C" size_t ign_gru_cell_ws_bytes(int f_in, int units) {
  return ign_gru_cell_tc_supported(f_in, units) ? ign_gru_cell_tc_ws(units) : 0;
}

extern "C" int ign_gru_cell(const float* x, const float* h, int64_t n, int f_in, int units, const float* kernel,
                            const float* recurrent_kernel, const float* bias, float* out, void* ws, size_t ws_bytes,
                            void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: gru_cell: negative size");
  int rc = check_gru_args("gru_cell", f_in, units, kernel, recurrent_kernel, bias);
  if (rc) return rc;
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(x && h && out, IGN_ERR_INVALID, "IGNNITION: gru_cell: null pointer");
  cudaStream_t st = ign_stream(stream);
  if (ws && ign_tensor_cores_enabled() && ign_gru_cell_tc_supported(f_in, units) &&
      ws_bytes >= ign_gru_cell_tc_ws(units) && n >= 128) {
    IGN_REQUIRE(out != h, IGN_ERR_INVALID, "IGNNITION: gru_cell: out must not alias h on the tensor-core path");
    return ign_gru_cell_tc_launch(x, h, n, units, kernel, recurrent_kernel, bias, out, ws, st);
  }
  return (ign_gru_use_small_tile(n) ? ign_gru_cell_fp32_small_tile : ign_gru_cell_fp32)(
      1, nullptr, nullptr, x, h, n, f_in, units, kernel, recurrent_kernel, bias, out, nullptr, st);
}

#endif  // IGN_GRU_SMALL_TILE
