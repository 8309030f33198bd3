// Backward twins of the fused GRU kernels (sm_100a): tf.gradients through GRUCell and through
// keras.layers.RNN(GRUCell) over every destination's message sequence (BPTT), as model_fn asks of
// TensorFlow at code/utils/generate_model.py:791.
//
// One CTA owns a tile of R destinations (same geometry as the forward, gru.cuh).  Per step:
//   A  recompute the gate pre-activations from (x_t, h_{t-1}) -- nothing but h is saved by the
//      forward -- and turn dL/dh_t into the gate gradients G = [d_az | d_ar | d_axh | d_ahh]
//      (shared memory) and the direct term dh_t * z;
//   B  dx_t = GX K^T,  dh_{t-1} = dh_t z + GH R^T   with K^T / R^T resident in shared memory;
//   C  dK += x_t^T GX, dR += h_{t-1}^T GH, db += colsum(G): register accumulators that live for
//      the CTA's whole life and are flushed once with atomics.
// GX = [d_az | d_ar | d_axh], GH = [d_az | d_ar | d_ahh].  Built for f_in == units in {16, 32}
// (RouteNet / Q-size); other shapes fail loudly.

#include "gru.cuh"

using namespace ign_gru;

namespace {

struct SrcPtrs {
  const float* p[IGN_MAX_SOURCES];
};
__device__ __forceinline__ const float* pick_src(const SrcPtrs& s, int k) {
  return k == 0 ? s.p[0] : k == 1 ? s.p[1] : k == 2 ? s.p[2] : s.p[3];
}

template <int U>
struct BSmem {
  using T = Tile<U>;
  static constexpr int FI = U;
  static constexpr int XS = U + 4, GS = 4 * U + 4;
  static constexpr int W_OFF = 0;                                      // K | Rk | bias (forward layout)
  static constexpr int KT_OFF = (WeightSmem<FI, U>::FLOATS + 3) / 4 * 4;   // K^T  [3U][FI]
  static constexpr int RT_OFF = KT_OFF + 3 * U * FI;                   // Rk^T [3U][U]
  static constexpr int X_OFF = RT_OFF + 3 * U * U;
  static constexpr int H_OFF = X_OFF + T::R * XS;                      // h_{t-1}
  static constexpr int D_OFF = H_OFF + T::R * XS;                      // dL/dh (running)
  static constexpr int G_OFF = D_OFF + T::R * XS;
  static constexpr int META_OFF = G_OFF + T::R * GS;
  static constexpr size_t BYTES = (size_t)(META_OFF + 3 * T::R + 4) * 4;
};

template <int U>
__device__ __forceinline__ void load_transposed(float* smem, const float* __restrict__ kernel,
                                                const float* __restrict__ rkernel) {
  using S = BSmem<U>;
  for (int i = threadIdx.x; i < U * 3 * U; i += THREADS) {
    const int k = i / (3 * U), c = i % (3 * U);
    smem[S::KT_OFF + c * U + k] = __ldg(kernel + i);
    smem[S::RT_OFF + c * U + k] = __ldg(rkernel + i);
  }
}

// weight-gradient accumulators of one thread: rows [kb*4, kb*4+4) of the stacked [x | h] dims,
// columns [cb*CB, cb*CB+CB) of the 3U gate columns, over the rows of its split of the tile
template <int U>
struct WAcc {
  static constexpr int CB = 3 * U / 8;
  static constexpr int TPS = (2 * U / 4) * 8;          // threads per split
  static constexpr int NSPLIT = THREADS / TPS;
  float a[4][CB];
};

// one backward step on the tile.  active[i] says whether row rg + i*NRG takes part.
template <int U, bool WRITE_DX>
__device__ __forceinline__ void bwd_tile_step(float* smem, const bool (&active)[Geo<U>::TR], WAcc<U>& wacc,
                                              float (&bacc)[2], float* const (&dx_row)[Geo<U>::TR]) {
  using S = BSmem<U>;
  using T = Tile<U>;
  constexpr int FI = U;
  constexpr int TU = T::TU, TR = T::TR, NUG = T::NUG, NRG = T::NRG, XS = S::XS, GS = S::GS, R = T::R;
  const int tid = threadIdx.x;
  const int ug = tid % NUG, rg = tid / NUG, u0 = ug * TU;
  float* sw = smem + S::W_OFF;
  float* X = smem + S::X_OFF;
  float* H = smem + S::H_OFF;
  float* D = smem + S::D_OFF;
  float* G = smem + S::G_OFF;

  // ---- A: gates and gate gradients
  {
    float az[TR][TU], ar[TR][TU], axh[TR][TU], ahh[TR][TU];
    gate_gemm<FI, U>(sw, X, H, rg, u0, az, ar, axh, ahh);
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      const int r = rg + i * NRG;
#pragma unroll
      for (int j = 0; j < TU; ++j) {
        float g_az = 0.f, g_ar = 0.f, g_xh = 0.f, g_hh = 0.f;
        if (active[i]) {
          const float z = sigmoid_f(az[i][j]);
          const float rr = sigmoid_f(ar[i][j]);
          const float hh = tanhf(fmaf(rr, ahh[i][j], axh[i][j]));
          const float hold = H[r * XS + u0 + j];
          const float dh = D[r * XS + u0 + j];
          g_xh = dh * (1.0f - z) * (1.0f - hh * hh);
          g_hh = g_xh * rr;
          g_ar = g_xh * ahh[i][j] * rr * (1.0f - rr);
          g_az = dh * (hold - hh) * z * (1.0f - z);
          D[r * XS + u0 + j] = dh * z;                    // direct term of dL/dh_{t-1}
        }
        G[r * GS + u0 + j] = g_az;
        G[r * GS + U + u0 + j] = g_ar;
        G[r * GS + 2 * U + u0 + j] = g_xh;
        G[r * GS + 3 * U + u0 + j] = g_hh;
      }
    }
  }
  __syncthreads();

  // ---- B: dx = GX K^T, dh += GH R^T   (thread owns rows rg+i*NRG, outputs u0..u0+TU)
  {
    float dx[TR][TU], dh[TR][TU];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < TU; ++j) { dx[i][j] = 0.f; dh[i][j] = 0.f; }
    const float* KT = smem + S::KT_OFF + u0;
    const float* RT = smem + S::RT_OFF + u0;
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
      const int gx_off = g * U;                       // GX: az | ar | axh
      const int gh_off = (g == 2) ? 3 * U : g * U;    // GH: az | ar | ahh
#pragma unroll 2
      for (int c = 0; c < U; c += 4) {
        float4 gx[TR], gh[TR];
#pragma unroll
        for (int i = 0; i < TR; ++i) {
          gx[i] = *reinterpret_cast<const float4*>(G + (rg + i * NRG) * GS + gx_off + c);
          gh[i] = (g == 2) ? *reinterpret_cast<const float4*>(G + (rg + i * NRG) * GS + gh_off + c) : gx[i];
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float wk[TU], wr[TU];
          VecLoad<TU>::ld(KT + (g * U + c + cc) * FI, wk);
          VecLoad<TU>::ld(RT + (g * U + c + cc) * U, wr);
#pragma unroll
          for (int i = 0; i < TR; ++i) {
            const float a = cc == 0 ? gx[i].x : cc == 1 ? gx[i].y : cc == 2 ? gx[i].z : gx[i].w;
            const float b = cc == 0 ? gh[i].x : cc == 1 ? gh[i].y : cc == 2 ? gh[i].z : gh[i].w;
#pragma unroll
            for (int j = 0; j < TU; ++j) {
              dx[i][j] = fmaf(a, wk[j], dx[i][j]);
              dh[i][j] = fmaf(b, wr[j], dh[i][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      if (!active[i]) continue;
      const int r = rg + i * NRG;
#pragma unroll
      for (int j = 0; j < TU; ++j) D[r * XS + u0 + j] += dh[i][j];
      if (WRITE_DX && dx_row[i]) {
#pragma unroll
        for (int j = 0; j < TU; ++j) dx_row[i][u0 + j] = dx[i][j];
      }
    }
  }

  // ---- C: weight gradients (rows of this thread's split), bias gradients (threads < 6U)
  {
    using WA = WAcc<U>;
    constexpr int CB = WA::CB;
    const int split = tid / WA::TPS, t = tid % WA::TPS;
    const int kb = t / 8, cb = t % 8;
    const bool hrow = kb * 4 >= FI;                         // stacked rows: [0,FI) -> x, [FI,FI+U) -> h
    const float* A = hrow ? (H + (kb * 4 - FI)) : (X + kb * 4);
    int gcol[CB / 2];
#pragma unroll
    for (int q = 0; q < CB / 2; ++q) {
      const int c = cb * CB + 2 * q;
      gcol[q] = (hrow && c >= 2 * U) ? c + U : c;
    }
    constexpr int RPS = R / WA::NSPLIT;
#pragma unroll 2
    for (int rr = 0; rr < RPS; ++rr) {
      const int r = split * RPS + rr;
      const float4 av = *reinterpret_cast<const float4*>(A + r * XS);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      float gv[CB];
#pragma unroll
      for (int q = 0; q < CB / 2; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(G + r * GS + gcol[q]);
        gv[2 * q] = v.x; gv[2 * q + 1] = v.y;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < CB; ++c) wacc.a[k][c] = fmaf(a4[k], gv[c], wacc.a[k][c]);
    }
    if (tid < 6 * U) {                                       // db0 = colsum(GX), db1 = colsum(GH)
      const int which = tid / (3 * U), c = tid % (3 * U);
      const int col = (which == 1 && c >= 2 * U) ? c + U : c;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
      for (int r = 0; r < R; r += 2) { s0 += G[r * GS + col]; s1 += G[(r + 1) * GS + col]; }
      bacc[0] += s0 + s1;
    }
  }
  __syncthreads();
}

template <int U>
__device__ __forceinline__ void flush_wacc(const WAcc<U>& wacc, const float (&bacc)[2], float* __restrict__ dK,
                                           float* __restrict__ dR, float* __restrict__ dB) {
  using WA = WAcc<U>;
  constexpr int FI = U, CB = WA::CB;
  const int tid = threadIdx.x;
  const int t = tid % WA::TPS, kb = t / 8, cb = t % 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = kb * 4 + k;
    float* base = row < FI ? dK + row * 3 * U : dR + (row - FI) * 3 * U;
#pragma unroll
    for (int c = 0; c < CB; ++c) atomicAdd(base + cb * CB + c, wacc.a[k][c]);
  }
  if (tid < 6 * U) atomicAdd(dB + tid, bacc[0]);
}

// ---------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(THREADS, 1) gru_cell_bwd_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ h, int64_t n,
                                                                  const float* __restrict__ kernel,
                                                                  const float* __restrict__ rkernel,
                                                                  const float* __restrict__ bias,
                                                                  const float* __restrict__ d_out,
                                                                  float* __restrict__ dx, float* __restrict__ dh,
                                                                  float* __restrict__ dK, float* __restrict__ dR,
                                                                  float* __restrict__ dB) {
  using S = BSmem<U>;
  using T = Tile<U>;
  constexpr int R = T::R, TR = T::TR, NUG = T::NUG, NRG = T::NRG, XS = S::XS;
  extern __shared__ float4 smem_f4[];
  float* smem = reinterpret_cast<float*>(smem_f4);
  load_weights<U, U>(smem + S::W_OFF, kernel, rkernel, bias);
  load_transposed<U>(smem, kernel, rkernel);
  const int tid = threadIdx.x, rg = tid / NUG;
  WAcc<U> wacc;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < WAcc<U>::CB; ++c) wacc.a[k][c] = 0.f;
  float bacc[2] = {0.f, 0.f};
  const int64_t ntiles = (n + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t d0 = tile * R;
    for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
      const int r = idx / (U / 4), c4 = idx % (U / 4);
      float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vh = vx, vd = vx;
      if (d0 + r < n) {
        vx = ldg_f4(x + (d0 + r) * U + c4 * 4);
        vh = ldg_f4(h + (d0 + r) * U + c4 * 4);
        vd = ldg_f4(d_out + (d0 + r) * U + c4 * 4);
      }
      st_f4(smem + S::X_OFF + r * XS + c4 * 4, vx);
      st_f4(smem + S::H_OFF + r * XS + c4 * 4, vh);
      st_f4(smem + S::D_OFF + r * XS + c4 * 4, vd);
    }
    __syncthreads();
    bool active[TR];
    float* dx_row[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      const int64_t d = d0 + rg + i * NRG;
      active[i] = d < n;
      dx_row[i] = (active[i] && dx) ? dx + d * U : nullptr;
    }
    bwd_tile_step<U, true>(smem, active, wacc, bacc, dx_row);
    if (dh) {
      for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
        const int r = idx / (U / 4), c4 = idx % (U / 4);
        if (d0 + r < n)
          st_f4(dh + (d0 + r) * U + c4 * 4, *reinterpret_cast<const float4*>(smem + S::D_OFF + r * XS + c4 * 4));
      }
    }
  }
  flush_wacc<U>(wacc, bacc, dK, dR, dB);
}

template <int U>
__global__ void __launch_bounds__(THREADS, 1) gru_seq_bwd_kernel(
    const int* __restrict__ steps_rowptr, const int* __restrict__ steps, const int* __restrict__ order,
    SrcPtrs srcs, const float* __restrict__ h0, const float* __restrict__ h_seq, int64_t num_dst,
    const float* __restrict__ kernel, const float* __restrict__ rkernel, const float* __restrict__ bias,
    const float* __restrict__ d_out, float* __restrict__ d_steps, float* __restrict__ dh0, float* __restrict__ dK,
    float* __restrict__ dR, float* __restrict__ dB) {
  using S = BSmem<U>;
  using T = Tile<U>;
  constexpr int R = T::R, TR = T::TR, NUG = T::NUG, NRG = T::NRG, XS = S::XS;
  extern __shared__ float4 smem_f4[];
  float* smem = reinterpret_cast<float*>(smem_f4);
  int* row_dst = reinterpret_cast<int*>(smem + S::META_OFF);
  int* row_lo = row_dst + R;
  int* row_len = row_lo + R;
  int* s_maxlen = row_len + R;
  load_weights<U, U>(smem + S::W_OFF, kernel, rkernel, bias);
  load_transposed<U>(smem, kernel, rkernel);
  const int tid = threadIdx.x, rg = tid / NUG;
  WAcc<U> wacc;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < WAcc<U>::CB; ++c) wacc.a[k][c] = 0.f;
  float bacc[2] = {0.f, 0.f};
  const int64_t ntiles = (num_dst + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
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
    for (int idx = tid; idx < R * (U / 4); idx += THREADS) {      // dL/dh_final
      const int r = idx / (U / 4), c4 = idx % (U / 4);
      const int d = row_dst[r];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d >= 0) v = ldg_f4(d_out + (int64_t)d * U + c4 * 4);
      st_f4(smem + S::D_OFF + r * XS + c4 * 4, v);
    }
    __syncthreads();
    const int maxlen = *s_maxlen;
    for (int t = maxlen - 1; t >= 0; --t) {
      // x_t (gathered message) and h_{t-1} of the rows still inside their sequence
      for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
        const int r = idx / (U / 4), c4 = idx % (U / 4);
        float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vh = vx;
        if (row_len[r] > t) {
          const int entry = steps[row_lo[r] + t];
          if (entry >= 0)
            vx = ldg_f4(pick_src(srcs, entry >> IGN_STEP_SRC_SHIFT) + (int64_t)(entry & IGN_STEP_ROW_MASK) * U + c4 * 4);
          vh = (t == 0) ? ldg_f4(h0 + (int64_t)row_dst[r] * U + c4 * 4)
                        : ldg_f4(h_seq + (int64_t)(row_lo[r] + t - 1) * U + c4 * 4);
        }
        st_f4(smem + S::X_OFF + r * XS + c4 * 4, vx);
        st_f4(smem + S::H_OFF + r * XS + c4 * 4, vh);
      }
      __syncthreads();
      bool active[TR];
      float* dx_row[TR];
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const int r = rg + i * NRG;
        active[i] = row_len[r] > t;
        dx_row[i] = (active[i] && d_steps) ? d_steps + (int64_t)(row_lo[r] + t) * U : nullptr;
      }
      bwd_tile_step<U, true>(smem, active, wacc, bacc, dx_row);
    }
    if (dh0) {
      for (int idx = tid; idx < R * (U / 4); idx += THREADS) {
        const int r = idx / (U / 4), c4 = idx % (U / 4);
        const int d = row_dst[r];
        if (d >= 0)
          st_f4(dh0 + (int64_t)d * U + c4 * 4, *reinterpret_cast<const float4*>(smem + S::D_OFF + r * XS + c4 * 4));
      }
    }
  }
  flush_wacc<U>(wacc, bacc, dK, dR, dB);
}

template <typename KernelT>
int bwd_grid(KernelT k, size_t smem, int64_t ntiles, int* grid) {
  // once per (kernel, device): kernels of one signature share this instantiation, and function attributes are per device
  static thread_local const void* done_fn[16];
  static thread_local int done_dev[16];
  const void* fn = reinterpret_cast<const void*>(k);
  int cur = 0;
  cudaGetDevice(&cur);
  bool seen = false;
  int slot = -1;
  for (int i = 0; i < 16; ++i) {
    if (done_fn[i] == fn && done_dev[i] == cur) { seen = true; break; }
    if (!done_fn[i] && slot < 0) slot = i;
  }
  if (!seen) {
    IGN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (slot >= 0) { done_fn[slot] = fn; done_dev[slot] = cur; }
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  *grid = (int)(ntiles < sms ? ntiles : sms);
  return IGN_OK;
}

// ---- generic GRU-cell backward (any f_in, units): the element-wise middle of a composition of Dense primitives.
//   zx = x K + b_in, zh = h R + b_rec                  (ign_dense, linear)
//   this kernel: zx <- GX = [d_az | d_ar | d_axh],  zh <- GH = [d_az | d_ar | d_ahh],  dh_direct = dL/dh' * z
//   dx = GX K^T, dK += x^T GX, db_in += colsum GX;  dh = dh_direct + GH R^T, dR += h^T GH, db_rec += colsum GH
//                                                      (ign_dense_bwd, linear)
// Same formulas as bwd_tile_step above (Keras GRUCell v2, reset_after).
__global__ void gru_gates_bwd_kernel(float* __restrict__ zx, float* __restrict__ zh, const float* __restrict__ h,
                                     const float* __restrict__ d_out, int64_t n, int U, float* __restrict__ dh_direct) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * U) return;
  const int64_t r = i / U;
  const int u = (int)(i - r * U);
  float* px = zx + r * 3 * U + u;
  float* ph = zh + r * 3 * U + u;
  const float z = sigmoid_f(px[0] + ph[0]);
  const float rr = sigmoid_f(px[U] + ph[U]);
  const float ahh = ph[2 * U];
  const float hh = tanhf(fmaf(rr, ahh, px[2 * U]));
  const float dh = d_out[i];
  const float g_xh = dh * (1.0f - z) * (1.0f - hh * hh);
  const float g_az = dh * (h[i] - hh) * z * (1.0f - z);
  const float g_ar = g_xh * ahh * rr * (1.0f - rr);
  px[0] = g_az; px[U] = g_ar; px[2 * U] = g_xh;
  ph[0] = g_az; ph[U] = g_ar; ph[2 * U] = g_xh * rr;
  dh_direct[i] = dh * z;
}

int check_bwd_shape(const char* who, int f_in, int units) {
  IGN_REQUIRE(f_in == units && (units == 16 || units == 32), IGN_ERR_UNSUPPORTED,
              "IGNNITION: %s: the backward pass is built for message width == units in {16, 32} "
              "(got %d, %d)", who, f_in, units);
  return IGN_OK;
}

}  // namespace

// fp32 launchers of this compilation's tile geometry (see gru.cuh)
int IGN_GRU_FN(ign_gru_cell_bwd_fp32)(const float* x, const float* h, int64_t n, int units, const float* kernel,
                                      const float* recurrent_kernel, const float* bias, const float* d_out, float* dx,
                                      float* dh, float* d_kernel, float* d_recurrent_kernel, float* d_bias,
                                      cudaStream_t st) {
  int grid = 0, rc;
  if (units == 32) {
    rc = bwd_grid(gru_cell_bwd_kernel<32>, BSmem<32>::BYTES, ign_cdiv(n, Tile<32>::R), &grid);
    if (rc) return rc;
    gru_cell_bwd_kernel<32><<<grid, THREADS, BSmem<32>::BYTES, st>>>(x, h, n, kernel, recurrent_kernel, bias, d_out,
                                                                      dx, dh, d_kernel, d_recurrent_kernel, d_bias);
  } else {
    rc = bwd_grid(gru_cell_bwd_kernel<16>, BSmem<16>::BYTES, ign_cdiv(n, Tile<16>::R), &grid);
    if (rc) return rc;
    gru_cell_bwd_kernel<16><<<grid, THREADS, BSmem<16>::BYTES, st>>>(x, h, n, kernel, recurrent_kernel, bias, d_out,
                                                                      dx, dh, d_kernel, d_recurrent_kernel, d_bias);
  }
  IGN_CHECK_LAUNCH("gru_cell_bwd");
  return IGN_OK;
}

int IGN_GRU_FN(ign_gru_seq_bwd_fp32)(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                                     const float* const* srcs, const float* h0, const float* h_seq, int64_t num_dst,
                                     int units, const float* kernel, const float* recurrent_kernel, const float* bias,
                                     const float* d_out, float* d_steps, float* dh0, float* d_kernel,
                                     float* d_recurrent_kernel, float* d_bias, cudaStream_t st) {
  SrcPtrs sp;
  for (int i = 0; i < IGN_MAX_SOURCES; ++i) sp.p[i] = i < n_src ? srcs[i] : nullptr;
  int grid = 0, rc;
  if (units == 32) {
    rc = bwd_grid(gru_seq_bwd_kernel<32>, BSmem<32>::BYTES, ign_cdiv(num_dst, Tile<32>::R), &grid);
    if (rc) return rc;
    gru_seq_bwd_kernel<32><<<grid, THREADS, BSmem<32>::BYTES, st>>>(steps_rowptr, steps, order, sp, h0, h_seq, num_dst,
                                                                     kernel, recurrent_kernel, bias, d_out, d_steps,
                                                                     dh0, d_kernel, d_recurrent_kernel, d_bias);
  } else {
    rc = bwd_grid(gru_seq_bwd_kernel<16>, BSmem<16>::BYTES, ign_cdiv(num_dst, Tile<16>::R), &grid);
    if (rc) return rc;
    gru_seq_bwd_kernel<16><<<grid, THREADS, BSmem<16>::BYTES, st>>>(steps_rowptr, steps, order, sp, h0, h_seq, num_dst,
                                                                     kernel, recurrent_kernel, bias, d_out, d_steps,
                                                                     dh0, d_kernel, d_recurrent_kernel, d_bias);
  }
  IGN_CHECK_LAUNCH("gru_seq_bwd");
  return IGN_OK;
}

#ifndef IGN_GRU_SMALL_TILE
int ign_gru_cell_bwd_fp32_small_tile(const float* x, const float* h, int64_t n, int units, const float* kernel,
                                     const float* recurrent_kernel, const float* bias, const float* d_out, float* dx,
                                     float* dh, float* d_kernel, float* d_recurrent_kernel, float* d_bias,
                                     cudaStream_t st);
int ign_gru_seq_bwd_fp32_small_tile(const int* steps_rowptr, const int* steps, const int* order, int n_src,
                                    const float* const* srcs, const float* h0, const float* h_seq, int64_t num_dst,
                                    int units, const float* kernel, const float* recurrent_kernel, const float* bias,
                                    const float* d_out, float* d_steps, float* dh0, float* d_kernel,
                                    float* d_recurrent_kernel, float* d_bias, cudaStream_t st);

extern "C" int ign_gru_cell_bwd(const float* x, const float* h, int64_t n, int f_in, int units,
                                const float* kernel, const float* recurrent_kernel, const float* bias,
                                const float* d_out, float* dx, float* dh, float* d_kernel,
                                float* d_recurrent_kernel, float* d_bias, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: gru_cell_bwd: negative size");
  int rc = check_bwd_shape("gru_cell_bwd", f_in, units);
  if (rc) return rc;
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(x && h && kernel && recurrent_kernel && bias && d_out && d_kernel && d_recurrent_kernel && d_bias,
              IGN_ERR_INVALID, "IGNNITION: gru_cell_bwd: null pointer");
  return (ign_gru_use_small_tile(n) ? ign_gru_cell_bwd_fp32_small_tile : ign_gru_cell_bwd_fp32)(
      x, h, n, units, kernel, recurrent_kernel, bias, d_out, dx, dh, d_kernel, d_recurrent_kernel, d_bias,
      ign_stream(stream));
}

extern "C" int ign_gru_gates_bwd(float* zx, float* zh, const float* h, const float* d_out, int64_t n, int units,
                                 float* dh_direct, void* stream) {
  IGN_REQUIRE(n >= 0 && units > 0, IGN_ERR_INVALID, "IGNNITION: gru_gates_bwd: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(zx && zh && h && d_out && dh_direct, IGN_ERR_INVALID, "IGNNITION: gru_gates_bwd: null pointer");
  gru_gates_bwd_kernel<<<(unsigned)ign_cdiv(n * units, 256), 256, 0, ign_stream(stream)>>>(zx, zh, h, d_out, n, units,
                                                                                          dh_direct);
  IGN_CHECK_LAUNCH("gru_gates_bwd");
  return IGN_OK;
}

extern "C" int ign_gru_seq_bwd(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* order, int n_src,
                               const float* const* srcs, int f_in, const float* h0, const float* h_seq,
                               int64_t num_dst, int units, const float* kernel, const float* recurrent_kernel,
                               const float* bias, const float* d_out, float* d_steps, float* dh0, float* d_kernel,
                               float* d_recurrent_kernel, float* d_bias, void* stream) {
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: gru_seq_bwd: negative size");
  IGN_REQUIRE(n_src >= 1 && n_src <= IGN_MAX_SOURCES && srcs, IGN_ERR_INVALID, "IGNNITION: gru_seq_bwd: bad sources");
  int rc = check_bwd_shape("gru_seq_bwd", f_in, units);
  if (rc) return rc;
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(steps_rowptr && steps && h0 && h_seq && kernel && recurrent_kernel && bias && d_out && d_kernel &&
                  d_recurrent_kernel && d_bias,
              IGN_ERR_INVALID, "IGNNITION: gru_seq_bwd: null pointer");
  return (ign_gru_use_small_tile(num_dst) ? ign_gru_seq_bwd_fp32_small_tile : ign_gru_seq_bwd_fp32)(
      steps_rowptr, steps, order, n_src, srcs, h0, h_seq, num_dst, units, kernel, recurrent_kernel, bias, d_out, d_steps,
      dh0, d_kernel, d_recurrent_kernel, d_bias, ign_stream(stream));
}

// step-synchronous tensor-core variant (gru_step_bwd_tc.cu)
size_t ign_gru_step_bwd_tc_ws(int64_t num_dst);
int ign_gru_step_bwd_tc_launch(int max_steps, const int* nt, const int* off, int64_t num_dst, const int* meta,
                               const int* steps_T, int n_src, const float* const* srcs, const float* h0,
                               const float* h_seq, const float* kernel, const float* rkernel, const float* bias,
                               const float* d_out, float* d_steps, float* dh0, float* dK, float* dR, float* dB,
                               void* ws, cudaStream_t st);

extern "C" size_t ign_gru_seq_bwd_steps_ws_bytes(int64_t num_dst) {
  return num_dst >= 0 ? ign_gru_step_bwd_tc_ws(num_dst) : 0;
}

extern "C" int ign_gru_seq_bwd_steps(int max_steps, const int32_t* nt, const int32_t* off, const int32_t* meta,
                                     const int32_t* steps_T, int n_src, const float* const* srcs, int f_in,
                                     const float* h0, const float* h_seq, int64_t num_dst, int units,
                                     const float* kernel, const float* recurrent_kernel, const float* bias,
                                     const float* d_out, float* d_steps, float* dh0, float* d_kernel,
                                     float* d_recurrent_kernel, float* d_bias, void* ws, size_t ws_bytes,
                                     void* stream) {
  IGN_REQUIRE(num_dst >= 0 && max_steps >= 1 && max_steps <= 1024, IGN_ERR_INVALID,
              "IGNNITION: gru_seq_bwd_steps: bad argument (1 <= max_steps <= 1024)");
  IGN_REQUIRE(n_src >= 1 && n_src <= IGN_MAX_SOURCES && srcs, IGN_ERR_INVALID,
              "IGNNITION: gru_seq_bwd_steps: bad sources");
  IGN_REQUIRE(f_in == 32 && units == 32, IGN_ERR_UNSUPPORTED,
              "IGNNITION: gru_seq_bwd_steps: built for 32-wide messages and states (got %d, %d)", f_in, units);
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(nt && off && meta && steps_T && h0 && h_seq && kernel && recurrent_kernel && bias && d_out && d_steps &&
                  dh0 && d_kernel && d_recurrent_kernel && d_bias,
              IGN_ERR_INVALID, "IGNNITION: gru_seq_bwd_steps: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= ign_gru_step_bwd_tc_ws(num_dst), IGN_ERR_WORKSPACE,
              "IGNNITION: gru_seq_bwd_steps: workspace too small (%zu < %zu)", ws_bytes,
              ign_gru_step_bwd_tc_ws(num_dst));
  return ign_gru_step_bwd_tc_launch(max_steps, nt, off, num_dst, meta, steps_T, n_src, srcs, h0, h_seq, kernel,
                                    recurrent_kernel, bias, d_out, d_steps, dh0, d_kernel, d_recurrent_kernel, d_bias,
                                    ws, ign_stream(stream));
}

#endif  // IGN_GRU_SMALL_TILE
