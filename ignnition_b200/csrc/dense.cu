// Dense layers y = act(x W + b) for the message / update / readout MLPs (sm_100a, fp32).
//
// Replaces the Keras functional Model of Dense layers the reference builds in
// Feed_forward_model.construct_tf_model (code/utils/auxilary_classes.py:918-975) and calls at
// code/utils/generate_model.py:468 (message), :600 (update) and :624 (readout).
//
// This file is the fp32 CUDA-core path: 128x128x16 shared-memory tiles, 8x8 register tiles,
// bias + activation fused in the epilogue (the pre-activation is optionally kept for the
// backward pass).  Layers with a handful of outputs (the 256 -> 1 readout head) use a
// warp-per-row dot product instead, which is HBM-bound.  The same tile kernel, with transposed
// operand reads, serves the backward products dX = dZ W^T and dW = X^T dZ.

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int64_t DENSE_TC_MIN_ROWS = 4096;     // fewer rows: the fp32 kernels (see ign_dense)
constexpr int LDA_S = BM + 4, LDB_S = BN + 4;

// C[M,N] (+)= op(A) op(B);  TA: A is stored [K,M] (read transposed); TB: B is stored [N,K].
// EPI 0: C = act(acc + bias[n]), optional pre-activation store   (forward)
// EPI 1: C = acc                                                (dX)
// EPI 2: atomicAdd(C, acc) over a split of the reduction dim     (dW; grid.z = splits)
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ C, int64_t M, int N, int64_t K,
                                                            const float* __restrict__ bias, int act,
                                                            float* __restrict__ pre, int64_t k_per_split) {
  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (EPI == 2) ? (int64_t)blockIdx.z * k_per_split : 0;
  const int64_t kend = (EPI == 2) ? min(K, kbeg + k_per_split) : K;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  // global -> register staging: 128x16 A tile and 16x128 B tile, 8 floats per thread each
  float ra[8], rb[8];
  auto load_tiles = [&](int64_t k0) {
    if (!TA) {          // A[m, k] row-major, lda = K: thread reads 4 consecutive k of 2 rows
      const int kq = (tid % 4) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t m = m0 + tid / 4 + 64 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t k = k0 + kq + j;
          ra[i * 4 + j] = (m < M && k < kend) ? __ldg(A + m * K + k) : 0.0f;
        }
      }
    } else {            // A stored [K, M]: thread reads 4 consecutive m of 2 k rows
      const int mq = (tid % 32) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t k = k0 + tid / 32 + 8 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t m = m0 + mq + j;
          ra[i * 4 + j] = (m < M && k < kend) ? __ldg(A + k * M + m) : 0.0f;
        }
      }
    }
    if (!TB) {          // B[k, n] row-major, ldb = N
      const int nq = (tid % 32) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t k = k0 + tid / 32 + 8 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + nq + j;
          rb[i * 4 + j] = (n < N && k < kend) ? __ldg(B + k * N + n) : 0.0f;
        }
      }
    } else {            // B stored [N, K]
      const int kq = (tid % 4) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n = n0 + tid / 4 + 64 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t k = k0 + kq + j;
          rb[i * 4 + j] = (n < N && k < kend) ? __ldg(B + (int64_t)n * K + k) : 0.0f;
        }
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (!TA) {
      const int kq = (tid % 4) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) As[buf][kq + j][tid / 4 + 64 * i] = ra[i * 4 + j];
    } else {
      const int mq = (tid % 32) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(&As[buf][tid / 32 + 8 * i][mq]) =
            make_float4(ra[i * 4], ra[i * 4 + 1], ra[i * 4 + 2], ra[i * 4 + 3]);
    }
    if (!TB) {
      const int nq = (tid % 32) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(&Bs[buf][tid / 32 + 8 * i][nq]) =
            make_float4(rb[i * 4], rb[i * 4 + 1], rb[i * 4 + 2], rb[i * 4 + 3]);
    } else {
      const int kq = (tid % 4) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[buf][kq + j][tid / 4 + 64 * i] = rb[i * 4 + j];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI == 0) {
        if (bias) v += bias[n];
        if (pre) pre[m * N + n] = v;
        C[m * N + n] = act_fwd(act, v);
      } else if (EPI == 1) {
        C[m * N + n] = v;
      } else {
        atomicAdd(C + m * N + n, v);
      }
    }
  }
}

// few outputs (N <= 8): one warp per row, HBM-bound
template <int NMAX>
__global__ void __launch_bounds__(256) dense_small_n_kernel(const float* __restrict__ x, int64_t M, int K,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            int N, int act, float* __restrict__ y,
                                                            float* __restrict__ pre) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.0f;
  for (int k = lane; k < K; k += 32) {
    const float xv = __ldg(x + row * K + k);
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) acc[n] = fmaf(xv, __ldg(w + (int64_t)k * N + n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < N) {
        float v = acc[n] + (bias ? bias[n] : 0.0f);
        if (pre) pre[row * N + n] = v;
        y[row * N + n] = act_fwd(act, v);
      }
    }
  }
}

// backward of a single-output layer (the readout head, K -> 1):  dx[m, k] = dz[m] w[k],  dw[k] += sum_m x[m, k] dz[m].
// Two HBM streams (x in, dx out); a warp owns whole rows, a lane the same float4 columns of every row, so the
// weight-gradient partial sums stay in registers until the CTA adds them once.
template <int KV>
__global__ void __launch_bounds__(256) dense_bwd_head_kernel(const float* __restrict__ x, int64_t M,
                                                             const float* __restrict__ w, const float* __restrict__ dz,
                                                             float* __restrict__ dx, float* __restrict__ dw,
                                                             int prev_act, float* __restrict__ db_prev) {
  // prev_act >= 0: x is the OUTPUT of the previous layer with that activation; dx is then written as that layer's dZ =
  // (dz w) * act'(x) (act' from the output, which this kernel reads anyway) and its bias gradient is summed here, so the
  // previous layer needs no activation pass of its own
  constexpr int K = KV * 128;
  __shared__ float4 red[8][KV * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 wv[KV], acc[KV], bacc[KV];
#pragma unroll
  for (int v = 0; v < KV; ++v) {
    wv[v] = ldg_f4(w + (v * 32 + lane) * 4);
    acc[v] = bacc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < M; r += nwarps) {
    const float g = __ldg(dz + r);
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dw || prev_act >= 0) xv = ld_stream_f4(x + r * K + (v * 32 + lane) * 4);
      if (dw) {
        acc[v].x = fmaf(xv.x, g, acc[v].x); acc[v].y = fmaf(xv.y, g, acc[v].y);
        acc[v].z = fmaf(xv.z, g, acc[v].z); acc[v].w = fmaf(xv.w, g, acc[v].w);
      }
      if (dx) {
        float4 d = make_float4(g * wv[v].x, g * wv[v].y, g * wv[v].z, g * wv[v].w);
        if (prev_act >= 0) {
          d.x *= act_bwd_from_output(prev_act, xv.x); d.y *= act_bwd_from_output(prev_act, xv.y);
          d.z *= act_bwd_from_output(prev_act, xv.z); d.w *= act_bwd_from_output(prev_act, xv.w);
          bacc[v].x += d.x; bacc[v].y += d.y; bacc[v].z += d.z; bacc[v].w += d.w;
        }
        st_f4(dx + r * K + (v * 32 + lane) * 4, d);
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int v = 0; v < KV; ++v) red[warp][v * 32 + lane] = acc[v];
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += 256) {
      float s = 0.0f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += reinterpret_cast<const float*>(&red[wq][0])[i];
      atomicAdd(dw + i, s);
    }
  }
  if (db_prev && prev_act >= 0) {
    __syncthreads();
#pragma unroll
    for (int v = 0; v < KV; ++v) red[warp][v * 32 + lane] = bacc[v];
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += 256) {
      float s = 0.0f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += reinterpret_cast<const float*>(&red[wq][0])[i];
      atomicAdd(db_prev + i, s);
    }
  }
}

// single output, K a multiple of 128 (the readout head on 256 features): a warp streams whole rows as float4, two rows in
// flight per iteration, the weights stay in registers
template <int KV>
__global__ void __launch_bounds__(256) dense_head_fwd_kernel(const float* __restrict__ x, int64_t M,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             int act, float* __restrict__ y, float* __restrict__ pre) {
  constexpr int K = KV * 128;
  const int lane = threadIdx.x & 31;
  float4 wv[KV];
#pragma unroll
  for (int v = 0; v < KV; ++v) wv[v] = ldg_f4(w + (v * 32 + lane) * 4);
  const float b = bias ? __ldg(bias) : 0.0f;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t r = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2; r < M; r += nwarps * 2) {
    float acc[2] = {0.f, 0.f};
    float4 xv[2][KV];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int v = 0; v < KV; ++v)
        xv[i][v] = (r + i < M) ? ld_stream_f4(x + (r + i) * K + (v * 32 + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int v = 0; v < KV; ++v) {
        acc[i] = fmaf(xv[i][v].x, wv[v].x, acc[i]); acc[i] = fmaf(xv[i][v].y, wv[v].y, acc[i]);
        acc[i] = fmaf(xv[i][v].z, wv[v].z, acc[i]); acc[i] = fmaf(xv[i][v].w, wv[v].w, acc[i]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane < 2 && r + lane < M) {
      const float v = (lane == 0 ? acc[0] : acc[1]) + b;
      if (pre) pre[r + lane] = v;
      y[r + lane] = act_fwd(act, v);
    }
  }
}

// dz = dy * act'(pre), in place; db += column sums of dz (per-CTA partial, then atomics)
__global__ void __launch_bounds__(256) act_bwd_bias_kernel(float* __restrict__ dy, const float* __restrict__ pre,
                                                           int64_t M, int N, int act, float* __restrict__ db,
                                                           bool from_out, int rows_per_cta) {
  // each CTA covers rows_per_cta rows (64, fewer for a small batch: the rows of a thread are one dependent chain of
  // loads); thread t handles columns t, t+256, ...
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float s = 0.0f;
#pragma unroll 4
    for (int64_t r = r0; r < r1; ++r) {
      float g = dy[r * N + n];
      if (act != IGN_ACT_LINEAR) {
        g *= from_out ? act_bwd_from_output(act, pre[r * N + n]) : act_bwd(act, pre[r * N + n]);
        dy[r * N + n] = g;
      }
      s += g;
    }
    if (db) atomicAdd(db + n, s);
  }
}

// the same for few columns (N <= 8, the readout head): one thread per row, one atomic per column and CTA
__global__ void __launch_bounds__(256) act_bwd_bias_small_kernel(float* __restrict__ dy, const float* __restrict__ pre,
                                                                 int64_t M, int N, int act, float* __restrict__ db,
                                                                 bool from_out) {
  __shared__ float red[8][8];
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    s[n] = 0.0f;
    if (n < N && r < M) {
      float g = dy[r * N + n];
      if (act != IGN_ACT_LINEAR) {
        g *= from_out ? act_bwd_from_output(act, pre[r * N + n]) : act_bwd(act, pre[r * N + n]);
        dy[r * N + n] = g;
      }
      s[n] = g;
    }
  }
  if (!db) return;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[n] += __shfl_xor_sync(0xffffffffu, s[n], o);
    if (lane == 0) red[warp][n] = s[n];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.0f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += red[wq][threadIdx.x];
    atomicAdd(db + threadIdx.x, t);
  }
}

__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ label, int64_t n,
                           float grad_scale, float* __restrict__ d_pred, double* __restrict__ sse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double e2 = 0.0;
  if (i < n) {
    const float e = pred[i] - label[i];
    if (d_pred) d_pred[i] = 2.0f * e * grad_scale;
    e2 = (double)e * (double)e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = e2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(sse, s);
  }
}

__global__ void l2_kernel(const float* __restrict__ w, int64_t n, float lambda, float* __restrict__ dw,
                          double* __restrict__ reg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < n) {
    const float v = w[i];
    if (dw) dw[i] += 2.0f * lambda * v;
    s = (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += part[k];
    if (reg) atomicAdd(reg, (double)lambda * t);
  }
}

__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float beta1, float beta2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// Keras losses by name (generate_model.py:745-751 resolves any tf.keras.losses class): per-element loss l_i, the mean
// over all predictions is taken by the caller (acc = sum l_i in fp64); d_pred = dl_i/dp * grad_scale.
__global__ void loss_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ label, int64_t n,
                            float grad_scale, float delta, float* __restrict__ d_pred, double* __restrict__ acc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (i < n) {
    const float p = pred[i], y = label[i], e = p - y;
    float d = 0.f;
    constexpr float EPS = 1e-7f;                              // keras.backend.epsilon()
    switch (kind) {
      case IGN_LOSS_MSE: l = e * e; d = 2.f * e; break;
      case IGN_LOSS_MAE: l = fabsf(e); d = e > 0.f ? 1.f : e < 0.f ? -1.f : 0.f; break;
      case IGN_LOSS_MAPE: {
        const float den = fmaxf(fabsf(y), EPS);
        l = 100.f * fabsf(e) / den; d = (e > 0.f ? 100.f : e < 0.f ? -100.f : 0.f) / den; break;
      }
      case IGN_LOSS_MSLE: {
        const float pc = fmaxf(p, EPS), yc = fmaxf(y, EPS);
        const float q = log1pf(pc) - log1pf(yc);
        l = q * q; d = p > EPS ? 2.f * q / (1.f + pc) : 0.f; break;
      }
      case IGN_LOSS_HUBER: {
        const float a = fabsf(e);
        if (a <= delta) { l = 0.5f * e * e; d = e; } else { l = delta * (a - 0.5f * delta); d = e > 0.f ? delta : -delta; }
        break;
      }
      case IGN_LOSS_LOGCOSH: {
        const float a = fabsf(e);                             // log cosh x = |x| + log1p(exp(-2|x|)) - log 2
        l = a + log1pf(expf(-2.f * a)) - 0.69314718f; d = tanhf(e); break;
      }
      default: {                                              // IGN_LOSS_BCE on probabilities, clipped like Keras
        const float pc = fminf(fmaxf(p, EPS), 1.f - EPS);
        l = -(y * logf(pc + EPS) + (1.f - y) * logf(1.f - pc + EPS));
        d = (p > EPS && p < 1.f - EPS) ? -(y / (pc + EPS) - (1.f - y) / (1.f - pc + EPS)) : 0.f; break;
      }
    }
    if (d_pred) d_pred[i] = d * grad_scale;
  }
  double v = (double)l;                                       // block sum, one atomic per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(acc, t);
  }
}

// Keras optimisers by name (generate_model.py:796-818) [TF-2.1 fused-op semantics]; s1 / s2 are the slot buffers
__global__ void optimizer_kernel(int kind, float* __restrict__ w, const float* __restrict__ g, float* __restrict__ s1,
                                 float* __restrict__ s2, int64_t n, float lr, float a, float b, float eps, int flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  switch (kind) {
    case IGN_OPT_SGD:                                         // a = momentum, flags & 1 = nesterov
      if (a == 0.f) { w[i] -= lr * gi; break; }
      { const float v = a * s1[i] - lr * gi; s1[i] = v; w[i] += (flags & 1) ? a * v - lr * gi : v; }
      break;
    case IGN_OPT_RMSPROP: {                                   // a = rho, b = momentum (ResourceApplyRMSProp)
      const float ms = a * s1[i] + (1.f - a) * gi * gi;
      s1[i] = ms;
      const float mom = b * s2[i] + lr * gi * rsqrtf(ms + eps);
      s2[i] = mom;
      w[i] -= mom;
      break;
    }
    case IGN_OPT_ADAGRAD: {                                   // s1 = accumulator (caller initialises it, Keras: 0.1)
      const float acc = s1[i] + gi * gi;
      s1[i] = acc;
      w[i] -= lr * gi / (sqrtf(acc) + eps);
      break;
    }
    default: {                                                // IGN_OPT_ADAMAX: a = beta1, b = beta2, lr = lr / (1 - beta1^t)
      const float m = a * s1[i] + (1.f - a) * gi;
      const float u = fmaxf(b * s2[i], fabsf(gi));
      s1[i] = m; s2[i] = u;
      w[i] -= lr * m / (u + eps);
      break;
    }
  }
}

// generic GRU step (any f_in, units) from zx = x K + b_in and zh = h R + b_rec: the element-wise end of ign_dense x 2
__global__ void gru_gates_fwd_kernel(const float* __restrict__ zx, const float* __restrict__ zh,
                                     const float* __restrict__ h, int64_t n, int U, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * U) return;
  const int64_t r = i / U;
  const int u = (int)(i - r * U);
  const float* px = zx + r * 3 * U + u;
  const float* ph = zh + r * 3 * U + u;
  const float z = 1.f / (1.f + expf(-(px[0] + ph[0])));
  const float rr = 1.f / (1.f + expf(-(px[U] + ph[U])));
  const float hh = tanhf(fmaf(rr, ph[2 * U], px[2 * U]));
  out[i] = fmaf(z, h[i] - hh, hh);
}

// ---- GRUCell with reset_after = False (the Keras v1 cell a model JSON can ask for through the cell parameters,
// auxilary_classes.py:740-750): the reset gate multiplies h BEFORE the candidate's recurrent product,
//   zx = x K + b [n, 3U]; zh2 = h R[:, :2U] [n, 2U]; z = s(zx_z + zh2_z); r = s(zx_r + zh2_r);
//   hh = tanh(zx_h + (r * h) R[:, 2U:]);  h' = z h + (1 - z) hh
// so one step is three Dense products with two element-wise kernels between them.
__global__ void gru_v1_reset_kernel(const float* __restrict__ zx, const float* __restrict__ zh2,
                                    const float* __restrict__ h, int64_t n, int U, float* __restrict__ rh) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * U) return;
  const int64_t r = i / U;
  const int u = (int)(i - r * U);
  rh[i] = sigmoid_f(zx[r * 3 * U + U + u] + zh2[r * 2 * U + U + u]) * h[i];
}

__global__ void gru_v1_out_kernel(const float* __restrict__ zx, const float* __restrict__ zh2,
                                  const float* __restrict__ zhh, const float* __restrict__ h, int64_t n, int U,
                                  float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * U) return;
  const int64_t r = i / U;
  const int u = (int)(i - r * U);
  const float z = sigmoid_f(zx[r * 3 * U + u] + zh2[r * 2 * U + u]);
  const float hh = tanhf(zx[r * 3 * U + 2 * U + u] + zhh[i]);
  out[i] = fmaf(z, h[i] - hh, hh);
}

// backward, first half: the z and candidate slices of zx / zh2 / zhh become their gradients in place (the r slices stay
// as they are for the second half), dh_direct = d_out * z
__global__ void gru_v1_bwd_out_kernel(float* __restrict__ zx, float* __restrict__ zh2, float* __restrict__ zhh,
                                      const float* __restrict__ h, const float* __restrict__ d_out, int64_t n, int U,
                                      float* __restrict__ dh_direct) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * U) return;
  const int64_t r = i / U;
  const int u = (int)(i - r * U);
  float* px = zx + r * 3 * U + u;
  float* ph = zh2 + r * 2 * U + u;
  const float z = sigmoid_f(px[0] + ph[0]);
  const float hh = tanhf(px[2 * U] + zhh[i]);
  const float dh = d_out[i];
  const float g_h = dh * (1.0f - z) * (1.0f - hh * hh);
  const float g_z = dh * (h[i] - hh) * z * (1.0f - z);
  px[0] = g_z; ph[0] = g_z;
  px[2 * U] = g_h; zhh[i] = g_h;
  dh_direct[i] = dh * z;
}

// second half: d(r * h) arrives from the candidate's recurrent product; the r slices become their gradient,
// dh_direct += d_rh * r
__global__ void gru_v1_bwd_reset_kernel(float* __restrict__ zx, float* __restrict__ zh2, const float* __restrict__ h,
                                        const float* __restrict__ d_rh, int64_t n, int U, float* __restrict__ dh_direct) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * U) return;
  const int64_t r = i / U;
  const int u = (int)(i - r * U);
  float* px = zx + r * 3 * U + U + u;
  float* ph = zh2 + r * 2 * U + U + u;
  const float rr = sigmoid_f(px[0] + ph[0]);
  const float g_r = d_rh[i] * h[i] * rr * (1.0f - rr);
  px[0] = g_r; ph[0] = g_r;
  dh_direct[i] += d_rh[i] * rr;
}

}  // namespace

bool ign_tensor_cores_enabled();
// tensor-core path (dense_tc.cu)
bool ign_dense_tc_supported(int k, int n);
size_t ign_dense_tc_ws(int k, int n);
int ign_dense_tc_launch(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                        float* y, float* pre_act, void* ws, cudaStream_t st, const float* head_w = nullptr,
                        const float* head_b = nullptr, float* head_out = nullptr, bool w_transposed = false);

extern "C" size_t ign_dense_ws_bytes(int k, int n) {
  return (k > 0 && n > 0 && ign_dense_tc_supported(k, n)) ? ign_dense_tc_ws(k, n) : 0;
}

extern "C" int ign_dense(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                         float* y, float* pre_act, void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(m >= 0 && k > 0 && n > 0, IGN_ERR_INVALID, "IGNNITION: dense: bad shape");
  IGN_REQUIRE(act >= IGN_ACT_LINEAR && act <= IGN_ACT_LEAKY_RELU, IGN_ERR_INVALID,
              "IGNNITION: dense: unknown activation %d", act);
  if (m == 0) return IGN_OK;
  IGN_REQUIRE(x && w && y, IGN_ERR_INVALID, "IGNNITION: dense: null pointer");
  cudaStream_t st = ign_stream(stream);
  // (below 4096 rows the tensor-core kernel's fixed latency -- weight image, TMEM, barriers: ~30 us -- is three times
  // the fp32 kernel's whole run: 546 rows x 256 x 256 take 8 us on the CUDA cores)
  if (ws && ign_tensor_cores_enabled() && ign_dense_tc_supported(k, n) && ws_bytes >= ign_dense_tc_ws(k, n) &&
      m >= DENSE_TC_MIN_ROWS)
    return ign_dense_tc_launch(x, m, k, w, bias, n, act, y, pre_act, ws, st);
  if (n <= 8) {
    if (n == 1 && (k == 128 || k == 256 || k == 512) && m >= 1024) {
      int sms = IGN_NUM_SMS, dev = 0;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int64_t grid = ign_cdiv(m, 16 * 8);
      if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
      if (k == 128) dense_head_fwd_kernel<1><<<(unsigned)grid, 256, 0, st>>>(x, m, w, bias, act, y, pre_act);
      else if (k == 256) dense_head_fwd_kernel<2><<<(unsigned)grid, 256, 0, st>>>(x, m, w, bias, act, y, pre_act);
      else dense_head_fwd_kernel<4><<<(unsigned)grid, 256, 0, st>>>(x, m, w, bias, act, y, pre_act);
      IGN_CHECK_LAUNCH("dense_head_fwd");
      return IGN_OK;
    }
    dense_small_n_kernel<8><<<(unsigned)ign_cdiv(m * 32, 256), 256, 0, st>>>(x, m, k, w, bias, n, act, y, pre_act);
    IGN_CHECK_LAUNCH("dense_small_n");
    return IGN_OK;
  }
  dim3 grid((unsigned)ign_cdiv(m, BM), (unsigned)ign_cdiv(n, BN), 1);
  gemm_kernel<false, false, 0><<<grid, GEMM_THREADS, 0, st>>>(x, w, y, m, n, k, bias, act, pre_act, 0);
  IGN_CHECK_LAUNCH("dense");
  return IGN_OK;
}

// fused two-hidden-layer readout (mlp_head_tc.cu)
bool ign_mlp_head_tc_supported(int k1, int n1, int n2);
size_t ign_mlp_head_tc_ws(int k1, int n1, int n2);
int ign_mlp_head_tc_launch(const float* x, int64_t m, int k1, const float* w1, const float* b1, int n1, int act1,
                           const float* w2, const float* b2, int n2, int act2, const float* w3, const float* b3,
                           float* out, void* ws, cudaStream_t st);

extern "C" size_t ign_mlp_head_ws_bytes(int k1, int n1, int n2) {
  return ign_mlp_head_tc_supported(k1, n1, n2) ? ign_mlp_head_tc_ws(k1, n1, n2) : 0;
}

extern "C" int ign_mlp_head(const float* x, int64_t m, int k1, const float* w1, const float* b1, int n1, int act1,
                            const float* w2, const float* b2, int n2, int act2, const float* w3, const float* b3,
                            float* out, void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(m >= 0 && k1 > 0 && n1 > 0 && n2 > 0, IGN_ERR_INVALID, "IGNNITION: mlp_head: bad shape");
  if (m == 0) return IGN_OK;
  IGN_REQUIRE(x && w1 && w2 && w3 && out && ws, IGN_ERR_INVALID, "IGNNITION: mlp_head: null pointer");
  IGN_REQUIRE(ign_mlp_head_tc_supported(k1, n1, n2) && ws_bytes >= ign_mlp_head_tc_ws(k1, n1, n2) && m >= 128 &&
                  ign_tensor_cores_enabled(),
              IGN_ERR_UNSUPPORTED,
              "IGNNITION: mlp_head: built for K1 %% 32 == 0, N1, N2 %% 32 == 0 (<= 256), M >= 128, tensor cores on");
  return ign_mlp_head_tc_launch(x, m, k1, w1, b1, n1, act1, w2, b2, n2, act2, w3, b3, out, ws, ign_stream(stream));
}

extern "C" int ign_dense_head(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                              const float* head_w, const float* head_b, float* out, void* ws, size_t ws_bytes,
                              void* stream) {
  IGN_REQUIRE(m >= 0 && k > 0 && n > 0, IGN_ERR_INVALID, "IGNNITION: dense_head: bad shape");
  if (m == 0) return IGN_OK;
  IGN_REQUIRE(x && w && head_w && out && ws, IGN_ERR_INVALID, "IGNNITION: dense_head: null pointer");
  IGN_REQUIRE(ign_dense_tc_supported(k, n) && ws_bytes >= ign_dense_tc_ws(k, n) && m >= 128, IGN_ERR_UNSUPPORTED,
              "IGNNITION: dense_head: built for the tensor-core shapes only (K %% 32 == 0, N %% 32 == 0, M >= 128)");
  return ign_dense_tc_launch(x, m, k, w, bias, n, act, nullptr, nullptr, ws, ign_stream(stream), head_w,
                             head_b, out);
}

// dX = dZ W^T is a Dense layer with the transposed kernel: K' = n, N' = k
// tensor-core weight gradient (dw_tc.cu)
bool ign_dw_tc_supported(int k, int n);
int ign_dw_tc_launch(const float* x, const float* dz, int64_t m, int k, int n, float* dw, cudaStream_t st);

extern "C" size_t ign_dense_bwd_ws_bytes(int k, int n) {
  return (k > 0 && n > 0 && ign_dense_tc_supported(n, k)) ? ign_dense_tc_ws(n, k) : 0;
}

extern "C" int ign_dense_bwd(const float* x, int64_t m, int k, const float* w, int n, int act,
                             const float* pre_act, float* dy, float* dx, float* dw, float* db, void* ws,
                             size_t ws_bytes, void* stream) {
  IGN_REQUIRE(m >= 0 && k > 0 && n > 0, IGN_ERR_INVALID, "IGNNITION: dense_bwd: bad shape");
  const bool from_out = (act & IGN_ACT_FROM_OUTPUT) != 0;    // pre_act holds act(x W + b)
  act &= ~IGN_ACT_FROM_OUTPUT;
  if (m == 0) return IGN_OK;
  IGN_REQUIRE(x && w && dy, IGN_ERR_INVALID, "IGNNITION: dense_bwd: null pointer");
  IGN_REQUIRE(act == IGN_ACT_LINEAR || pre_act, IGN_ERR_INVALID, "IGNNITION: dense_bwd: pre-activation needed");
  cudaStream_t st = ign_stream(stream);
  if (act != IGN_ACT_LINEAR || db) {
    if (n <= 8) act_bwd_bias_small_kernel<<<(unsigned)ign_cdiv(m, 256), 256, 0, st>>>(dy, pre_act, m, n, act, db, from_out);
    else {
      int64_t rows = ign_cdiv(m, 2 * IGN_NUM_SMS);          // two CTAs per SM before a CTA takes more rows
      rows = rows < 4 ? 4 : rows > 64 ? 64 : rows;
      act_bwd_bias_kernel<<<(unsigned)ign_cdiv(m, rows), 256, 0, st>>>(dy, pre_act, m, n, act, db, from_out, (int)rows);
    }
    IGN_CHECK_LAUNCH("act_bwd_bias");
  }
  if (!dx && !dw) return IGN_OK;                             // only the activation / bias step was asked for
  if (n == 1 && (k == 128 || k == 256 || k == 512) && m >= 1024) {
    // single-output head: both gradients in one streaming pass
    int sms = IGN_NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = ign_cdiv(m, 8 * 16);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    if (k == 128) dense_bwd_head_kernel<1><<<(unsigned)grid, 256, 0, st>>>(x, m, w, dy, dx, dw, -1, nullptr);
    else if (k == 256) dense_bwd_head_kernel<2><<<(unsigned)grid, 256, 0, st>>>(x, m, w, dy, dx, dw, -1, nullptr);
    else dense_bwd_head_kernel<4><<<(unsigned)grid, 256, 0, st>>>(x, m, w, dy, dx, dw, -1, nullptr);
    IGN_CHECK_LAUNCH("dense_bwd_head");
    return IGN_OK;
  }
  if (dx && ws && ign_tensor_cores_enabled() && ign_dense_tc_supported(n, k) && ws_bytes >= ign_dense_tc_ws(n, k) &&
      m >= DENSE_TC_MIN_ROWS) {
    // tensor cores (3xTF32): y[m, k] = dZ[m, n] . W^T[n, k]
    int rc = ign_dense_tc_launch(dy, m, n, w, nullptr, k, IGN_ACT_LINEAR, dx, nullptr, ws, st, nullptr, nullptr,
                                 nullptr, true);
    if (rc != IGN_OK) return rc;
  } else if (dx) {   // dX[m,k] = dZ[m,n] W^T : B = W stored [k,n] = [N',K'] with N'=k, K'=n
    dim3 grid((unsigned)ign_cdiv(m, BM), (unsigned)ign_cdiv(k, BN), 1);
    gemm_kernel<false, true, 1><<<grid, GEMM_THREADS, 0, st>>>(dy, w, dx, m, k, n, nullptr, 0, nullptr, 0);
    IGN_CHECK_LAUNCH("dense_bwd_dx");
  }
  if (dw && ign_tensor_cores_enabled() && ign_dw_tc_supported(k, n) && m >= 4096) {
    // tensor cores (3xTF32, MN-major operands, accumulators resident in TMEM): csrc/dw_tc.cu
    int rc = ign_dw_tc_launch(x, dy, m, k, n, dw, st);
    if (rc != IGN_OK) return rc;
  } else if (dw) {   // dW[k,n] += X^T[k,m] dZ[m,n] : A = X stored [m,k] = [K',M'] with M'=k, K'=m
    // a split per 4096 rows -- or, for a small batch whose few output tiles would leave the machine idle while each
    // walks all m rows, enough splits for one CTA per SM (down to 2 BK rows per split)
    int64_t splits = ign_cdiv(m, 4096);
    const int64_t tiles = ign_cdiv(k, BM) * ign_cdiv(n, BN);
    const int64_t fill = ign_cdiv(IGN_NUM_SMS, tiles) < ign_cdiv(m, 2 * BK) ? ign_cdiv(IGN_NUM_SMS, tiles) : ign_cdiv(m, 2 * BK);
    if (splits < fill) splits = fill;
    if (splits > 1024) splits = 1024;
    const int64_t per = ign_cdiv(ign_cdiv(m, splits), BK) * BK;
    splits = ign_cdiv(m, per);
    dim3 grid((unsigned)ign_cdiv(k, BM), (unsigned)ign_cdiv(n, BN), (unsigned)splits);
    gemm_kernel<true, false, 2><<<grid, GEMM_THREADS, 0, st>>>(x, dy, dw, k, n, m, nullptr, 0, nullptr, per);
    IGN_CHECK_LAUNCH("dense_bwd_dw");
  }
  return IGN_OK;
}

extern "C" int ign_dense_head_bwd_chain(const float* x, int64_t m, int k, const float* w, const float* dz, int prev_act,
                                        float* dz_prev, float* dw, float* db_prev, void* stream) {
  IGN_REQUIRE(m >= 0 && (k == 128 || k == 256 || k == 512), IGN_ERR_UNSUPPORTED,
              "IGNNITION: dense_head_bwd_chain: built for 128, 256 or 512 inputs (got %d)", k);
  IGN_REQUIRE(prev_act >= IGN_ACT_LINEAR && prev_act <= IGN_ACT_LEAKY_RELU, IGN_ERR_INVALID,
              "IGNNITION: dense_head_bwd_chain: unknown activation %d", prev_act);
  if (m == 0) return IGN_OK;
  IGN_REQUIRE(x && w && dz && dz_prev && dw, IGN_ERR_INVALID, "IGNNITION: dense_head_bwd_chain: null pointer");
  cudaStream_t st = ign_stream(stream);
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = ign_cdiv(m, 8 * 16);
  if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
  if (grid < 1) grid = 1;
  if (k == 128) dense_bwd_head_kernel<1><<<(unsigned)grid, 256, 0, st>>>(x, m, w, dz, dz_prev, dw, prev_act, db_prev);
  else if (k == 256) dense_bwd_head_kernel<2><<<(unsigned)grid, 256, 0, st>>>(x, m, w, dz, dz_prev, dw, prev_act, db_prev);
  else dense_bwd_head_kernel<4><<<(unsigned)grid, 256, 0, st>>>(x, m, w, dz, dz_prev, dw, prev_act, db_prev);
  IGN_CHECK_LAUNCH("dense_head_bwd_chain");
  return IGN_OK;
}

extern "C" int ign_mse_loss(const float* pred, const float* label, int64_t n, float grad_scale, float* d_pred,
                            double* sse, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: mse_loss: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(pred && label && sse, IGN_ERR_INVALID, "IGNNITION: mse_loss: null pointer");
  mse_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(pred, label, n, grad_scale, d_pred, sse);
  IGN_CHECK_LAUNCH("mse_loss");
  return IGN_OK;
}

extern "C" int ign_loss(int kind, const float* pred, const float* label, int64_t n, float grad_scale, float delta,
                        float* d_pred, double* acc, void* stream) {
  IGN_REQUIRE(kind >= IGN_LOSS_MSE && kind <= IGN_LOSS_BCE, IGN_ERR_INVALID, "IGNNITION: loss: unknown loss %d", kind);
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: loss: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(pred && label && acc, IGN_ERR_INVALID, "IGNNITION: loss: null pointer");
  loss_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(kind, pred, label, n, grad_scale, delta, d_pred, acc);
  IGN_CHECK_LAUNCH("loss");
  return IGN_OK;
}

extern "C" int ign_optimizer_step(int kind, float* w, const float* g, float* s1, float* s2, int64_t n, float lr, float a,
                                  float b, float eps, int flags, void* stream) {
  IGN_REQUIRE(kind >= IGN_OPT_SGD && kind <= IGN_OPT_ADAMAX, IGN_ERR_INVALID, "IGNNITION: optimizer_step: unknown optimiser %d", kind);
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: optimizer_step: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(w && g && s1 && s2, IGN_ERR_INVALID, "IGNNITION: optimizer_step: null pointer");
  optimizer_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(kind, w, g, s1, s2, n, lr, a, b, eps, flags);
  IGN_CHECK_LAUNCH("optimizer_step");
  return IGN_OK;
}

extern "C" int ign_gru_gates_fwd(const float* zx, const float* zh, const float* h, int64_t n, int units, float* out,
                                 void* stream) {
  IGN_REQUIRE(n >= 0 && units > 0, IGN_ERR_INVALID, "IGNNITION: gru_gates_fwd: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(zx && zh && h && out, IGN_ERR_INVALID, "IGNNITION: gru_gates_fwd: null pointer");
  gru_gates_fwd_kernel<<<(unsigned)ign_cdiv(n * units, 256), 256, 0, ign_stream(stream)>>>(zx, zh, h, n, units, out);
  IGN_CHECK_LAUNCH("gru_gates_fwd");
  return IGN_OK;
}

extern "C" int ign_gru_v1_reset(const float* zx, const float* zh2, const float* h, int64_t n, int units, float* rh,
                                void* stream) {
  IGN_REQUIRE(n >= 0 && units > 0, IGN_ERR_INVALID, "IGNNITION: gru_v1_reset: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(zx && zh2 && h && rh, IGN_ERR_INVALID, "IGNNITION: gru_v1_reset: null pointer");
  gru_v1_reset_kernel<<<(unsigned)ign_cdiv(n * units, 256), 256, 0, ign_stream(stream)>>>(zx, zh2, h, n, units, rh);
  IGN_CHECK_LAUNCH("gru_v1_reset");
  return IGN_OK;
}

extern "C" int ign_gru_v1_out(const float* zx, const float* zh2, const float* zhh, const float* h, int64_t n, int units,
                              float* out, void* stream) {
  IGN_REQUIRE(n >= 0 && units > 0, IGN_ERR_INVALID, "IGNNITION: gru_v1_out: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(zx && zh2 && zhh && h && out, IGN_ERR_INVALID, "IGNNITION: gru_v1_out: null pointer");
  gru_v1_out_kernel<<<(unsigned)ign_cdiv(n * units, 256), 256, 0, ign_stream(stream)>>>(zx, zh2, zhh, h, n, units, out);
  IGN_CHECK_LAUNCH("gru_v1_out");
  return IGN_OK;
}

extern "C" int ign_gru_v1_bwd_out(float* zx, float* zh2, float* zhh, const float* h, const float* d_out, int64_t n,
                                  int units, float* dh_direct, void* stream) {
  IGN_REQUIRE(n >= 0 && units > 0, IGN_ERR_INVALID, "IGNNITION: gru_v1_bwd_out: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(zx && zh2 && zhh && h && d_out && dh_direct, IGN_ERR_INVALID, "IGNNITION: gru_v1_bwd_out: null pointer");
  gru_v1_bwd_out_kernel<<<(unsigned)ign_cdiv(n * units, 256), 256, 0, ign_stream(stream)>>>(zx, zh2, zhh, h, d_out, n,
                                                                                           units, dh_direct);
  IGN_CHECK_LAUNCH("gru_v1_bwd_out");
  return IGN_OK;
}

extern "C" int ign_gru_v1_bwd_reset(float* zx, float* zh2, const float* h, const float* d_rh, int64_t n, int units,
                                    float* dh_direct, void* stream) {
  IGN_REQUIRE(n >= 0 && units > 0, IGN_ERR_INVALID, "IGNNITION: gru_v1_bwd_reset: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(zx && zh2 && h && d_rh && dh_direct, IGN_ERR_INVALID, "IGNNITION: gru_v1_bwd_reset: null pointer");
  gru_v1_bwd_reset_kernel<<<(unsigned)ign_cdiv(n * units, 256), 256, 0, ign_stream(stream)>>>(zx, zh2, h, d_rh, n, units,
                                                                                             dh_direct);
  IGN_CHECK_LAUNCH("gru_v1_bwd_reset");
  return IGN_OK;
}

extern "C" int ign_l2_reg(const float* w, int64_t n, float lambda, float* dw, double* reg, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: l2_reg: negative size");
  if (n == 0 || lambda == 0.0f) return IGN_OK;
  IGN_REQUIRE(w, IGN_ERR_INVALID, "IGNNITION: l2_reg: null pointer");
  l2_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(w, n, lambda, dw, reg);
  IGN_CHECK_LAUNCH("l2_reg");
  return IGN_OK;
}

extern "C" int ign_adam_step(float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int64_t step, void* stream) {
  IGN_REQUIRE(n >= 0 && step >= 1, IGN_ERR_INVALID, "IGNNITION: adam_step: bad argument");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(w && g && m && v, IGN_ERR_INVALID, "IGNNITION: adam_step: null pointer");
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  adam_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(w, g, m, v, n, (float)lr_t, beta1, beta2, eps);
  IGN_CHECK_LAUNCH("adam_step");
  return IGN_OK;
}
