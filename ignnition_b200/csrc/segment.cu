// Gather + segmented aggregation over a CSR-by-destination (sm_100a), plus the small layout
// kernels around it (initial hidden state, gathered concat).
//
// ign_segment_reduce replaces tf.gather -> tf.scatter_nd into the padded [num_dst,max_len,F]
// tensor -> tf.reduce_sum(axis=1)  (reference code/utils/generate_model.py:432,479-490 and
// code/utils/auxilary_classes.py:254-262).  One sub-warp of G = F/4 lanes owns one destination,
// walks its slots in seq order with a single accumulator (so the sum order is the reference's
// column order: deterministic, no atomics), stages G column indices per coalesced load and
// broadcasts them by shuffle, and keeps UNROLL independent 16-byte row loads in flight per lane.
// HBM-bound: algorithmic bytes = E*(4 + 4F) + num_dst*(4F + 4)  (DESIGN.md).

#include "common.cuh"

namespace {

constexpr int SEG_THREADS = 256;
constexpr int SEG_UNROLL = 4;

template <int OP>
__device__ __forceinline__ void seg_acc(float4& a, const float4& v) {
  if (OP == IGN_OP_MAX) {
    a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
  } else {
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
}

// G lanes per destination (power of two <= 32), V float4 columns per lane: F <= 4*G*V
template <int G, int V, int OP>
__global__ void __launch_bounds__(SEG_THREADS) segment_reduce_kernel(const int* __restrict__ rowptr,
                                                                     const int* __restrict__ col,
                                                                     const float* __restrict__ src, int F,
                                                                     int64_t num_dst, float* __restrict__ out) {
  const int64_t gid = ((int64_t)blockIdx.x * SEG_THREADS + threadIdx.x) / G;
  if (gid >= num_dst) return;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);                                   // lane inside the group
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const int lo = rowptr[gid], hi = rowptr[gid + 1];
  const float init = (OP == IGN_OP_MAX) ? -INFINITY : 0.0f;
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(init, init, init, init);
  bool colok[V];
#pragma unroll
  for (int v = 0; v < V; ++v) colok[v] = (gl + v * G) * 4 < F;

  for (int e = lo; e < hi; e += G) {
    const int mine = (e + gl < hi) ? (col ? col[e + gl] : e + gl) : -1;
    const int cnt = min(G, hi - e);
    for (int j = 0; j < cnt; j += SEG_UNROLL) {
      int c[SEG_UNROLL];
#pragma unroll
      for (int u = 0; u < SEG_UNROLL; ++u) c[u] = __shfl_sync(gmask, mine, (j + u) & (G - 1), G);
      float4 val[SEG_UNROLL][V];
#pragma unroll
      for (int u = 0; u < SEG_UNROLL; ++u) {
        const bool ok = (j + u < cnt);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (ok && colok[v])     // a negative column is a zero row (slot no edge claimed, csr_build.cu)
            val[u][v] = c[u] >= 0 ? ldg_f4(src + (int64_t)c[u] * F + (gl + v * G) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < SEG_UNROLL; ++u) {
        if (j + u < cnt) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (colok[v]) seg_acc<OP>(acc[v], val[u][v]);
        }
      }
    }
  }
  const int len = hi - lo;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    if (!colok[v]) continue;
    float4 r = acc[v];
    if (OP == IGN_OP_MEAN) {
      const float inv = 1.0f / (float)max(len, 1);
      r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
    }
    if (OP == IGN_OP_MAX && len == 0) r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (OP == IGN_OP_SUM_ADD) {
      const float4 o = *reinterpret_cast<const float4*>(out + gid * F + (gl + v * G) * 4);
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    st_f4(out + gid * F + (gl + v * G) * 4, r);
  }
}

template <int G, int V>
int launch_segment(int op, const int* rowptr, const int* col, const float* src, int F, int64_t num_dst,
                   float* out, cudaStream_t st) {
  const int64_t threads = num_dst * G;
  const unsigned grid = (unsigned)ign_cdiv(threads, SEG_THREADS);
  switch (op) {
    case IGN_OP_SUM:
      segment_reduce_kernel<G, V, IGN_OP_SUM><<<grid, SEG_THREADS, 0, st>>>(rowptr, col, src, F, num_dst, out);
      break;
    case IGN_OP_MEAN:
      segment_reduce_kernel<G, V, IGN_OP_MEAN><<<grid, SEG_THREADS, 0, st>>>(rowptr, col, src, F, num_dst, out);
      break;
    case IGN_OP_SUM_ADD:
      segment_reduce_kernel<G, V, IGN_OP_SUM_ADD><<<grid, SEG_THREADS, 0, st>>>(rowptr, col, src, F, num_dst, out);
      break;
    default:
      segment_reduce_kernel<G, V, IGN_OP_MAX><<<grid, SEG_THREADS, 0, st>>>(rowptr, col, src, F, num_dst, out);
      break;
  }
  IGN_CHECK_LAUNCH("segment_reduce");
  return IGN_OK;
}

constexpr int MAX_FEATS = 8;
struct FeatList {
  const float* ptr[MAX_FEATS];
  int size[MAX_FEATS];
  int n;
};

__device__ __forceinline__ float init_state_value(const FeatList& f, int64_t row, int c) {
  for (int k = 0; k < f.n; ++k) {
    if (c < f.size[k]) return f.ptr[k][row * f.size[k] + c];
    c -= f.size[k];
  }
  return 0.0f;
}
// one thread per row: the feature columns (a handful) then zeros, all as 16-byte stores; rows are
// hidden * 4 bytes apart, so a warp writes 32 consecutive rows = one contiguous span (hidden % 4 == 0)
__global__ void init_state_kernel(FeatList f, int64_t n, int hidden, int total_feat, float* __restrict__ state) {
  const int q = hidden / 4;                                    // float4 columns per row
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // lane l of a warp takes float4 column (l % q), (l % q) + 32 ... of row (warp rows) -- simple and
  // coalesced: 32 lanes cover 32 consecutive float4 of the flat [n * q] array, the row index costs one
  // 32-bit division per lane instead of a 64-bit one per float4
  const int64_t total4 = n * q;
  for (int64_t i = gid; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = (i >> 31) == 0 ? (int64_t)((uint32_t)i / (uint32_t)q) : i / q;
    const int c = (int)(i - row * q) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < total_feat) {
      v.x = init_state_value(f, row, c);
      v.y = init_state_value(f, row, c + 1);
      v.z = init_state_value(f, row, c + 2);
      v.w = init_state_value(f, row, c + 3);
    }
    st_f4(state + row * hidden + c, v);
  }
}
__global__ void init_state_scalar_kernel(FeatList f, int64_t n, int hidden, float* __restrict__ state) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * hidden) return;
  const int64_t row = i / hidden;
  state[i] = init_state_value(f, row, (int)(i - row * hidden));
}

struct ConcatParts {
  const float* ptr[4];
  const int* idx[4];
  int width[4];
  int n;
  int total;
};

__global__ void gather_concat_kernel(ConcatParts p, int64_t rows, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * p.total) return;
  const int64_t row = i / p.total;
  int c = (int)(i - row * p.total);
  float v = 0.0f;
  for (int k = 0; k < p.n; ++k) {
    if (c < p.width[k]) {
      const int64_t r = p.idx[k] ? (int64_t)p.idx[k][row] : row;
      if (r >= 0) v = p.ptr[k][r * p.width[k] + c];        // negative index: a zero (padded) row
      break;
    }
    c -= p.width[k];
  }
  out[i] = v;
}

__global__ void axpy_kernel(int64_t n, float a, const float* __restrict__ x, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}

// Concat_aggr along the feature axis (generate_model.py:496-505, concat_axis = 2): position j of the
// first source's CSR (destination d, column s = j - rowptr[d]) pairs with column s of another source's
// padded block: that source's row index, or -1 where its block holds zeros.
__global__ void partner_index_kernel(const int* __restrict__ rowptr0, const int* __restrict__ rowptr1,
                                     const int* __restrict__ idx1, int64_t num_dst, int* __restrict__ out) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= num_dst) return;
  const int lo0 = rowptr0[d], n0 = rowptr0[d + 1] - lo0;
  const int lo1 = rowptr1[d], n1 = rowptr1[d + 1] - lo1;
  for (int s = 0; s < n0; ++s) out[lo0 + s] = s < n1 ? idx1[lo1 + s] : -1;
}

__global__ void mul_kernel(int64_t n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

// Conv_aggr tail (auxilary_classes.py:388-401): act((neighbour_sum + self) / degree)
__global__ void conv_finish_kernel(const float* __restrict__ nsum, const float* __restrict__ self,
                                   const int* __restrict__ rowptr, int F, int64_t n, int act, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * F) return;
  const int64_t d = i / F;
  const float deg = (float)(rowptr[d + 1] - rowptr[d]);
  out[i] = act_fwd(act, (nsum[i] + self[i]) / deg);         // degree 0 -> inf / nan, as in the reference
}

}  // namespace

extern "C" int ign_partner_index(const int32_t* rowptr0, const int32_t* rowptr1, const int32_t* idx1, int64_t num_dst,
                                 int32_t* out, void* stream) {
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: partner_index: negative size");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr0 && rowptr1, IGN_ERR_INVALID, "IGNNITION: partner_index: null pointer");
  partner_index_kernel<<<(unsigned)ign_cdiv(num_dst, 128), 128, 0, ign_stream(stream)>>>(rowptr0, rowptr1, idx1, num_dst, out);
  IGN_CHECK_LAUNCH("partner_index");
  return IGN_OK;
}

extern "C" int ign_mul(int64_t n, const float* a, const float* b, float* out, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: mul: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(a && b && out, IGN_ERR_INVALID, "IGNNITION: mul: null pointer");
  mul_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(n, a, b, out);
  IGN_CHECK_LAUNCH("mul");
  return IGN_OK;
}

extern "C" int ign_conv_finish(const float* nsum, const float* self, const int32_t* rowptr, int F, int64_t n, int act,
                               float* out, void* stream) {
  IGN_REQUIRE(n >= 0 && F > 0, IGN_ERR_INVALID, "IGNNITION: conv_finish: bad size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(nsum && self && rowptr && out, IGN_ERR_INVALID, "IGNNITION: conv_finish: null pointer");
  conv_finish_kernel<<<(unsigned)ign_cdiv(n * F, 256), 256, 0, ign_stream(stream)>>>(nsum, self, rowptr, F, n, act, out);
  IGN_CHECK_LAUNCH("conv_finish");
  return IGN_OK;
}

extern "C" int ign_segment_reduce(int op, const int32_t* rowptr, const int32_t* col, const float* src_states,
                                  int F, int64_t num_dst, float* out, void* stream) {
  IGN_REQUIRE(op == IGN_OP_SUM || op == IGN_OP_MEAN || op == IGN_OP_MAX || op == IGN_OP_SUM_ADD, IGN_ERR_INVALID,
              "IGNNITION: segment_reduce: unknown aggregation %d", op);
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: segment_reduce: negative size");
  IGN_REQUIRE(F > 0 && F % 4 == 0 && F <= 256, IGN_ERR_UNSUPPORTED,
              "IGNNITION: segment_reduce: message width %d unsupported (multiple of 4, <= 256)", F);
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && src_states && out, IGN_ERR_INVALID, "IGNNITION: segment_reduce: null pointer");
  cudaStream_t st = ign_stream(stream);
  const int q = F / 4;
  if (q <= 1) return launch_segment<1, 1>(op, rowptr, col, src_states, F, num_dst, out, st);
  if (q <= 2) return launch_segment<2, 1>(op, rowptr, col, src_states, F, num_dst, out, st);
  if (q <= 4) return launch_segment<4, 1>(op, rowptr, col, src_states, F, num_dst, out, st);
  if (q <= 8) return launch_segment<8, 1>(op, rowptr, col, src_states, F, num_dst, out, st);
  if (q <= 16) return launch_segment<16, 1>(op, rowptr, col, src_states, F, num_dst, out, st);
  if (q <= 32) return launch_segment<32, 1>(op, rowptr, col, src_states, F, num_dst, out, st);
  return launch_segment<32, 2>(op, rowptr, col, src_states, F, num_dst, out, st);
}

extern "C" int ign_init_state(int n_feat, const float* const* feats, const int32_t* feat_size, int64_t n,
                              int hidden, float* state, void* stream) {
  IGN_REQUIRE(n_feat >= 0 && n_feat <= MAX_FEATS, IGN_ERR_UNSUPPORTED,
              "IGNNITION: init_state: at most %d features per entity", MAX_FEATS);
  IGN_REQUIRE(n >= 0 && hidden > 0, IGN_ERR_INVALID, "IGNNITION: init_state: bad size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(state && (n_feat == 0 || (feats && feat_size)), IGN_ERR_INVALID, "IGNNITION: init_state: null pointer");
  FeatList f;
  f.n = n_feat;
  int total = 0;
  for (int k = 0; k < MAX_FEATS; ++k) {
    f.ptr[k] = k < n_feat ? feats[k] : nullptr;
    f.size[k] = k < n_feat ? feat_size[k] : 0;
    IGN_REQUIRE(k >= n_feat || (f.ptr[k] && f.size[k] > 0), IGN_ERR_INVALID, "IGNNITION: init_state: bad feature");
    total += f.size[k];
  }
  IGN_REQUIRE(total <= hidden, IGN_ERR_INVALID,
              "IGNNITION: init_state: features (%d) wider than the hidden state (%d)", total, hidden);
  if (hidden % 4 == 0)
  {
    const int64_t blocks = ign_cdiv(n * (hidden / 4), 256 * 4);       // 4 float4 per thread
    init_state_kernel<<<(unsigned)(blocks > 0 ? blocks : 1), 256, 0, ign_stream(stream)>>>(f, n, hidden, total, state);
  }
  else
    init_state_scalar_kernel<<<(unsigned)ign_cdiv(n * hidden, 256), 256, 0, ign_stream(stream)>>>(f, n, hidden, state);
  IGN_CHECK_LAUNCH("init_state");
  return IGN_OK;
}

extern "C" int ign_gather_concat(int n_parts, const float* const* parts, const int32_t* const* idx,
                                 const int32_t* widths, int64_t rows, float* out, void* stream) {
  IGN_REQUIRE(n_parts >= 1 && n_parts <= 4, IGN_ERR_UNSUPPORTED, "IGNNITION: gather_concat: 1..4 parts");
  IGN_REQUIRE(rows >= 0 && parts && widths && out, IGN_ERR_INVALID, "IGNNITION: gather_concat: bad argument");
  if (rows == 0) return IGN_OK;
  ConcatParts p;
  p.n = n_parts;
  p.total = 0;
  for (int k = 0; k < 4; ++k) {
    p.ptr[k] = k < n_parts ? parts[k] : nullptr;
    p.idx[k] = (k < n_parts && idx) ? idx[k] : nullptr;
    p.width[k] = k < n_parts ? widths[k] : 0;
    IGN_REQUIRE(k >= n_parts || (p.ptr[k] && p.width[k] > 0), IGN_ERR_INVALID, "IGNNITION: gather_concat: bad part");
    p.total += p.width[k];
  }
  gather_concat_kernel<<<(unsigned)ign_cdiv(rows * p.total, 256), 256, 0, ign_stream(stream)>>>(p, rows, out);
  IGN_CHECK_LAUNCH("gather_concat");
  return IGN_OK;
}

// out[r, c] = x[r, col0 + c]: the column block of one concatenated input (backward of the concat in ign_gather_concat)
__global__ void slice_cols_kernel(const float* __restrict__ x, int64_t rows, int ld, int col0, int w, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * w) return;
  const int64_t r = i / w;
  const int c = (int)(i % w);
  out[i] = x[r * ld + col0 + c];
}

extern "C" int ign_slice_cols(const float* x, int64_t rows, int ld, int col0, int width, float* out, void* stream) {
  IGN_REQUIRE(rows >= 0 && ld > 0 && col0 >= 0 && width > 0 && col0 + width <= ld, IGN_ERR_INVALID,
              "IGNNITION: slice_cols: bad shape");
  if (rows == 0) return IGN_OK;
  IGN_REQUIRE(x && out, IGN_ERR_INVALID, "IGNNITION: slice_cols: null pointer");
  slice_cols_kernel<<<(unsigned)ign_cdiv(rows * width, 256), 256, 0, ign_stream(stream)>>>(x, rows, ld, col0, width, out);
  IGN_CHECK_LAUNCH("slice_cols");
  return IGN_OK;
}

extern "C" int ign_axpy(int64_t n, float a, const float* x, float* y, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: axpy: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(x && y, IGN_ERR_INVALID, "IGNNITION: axpy: null pointer");
  axpy_kernel<<<(unsigned)ign_cdiv(n, 256), 256, 0, ign_stream(stream)>>>(n, a, x, y);
  IGN_CHECK_LAUNCH("axpy");
  return IGN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the mean / max aggregations (north_star extensions; TensorFlow semantics of reduce_mean / of
// unsorted_segment_max: the gradient of a maximum is split evenly among the slots that attain it).
namespace {

// d[r, :] *= 1 / max(deg(r), 1): turns dL/d(mean) into dL/d(sum)
__global__ void scale_rows_inv_degree_kernel(float* __restrict__ d, const int* __restrict__ rowptr, int64_t n, int width) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * (width / 4)) return;
  const int64_t r = i / (width / 4);
  const int deg = rowptr[r + 1] - rowptr[r];
  if (deg <= 1) return;
  const float inv = 1.0f / (float)deg;
  float4 v = reinterpret_cast<float4*>(d)[i];
  v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
  reinterpret_cast<float4*>(d)[i] = v;
}

// d_msg[pos(slot), :] = (rows[idx[slot], f] == agg[d, f]) ? d_agg[d, f] / #ties : 0 for every slot of destination d;
// pos(slot) = perm[slot] (input edge position) or the slot itself.  G = width / 4 lanes per destination.
__global__ void segment_max_bwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ idx,
                                       const int* __restrict__ perm, const float* __restrict__ rows,
                                       const float* __restrict__ agg, const float* __restrict__ d_agg, int64_t n,
                                       int width, float* __restrict__ d_msg) {
  const int q = width / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * q) return;
  const int64_t d = i / q;
  const int c = (int)(i - d * q) * 4;
  const int lo = rowptr[d], hi = rowptr[d + 1];
  if (hi <= lo) return;
  const float4 m = *reinterpret_cast<const float4*>(agg + d * width + c);
  const float4 g = *reinterpret_cast<const float4*>(d_agg + d * width + c);
  float4 cnt = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = lo; s < hi; ++s) {
    const int r = idx[s];
    const float4 v = r >= 0 ? *reinterpret_cast<const float4*>(rows + (int64_t)r * width + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    cnt.x += v.x == m.x; cnt.y += v.y == m.y; cnt.z += v.z == m.z; cnt.w += v.w == m.w;
  }
  const float4 w = make_float4(g.x / fmaxf(cnt.x, 1.f), g.y / fmaxf(cnt.y, 1.f), g.z / fmaxf(cnt.z, 1.f), g.w / fmaxf(cnt.w, 1.f));
  for (int s = lo; s < hi; ++s) {
    const int r = idx[s];
    const float4 v = r >= 0 ? *reinterpret_cast<const float4*>(rows + (int64_t)r * width + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t pos = perm ? perm[s] : s;
    *reinterpret_cast<float4*>(d_msg + pos * width + c) =
        make_float4(v.x == m.x ? w.x : 0.f, v.y == m.y ? w.y : 0.f, v.z == m.z ? w.z : 0.f, v.w == m.w ? w.w : 0.f);
  }
}

// out[r, :] = rows[s, :] for every row r of segment s (rowptr[s] <= r < rowptr[s + 1]): backward of a per-sample sum
__global__ void segment_broadcast_kernel(const int* __restrict__ rowptr, const float* __restrict__ rows, int64_t n_seg,
                                         int width, float* __restrict__ out) {
  const int q = width / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg * q) return;
  const int64_t s = i / q;
  const int c = (int)(i - s * q) * 4;
  const float4 v = *reinterpret_cast<const float4*>(rows + s * width + c);
  for (int64_t r = rowptr[s]; r < rowptr[s + 1]; ++r) *reinterpret_cast<float4*>(out + r * width + c) = v;
}

}  // namespace

extern "C" int ign_segment_broadcast(const int32_t* rowptr, const float* rows, int64_t n_seg, int width, float* out,
                                     void* stream) {
  IGN_REQUIRE(n_seg >= 0 && width > 0 && width % 4 == 0, IGN_ERR_INVALID, "IGNNITION: segment_broadcast: bad shape");
  if (n_seg == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && rows && out, IGN_ERR_INVALID, "IGNNITION: segment_broadcast: null pointer");
  segment_broadcast_kernel<<<(unsigned)ign_cdiv(n_seg * (width / 4), 256), 256, 0, ign_stream(stream)>>>(rowptr, rows, n_seg,
                                                                                                      width, out);
  IGN_CHECK_LAUNCH("segment_broadcast");
  return IGN_OK;
}

extern "C" int ign_scale_rows_inv_degree(float* d, const int32_t* rowptr, int64_t n, int width, void* stream) {
  IGN_REQUIRE(n >= 0 && width > 0 && width % 4 == 0, IGN_ERR_INVALID, "IGNNITION: scale_rows_inv_degree: bad shape");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(d && rowptr, IGN_ERR_INVALID, "IGNNITION: scale_rows_inv_degree: null pointer");
  scale_rows_inv_degree_kernel<<<(unsigned)ign_cdiv(n * (width / 4), 256), 256, 0, ign_stream(stream)>>>(d, rowptr, n, width);
  IGN_CHECK_LAUNCH("scale_rows_inv_degree");
  return IGN_OK;
}

extern "C" int ign_segment_max_bwd(const int32_t* rowptr, const int32_t* idx, const int32_t* perm, const float* rows,
                                   const float* agg, const float* d_agg, int64_t num_dst, int width, float* d_msg,
                                   void* stream) {
  IGN_REQUIRE(num_dst >= 0 && width > 0 && width % 4 == 0, IGN_ERR_INVALID, "IGNNITION: segment_max_bwd: bad shape");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && idx && rows && agg && d_agg && d_msg, IGN_ERR_INVALID, "IGNNITION: segment_max_bwd: null pointer");
  segment_max_bwd_kernel<<<(unsigned)ign_cdiv(num_dst * (width / 4), 256), 256, 0, ign_stream(stream)>>>(
      rowptr, idx, perm, rows, agg, d_agg, num_dst, width, d_msg);
  IGN_CHECK_LAUNCH("segment_max_bwd");
  return IGN_OK;
}
