// Attention aggregation of the generated model (reference: Attention_aggr.calculate_input,
// auxilary_classes.py:278-344), on the CSR by destination.
//
// What the reference computes, restated on the CSR (edge j of destination d sits at column
// s = j - rowptr[d] of the padded [num_dst, max_len, 1] tensor):
//   a_j       = leaky_relu([m_j . kernel1 | h_d . kernel2] . attn_kernel, 0.2)
//             = leaky_relu(m_j . (kernel1 . ak_top) + h_d . (kernel2 . ak_bottom))      (associativity)
//   aux[d, s] = a_j, zero where destination d has no message at column s
//   coef      = softmax(aux, axis = 0): over the DESTINATIONS of one sample, per column s, zero pads
//               included (SURVEY section 8f rank 3 "reproduce as-is")
//   out[d]    = sum_j coef[d, s_j] * m_j
// The two matrix-vector folds and the per-row scores are ign_dense calls on the host side; this file
// does the column softmax and the weighted segment sum.  Column statistics: max by integer atomicMax
// on an order-preserving encoding, sum by fp64 atomicAdd (order-independent after rounding to fp32).
//
// Several sources (generate_model.py:525-543): the edges of all sources form one list, the column of an edge of
// source k > 0 is seq + (number of edges of source k into its destination) -- SURVEY quirk 7 -- so two edges of
// one destination can land on the same column; scatter_nd then ADDS their activated scores, and both edges read
// the one coefficient of that cell.  Here: one CSR over the combined list, an explicit column per slot
// (`slot_col`, ign_attention_combine), and per slot the first slot of its row with the same column (`rep`), which
// carries the cell's summed score.  Rows are short; the collisions are found by scanning the row.
#include "common.cuh"

namespace {

__device__ __forceinline__ int enc_f(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float dec_f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct AttnWs {
  int* colmax;      // [n_cols] encoded float
  int* cnt;         // [n_cols] destinations with a message at this column
  double* colsum;   // [n_cols]
  float* a;         // [E] activated score of every slot
  float* x;         // [E] score of the padded cell, valid at the cell's first slot (== a without slot columns)
  int* rep;         // [E] first slot of the row on the same column (only with slot columns)
};

__device__ __forceinline__ int cell_col(const int* __restrict__ slot_col, int j, int lo) {
  return slot_col ? slot_col[j] : j - lo;
}

__device__ __forceinline__ int sample_of(const int* __restrict__ off, int n_samples, int d) {
  int lo = 0, hi = n_samples;                 // off[lo] <= d < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= d) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void attn_init_kernel(int64_t n_cols, AttnWs w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cols) return;
  w.colmax[i] = enc_f(-INFINITY);
  w.cnt[i] = 0;
  w.colsum[i] = 0.0;
}

__global__ void attn_score_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                  const int* __restrict__ slot_col, const float* __restrict__ src_score,
                                  const float* __restrict__ dst_score, const int* __restrict__ off, int n_samples,
                                  int64_t num_dst, int max_len, AttnWs w) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= num_dst) return;
  const int lo = rowptr[d], hi = rowptr[d + 1];
  if (hi == lo) return;
  const int64_t base = (int64_t)sample_of(off, n_samples, (int)d) * max_len;
  const float q = dst_score[d];
  for (int j = lo; j < hi; ++j) {
    float a = src_score[col[j]] + q;
    a = a > 0.f ? a : 0.2f * a;               // tf.keras.layers.LeakyReLU(alpha=0.2), auxilary_classes.py:319
    w.a[j] = a;
    if (!slot_col) {
      atomicMax(&w.colmax[base + (j - lo)], enc_f(a));
      atomicAdd(&w.cnt[base + (j - lo)], 1);
    }
  }
  if (!slot_col) return;
  for (int j = lo; j < hi; ++j) {             // cells: scatter_nd adds the scores that share (destination, column)
    const int c = slot_col[j];
    int first = j;
    float sum = 0.f;
    for (int j2 = hi - 1; j2 >= lo; --j2)
      if (slot_col[j2] == c) { first = j2; }
    w.rep[j] = first;
    if (first != j) continue;
    for (int j2 = lo; j2 < hi; ++j2)
      if (slot_col[j2] == c) sum += w.a[j2];
    w.x[j] = sum;
    atomicMax(&w.colmax[base + c], enc_f(sum));
    atomicAdd(&w.cnt[base + c], 1);
  }
}

__device__ __forceinline__ float col_shift(const AttnWs& w, int64_t c, int n_in_sample) {
  float m = dec_f(w.colmax[c]);
  if (w.cnt[c] < n_in_sample) m = fmaxf(m, 0.f);       // zero pads take part in the softmax
  return m;
}

__global__ void attn_sum_kernel(const int* __restrict__ rowptr, const int* __restrict__ slot_col,
                                const int* __restrict__ off, int n_samples, int64_t num_dst, int max_len, AttnWs w) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= num_dst) return;
  const int lo = rowptr[d], hi = rowptr[d + 1];
  if (hi == lo) return;
  const int s = sample_of(off, n_samples, (int)d);
  const int n_in = off[s + 1] - off[s];
  const int64_t base = (int64_t)s * max_len;
  for (int j = lo; j < hi; ++j) {
    if (slot_col && w.rep[j] != j) continue;              // one term per padded cell
    const int64_t c = base + cell_col(slot_col, j, lo);
    atomicAdd(&w.colsum[c], (double)expf(w.x[j] - col_shift(w, c, n_in)));
  }
}

// 8 lanes per destination, float4 columns strided by 8 lanes
__global__ void attn_apply_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                  const int* __restrict__ slot_col, const float* __restrict__ rows, int F,
                                  const int* __restrict__ off, int n_samples,
                                  int64_t num_dst, int max_len, AttnWs w, float* __restrict__ out) {
  const int64_t d = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  if (d >= num_dst) return;
  const int lo = rowptr[d], hi = rowptr[d + 1];
  const int s = sample_of(off, n_samples, (int)d);
  const int n_in = off[s + 1] - off[s];
  const int64_t base = (int64_t)s * max_len;
  for (int f = gl * 4; f < F; f += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = lo; j < hi; ++j) {
      const int64_t c = base + cell_col(slot_col, j, lo);
      const float m = col_shift(w, c, n_in);
      const double den = w.colsum[c] + (double)(n_in - w.cnt[c]) * (double)expf(-m);
      const float coef = (float)((double)expf(w.x[slot_col ? w.rep[j] : j] - m) / den);
      const float4 v = *reinterpret_cast<const float4*>(rows + (int64_t)col[j] * F + f);
      acc.x += coef * v.x; acc.y += coef * v.y; acc.z += coef * v.z; acc.w += coef * v.w;
    }
    *reinterpret_cast<float4*>(out + d * F + f) = acc;
  }
}

// ---- backward (tf.gradients through Attention_aggr).  With coef_j = exp(a_j - m_c) / den_c per slot j of column c:
//   d_coef_j = g_out[d] . rows[idx_j]                    S_c = sum_{j in c} coef_j d_coef_j   (zero pads: d_coef = 0)
//   d_a_j    = coef_j (d_coef_j - S_c)                   d_pre_j = d_a_j (a_j > 0 ? 1 : 0.2)  (LeakyReLU 0.2)
//   d_msg[pos_j] = coef_j g_out[d]   (the part of dL/d rows that comes through the weighted sum, per edge)
//   d_pre4[pos_j] = (d_pre_j, 0, 0, 0),   d_ds[d] = sum_j d_pre_j
// pos_j = perm[j] (input edge position) or j.  The caller reduces d_msg / d_pre4 per row of `rows` and pushes d_pre
// through the two score products (ign_dense_bwd).  Pass 1: 8 lanes per destination, d_coef into dcoef[], S_c by fp64 atomics.
__device__ __forceinline__ float attn_coef(const AttnWs& w, int64_t c, int n_in, float a) {
  const float m = col_shift(w, c, n_in);
  const double den = w.colsum[c] + (double)(n_in - w.cnt[c]) * (double)expf(-m);
  return (float)((double)expf(a - m) / den);
}

__global__ void attn_bwd_dot_kernel(const int* __restrict__ rowptr, const int* __restrict__ idx,
                                    const int* __restrict__ slot_col, const float* __restrict__ rows, int F,
                                    const float* __restrict__ g_out,
                                    const int* __restrict__ off, int n_samples, int64_t num_dst, int max_len, AttnWs w,
                                    float* __restrict__ dcoef, double* __restrict__ colS) {
  const int64_t d = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  if (d >= num_dst) return;                               // whole 8-lane groups leave together
  const int lo = rowptr[d], hi = rowptr[d + 1];
  if (hi == lo) return;
  const int s = sample_of(off, n_samples, (int)d);
  const int n_in = off[s + 1] - off[s];
  const int64_t base = (int64_t)s * max_len;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  for (int j = lo; j < hi; ++j) {
    float dot = 0.f;
    for (int f = gl * 4; f < F; f += 32) {
      const float4 v = *reinterpret_cast<const float4*>(rows + (int64_t)idx[j] * F + f);
      const float4 g = *reinterpret_cast<const float4*>(g_out + d * F + f);
      dot += v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w;
    }
    dot += __shfl_xor_sync(gmask, dot, 1);
    dot += __shfl_xor_sync(gmask, dot, 2);
    dot += __shfl_xor_sync(gmask, dot, 4);
    if (gl == 0) {
      const int64_t c = base + cell_col(slot_col, j, lo);
      dcoef[j] = dot;           // the slots of one cell share its coefficient: their terms add up to coef * d_coef(cell)
      atomicAdd(&colS[c], (double)attn_coef(w, c, n_in, w.x[slot_col ? w.rep[j] : j]) * (double)dot);
    }
  }
}

__global__ void attn_bwd_apply_kernel(const int* __restrict__ rowptr, const int* __restrict__ perm,
                                      const int* __restrict__ slot_col, int F,
                                      const float* __restrict__ g_out, const int* __restrict__ off, int n_samples,
                                      int64_t num_dst, int max_len, AttnWs w, const float* __restrict__ dcoef,
                                      const double* __restrict__ colS, float* __restrict__ d_msg,
                                      float* __restrict__ d_pre4, float* __restrict__ d_ds) {
  const int64_t d = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  if (d >= num_dst) return;
  const int lo = rowptr[d], hi = rowptr[d + 1];
  float sum_pre = 0.f;
  if (hi > lo) {
    const int s = sample_of(off, n_samples, (int)d);
    const int n_in = off[s + 1] - off[s];
    const int64_t base = (int64_t)s * max_len;
    for (int j = lo; j < hi; ++j) {
      const int64_t c = base + cell_col(slot_col, j, lo);
      const float a = w.a[j];
      float dc = dcoef[j];
      float x = a;
      if (slot_col) {                                      // d_coef of the cell: over the slots that share it
        const int r = w.rep[j];
        x = w.x[r];
        dc = 0.f;
        for (int j2 = lo; j2 < hi; ++j2)
          if (w.rep[j2] == r) dc += dcoef[j2];
      }
      const float coef = attn_coef(w, c, n_in, x);
      const float d_pre = coef * (dc - (float)colS[c]) * (a > 0.f ? 1.f : 0.2f);
      sum_pre += d_pre;
      const int64_t pos = perm ? perm[j] : j;
      for (int f = gl * 4; f < F; f += 32) {
        const float4 g = *reinterpret_cast<const float4*>(g_out + d * F + f);
        *reinterpret_cast<float4*>(d_msg + pos * F + f) = make_float4(coef * g.x, coef * g.y, coef * g.z, coef * g.w);
      }
      if (gl == 0) *reinterpret_cast<float4*>(d_pre4 + pos * 4) = make_float4(d_pre, 0.f, 0.f, 0.f);
    }
  }
  if (gl == 0) d_ds[d] = sum_pre;
}

AttnWs carve(void* ws, int64_t n_edges, int64_t n_cols, bool with_cols) {
  AttnWs w;
  char* p = static_cast<char*>(ws);
  w.colsum = reinterpret_cast<double*>(p); p += ign_align_up(n_cols * sizeof(double), 256);
  w.colmax = reinterpret_cast<int*>(p);    p += ign_align_up(n_cols * sizeof(int), 256);
  w.cnt = reinterpret_cast<int*>(p);       p += ign_align_up(n_cols * sizeof(int), 256);
  const size_t eb = ign_align_up((n_edges > 0 ? n_edges : 0) * sizeof(float), 256);
  w.a = reinterpret_cast<float*>(p);       p += eb;
  w.x = with_cols ? reinterpret_cast<float*>(p) : w.a;   p += eb;
  w.rep = reinterpret_cast<int*>(p);
  return w;
}

struct ColSources {
  const int* rowptr[8];
  const int* perm[8];
  int64_t edge_off[9];
  int n;
};

// One CSR over the edge lists of several sources (generate_model.py:523-543).  Row d = the rows of the sources one
// after the other; cperm = position of the slot's edge in the concatenation of the sources' edge lists; column = the
// edge's seq (its slot in the source's own row) for the first source, seq + the destination's edge count IN THAT
// SOURCE for the others (the reference gathers `lens` of the current source, not the running total).
__global__ void attn_combine_kernel(ColSources cs, int64_t num_dst, int max_len, int* __restrict__ crowptr,
                                    int* __restrict__ cperm, int* __restrict__ slot_col) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d > num_dst) return;
  int start = 0;
  for (int k = 0; k < cs.n; ++k) start += cs.rowptr[k][d];
  crowptr[d] = start;
  if (d == num_dst) return;
  for (int k = 0; k < cs.n; ++k) {
    const int lo = cs.rowptr[k][d], hi = cs.rowptr[k][d + 1];
    for (int j = lo; j < hi; ++j) {
      const int c = (j - lo) + (k > 0 ? hi - lo : 0);
      cperm[start] = (int)(cs.edge_off[k] + (cs.perm[k] ? cs.perm[k][j] : j));
      slot_col[start] = min(c, max_len - 1);
      ++start;
    }
  }
}

}  // namespace

extern "C" size_t ign_attention_ws_bytes(int64_t n_edges, int64_t n_samples, int max_len) {
  const int64_t n_cols = (n_samples > 0 ? n_samples : 0) * (int64_t)(max_len > 0 ? max_len : 0);
  return ign_align_up(n_cols * sizeof(double), 256) + 2 * ign_align_up(n_cols * sizeof(int), 256) +
         3 * ign_align_up((n_edges > 0 ? n_edges : 0) * sizeof(float), 256) + 256;
}

extern "C" int ign_attention_combine(int n_sources, const int32_t* const* src_rowptr, const int32_t* const* src_perm,
                                     const int64_t* edge_counts, int64_t num_dst, int max_len, int32_t* rowptr,
                                     int32_t* perm, int32_t* slot_col, void* stream) {
  IGN_REQUIRE(n_sources >= 1 && n_sources <= 8, IGN_ERR_UNSUPPORTED, "IGNNITION: attention_combine: 1..8 sources");
  IGN_REQUIRE(num_dst >= 0 && max_len >= 1, IGN_ERR_INVALID, "IGNNITION: attention_combine: bad size");
  IGN_REQUIRE(src_rowptr && src_perm && edge_counts && rowptr, IGN_ERR_INVALID, "IGNNITION: attention_combine: null pointer");
  ColSources cs;
  cs.n = n_sources;
  cs.edge_off[0] = 0;
  for (int k = 0; k < n_sources; ++k) {
    IGN_REQUIRE(edge_counts[k] >= 0 && src_rowptr[k], IGN_ERR_INVALID, "IGNNITION: attention_combine: null source array");
    cs.rowptr[k] = src_rowptr[k];
    cs.perm[k] = src_perm[k];
    cs.edge_off[k + 1] = cs.edge_off[k] + edge_counts[k];
  }
  IGN_REQUIRE(cs.edge_off[n_sources] < (int64_t)INT32_MAX, IGN_ERR_UNSUPPORTED, "IGNNITION: attention_combine: int32 edge positions");
  IGN_REQUIRE(cs.edge_off[n_sources] == 0 || (perm && slot_col), IGN_ERR_INVALID, "IGNNITION: attention_combine: null pointer");
  attn_combine_kernel<<<(unsigned)ign_cdiv(num_dst + 1, 128), 128, 0, ign_stream(stream)>>>(cs, num_dst, max_len, rowptr, perm,
                                                                                         slot_col);
  IGN_CHECK_LAUNCH("attn_combine");
  return IGN_OK;
}

extern "C" int ign_attention_aggregate(const int32_t* rowptr, const int32_t* col, const int32_t* slot_col,
                                       const float* rows, int F,
                                       const float* src_score, const float* dst_score,
                                       const int32_t* sample_offsets, int64_t n_samples, int64_t num_dst,
                                       int64_t n_edges, int max_len, float* out, void* ws, size_t ws_bytes,
                                       void* stream) {
  IGN_REQUIRE(num_dst >= 0 && n_edges >= 0 && n_samples >= 0 && max_len >= 0, IGN_ERR_INVALID,
              "IGNNITION: attention: negative size");
  IGN_REQUIRE(F > 0 && F % 4 == 0, IGN_ERR_UNSUPPORTED, "IGNNITION: attention: message width must be a multiple of 4");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && out && sample_offsets && dst_score, IGN_ERR_INVALID, "IGNNITION: attention: null pointer");
  IGN_REQUIRE(n_edges == 0 || (col && rows && src_score), IGN_ERR_INVALID, "IGNNITION: attention: null pointer");
  IGN_REQUIRE(ws_bytes >= ign_attention_ws_bytes(n_edges, n_samples, max_len) && (ws || ws_bytes == 0),
              IGN_ERR_INVALID, "IGNNITION: attention: workspace too small");
  cudaStream_t st = ign_stream(stream);
  const int64_t n_cols = n_samples * (int64_t)max_len;
  AttnWs w = carve(ws, n_edges, n_cols, slot_col != nullptr);
  if (n_cols > 0) {
    attn_init_kernel<<<(unsigned)ign_cdiv(n_cols, 256), 256, 0, st>>>(n_cols, w);
    IGN_CHECK_LAUNCH("attn_init");
    attn_score_kernel<<<(unsigned)ign_cdiv(num_dst, 128), 128, 0, st>>>(rowptr, col, slot_col, src_score, dst_score,
                                                                        sample_offsets, (int)n_samples, num_dst, max_len, w);
    IGN_CHECK_LAUNCH("attn_score");
    attn_sum_kernel<<<(unsigned)ign_cdiv(num_dst, 128), 128, 0, st>>>(rowptr, slot_col, sample_offsets, (int)n_samples,
                                                                      num_dst, max_len, w);
    IGN_CHECK_LAUNCH("attn_sum");
  }
  attn_apply_kernel<<<(unsigned)ign_cdiv(num_dst * 8, 256), 256, 0, st>>>(rowptr, col, slot_col, rows, F, sample_offsets,
                                                                          (int)n_samples, num_dst, max_len, w, out);
  IGN_CHECK_LAUNCH("attn_apply");
  return IGN_OK;
}

extern "C" size_t ign_attention_bwd_ws_bytes(int64_t n_edges, int64_t n_samples, int max_len) {
  const int64_t n_cols = (n_samples > 0 ? n_samples : 0) * (int64_t)(max_len > 0 ? max_len : 0);
  return ign_align_up(n_cols * sizeof(double), 256) + ign_align_up((n_edges > 0 ? n_edges : 0) * sizeof(float), 256) + 256;
}

extern "C" int ign_attention_aggregate_bwd(const int32_t* rowptr, const int32_t* idx, const int32_t* perm,
                                           const int32_t* slot_col, const float* rows, int F, const float* g_out,
                                           const int32_t* sample_offsets, int64_t n_samples, int64_t num_dst,
                                           int64_t n_edges, int max_len, const void* fwd_ws, float* d_msg, float* d_pre4,
                                           float* d_ds, void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(num_dst >= 0 && n_edges >= 0 && n_samples >= 0 && max_len >= 0, IGN_ERR_INVALID,
              "IGNNITION: attention_bwd: negative size");
  IGN_REQUIRE(F > 0 && F % 4 == 0, IGN_ERR_UNSUPPORTED, "IGNNITION: attention_bwd: message width must be a multiple of 4");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && g_out && sample_offsets && d_ds && fwd_ws, IGN_ERR_INVALID, "IGNNITION: attention_bwd: null pointer");
  IGN_REQUIRE(n_edges == 0 || (idx && rows && d_msg && d_pre4), IGN_ERR_INVALID, "IGNNITION: attention_bwd: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= ign_attention_bwd_ws_bytes(n_edges, n_samples, max_len), IGN_ERR_WORKSPACE,
              "IGNNITION: attention_bwd: workspace too small");
  cudaStream_t st = ign_stream(stream);
  const int64_t n_cols = n_samples * (int64_t)max_len;
  AttnWs w = carve(const_cast<void*>(fwd_ws), n_edges, n_cols, slot_col != nullptr);      // column statistics and scores of the forward pass
  double* colS = reinterpret_cast<double*>(ws);
  float* dcoef = reinterpret_cast<float*>(static_cast<char*>(ws) + ign_align_up(n_cols * sizeof(double), 256));
  IGN_CUDA(cudaMemsetAsync(colS, 0, (size_t)n_cols * sizeof(double), st));
  if (n_edges > 0) {
    attn_bwd_dot_kernel<<<(unsigned)ign_cdiv(num_dst * 8, 256), 256, 0, st>>>(rowptr, idx, slot_col, rows, F, g_out,
                                                                            sample_offsets, (int)n_samples, num_dst, max_len,
                                                                            w, dcoef, colS);
    IGN_CHECK_LAUNCH("attn_bwd_dot");
  }
  attn_bwd_apply_kernel<<<(unsigned)ign_cdiv(num_dst * 8, 256), 256, 0, st>>>(rowptr, perm, slot_col, F, g_out,
                                                                            sample_offsets, (int)n_samples, num_dst, max_len,
                                                                            w, dcoef, colS, d_msg, d_pre4, d_ds);
  IGN_CHECK_LAUNCH("attn_bwd_apply");
  return IGN_OK;
}
