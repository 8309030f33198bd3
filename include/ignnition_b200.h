/*
 * ignnition_b200 -- C-ABI of the B200-native message-passing engine.
 *
 * The reference (zhangbiqiong/ignnition) has no FFI: its generated model calls TensorFlow ops from
 * Python.  Each entry point below replaces the TF op sequence named in its comment (reference
 * file:line, relative to the reference root) and is what a ctypes binding in the reference's
 * ComnetModel.call would bind.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.
 *  - every pointer is a DEVICE pointer owned by the caller unless the comment says "host".
 *    The library never allocates or frees device memory: scratch is passed in, its size is
 *    queried with the matching *_ws_bytes function.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), does no host
 *    synchronisation and no allocation, so it can be captured in a CUDA graph.
 *  - return value: 0 = OK, <0 = invalid argument (checked on the host before any launch),
 *    >0 = cudaError_t of the failed launch.  ign_last_error() returns the thread-local message.
 *  - states / messages / weights are fp32 row-major; indices are int32; weights are in Keras
 *    layout: Dense kernel[in,out], bias[out]; GRUCell (v2, reset_after=True) kernel[in,3u],
 *    recurrent_kernel[u,3u], bias[2,3u], gate order z|r|h.
 */
#ifndef IGNNITION_B200_H
#define IGNNITION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGN_OK 0
#define IGN_ERR_INVALID (-1)
#define IGN_ERR_UNSUPPORTED (-2)
#define IGN_ERR_WORKSPACE (-3)

/* aggregation operators of ign_segment_reduce */
#define IGN_OP_SUM 0
#define IGN_OP_MEAN 1
#define IGN_OP_MAX 2
#define IGN_OP_SUM_ADD 3 /* out += segment sum: partial sums over buckets of edges (partitioned graphs), gradient accumulation */

/* activations (tf.keras.activations names, auxilary_classes.py:839-865) */
#define IGN_ACT_LINEAR 0
#define IGN_ACT_RELU 1
#define IGN_ACT_SELU 2
#define IGN_ACT_SIGMOID 3
#define IGN_ACT_TANH 4
#define IGN_ACT_ELU 5
#define IGN_ACT_SOFTPLUS 6
#define IGN_ACT_LEAKY_RELU 7
/* ign_dense_bwd only: OR-ed into `act` when the `pre_act` argument holds the layer's OUTPUT act(x W + b) instead of the
 * pre-activation.  Every activation above has a derivative that is a function of its output (selu: y > 0 ? scale :
 * y + scale alpha), so the train step saves one tensor per layer instead of two. */
#define IGN_ACT_FROM_OUTPUT 0x100

/* step-table entries of ign_gru_seq: (source id << 28) | row, or IGN_STEP_ZERO for a zero message */
#define IGN_STEP_SRC_SHIFT 28
#define IGN_STEP_ROW_MASK 0x0FFFFFFF
#define IGN_STEP_ZERO (-1)
#define IGN_MAX_SOURCES 4
/* state buffers one fused update can write (the caller's own + every peer GPU of a partitioned graph) */
#define IGN_MAX_PEERS 8

int ign_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf; returns its length */
int ign_last_error(char* buf, size_t n);
/* number of CUDA kernels this library has launched in the process so far (for bench.py's gpu_launches) */
int64_t ign_launch_count(void);
/* Process-wide switch: 1 (default) = kernels that have a tcgen05 3xTF32 variant use it (ign_gru_seq
 * for 32-wide states, ign_dense when given a workspace); 0 = always the fp32 CUDA-core twins.
 * Returns the previous setting.  Both variants meet the 1e-5 fp32 parity bar. */
int ign_set_tensor_cores(int enable);

/* ---------------------------------------------------------------------------------------------
 * Adjacency -> CSR by destination.
 * Replaces the per-edge host loop + padded scatter of the reference:
 *   generator_std_to_framework.py:134-185 (src_idx/dst_idx/seq), generate_model.py:479-490
 *   (lens = unsorted_segment_sum(1, dst), scatter_nd into [num_dst, max_len, F]).
 * Output: rowptr[num_dst+1] (exclusive scan of in-degrees), col[E] = src of the edge at slot
 * rowptr[d]+seq, perm[E] = position of that edge in the input arrays (nullable).
 * seq == NULL : stable LSD radix sort of the edges by dst (seq is then the rank in input order).
 * seq != NULL and mode == IGN_CSR_RANK: histogram + scan + placement at rowptr[dst]+seq.
 * seq != NULL and mode == IGN_CSR_SORT: radix sort as above (edges of one destination must appear
 *   in ascending seq order, which the reference generator guarantees, :153).
 * status (nullable, int32[2] device): [0] = number of slots whose seq disagrees with its slot
 *   (0 for valid input), [1] = max in-degree.
 */
#define IGN_CSR_SORT 0
#define IGN_CSR_RANK 1
size_t ign_csr_build_ws_bytes(int64_t n_edges, int64_t num_dst);
int ign_csr_build(const int32_t* dst, const int32_t* src, const int32_t* seq, int64_t n_edges,
                  int64_t num_dst, int mode, int32_t* rowptr, int32_t* col, int32_t* perm,
                  int32_t* status, void* ws, size_t ws_bytes, void* stream);

/* Destinations ordered by descending in-degree (stable), so that one CTA of ign_gru_seq walks
 * sequences of equal length (replaces the dense right-padding + tf.sequence_mask of
 * generate_model.py:484-490 / auxilary_classes.py:785-790).  order[num_dst]. */
size_t ign_length_order_ws_bytes(int64_t num_dst);
int ign_length_order(const int32_t* rowptr, int64_t num_dst, int32_t* order, void* ws,
                     size_t ws_bytes, void* stream);

/* Walk plan of ign_gru_seq: meta[4*i .. 4*i+3] = (destination, first step, number of steps, first
 * step entry) of the i-th destination in `order` (identity when order == NULL).  One 16-byte load per
 * destination replaces the order -> rowptr -> steps pointer chase at the start of every tile. */
int ign_seq_meta(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* order,
                 int64_t num_dst, int32_t* meta, void* stream);

/* Step table for multi-source ordered / interleave aggregation (generate_model.py:507-543,
 * auxilary_classes.py:421-440).  For destination d of sample s = dst_sample[d], position t of the
 * reference's padded sequence is source pos_src[pos_off[s]+t], column pos_col[pos_off[s]+t];
 * it holds a real message iff column < in-degree of d in that source's CSR, else zeros.
 * final_len(d) = sum of in-degrees.  Outputs: steps_rowptr[num_dst+1], steps[sum final_len]
 * (entries as IGN_STEP_*).  Call with steps == NULL to compute only steps_rowptr. */
size_t ign_steps_build_ws_bytes(int64_t num_dst);
int ign_steps_build(int n_src, const int32_t* const* rowptrs /*host array of device ptrs*/,
                    const int32_t* const* cols /*host array of device ptrs*/,
                    const int32_t* dst_sample, const int32_t* pos_off, const int32_t* pos_src,
                    const int32_t* pos_col, int64_t num_dst, int32_t* steps_rowptr, int32_t* steps,
                    void* ws, size_t ws_bytes, void* stream);

/* Sort keys for the transposed (by source row) view of a step table, used by the backward pass to
 * reduce d_steps per source row without atomics: keys[i] = row of step i if it belongs to source
 * src_id, else n_rows (a dummy bucket).  Feed keys to ign_csr_build(seq = NULL, num_dst = n_rows+1). */
int ign_steps_keys(const int32_t* steps, int64_t n, int src_id, int64_t n_rows, int32_t* keys,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Entity.calculate_hs (auxilary_classes.py:128-160): state[n, hidden] = [features | zeros].
 * feats: host array of n_feat device pointers, each [n, feat_size[i]] fp32. */
int ign_init_state(int n_feat, const float* const* feats, const int32_t* feat_size /*host*/,
                   int64_t n, int hidden, float* state, void* stream);

/* gather + aggregation: out[d] = op_{e in [rowptr[d], rowptr[d+1])} src_states[col[e]]  in slot order.
 * Replaces tf.gather (generate_model.py:432) + scatter_nd (:490) + reduce_sum(axis=1)
 * (auxilary_classes.py:254-262).  col == NULL means identity (messages already in slot order).
 * mean / max are north-star extensions (empty destination -> 0). F % 4 == 0, F <= 256. */
int ign_segment_reduce(int op, const int32_t* rowptr, const int32_t* col, const float* src_states,
                       int F, int64_t num_dst, float* out, void* stream);

/* One GRU step for every destination, x = aggregated messages, h = old state:
 * Recurrent_Cell.perform_unsorted_update (auxilary_classes.py:752-765).
 * With a workspace of ign_gru_cell_ws_bytes(f_in, units) > 0 bytes (f_in == units in {32, 64}) the gate
 * GEMMs run as 3xTF32 on tcgen05 and out must NOT alias h; without it (ws == NULL) the fp32 CUDA-core
 * kernel runs and out may alias h. */
size_t ign_gru_cell_ws_bytes(int f_in, int units);
int ign_gru_cell(const float* x, const float* h, int64_t n, int f_in, int units,
                 const float* kernel, const float* recurrent_kernel, const float* bias,
                 float* out, void* ws, size_t ws_bytes, void* stream);

/* Fused gather + sum aggregation + GRU update (RouteNet stage 2, Q-size step 2, config 5):
 * generate_model.py:432,490 + Sum_aggr (auxilary_classes.py:254-262) + perform_unsorted_update
 * (:752-765).  The aggregated message never leaves the SM.  agg_out (nullable) receives the
 * aggregated messages [num_dst, f_in] (saved for the backward pass).  out must not alias h_dst
 * when src_states == h_dst. */
int ign_agg_gru_cell(const int32_t* rowptr, const int32_t* col, const float* src_states, int f_in,
                     const float* h_dst, int64_t num_dst, int units, const float* kernel,
                     const float* recurrent_kernel, const float* bias, float* out, float* agg_out,
                     void* stream);

/* Ordered / interleave aggregation + recurrent update: for every destination d,
 *   h <- h0[d]; for t in 0..len(d)-1: h <- GRU(x = message(steps[steps_rowptr[d]+t]), h); out[d] = h
 * Replaces keras RNN(GRUCell)(padded, initial_state, mask=sequence_mask(len)) + gather_nd
 * (auxilary_classes.py:767-796) and Interleave_aggr (:421-440).  A destination with len 0 keeps
 * its state.  srcs: host array of n_src device pointers [*, f_in].  order (nullable) = output of
 * ign_length_order.  h_seq (nullable): [sum len, units] hidden state after every step (saved for
 * the backward pass).  out must not alias h0 unless no source aliases h0. */
int ign_gru_seq(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* order, int n_src,
                const float* const* srcs, int f_in, const float* h0, int64_t num_dst, int units,
                const float* kernel, const float* recurrent_kernel, const float* bias, float* out,
                float* h_seq, const int32_t* meta /* nullable: output of ign_seq_meta for this order */,
                void* stream);

/* The same update with the input projection hoisted out of the walk (csrc/gru_seq_proj_tc.cu): x K + b is computed
 * once per SOURCE row into a 96-wide table (workspace), the walk gathers table rows and runs only the recurrent
 * GEMM per step on tcgen05 with the state operand in tensor memory.  Pays when a source row is walked over more
 * than once (RouteNet: 303 k link rows, 6.1 M steps).  f_in == units == 32; meta = the plan of ign_seq_meta
 * (required); src_rows: HOST array, rows of every source array.  Results equal ign_gru_seq to fp32 rounding. */
size_t ign_gru_seq_proj_ws_bytes(int n_src, const int64_t* src_rows /*host*/, int f_in, int units);
int ign_gru_seq_proj(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* meta, int n_src,
                     const float* const* srcs /*host array of device ptrs*/, const int64_t* src_rows /*host*/,
                     int f_in, const float* h0, int64_t num_dst, int units, const float* kernel,
                     const float* recurrent_kernel, const float* bias, float* out, float* h_seq, void* ws,
                     size_t ws_bytes, void* stream);


/* Step-synchronous form of ign_gru_seq for short sequences (RouteNet paths: <= 6 links): launch t
 * executes step t of every destination that has one.  Needs the destinations sorted by descending
 * length (ign_length_order), their walk plan (ign_seq_meta) and the step-major plan of
 * ign_seq_step_plan:  nt[t] = destinations with more than t steps (a prefix of the sorted order),
 * off[t] = sum_{u<t} nt[u], steps_T[off[t] + i] = step t of the i-th sorted destination
 * (steps_T must hold max(sum len, max_steps + 1) ints; max_steps >= the longest sequence).
 * hs[num_dst, units] is the running state in sorted order between launches; the launch that runs a
 * destination's last step writes out[d]; launch 0 also copies the state of destinations without any
 * step.  Call for t = 0 .. max_steps-1.  Same results as ign_gru_seq.  32-wide, tcgen05 3xTF32. */
int ign_seq_step_plan(const int32_t* meta, const int32_t* steps, int64_t num_dst, int max_steps,
                      int32_t* nt, int32_t* off, int32_t* steps_T, void* stream);
int ign_gru_seq_step(int t, const int32_t* nt, const int32_t* off, const int32_t* meta,
                     const int32_t* steps_T, int n_src, const float* const* srcs, int f_in,
                     const float* h0, float* hs, int64_t num_dst, int units, const float* kernel,
                     const float* recurrent_kernel, const float* bias, float* out, float* h_seq,
                     void* stream);

/* Dense layer y = act(x W + b): Feed_forward_Layer (auxilary_classes.py:800-866), used by the
 * message MLP (generate_model.py:448-473), the FF update (:594-600) and the readout (:607-629).
 * bias nullable.  pre_act (nullable) receives x W + b (saved for the backward pass).
 * When ws holds at least ign_dense_ws_bytes(k, n) > 0 bytes the layer runs on the tcgen05 tensor
 * cores as 3xTF32 (fp32-accurate split products, fp32 accumulation in TMEM); otherwise, or for
 * shapes the tensor-core path is not built for (ign_dense_ws_bytes == 0), on the fp32 CUDA cores. */
size_t ign_dense_ws_bytes(int k, int n);
int ign_dense(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
              float* y, float* pre_act, void* ws, size_t ws_bytes, void* stream);

/* Dense layer fused with a following single-output linear layer (the readout's 256 -> 1 head,
 * examples/Routenet/model_description.json:119-141):  out[m] = act(x W + b) . head_w + head_b.
 * The hidden activations never go to memory.  Tensor-core shapes only (ign_dense_ws_bytes(k, n) > 0,
 * m >= 128); head_b points to one float on the device (nullable = 0). */
int ign_dense_head(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                   const float* head_w, const float* head_b, float* out, void* ws,
                   size_t ws_bytes, void* stream);

/* The whole predict stack with two hidden layers and one linear output in ONE kernel:
 *   out[m] = act2( act1(x W1 + b1) W2 + b2 ) . w3 + b3
 * (RouteNet / Q-size readout 32 -> 256 -> 256 -> 1, generate_model.py:612-629).  Both hidden activations
 * stay on chip (TMEM -> shared-memory operand image -> next GEMM).  Inference only; tensor-core shapes
 * only (ign_mlp_head_ws_bytes > 0, m >= 128).  b1, b2, b3 nullable; b3 points to one device float. */
size_t ign_mlp_head_ws_bytes(int k1, int n1, int n2);
int ign_mlp_head(const float* x, int64_t m, int k1, const float* w1, const float* b1, int n1, int act1,
                 const float* w2, const float* b2, int n2, int act2, const float* w3, const float* b3,
                 float* out, void* ws, size_t ws_bytes, void* stream);

/* Row-wise concatenation of up to 4 blocks, each optionally gathered by an index:
 * tf.concat([hs_source, hs_dest, edge_params], axis=1) after tf.gather (generate_model.py:432-465)
 * and tf.concat([agg, old_state], 1) of the FF update (:599).  idx[i] nullable = identity. */
int ign_gather_concat(int n_parts, const float* const* parts, const int32_t* const* idx,
                      const int32_t* widths /*host*/, int64_t rows, float* out, void* stream);

/* First Dense layer of a message network with the gather and the concat FUSED into the A-operand loaders of the
 * tcgen05 GEMM (generate_model.py:432-475): y[r, :] = act(concat_k srcs[k][idx[k] ? idx[k][r] : r, :] W + b); the
 * [rows, sum widths] input is never written (at 200 M edges and three 64-wide parts that is a 154 GB tensor).  Built
 * for widths that are multiples of 32 with sum <= 256, 32 <= n <= 256 (n % 32 == 0), rows >= 128; anything else
 * returns IGN_ERR_UNSUPPORTED and the caller runs ign_gather_concat + ign_dense.  idx entries may be NULL
 * (identity) and negative indices give zero rows. */
size_t ign_gather_dense_ws_bytes(int n_src, const int32_t* widths /*host*/, int n);
int ign_gather_dense(int n_src, const float* const* srcs /*host array of device ptrs*/,
                     const int32_t* const* idx /*host array of device ptrs, entries nullable*/,
                     const int32_t* widths /*host*/, int64_t rows, const float* w, const float* bias, int n, int act,
                     float* y, void* ws, size_t ws_bytes, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Train step (model_fn, generate_model.py:697-830): backward twins + loss + Adam.
 */
/* loss = mean((y - pred)^2) over all n predictions (MeanSquaredError, :745-751).
 * Writes d_pred = 2 (pred - y) * grad_scale  (grad_scale = 1/global_n for data-parallel runs) and
 * accumulates sum of squared errors into sse[0] (fp64, caller zeroes it). */
int ign_mse_loss(const float* pred, const float* label, int64_t n, float grad_scale, float* d_pred,
                 double* sse, void* stream);

/* dx = (dy * act'(pre)) W^T ; dW += x^T (dy*act') ; db += colsum(dy*act').
 * dx nullable.  dy is overwritten with dy*act'(pre).  dW/db are ACCUMULATED (caller zeroes).
 * ws (nullable, ign_dense_bwd_ws_bytes(k, n) bytes, 0 = shape not built for tensor cores): with it dx
 * runs as a 3xTF32 tcgen05 GEMM with the transposed kernel; dW stays an fp32 CUDA-core GEMM. */
size_t ign_dense_bwd_ws_bytes(int k, int n);
int ign_dense_bwd(const float* x, int64_t m, int k, const float* w, int n, int act,
                  const float* pre_act, float* dy, float* dx, float* dw, float* db, void* ws,
                  size_t ws_bytes, void* stream);

/* backward of ign_gru_cell: given d_out [n,units] computes dx [n,f_in], dh [n,units] (both nullable)
 * and ACCUMULATES d_kernel, d_recurrent_kernel, d_bias. */
int ign_gru_cell_bwd(const float* x, const float* h, int64_t n, int f_in, int units,
                     const float* kernel, const float* recurrent_kernel, const float* bias,
                     const float* d_out, float* dx, float* dh, float* d_kernel,
                     float* d_recurrent_kernel, float* d_bias, void* stream);

/* Element-wise middle of the GENERIC GRU-cell backward (any f_in, units; the fused ign_gru_cell_bwd is built for
 * f_in == units in {16, 32}): given zx = x K + b_in and zh = h R + b_rec ([n, 3 units], from ign_dense), overwrites
 * zx with GX = [d_az | d_ar | d_axh], zh with GH = [d_az | d_ar | d_ahh] and writes dh_direct = d_out * z; the caller
 * finishes with ign_dense_bwd(x, K, GX) and ign_dense_bwd(h, R, GH) (tf.gradients through GRUCell,
 * generate_model.py:791, auxilary_classes.py:752-765). */
int ign_gru_gates_bwd(float* zx, float* zh, const float* h, const float* d_out, int64_t n, int units,
                      float* dh_direct, void* stream);
/* Backward of the mean / max segment aggregations (north_star extensions): d[r, :] /= max(deg r, 1) turns dL/d(mean)
 * into dL/d(sum); segment_max_bwd writes the per-slot message gradients of a max (ties share the gradient evenly, as
 * TensorFlow's unsorted_segment_max does) at row perm[slot] (or slot when perm is NULL) of d_msg [E, width]. */
int ign_scale_rows_inv_degree(float* d, const int32_t* rowptr, int64_t n, int width, void* stream);
/* out[r, :] = rows[s, :] for every r in [rowptr[s], rowptr[s + 1]): backward of a per-sample sum / mean pooling
 * (auxilary_classes.py:1165-1185) */
int ign_segment_broadcast(const int32_t* rowptr, const float* rows, int64_t n_seg, int width, float* out, void* stream);
int ign_segment_max_bwd(const int32_t* rowptr, const int32_t* idx, const int32_t* perm, const float* rows,
                        const float* agg, const float* d_agg, int64_t num_dst, int width, float* d_msg, void* stream);


/* backward of ign_gru_seq (BPTT over every destination's sequence).  h_seq is the saved output of
 * the forward call.  d_steps [sum len, f_in] receives the gradient w.r.t. each step's message
 * (the caller reduces it per source row with ign_segment_reduce over the transposed CSR);
 * dh0 [num_dst, units] the gradient w.r.t. the initial state.  Weight gradients are ACCUMULATED. */
int ign_gru_seq_bwd(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* order,
                    int n_src, const float* const* srcs, int f_in, const float* h0,
                    const float* h_seq, int64_t num_dst, int units, const float* kernel,
                    const float* recurrent_kernel, const float* bias, const float* d_out,
                    float* d_steps, float* dh0, float* d_kernel, float* d_recurrent_kernel,
                    float* d_bias, void* stream);

/* Step-synchronous form of ign_gru_seq_bwd on the tcgen05 tensor cores (3xTF32; 32-wide messages and
 * states): BPTT step t of every destination that has one is a pair of launches (gate recomputation +
 * [dx | dh] GEMM; weight-gradient GEMM with TMEM-resident accumulators), t = max_steps-1 .. 0.  Needs the
 * plan of ign_seq_step_plan over destinations sorted by descending length (nt, off, steps_T, meta; see
 * ign_gru_seq_step) and h_seq of the forward call.  Same outputs as ign_gru_seq_bwd; weight gradients are
 * ACCUMULATED.  ws: ign_gru_seq_bwd_steps_ws_bytes(num_dst) bytes. */
size_t ign_gru_seq_bwd_steps_ws_bytes(int64_t num_dst);
int ign_gru_seq_bwd_steps(int max_steps, const int32_t* nt, const int32_t* off, const int32_t* meta,
                          const int32_t* steps_T, int n_src, const float* const* srcs, int f_in,
                          const float* h0, const float* h_seq, int64_t num_dst, int units,
                          const float* kernel, const float* recurrent_kernel, const float* bias,
                          const float* d_out, float* d_steps, float* dh0, float* d_kernel,
                          float* d_recurrent_kernel, float* d_bias, void* ws, size_t ws_bytes,
                          void* stream);

/* Backward of a linear single-output layer (the readout head, k -> 1) chained with the activation of the layer below it:
 * x [m, k] is that layer's OUTPUT (activation prev_act), dz [m] the gradient w.r.t. the head's output.
 *   dw[k] += sum_m x[m, k] dz[m];   dz_prev[m, k] = dz[m] w[k] act'(x[m, k]);   db_prev[k] += colsum(dz_prev) (nullable)
 * One streaming pass; the layer below then calls ign_dense_bwd with IGN_ACT_LINEAR and d_bias = NULL.  k in {128, 256, 512}. */
int ign_dense_head_bwd_chain(const float* x, int64_t m, int k, const float* w, const float* dz, int prev_act,
                             float* dz_prev, float* dw, float* db_prev, void* stream);

/* l2 regulariser: reg[0] += lambda * sum(w^2) (fp64), dw += 2 lambda w (auxilary_classes.py:834). */
int ign_l2_reg(const float* w, int64_t n, float lambda, float* dw, double* reg, void* stream);

/* Keras losses by name (generate_model.py:745-751 takes any tf.keras.losses class): acc += sum_i l_i (fp64, caller
 * zeroes; the caller divides by the global count), d_pred = dl_i/dp * grad_scale (nullable).  delta: Huber only. */
#define IGN_LOSS_MSE 0      /* MeanSquaredError */
#define IGN_LOSS_MAE 1      /* MeanAbsoluteError */
#define IGN_LOSS_MAPE 2     /* MeanAbsolutePercentageError */
#define IGN_LOSS_MSLE 3     /* MeanSquaredLogarithmicError */
#define IGN_LOSS_HUBER 4    /* Huber(delta) */
#define IGN_LOSS_LOGCOSH 5  /* LogCosh */
#define IGN_LOSS_BCE 6      /* BinaryCrossentropy on probabilities */
int ign_loss(int kind, const float* pred, const float* label, int64_t n, float grad_scale, float delta,
             float* d_pred, double* acc, void* stream);
/* Keras optimisers other than Adam (generate_model.py:796-818), TF-2.1 fused-op formulas on the flat buffers; s1 / s2
 * are the slot buffers (SGD: momentum | -; RMSprop: rms | momentum; Adagrad: accumulator | -; Adamax: m | u).
 * a, b: SGD momentum, -; RMSprop rho, momentum; Adamax beta1, beta2 (lr already divided by 1 - beta1^t).
 * flags & 1: Nesterov momentum (SGD). */
#define IGN_OPT_SGD 1
#define IGN_OPT_RMSPROP 2
#define IGN_OPT_ADAGRAD 3
#define IGN_OPT_ADAMAX 4
int ign_optimizer_step(int kind, float* w, const float* g, float* s1, float* s2, int64_t n, float lr, float a,
                       float b, float eps, int flags, void* stream);
/* Element-wise end of the GENERIC GRU step (any f_in, units): out = GRU gates of zx = x K + b_in, zh = h R + b_rec
 * (two ign_dense calls) and the old state h (auxilary_classes.py:752-765 for widths the fused kernels do not cover) */
int ign_gru_gates_fwd(const float* zx, const float* zh, const float* h, int64_t n, int units, float* out, void* stream);

/* GRUCell with reset_after = False (the Keras v1 cell; a model JSON reaches it through the recurrent network's cell
 * parameters, auxilary_classes.py:740-750, bias [3 units]): the reset gate multiplies h before the candidate's recurrent
 * product, so one step is zx = x K + b, zh2 = h R[:, :2U], rh = sigmoid(zx_r + zh2_r) * h (ign_gru_v1_reset),
 * zhh = rh R[:, 2U:], out = z h + (1 - z) tanh(zx_h + zhh) (ign_gru_v1_out) -- three ign_dense calls around two
 * element-wise kernels.  Backward: ign_gru_v1_bwd_out turns the z and candidate slices of zx / zh2 and zhh into their
 * gradients in place and writes dh_direct = d_out * z; after d_rh = d_zhh R[:, 2U:]^T (ign_dense_bwd) ign_gru_v1_bwd_reset
 * turns the r slices into theirs and adds d_rh * r to dh_direct; ign_dense_bwd on (x, K) and (h, R[:, :2U]) finish. */
int ign_gru_v1_reset(const float* zx, const float* zh2, const float* h, int64_t n, int units, float* rh, void* stream);
int ign_gru_v1_out(const float* zx, const float* zh2, const float* zhh, const float* h, int64_t n, int units, float* out,
                   void* stream);
int ign_gru_v1_bwd_out(float* zx, float* zh2, float* zhh, const float* h, const float* d_out, int64_t n, int units,
                       float* dh_direct, void* stream);
int ign_gru_v1_bwd_reset(float* zx, float* zh2, const float* h, const float* d_rh, int64_t n, int units,
                         float* dh_direct, void* stream);

/* Keras Adam step on a flat parameter buffer (generate_model.py:796-818) [TF-2.1 semantics]:
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v moments; w -= lr_t*m/(sqrt(v)+eps).  step is 1-based. */
int ign_adam_step(float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, void* stream);

/* Attention_aggr.calculate_input (auxilary_classes.py:278-344) on the CSR by destination:
 * a_j = leaky_relu(src_score[col[j]] + dst_score[d], 0.2); softmax over the destinations of one sample
 * per padded column s = j - rowptr[d] (zero pads included, as the reference's softmax over axis 0 does);
 * out[d] = sum_j coef_j * rows[col[j]].  src_score = rows . (kernel1 . attn_kernel[:F]) and
 * dst_score = h_dst . (kernel2 . attn_kernel[F:]) are ign_dense calls.  sample_offsets[n_samples+1] are
 * the destination entity's per-sample row offsets; max_len >= the longest destination list.
 * slot_col (NULL for one source): the padded column of every CSR slot when several sources feed the aggregation
 * (generate_model.py:525-543, ign_attention_combine); slots of one destination on the same column form one cell whose
 * score is the SUM of their activated scores (tf.scatter_nd adds) and share its coefficient. */
size_t ign_attention_ws_bytes(int64_t n_edges, int64_t n_samples, int max_len);
int ign_attention_aggregate(const int32_t* rowptr, const int32_t* col, const int32_t* slot_col, const float* rows, int F,
                            const float* src_score, const float* dst_score, const int32_t* sample_offsets,
                            int64_t n_samples, int64_t num_dst, int64_t n_edges, int max_len, float* out,
                            void* ws, size_t ws_bytes, void* stream);

/* Backward of ign_attention_aggregate (tf.gradients through Attention_aggr).  fwd_ws = the workspace the forward call
 * filled (scores and column statistics; keep it).  Outputs, per edge at its input position perm[slot] (slot when perm is
 * NULL): d_msg [E, F] = coef g_out[d] (the weighted-sum part of dL/d rows) and d_pre4 [E, 4] = (dL/d(src_score +
 * dst_score before the LeakyReLU), 0, 0, 0); per destination d_ds [num_dst] = dL/d dst_score.  The caller reduces the
 * per-edge arrays per row of `rows` (ign_segment_reduce over the transposed adjacency) and finishes the two score
 * products with ign_dense_bwd. */
size_t ign_attention_bwd_ws_bytes(int64_t n_edges, int64_t n_samples, int max_len);
int ign_attention_aggregate_bwd(const int32_t* rowptr, const int32_t* idx, const int32_t* perm, const int32_t* slot_col,
                                const float* rows, int F, const float* g_out, const int32_t* sample_offsets, int64_t n_samples, int64_t num_dst,
                                int64_t n_edges, int max_len, const void* fwd_ws, float* d_msg, float* d_pre4,
                                float* d_ds, void* ws, size_t ws_bytes, void* stream);

/* One CSR by destination over the edge lists of several sources (generate_model.py:523-543): row d = the sources'
 * rows one after the other; perm[slot] = position of the slot's edge in the concatenation of the sources' input edge
 * lists; slot_col[slot] = the edge's seq (its slot in the source's own row) for the first source, seq + the
 * destination's edge count in that source for the others -- the reference gathers `lens` of the CURRENT source
 * (generate_model.py:538-539), reproduced as it is.  src_rowptr[k] / src_perm[k] = CSR of source k alone (perm NULL =
 * identity), edge_counts[k] on the host; rowptr [num_dst + 1], perm and slot_col [sum of edge_counts]. */
int ign_attention_combine(int n_sources, const int32_t* const* src_rowptr, const int32_t* const* src_perm,
                          const int64_t* edge_counts, int64_t num_dst, int max_len, int32_t* rowptr, int32_t* perm,
                          int32_t* slot_col, void* stream);

/* The whole message-passing loop (generate_model.py:405-602) of a SMALL graph in ONE launch: a persistent grid keeps
 * every stage's GRU weights in shared memory for all `iterations`, a warp owns a destination row, a grid-wide barrier
 * replaces the launch boundary between stages (csrc/small_graph.cu).  For the reference's own batch sizes (3 samples,
 * code/train_options.ini:26) where the per-stage launches are all gaps.  fp32, bit-identical to the per-stage fp32
 * kernels (ign_gru_seq / ign_agg_gru_cell).
 *   units            state AND message width of every entity: 16 or 32
 *   rows[e], buf0[e], buf1[e]   per entity: row count and two state buffers [rows, units]; buf0 holds the initial state
 *   op_kind[o]       0 = ordered walk over step entries (IGN_STEP_* encoding, op_src = entity of each source id),
 *                    1 = sum over the CSR row of op_src[o*4] + one GRU step
 *   op_dst[o], op_src[o*4 .. o*4+3] (-1 = unused), op_rowptr[o], op_idx[o], op_kernel/rkernel/bias[o] (Keras GRU v2)
 *   step_out / step_hseq / step_agg   NULL for inference.  Training (tf.gradients needs every intermediate,
 *                    generate_model.py:791): arrays of iterations * n_ops pointers, stage s = iteration * n_ops + o writes
 *                    its new states to step_out[s] ([rows, units], instead of the entity's other buffer), the state after
 *                    every step of a walk to step_hseq[s] ([step entries, units], as ign_gru_seq's h_seq) and the neighbour
 *                    sum of a kind-1 stage to step_agg[s] ([rows, units]); hseq / agg entries may be NULL.  At most 64 stages.
 *   final_buffer[e]  (host, out) which of the two buffers holds entity e's state after the last iteration (inference)
 *   ws               ign_small_graph_ws_bytes() of device memory (the barrier counter)
 * All stages of one iteration run in order o = 0 .. n_ops-1, each seeing the states the previous ones wrote. */
size_t ign_small_graph_ws_bytes(void);
int ign_small_graph_forward(int units, int n_entities, const int64_t* rows, float* const* buf0, float* const* buf1,
                            int n_ops, const int32_t* op_kind, const int32_t* op_dst, const int32_t* op_src,
                            const int32_t* const* op_rowptr, const int32_t* const* op_idx,
                            const float* const* op_kernel, const float* const* op_rkernel,
                            const float* const* op_bias, int iterations, float* const* step_out,
                            float* const* step_hseq, float* const* step_agg, int32_t* final_buffer, void* ws,
                            size_t ws_bytes, void* stream);

/* ign_csr_build for all adjacencies of a SMALL graph in one launch, one CTA per adjacency (at most 8 adjacencies, 11000
 * destinations each): seq[k] given = every edge at rowptr[dst] + seq (IGN_CSR_RANK; unclaimed slots = -1, a zero row),
 * seq[k] NULL = stable order of the input (IGN_CSR_SORT).  perm[k] may be NULL.  Same arrays as ign_csr_build. */
int ign_csr_build_small(int n_adj, const int32_t* const* dst, const int32_t* const* src, const int32_t* const* seq,
                        const int64_t* n_edges, const int64_t* num_dst, int32_t* const* rowptr, int32_t* const* col,
                        int32_t* const* perm, void* stream);

/* Concat_aggr with concat_axis = 2 (generate_model.py:496-505): for CSR position j of the first source
 * (destination d, padded column s = j - rowptr0[d]) the row of another source sitting at the same padded
 * column: out[j] = idx1[rowptr1[d] + s], or -1 where that source's block is zero padding.  Feed the index
 * lists to ign_gather_concat (a negative index gathers a zero row). */
int ign_partner_index(const int32_t* rowptr0, const int32_t* rowptr1, const int32_t* idx1, int64_t num_dst,
                      int32_t* out, void* stream);

/* out[rows, width] = x[:, col0 : col0 + width] of a row-major [rows, ld] matrix: the gradient of one input of
 * the concat at generate_model.py:465 / :599 (backward of ign_gather_concat; the caller reduces gathered
 * parts per row with ign_segment_reduce). */
int ign_slice_cols(const float* x, int64_t rows, int ld, int col0, int width, float* out, void* stream);

/* Product_operation 'element_wise' (auxilary_classes.py:1085-1086): out = a * b. */
int ign_mul(int64_t n, const float* a, const float* b, float* out, void* stream);

/* Tail of Conv_aggr.calculate_input (auxilary_classes.py:388-401):
 * out[d] = act((nsum[d] + self[d]) / in-degree(d)), in-degree from the CSR rowptr (0 -> inf / nan as in TF).
 * nsum = (sum of messages) . conv_kernel = ign_segment_reduce + ign_dense (the matrix product is linear). */
int ign_conv_finish(const float* nsum, const float* self, const int32_t* rowptr, int F, int64_t n, int act,
                    float* out, void* stream);

/* y[i] += x[idx[i]] row-wise helper and elementwise utilities used by the backward pass */
int ign_axpy(int64_t n, float a, const float* x, float* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Native dataset ingest (host only): the text of a data.json -> the block-diagonal batch arrays.
 * Replaces the per-edge Python loops of generator (generator_std_to_framework.py:32-50, :102-107,
 * :134-190) and the per-sample concatenation; results are identical array for array.
 *
 * The model side is given as flat tables: entity type names; features (name, index of the entity
 * type they belong to); adjacencies (name, source / destination entity type index, 1 if the model
 * reads edge parameters); the label name (nullable: inference).  ign_ingest_parse appends every sample
 * of `json` (one sample object, or an array of them, at most max_samples if >= 0) to the batch under
 * construction and returns how many it appended (< 0: error, message in ign_last_error).  Source and
 * destination indices already carry the per-sample offsets; seq is the position inside the
 * destination's list; edge parameters are truncated towards zero like the reference's int64 cast.
 * The result pointers stay valid until the next parse / reset / destroy.  A handle is single-threaded;
 * different handles may parse different files concurrently (the call holds no global state).
 * ------------------------------------------------------------------------------------------- */
typedef struct ign_ingest ign_ingest_t;
ign_ingest_t* ign_ingest_create(int n_entities, const char* const* entity_names, int n_features,
                                const char* const* feature_names, const int32_t* feature_entity, int n_adj,
                                const char* const* adj_names, const int32_t* adj_src, const int32_t* adj_dst,
                                const int32_t* adj_params, const char* label_name);
/* A multi-source ordered / concat / interleave message passing (generate_model.py:496-543): registers the
 * adjacencies that feed it, in source order; interleave = 1 reads the pattern (a list of entity type names)
 * from the sample key pattern_key (generator_std_to_framework.py:193-219).  Returns the sequence id.
 * ign_ingest_sequence gives, per sample (pos_off[n_samples + 1]), which (source, column) of the reference's
 * concatenated padded tensor sits at sequence position p: the input of ign_steps_build. */
int ign_ingest_add_sequence(ign_ingest_t* g, int n_adj, const int32_t* adj_indices, int interleave,
                            const char* pattern_key);
int64_t ign_ingest_sequence(const ign_ingest_t* g, int seq, const int32_t** pos_off, const int32_t** pos_src,
                            const int32_t** pos_col);
void ign_ingest_destroy(ign_ingest_t* g);
void ign_ingest_reset(ign_ingest_t* g);
int64_t ign_ingest_parse(ign_ingest_t* g, const char* json, size_t len, int64_t max_samples);
int64_t ign_ingest_n_samples(const ign_ingest_t* g);
const int64_t* ign_ingest_offsets(const ign_ingest_t* g, int entity);          /* [n_samples + 1] */
int64_t ign_ingest_feature(const ign_ingest_t* g, int feature, const float** data);
int64_t ign_ingest_adjacency(const ign_ingest_t* g, int adj, const int32_t** src, const int32_t** dst,
                             const int32_t** seq, const float** params, int32_t* params_width);
int64_t ign_ingest_labels(const ign_ingest_t* g, const float** data);

/* ---------------------------------------------------------------------------------------------
 * Fused message passing of an aggregating update on the tensor cores, with the state exchange of a
 * destination-partitioned graph in its epilogue (north_star: "gather ...
 * aggregation ... update", one kernel).  Replaces, for one message passing whose sources send their
 * states (direct_assignation) into a sum / mean / max aggregation with a GRU update:
 *   tf.gather + tf.scatter_nd + reduce_sum   generate_model.py:432, 479-490, auxilary_classes.py:254-262
 *   GRUCell step                              auxilary_classes.py:752-765
 *   state write-back                          generate_model.py:602
 * rowptr[num_dst + 1] / col[E] index rows of src_states [*, f_in] (col < 0 = zero row); h_dst
 * [num_dst, units] are the old states of the destinations.  The new states are written to rows
 * [out_row0, out_row0 + num_dst) of EVERY buffer outs[k] (host array of n_out device pointers, each
 * the base of a [>= out_row0 + num_dst, units] array): the caller's own state array and, on a
 * partitioned graph, the peer-mapped copies of it on the other GPUs (ign_peer_open), so the
 * per-iteration all-gather happens tile by tile from the kernel that computes the states (TMA
 * tensor stores over NVLink).  No buffer may alias src_states / h_dst rows that are still read.
 * agg_out (nullable): the aggregated messages [num_dst, f_in] (saved for the backward pass).
 * Built for f_in == units in {32, 64} (ws_bytes == 0 otherwise: use ign_segment_reduce + ign_gru_cell).
 */
size_t ign_agg_gru_cell_tc_ws_bytes(int f_in, int units);
int ign_agg_gru_cell_tc(int op, const int32_t* rowptr, const int32_t* col, const float* src_states, int f_in,
                        const float* h_dst, int64_t num_dst, int units, const float* kernel,
                        const float* recurrent_kernel, const float* bias, int n_out,
                        float* const* outs /*host array of device ptrs*/, int64_t out_row0, float* agg_out,
                        void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Destination-partitioned graphs: peer-mapped state buffers and the builder's routing helpers.
 * The reference is single-process; these have no counterpart there.
 * ign_peer_* are the ONLY calls of the library that allocate: a CUDA IPC handle covers one whole
 * cudaMalloc allocation, so an exchanged buffer cannot come out of the caller's memory pool.
 *   alloc / free     : cudaMalloc / cudaFree on the current device
 *   export           : 64-byte handle (host memory) to send to the other processes of the node
 *   open / close     : map / unmap a peer's buffer into this process (peer access enabled lazily)
 * ------------------------------------------------------------------------------------------- */
#define IGN_PEER_HANDLE_BYTES 64
int ign_peer_alloc(size_t bytes, void** ptr);
int ign_peer_free(void* ptr);
int ign_peer_export(void* ptr, void* handle64 /*host*/);
int ign_peer_open(const void* handle64 /*host*/, void** ptr);
int ign_peer_close(void* ptr);
/* owner[i] = rank whose row range [bounds[r], bounds[r+1]) holds dst[i]; bounds: HOST array of world + 1 ints */
int ign_edge_owner(const int32_t* dst, int64_t n_edges, const int32_t* bounds /*host*/, int world, int32_t* owner,
                   void* stream);
/* out[i] = in[perm[i]] + add (perm nullable = identity): edge arrays in routed order, global -> local row ids */
int ign_gather_int(const int32_t* in, const int32_t* perm, int64_t n, int add, int32_t* out, void* stream);
/* flags[col[e]] = 1: the source rows an adjacency reads (boundary-only exchange) */
int ign_mark_rows(const int32_t* col, int64_t n, int32_t* flags, void* stream);
/* out[0 .. *count) = add + i for every i with flags[i] != 0, ascending; count: device int */
size_t ign_flag_compact_ws_bytes(int64_t n);
int ign_flag_compact(const int32_t* flags, int64_t n, int add, int32_t* out, int32_t* count, void* ws,
                     size_t ws_bytes, void* stream);
/* dst[rows[i], :] = src[rows[i], :] for a [*, width] fp32 array; dst may be a peer-mapped buffer */
int ign_rows_put(const float* src, const int32_t* rows, int64_t n_rows, int width, float* dst, void* stream);
/* dst[rows[i], :] = packed[i, :]: the receiving side of the PACKED boundary exchange (the sender gathers the rows a
 * peer reads into one contiguous block with ign_gather_concat, the copy engine moves the block with ign_peer_copy
 * into the peer's inbox, the peer scatters it here) */
int ign_rows_unpack(const float* packed, const int32_t* rows, int64_t n_rows, int width, float* dst, void* stream);
/* dst[0 .. bytes) = src[0 .. bytes) by the COPY ENGINE (cudaMemcpyAsync), dst usually a peer-mapped buffer: the
 * exchange of finished row chunks while the update kernel works on the next chunk.  Stores issued by SMs whose
 * memory pipes are busy with a gather reach a peer at 30-70 GB/s, the copy engine at 750 GB/s and without slowing
 * the gather (profiles/r2_p2p_rate.md). */
int ign_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
/* *bad (device int, caller-zeroed) += number of idx[i] outside [0, bound) */
int ign_index_range_check(const int32_t* idx, int64_t n, int64_t bound, int32_t* bad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IGNNITION_B200_H */
