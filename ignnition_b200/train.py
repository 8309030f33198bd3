"""Train step of the generated model: loss, backward, optimiser.

Mirrors ``model_fn`` (reference ``code/utils/generate_model.py:697-830``):
  loss = MeanSquaredError()(labels, predictions) over ALL predictions of the batch (:745-751)
         + sum(model.losses)  (l2(lambda) = lambda * sum(w^2) per regularised Dense kernel, :749)
  grads = tf.gradients(total_loss, variables) (:791);  Adam(ExponentialDecay(...)) (:796-818).

The backward pass walks the forward tape in reverse and calls the backward twins of the kernels
(``ign_dense_bwd``, ``ign_gru_cell_bwd``, ``ign_gru_seq_bwd``); gradients w.r.t. source states are
reduced per source row with ``ign_segment_reduce`` over the TRANSPOSED adjacency (built once per
batch by the same device radix sort), so no atomics touch state gradients.  Data-parallel runs
all-reduce the flat gradient buffer over NCCL and divide by the GLOBAL prediction count, which
reproduces the reference's mean over all predictions.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops


class LearningRate:
    """tf.keras.optimizers.schedules by name (generate_model.py:802-809) [TF-2.1 semantics]."""

    def __init__(self, optimizer: dict):
        self.const = float(optimizer.get("learning_rate", 0.001))
        self.schedule = optimizer.get("schedule")
        if self.schedule is not None:
            t = self.schedule.get("type")
            if t not in ("ExponentialDecay", "InverseTimeDecay", "PolynomialDecay", "PiecewiseConstantDecay"):
                raise RuntimeError("IGNNITION: learning-rate schedule %s is not built" % t)

    def __call__(self, step: int) -> float:
        s = self.schedule
        if s is None:
            return self.const
        if s["type"] == "PiecewiseConstantDecay":
            k = sum(1 for b in s["boundaries"] if step > float(b))       # values[i] while step <= boundaries[i]
            return float(s["values"][k])
        if s["type"] == "PolynomialDecay":
            lr0, ds = float(s["initial_learning_rate"]), float(s["decay_steps"])
            end, power = float(s.get("end_learning_rate", 0.0001)), float(s.get("power", 1.0))
            if s.get("cycle"):
                ds = ds * max(1.0, math.ceil(step / ds))
            return (lr0 - end) * (1.0 - min(step, ds) / ds) ** power + end
        lr0, ds, dr = float(s["initial_learning_rate"]), float(s["decay_steps"]), float(s["decay_rate"])
        p = step / ds
        if s.get("staircase"):            # any truthy value, e.g. the string "True" of the Q-size example
            p = math.floor(p)
        if s["type"] == "ExponentialDecay":
            return lr0 * dr ** p
        return lr0 / (1.0 + dr * p)


def _sorted_csr(graph, dst, src, num_dst: int, want_perm: bool):
    """(rowptr, col, perm) of the stable sort by ``dst``: one launch for small graphs (ign_csr_build_small), the radix
    sort otherwise."""
    if getattr(graph, "small_rows", False) and num_dst <= 11000:
        return ops.csr_build_small([(dst, src, None, num_dst, want_perm)])[0]
    return ops.csr_build(dst, src, None, num_dst, ops.CSR_SORT, want_perm=want_perm)[:3]


class Trainer:
    """Owns gradients, Adam moments and the step counter of one Engine."""

    def __init__(self, engine, world_size: int = 1, process_group=None):
        self.e = engine
        self.world = world_size
        self.pg = process_group
        opt = engine.model.get_optimizer()
        kind = opt.get("type", "Adam")
        if kind != "Adam" and kind not in ops.OPTIMIZERS:
            raise RuntimeError("IGNNITION: optimizer %s is not built in the B200 engine (one of Adam, %s)"
                               % (kind, ", ".join(sorted(ops.OPTIMIZERS))))
        if getattr(engine, "has_dropout", False):
            raise RuntimeError("IGNNITION: training through Dropout layers (a random mask per step) is not built in the "
                               "B200 engine; inference treats them as the identity, as Keras does")
        loss_name = engine.model.get_loss()
        if loss_name not in ops.LOSSES:
            raise RuntimeError("IGNNITION: loss %s is not built in the B200 engine (one of %s)"
                               % (loss_name, ", ".join(sorted(ops.LOSSES))))
        self.opt_kind = kind
        self.opt = opt
        self.loss_kind = ops.LOSSES[loss_name]
        self.huber_delta = float(engine.model.get_loss_options().get("delta", 1.0)) \
            if hasattr(engine.model, "get_loss_options") else 1.0
        self.beta1 = float(opt.get("beta_1", 0.9))
        self.beta2 = float(opt.get("beta_2", 0.999))
        self.eps = float(opt.get("epsilon", 1e-7))
        self.lr = LearningRate(opt)
        n = engine.weights.numel()
        dev = engine.device
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.scalars = torch.zeros(4, dtype=torch.float64, device=dev)    # [sum of losses, reg, n_pred, unused]
        self.step = 0
        if kind == "Adagrad":                       # Keras: initial_accumulator_value = 0.1
            self.m.fill_(float(opt.get("initial_accumulator_value", 0.1)))

    def g(self, name: str) -> torch.Tensor:
        return self.e._pview(self.grads, name)

    # ------------------------------------------------------------------ transposed adjacencies
    def build_transposed(self, graph):
        e = self.e
        for stage in e.plans:
            for p in stage:
                if p.kind == "seq_gru":
                    rowptr_s, steps = graph.steps[p.key]
                    for k, a in enumerate(p.adjs):
                        key = "%s/%d" % (p.key, k)
                        if key in graph.csr_t:
                            continue
                        # rows a step entry can name: the source entity's, or one per edge when the walk reads the
                        # rows of a message network (entries are edge positions then)
                        n_rows = int(graph.t["src_" + a.name].numel()) if p.msg_src[k] else graph.num[a.src]
                        if p.concat2:
                            # concat along the features: row i of the walked array is [src_0[partner_0[i]] | src_1[...]],
                            # so the transposed view of source k groups the rows by partner_k (-1 = zero block)
                            keys = ops.steps_keys(graph.partner[p.key][k], 0, n_rows)
                        else:
                            keys = ops.steps_keys(steps, k, n_rows)
                        rp, _, perm = _sorted_csr(graph, keys, keys, n_rows + 1, True)
                        graph.csr_t[key] = (rp[:n_rows + 1], perm)
                        if p.msg_src[k] and a.name not in graph.csr_t:   # the message network's own backward (msg_ff)
                            rp2, col_t, perm_t = _sorted_csr(graph, graph.t["src_" + a.name], graph.t["dst_" + a.name],
                                                             graph.num[a.src], True)
                            graph.csr_t[a.name] = (rp2, col_t, perm_t)
                else:
                    for a in p.adjs:
                        if a.name in graph.csr_t:
                            continue
                        # perm (edge position per slot) only where per-edge rows are reduced: message networks
                        rp, col_t, perm_t = _sorted_csr(graph, graph.t["src_" + a.name], graph.t["dst_" + a.name],
                                                        graph.num[a.src], a.name in e._needs_perm)
                        graph.csr_t[a.name] = (rp, col_t, perm_t)

        for _, op in e.readout:
            if op.type == "extend_adjacencies" and op.adj_list not in graph.csr_t:
                a = [x for x in e.adjacencies if x.name == op.adj_list][0]
                rp, col_t, perm_t = _sorted_csr(graph, graph.t["src_" + a.name], graph.t["dst_" + a.name],
                                                graph.num[a.src], True)
                graph.csr_t[a.name] = (rp, col_t, perm_t)

    # ------------------------------------------------------------------ backward
    def backward(self, graph, tape: list, d_pred: torch.Tensor):
        e = self.e
        dev = e.device
        # dL/d(state) per entity, and per derived readout name (outputs of pooling / product / extend_adjacencies /
        # neural_network operations) while the readout is walked backwards
        gstate: Dict[str, Optional[torch.Tensor]] = {n: None for n in e.entities}

        def add_grad(ent: str, contrib: torch.Tensor):
            if gstate.get(ent) is None:
                gstate[ent] = contrib
            else:
                ops.axpy(1.0, contrib, gstate[ent])

        def reduce_into(ent: str, rowptr, idx, rows):
            """gstate[ent] (+)= per-row sums of `rows` over the segments of (rowptr, idx): the accumulation happens in
            the reduction kernel (IGN_OP_SUM_ADD), no temporary and no second pass"""
            if gstate.get(ent) is None:
                gstate[ent] = ops.segment_reduce(ops.OP_SUM, rowptr, idx, rows)
            else:
                ops.segment_reduce(ops.OP_SUM_ADD, rowptr, idx, rows, out=gstate[ent])

        pending: Dict[tuple, torch.Tensor] = {}          # (mp key, source) -> dL/d(per-edge message), input edge order

        def dense_chain_bwd(saves, dy):
            dz_ready = False          # dy already holds dZ of the layer (act' and bias gradient applied by the layer above)
            for j in range(len(saves) - 1, -1, -1):
                prefix, layer, x, out = saves[j]               # out = the layer's output (act' is taken from it)
                w = e.param("%s/%s/kernel" % (prefix, layer.name))
                dw = self.g("%s/%s/kernel" % (prefix, layer.name))
                db = self.g("%s/%s/bias" % (prefix, layer.name)) if layer.use_bias else None
                dx = torch.empty_like(x)
                act = ops.ACTIVATIONS["linear"] if dz_ready else e._act(layer.activation)
                below = saves[j - 1] if j > 0 else None
                if (below is not None and below[3] is x and act == 0
                        and ops.dense_head_bwd_chain_supported(x.shape[0], x.shape[1], w.shape[1])):
                    # linear single-output head: its dX is written as dZ of the layer below in the same streaming pass
                    if db is not None and not dz_ready:
                        ops.dense_bwd(x, w, 0, None, dy, None, None, db)         # bias gradient of the head itself
                    bl = below[1]
                    db_below = self.g("%s/%s/bias" % (below[0], bl.name)) if bl.use_bias else None
                    ops.dense_head_bwd_chain(x, w.reshape(-1), dy.reshape(-1), e._act(bl.activation), dx, dw.reshape(-1),
                                             db_below)
                    dz_ready = True
                else:
                    ops.dense_bwd(x, w, act | ops.ACT_FROM_OUTPUT, out, dy, dx, dw, None if dz_ready else db)
                    dz_ready = False
                dy = dx
            return dy

        def route_aggregate_grad(p, has_msg, d_agg, agg=None, max_src=None):
            """dL/d(aggregated messages per destination) -> source states, or -> per-edge messages of a message
            network.  mean: dL/d(sum) = dL/d(mean) / degree; max: the slots that attain the maximum share the gradient
            (TensorFlow's unsorted_segment_max), written per edge and reduced per source row."""
            if p.attn:
                # Attention_aggr (auxilary_classes.py:278-344): through the column softmax and the LeakyReLU to the two
                # score products, and through the weighted sum to the messages
                _, rows, idx, v1, v2, attn_ws, h_dst, max_len = max_src
                F = p.msg_dim
                multi = len(p.adjs) > 1
                if multi:        # one CSR over all sources' edges; rows = their per-edge messages, one list
                    rowptr, perm, slot_col, _ = graph.attn_comb[p.key]
                else:
                    rowptr, _, perm = graph.csr[p.adjs[0].name]
                    slot_col = None
                d_msg, d_pre4, d_ds = ops.attention_aggregate_bwd(rowptr, idx, perm, rows, d_agg,
                                                                  graph.t["offsets_" + p.dst], max_len, attn_ws,
                                                                  slot_col=slot_col)
                d_v1, d_v2 = torch.zeros_like(v1), torch.zeros_like(v2)
                tmp = torch.empty_like(h_dst)
                ops.dense_bwd(h_dst, v2, 0, None, d_ds, tmp, d_v2, None)        # dst_score = h v2
                add_grad(p.dst, tmp)
                if multi or has_msg[0]:                                         # rows = per-edge messages (edge order)
                    if rows.shape[0]:
                        tmp = torch.empty_like(rows)
                        ops.dense_bwd(rows, v1, 0, None, ops.slice_cols(d_pre4, 0, 1), tmp, d_v1, None)
                        ops.axpy(1.0, tmp, d_msg)
                    off = 0
                    for k, a in enumerate(p.adjs):
                        n_k = int(graph.t["dst_" + a.name].numel())
                        part = d_msg[off:off + n_k]
                        off += n_k
                        if has_msg[k]:
                            pending[(p.key, k)] = part
                        elif n_k:                                               # gathered source states: per source row
                            rp_t, _, perm_t = graph.csr_t[a.name]
                            add_grad(a.src, ops.segment_reduce(ops.OP_SUM, rp_t, perm_t, part))
                else:                                                           # rows = source states: reduce per source row
                    a = p.adjs[0]
                    rp_t, _, perm_t = graph.csr_t[a.name]
                    d_rows = ops.segment_reduce(ops.OP_SUM, rp_t, perm_t, d_msg)
                    d_ss = ops.slice_cols(ops.segment_reduce(ops.OP_SUM, rp_t, perm_t, d_pre4), 0, 1)
                    tmp = torch.empty_like(rows)
                    ops.dense_bwd(rows, v1, 0, None, d_ss, tmp, d_v1, None)     # src_score = rows v1
                    ops.axpy(1.0, tmp, d_rows)
                    add_grad(a.src, d_rows)
                ak, d_ak = e.param(p.dst + "_attention/attn_kernel"), self.g(p.dst + "_attention/attn_kernel")
                for kname, d_v, lo_, hi_ in (("kernel1", d_v1, 0, F), ("kernel2", d_v2, F, ak.shape[0])):
                    kmat = e.param("%s_attention/%s" % (p.dst, kname))          # v = kernel . attn_kernel[lo:hi]
                    dk = torch.empty_like(kmat)
                    ops.dense_bwd(kmat, ak[lo_:hi_], 0, None, d_v, dk, d_ak[lo_:hi_], None)
                    ops.axpy(1.0, dk, self.g("%s_attention/%s" % (p.dst, kname)))
                return
            if p.conv:
                # Conv_aggr (auxilary_classes.py:366-401): out = act((sum W + h) / deg).  dL/d(sum W + h) = act'(out) dL/dout
                # / deg goes to the destination's own state as it is, and through the kernel product to the neighbour sum
                _, agg_sum = max_src
                ck = e.param(p.dst + "_convolution/conv_kernel")
                ops.scale_rows_inv_degree(d_agg, graph.csr[p.adjs[0].name][0])
                d_sum = torch.empty_like(agg_sum)
                act = e._act(p.mp.aggregation.activation_function)
                ops.dense_bwd(agg_sum, ck, act | ops.ACT_FROM_OUTPUT, agg, d_agg, d_sum,
                              self.g(p.dst + "_convolution/conv_kernel"), None)     # d_agg becomes dZ in place
                add_grad(p.dst, d_agg)
                d_agg = d_sum
            elif p.op == ops.OP_MEAN:
                ops.scale_rows_inv_degree(d_agg, graph.csr[p.adjs[0].name][0])
            elif p.op == ops.OP_MAX:
                a = p.adjs[0]
                rowptr, _, perm = graph.csr[a.name]
                rows, idx = max_src
                d_msg = ops.segment_max_bwd(rowptr, idx, perm, rows, agg, d_agg, int(graph.t["dst_" + a.name].numel()))
                if has_msg[0]:
                    pending[(p.key, 0)] = d_msg                       # input edge order
                else:
                    rp_t, _, perm_t = graph.csr_t[a.name]
                    reduce_into(a.src, rp_t, perm_t, d_msg)
                return
            for k, a in enumerate(p.adjs):
                if has_msg[k]:       # every edge's message received d_agg of its destination
                    pending[(p.key, k)] = ops.gather_concat([d_agg], [graph.t["dst_" + a.name]],
                                                            int(graph.t["dst_" + a.name].numel()))
                else:
                    rp_t, col_t, _ = graph.csr_t[a.name]
                    reduce_into(a.src, rp_t, col_t, d_agg)

        for entry in reversed(tape):
            kind = entry[0]
            if kind == "msg_ff":
                _, p, k, inputs, widths, saves = entry
                dy = pending.pop((p.key, k), None)
                if dy is None:
                    continue
                a = p.adjs[k]
                dx = dense_chain_bwd(saves, dy)                   # [E, sum(widths)], input edge order
                off = 0
                for name, w_ in zip(inputs, widths):
                    if name == "hs_source":                       # rows gathered by src index: reduce per source row
                        rp_t, _, perm_t = graph.csr_t[a.name]
                        reduce_into(a.src, rp_t, perm_t, ops.slice_cols(dx, off, w_))
                    elif name == "hs_dest":                       # rows gathered by dst index: reduce per destination
                        rowptr, _, perm = graph.csr[a.name]
                        reduce_into(p.dst, rowptr, perm, ops.slice_cols(dx, off, w_))
                    off += w_                                     # edge_params are inputs, not variables
            elif kind == "agg_ff":
                _, p, has_msg, msg_dim, saves, agg, max_src = entry
                g_new = gstate[p.dst]
                if g_new is None:
                    continue
                dx = dense_chain_bwd(saves, g_new)                # [n, msg_dim + hidden]: concat([agg, h], 1)
                gstate[p.dst] = ops.slice_cols(dx, msg_dim, dx.shape[1] - msg_dim)
                route_aggregate_grad(p, has_msg, ops.slice_cols(dx, 0, msg_dim), agg, max_src)
            elif kind == "readout":
                _, op, saves, widths = entry
                dy = d_pred if op.type == "predict" else gstate.pop(op.output_name, None)
                if dy is None:                          # an intermediate network nothing downstream reads
                    continue
                dx = dense_chain_bwd(saves, dy)
                off = 0
                for name, w_ in zip(op.input, widths):  # inputs were concatenated along the features
                    add_grad(name, dx if len(widths) == 1 else ops.slice_cols(dx, off, w_))
                    off += w_
            elif kind == "product":                     # element-wise product (auxilary_classes.py:1072-1094)
                _, op, a_in, b_in = entry
                dy = gstate.pop(op.output_name, None)
                if dy is not None:
                    add_grad(op.input[0], ops.mul(dy, b_in))
                    add_grad(op.input[1], ops.mul(dy, a_in))
            elif kind == "outer":                       # dot_product = tf.tensordot(axes=0), per sample vec(a) b^T
                _, op, a_in, b_in, oa, ob = entry
                dy = gstate.pop(op.output_name, None)
                if dy is not None:
                    fa = int(a_in.shape[1])
                    da, db = torch.zeros_like(a_in), torch.zeros_like(b_in)
                    pos = 0
                    for i in range(len(oa) - 1):
                        na, nb = oa[i + 1] - oa[i], ob[i + 1] - ob[i]
                        if na * nb:
                            ops.dense_bwd(a_in[oa[i]:oa[i + 1]].view(na * fa, 1), b_in[ob[i]:ob[i + 1]].view(1, nb), 0, None,
                                          dy[pos:pos + na * fa * nb].view(na * fa, nb),
                                          da[oa[i]:oa[i + 1]].view(na * fa, 1), db[ob[i]:ob[i + 1]].view(1, nb), None)
                        pos += na * fa * nb
                    add_grad(op.input[0], da)
                    add_grad(op.input[1], db)
            elif kind == "pool":                        # per-sample pooling (auxilary_classes.py:1165-1185)
                _, op, red, offsets, x_in, pooled = entry
                dy = gstate.pop(op.output_name, None)
                if dy is not None:
                    n_rows = x_in.shape[0]
                    if red == ops.OP_MAX:               # ties share the gradient (tf.reduce_max)
                        ident = torch.arange(n_rows, dtype=torch.int32, device=dev)
                        add_grad(op.input[0], ops.segment_max_bwd(offsets, ident, None, x_in, pooled, dy, n_rows))
                    else:
                        if red == ops.OP_MEAN:
                            dy = ops.scale_rows_inv_degree(dy.clone(), offsets)
                        add_grad(op.input[0], ops.segment_broadcast(offsets, dy, n_rows))
            elif kind == "extend":                      # rows gathered by an adjacency (auxilary_classes.py:1236-1265)
                _, op = entry
                a = [x for x in e.adjacencies if x.name == op.adj_list][0]
                d_src = gstate.pop(op.output_name[0], None)
                d_dst = gstate.pop(op.output_name[1], None)
                if d_src is not None:
                    rp_t, _, perm_t = graph.csr_t[a.name]
                    reduce_into(op.input[0], rp_t, perm_t, d_src)
                if d_dst is not None:
                    rowptr, _, perm = graph.csr[a.name]
                    reduce_into(op.input[1], rowptr, perm, d_dst)
            elif kind == "seq_gru":
                _, p, src_states, h_old, h_seq, concat_widths = entry
                g_new = gstate[p.dst]
                if g_new is None:                       # state never reaches the loss
                    continue
                rowptr_s, steps = graph.steps[p.key]
                d_steps = torch.empty(steps.numel(), p.msg_dim, dtype=torch.float32, device=dev)
                dh0 = torch.empty_like(h_old)
                if p.key in graph.step_plan_bwd:
                    plan, max_steps = graph.step_plan_bwd[p.key]
                    ops.gru_seq_bwd_steps(plan, graph.meta[p.key], max_steps, src_states, h_old, h_seq,
                                          e.param(p.dst + "_update/kernel"), e.param(p.dst + "_update/recurrent_kernel"),
                                          e.param(p.dst + "_update/bias"), g_new, d_steps, dh0,
                                          self.g(p.dst + "_update/kernel"), self.g(p.dst + "_update/recurrent_kernel"),
                                          self.g(p.dst + "_update/bias"))
                else:
                    ops.gru_seq_bwd(rowptr_s, steps, graph.order.get(p.key), src_states, h_old, h_seq,
                                    e.param(p.dst + "_update/kernel"), e.param(p.dst + "_update/recurrent_kernel"),
                                    e.param(p.dst + "_update/bias"), g_new, d_steps, dh0,
                                    self.g(p.dst + "_update/kernel"), self.g(p.dst + "_update/recurrent_kernel"),
                                    self.g(p.dst + "_update/bias"))
                gstate[p.dst] = dh0
                col_off = 0
                for k, a in enumerate(p.adjs):
                    rp_t, perm_t = graph.csr_t["%s/%d" % (p.key, k)]
                    d_k = d_steps
                    if concat_widths is not None:     # concat along the features: source k owns a block of columns
                        d_k = ops.slice_cols(d_steps, col_off, concat_widths[k])
                        col_off += concat_widths[k]
                    if p.msg_src[k]:        # one step per message row: its gradient goes back into the message network
                        pending[(p.key, k)] = ops.segment_reduce(ops.OP_SUM, rp_t, perm_t, d_k)
                    else:
                        reduce_into(a.src, rp_t, perm_t, d_k)
            elif kind in ("agg_gru", "agg_gru_unfused"):
                _, p, has_msg, h_old, agg, max_src = entry
                g_new = gstate[p.dst]
                if g_new is None:
                    continue
                d_agg = torch.empty_like(agg)
                dh = torch.empty_like(h_old)
                n_dst = h_old.shape[0]
                if (ops.tensor_cores_enabled() and agg.shape[1] == 32 and h_old.shape[1] == 32 and n_dst >= 4096
                        and e.max_bwd_step_launches >= 1 and not p.v1):
                    # one GRU step = a sequence of length 1 whose message is the aggregate: the tensor-core BPTT kernels
                    # with the identity plan (destination i, step i), built once per batch and entity
                    key = "cell/" + p.dst
                    if key not in graph.step_plan_bwd:
                        rp = torch.arange(n_dst + 1, dtype=torch.int32, device=dev)
                        st = torch.arange(n_dst, dtype=torch.int32, device=dev)
                        meta = ops.seq_meta(rp, st, None)
                        graph.step_plan_bwd[key] = (ops.seq_step_plan(meta, st, 1), meta)
                    plan, meta = graph.step_plan_bwd[key]
                    ops.gru_seq_bwd_steps(plan, meta, 1, [agg], h_old, agg, e.param(p.dst + "_update/kernel"),
                                          e.param(p.dst + "_update/recurrent_kernel"), e.param(p.dst + "_update/bias"),
                                          g_new, d_agg, dh, self.g(p.dst + "_update/kernel"),
                                          self.g(p.dst + "_update/recurrent_kernel"), self.g(p.dst + "_update/bias"))
                else:
                    ops.gru_cell_bwd(agg, h_old, e.param(p.dst + "_update/kernel"),
                                     e.param(p.dst + "_update/recurrent_kernel"), e.param(p.dst + "_update/bias"),
                                     g_new, d_agg, dh, self.g(p.dst + "_update/kernel"),
                                     self.g(p.dst + "_update/recurrent_kernel"), self.g(p.dst + "_update/bias"))
                gstate[p.dst] = dh
                route_aggregate_grad(p, has_msg, d_agg, agg, max_src)
            else:
                raise RuntimeError("IGNNITION: training through '%s' is not built" % kind)

    # ------------------------------------------------------------------ one step
    def loss_and_grads(self, graph, global_n: Optional[int] = None):
        """Forward + backward.  Returns (pred, local sum of squared errors tensor[fp64], n_local)."""
        e = self.e
        if "labels" not in graph.t:
            raise RuntimeError("IGNNITION: the batch has no labels")
        self.build_transposed(graph)
        self.grads.zero_()
        self.scalars.zero_()
        tape: list = []
        pred = e.forward(graph, training=True, tape=tape)
        n_local = pred.numel()
        n_glob = global_n if global_n is not None else n_local * self.world
        d_pred = torch.empty_like(pred)
        if self.loss_kind == 0:
            ops.mse_loss(pred, graph.t["labels"], 1.0 / float(n_glob), d_pred, self.scalars[0:1])
        else:                                       # any other Keras loss by name: mean over all predictions (:745-751)
            ops.loss(self.loss_kind, pred, graph.t["labels"], 1.0 / float(n_glob), d_pred, self.scalars[0:1],
                     self.huber_delta)
        self.backward(graph, tape, d_pred)
        return pred, n_local

    def apply(self):
        e = self.e
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grads, group=self.pg)              # NCCL over NVLink: sum of rank gradients
            dist.all_reduce(self.scalars[0:1], group=self.pg)
        for name, lam in e._reg.items():                            # added once, identical on every rank
            ops.l2_reg(e.param(name), lam, self.g(name), self.scalars[1:2])
        self.step += 1
        lr = self.lr(self.step - 1)
        o = self.opt
        if self.opt_kind == "Adam":
            ops.adam_step(e.weights, self.grads, self.m, self.v, lr, self.beta1, self.beta2, self.eps, self.step)
        elif self.opt_kind == "SGD":
            ops.optimizer_step(ops.OPTIMIZERS["SGD"], e.weights, self.grads, self.m, self.v, lr,
                               float(o.get("momentum", 0.0)), 0.0, 0.0, 1 if o.get("nesterov") in (True, "True") else 0)
        elif self.opt_kind == "RMSprop":
            ops.optimizer_step(ops.OPTIMIZERS["RMSprop"], e.weights, self.grads, self.m, self.v, lr,
                               float(o.get("rho", 0.9)), float(o.get("momentum", 0.0)), self.eps)
        elif self.opt_kind == "Adagrad":
            ops.optimizer_step(ops.OPTIMIZERS["Adagrad"], e.weights, self.grads, self.m, self.v, lr, 0.0, 0.0, self.eps)
        else:                                       # Adamax
            ops.optimizer_step(ops.OPTIMIZERS["Adamax"], e.weights, self.grads, self.m, self.v,
                               lr / (1.0 - self.beta1 ** self.step), self.beta1, self.beta2, self.eps)

    def train_step(self, graph, global_n: Optional[int] = None):
        """One optimiser step on a prepared batch.  Returns (loss, regularisation) as device scalars
        in ``self.scalars`` -- read them with ``losses()`` (a host sync)."""
        pred, n_local = self.loss_and_grads(graph, global_n)
        self._n_glob = global_n if global_n is not None else n_local * self.world
        self.apply()
        return pred

    def train_step_graphed(self, batch, pinned=None, global_n: Optional[int] = None, copy: bool = True):
        """``train_step`` for the reference's own batch sizes (train_options.ini: batch_size 3 .. 32), where a step
        is ~150 launches of a few microseconds each: host -> device copy into a static staging buffer, then ONE captured
        CUDA graph per batch shape for adjacency build + transposes + forward + loss + backward.  The all-reduce, the
        regulariser and the optimiser update run after the replay (their step count and learning rate are host
        scalars that change every step).  Gradients are the same bits as ``train_step``'s kernels produce."""
        e = self.e
        if pinned is None:
            pinned = e.pack(batch)
        buf, layout = pinned
        n_local = int(batch.arrays["labels"].size)
        n_glob = global_n if global_n is not None else n_local * self.world
        sig = (batch.n_samples, tuple(sorted(batch.num.items())), tuple(sorted(batch.max_seq.items())),
               tuple(sorted(batch.dst_sorted.items())), n_glob,
               tuple((k, off, str(np.dtype(dt)), tuple(shape)) for k, (off, dt, shape) in layout.items()))
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        entry = self._graphs.get(sig)
        if entry is None:
            dev = e.device
            stage = torch.empty(buf.numel(), dtype=torch.uint8, device=dev)
            dg = e.upload(batch, pinned, out=stage)
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # eager pass first: one-time kernel attributes, workspace sizes
                e.build_graph(dg, training=True)
                self.loss_and_grads(dg, n_glob)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            l0 = self.e.gpu_launches
            with torch.cuda.graph(graph):
                dg.csr_t.clear()
                e.build_graph(dg, training=True)
                pred, _ = self.loss_and_grads(dg, n_glob)
            entry = (graph, stage, dg, pred, self.e.gpu_launches - l0)
            self._graphs[sig] = entry
            copy = True
        graph, stage, dg, pred, _ = entry
        if copy:
            stage.copy_(buf, non_blocking=True)
        graph.replay()
        self._n_glob = n_glob
        self.apply()
        return pred

    def step_batch(self, batch, global_n: Optional[int] = None, graph_rows: int = 8192, max_graphs: int = 32):
        """One optimiser step on a host batch, picking the path: small batches whose shape was seen twice before go
        through ``train_step_graphed`` (capture on the third occurrence, replay afterwards; at most ``max_graphs``
        shapes are kept), everything else through upload + build + ``train_step``."""
        e = self.e
        if graph_rows > 0 and max(batch.num.values()) <= graph_rows:
            pinned = e.pack(batch)
            key = (batch.n_samples, tuple(sorted(batch.num.items())), tuple(sorted(batch.max_seq.items())),
                   tuple(sorted(batch.dst_sorted.items())), global_n, pinned[0].numel())
            if not hasattr(self, "_shape_seen"):
                self._shape_seen, self._shape_captured = {}, set()
            self._shape_seen[key] = self._shape_seen.get(key, 0) + 1
            if key in self._shape_captured or (self._shape_seen[key] >= 3 and len(self._shape_captured) < max_graphs):
                self._shape_captured.add(key)
                return self.train_step_graphed(batch, pinned, global_n)
            return self.train_step(e.build_graph(e.upload(batch, pinned), training=True), global_n=global_n)
        return self.train_step(e.build_graph(e.upload(batch), training=True), global_n=global_n)

    def losses(self):
        s = self.scalars.cpu().numpy()
        mse = float(s[0]) / float(self._n_glob)
        return {"loss": mse, "regularization_loss": float(s[1]), "total_loss": mse + float(s[1])}
