"""CPU ORACLE, differentiable twin (test infrastructure, not product code).

The same op-for-op restatement of ``ComnetModel.call`` as ``oracle/ignnition_oracle.py`` (dense
right-padded message tensor, masked RNN, Keras GRU / Dense / SELU formulas), written on torch CPU
tensors so that ``torch.autograd`` gives the reference for ``tf.gradients(total_loss, variables)``
(``code/utils/generate_model.py:791``).  Covers what the BASELINE configs train (direct_assignation
messages, sum / ordered / interleave aggregation, GRU update, Dense readout) plus message neural networks on
``hs_source | hs_dest | edge_params`` (:440-475) and the feed-forward update (:594-600).  Cross-checked against
the NumPy oracle in ``tests/test_host.py``.  float parity vs TensorFlow itself is unpinned (see the
NumPy oracle's header).
"""

from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .ignnition_oracle import SELU_ALPHA, SELU_SCALE, Oracle


def _act(name, x):
    if name is None or name in ("None", "linear"):
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "selu":
        return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * (torch.exp(torch.clamp(x, max=0)) - 1))
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    raise ValueError("torch oracle: unsupported activation " + str(name))


def gru_cell(x, h, K, R, b):
    u = h.shape[1]
    if b.dim() == 1:                    # GRUCell(reset_after=False): one bias, r applied before the recurrent product
        mx = x @ K + b
        z = torch.sigmoid(mx[:, :u] + h @ R[:, :u])
        r = torch.sigmoid(mx[:, u:2 * u] + h @ R[:, u:2 * u])
        hh = torch.tanh(mx[:, 2 * u:] + (r * h) @ R[:, 2 * u:])
        return z * h + (1 - z) * hh
    mx = x @ K + b[0]
    mh = h @ R + b[1]
    z = torch.sigmoid(mx[:, :u] + mh[:, :u])
    r = torch.sigmoid(mx[:, u:2 * u] + mh[:, u:2 * u])
    hh = torch.tanh(mx[:, 2 * u:] + r * mh[:, 2 * u:])
    return z * h + (1 - z) * hh


def dense_stack(x, layers, prefix, w):
    """Keras functional Model of Dense layers (auxilary_classes.py:918-975) on torch tensors."""
    for l in layers:
        x = x @ w[prefix + "/" + l["name"] + "/kernel"]
        if (prefix + "/" + l["name"] + "/bias") in w:
            x = x + w[prefix + "/" + l["name"] + "/bias"]
        a = l.get("activation")
        x = _act(None if a == "None" else a, x)
    return x


class TorchOracle:
    def __init__(self, model_json: dict, dims: Dict[str, int], dtype=torch.float64):
        self.np_oracle = Oracle(model_json, dims, dtype=np.float64)
        self.m = self.np_oracle.m
        self.dtype = dtype

    def params(self, w: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
        return {k: torch.tensor(np.asarray(v, dtype=np.float64), dtype=self.dtype, requires_grad=True)
                for k, v in w.items()}

    def forward(self, inp: dict, w: Dict[str, torch.Tensor]) -> torch.Tensor:
        o = self.np_oracle
        dt = self.dtype
        state = {e["name"]: torch.tensor(o.initial_state(e, inp), dtype=dt) for e in self.m["entities"]}
        for _ in range(o.T):
            for st in self.m["message_passing"]["stages"]:
                for mp in st["stage_mp"]:
                    dst = mp["destination_entity"]
                    state[dst] = self._mp(mp, dst, state, inp, w)
        for k, op in enumerate(self.m["readout"]):                               # generate_model.py:607-656
            if op["type"] in ("predict", "neural_network"):
                x = torch.cat([state[i] for i in op["input"]], dim=1)
                for l in o.layer_names(op["nn_name"], "readout"):
                    x = x @ w["readout_model_%d/%s/kernel" % (k, l["name"])] + w["readout_model_%d/%s/bias" % (k, l["name"])]
                    a = l.get("activation")
                    x = _act(None if a == "None" else a, x)
                if op["type"] == "predict":
                    return x
                state[op["output_name"]] = x
            elif op["type"] == "pooling":                                        # auxilary_classes.py:1165-1185
                x = state[op["input"][0]]
                red = {"sum": lambda t: t.sum(dim=0), "mean": lambda t: t.mean(dim=0),
                       "max": lambda t: t.amax(dim=0)}[op["type_pooling"]]
                state[op["output_name"]] = red(x).reshape(1, -1)
            elif op["type"] == "product":                                        # auxilary_classes.py:1072-1088
                a_, b_ = state[op["input"][0]], state[op["input"][1]]
                state[op["output_name"]] = (torch.tensordot(a_, b_, dims=0) if op["type_product"] == "dot_product"
                                            else a_ * b_)      # tf.tensordot(axes=0): the outer product
            elif op["type"] == "extend_adjacencies":                             # auxilary_classes.py:1236-1265
                si = torch.as_tensor(np.asarray(inp["src_" + op["adj_list"]], dtype=np.int64))
                di = torch.as_tensor(np.asarray(inp["dst_" + op["adj_list"]], dtype=np.int64))
                state[op["output_name_src"]] = state[op["input"][0]][si]
                state[op["output_name_dst"]] = state[op["input"][1]][di]
        raise ValueError("torch oracle: no predict operation")

    def _mp(self, mp, dst, state, inp, w):
        dt = self.dtype
        num_dst = int(inp["num_" + dst])
        agg = mp["aggregation"]["type"]
        blocks, lens_all, idx_all, edge_lists = [], [], [], []
        for src in mp["source_entities"]:
            src_idx = torch.as_tensor(np.asarray(inp["src_" + src["adj_vector"]], dtype=np.int64))
            dst_idx = torch.as_tensor(np.asarray(inp["dst_" + src["adj_vector"]], dtype=np.int64))
            seq = torch.as_tensor(np.asarray(inp["seq_" + src["name"] + "_" + dst], dtype=np.int64))
            msgs = state[src["name"]][src_idx]                                   # tf.gather
            src_messages, dst_messages = msgs, state[dst][dst_idx]
            for k, op in enumerate(src["message"]):                              # generate_model.py:440-475
                if op["type"] != "neural_network":
                    continue
                parts = []
                for i in op["input"]:
                    if i == "hs_source":
                        parts.append(src_messages)
                    elif i == "hs_dest":
                        parts.append(dst_messages)
                    elif i == "edge_params":                                     # int64 then cast: truncation (:149)
                        if len(src_idx) == 0:                                    # no edges: no parameter rows
                            parts.append(torch.zeros(0, int(self.np_oracle.dims.get(src["adj_vector"], 0)), dtype=dt))
                            continue
                        pr = np.trunc(np.asarray(inp["params_" + src["adj_vector"]], dtype=np.float64))
                        parts.append(torch.tensor(pr.reshape(len(src_idx), -1), dtype=dt))
                    else:
                        raise ValueError("torch oracle: named message inputs are broken in the reference")
                msgs = dense_stack(torch.cat(parts, dim=1),
                                   self.np_oracle.layer_names(op["nn_name"], "message_creation_%d" % k),
                                   "%s_to_%s_message_creation_%d" % (src["name"], dst, k), w)
            max_len = int(seq.max()) + 1 if len(seq) else 0
            s = torch.zeros(num_dst, max_len, msgs.shape[1], dtype=dt).index_put((dst_idx, seq), msgs)  # scatter_nd
            blocks.append(s)
            lens_all.append(torch.bincount(dst_idx, minlength=num_dst))
            edge_lists.append((msgs, dst_idx, seq))
            if agg == "interleave":
                idx_all.append(np.asarray(inp["indices_" + src["name"] + "_to_" + dst], dtype=np.int64))
        if agg == "concat" and int(mp["aggregation"].get("concat_axis", 1)) == 2:   # generate_model.py:496-505
            src_input = torch.cat(blocks, dim=2)          # bigger messages, the first source's lengths
            final_len = lens_all[0]
        else:
            src_input = torch.cat(blocks, dim=1)
            final_len = sum(lens_all)
        h = state[dst]
        if mp["update"]["type"] != "recurrent_neural_network":                  # feed-forward update (:594-600)
            if agg != "sum":
                raise ValueError("torch oracle: feed-forward update is restated for the sum aggregation")
            ls = self.np_oracle.layer_names(mp["update"]["nn_name"], "update")
            return dense_stack(torch.cat([src_input.sum(dim=1), h], dim=1), ls, dst + "_ff_update", w)
        K, R, b = w[dst + "_update/kernel"], w[dst + "_update/recurrent_kernel"], w[dst + "_update/bias"]
        if agg == "sum":
            return gru_cell(src_input.sum(dim=1), h, K, R, b)
        if agg == "attention":              # Attention_aggr (auxilary_classes.py:278-344), as it is: softmax over the
            # DESTINATIONS per padded column, zero pads included; several sources = one edge list, the columns of source
            # k > 0 shifted by that source's own edge count per destination, colliding cells added by scatter_nd
            # (generate_model.py:523-543)
            comb = torch.cat([m for m, _, _ in edge_lists], dim=0)
            dst_idx = torch.cat([d for _, d, _ in edge_lists])
            seq = torch.cat([q if k == 0 else q + lens_all[k][d] for k, (_, d, q) in enumerate(edge_lists)])
            k1, k2, ak = w[dst + "_attention/kernel1"], w[dst + "_attention/kernel2"], w[dst + "_attention/attn_kernel"]
            a_in = torch.cat([comb @ k1, h[dst_idx] @ k2], dim=1) @ ak
            a_in = torch.where(a_in > 0, a_in, 0.2 * a_in)
            mx = int(seq.max()) + 1
            aux = torch.zeros(num_dst, mx, 1, dtype=dt).index_put((dst_idx, seq), a_in, accumulate=True)
            coef = torch.softmax(aux, dim=0)
            red = torch.zeros(num_dst, comb.shape[1], dtype=dt).index_add(0, dst_idx, comb * coef[dst_idx, seq])
            return gru_cell(red, h, K, R, b)
        if agg == "convolution":            # Conv_aggr (auxilary_classes.py:366-401), single source
            ck = w[dst + "_convolution/conv_kernel"]
            deg = final_len.to(dt)[:, None]
            pre = (src_input.sum(dim=1) @ ck + h) / deg
            a = mp["aggregation"].get("activation_function", "relu")
            return gru_cell(_act(None if a == "None" else a, pre), h, K, R, b)
        if agg in ("mean", "max"):          # north-star extensions: the definitions of ignnition_oracle.py, on torch
            valid = (torch.arange(src_input.shape[1])[None, :] < final_len[:, None])[:, :, None]
            if agg == "mean":
                red = src_input.sum(dim=1) / torch.clamp(final_len, min=1)[:, None].to(dt)
            elif src_input.shape[1] == 0:
                red = torch.zeros(num_dst, src_input.shape[2], dtype=dt)
            else:
                neg = torch.where(valid, src_input, torch.full_like(src_input, -float("inf")))
                red = torch.where((final_len > 0)[:, None], neg.amax(dim=1), torch.zeros((), dtype=dt))
            return gru_cell(red, h, K, R, b)
        if agg == "interleave":
            idx = torch.as_tensor(np.concatenate(idx_all))
            tr = src_input.transpose(0, 1)
            src_input = torch.zeros_like(tr).index_put((idx,), tr).transpose(0, 1)
        elif agg not in ("ordered", "concat"):
            raise ValueError("torch oracle: aggregation " + agg + " is not restated here")
        out_prev = torch.zeros_like(h)
        outs = []
        for t in range(src_input.shape[1]):                                      # masked K.rnn
            m = (t < final_len)[:, None]
            nh = gru_cell(src_input[:, t], h, K, R, b)
            out_prev = torch.where(m, nh, out_prev)
            h = torch.where(m, nh, h)
            outs.append(out_prev)
        outputs = torch.stack(outs, dim=1)
        return outputs[torch.arange(num_dst), final_len - 1]                     # gather_nd

    def loss_and_grads(self, samples: List[dict], labels: List[np.ndarray], w_np: Dict[str, np.ndarray]):
        """model_fn: MSE over all predictions of all samples + sum of l2 regularisers; gradients."""
        w = self.params(w_np)
        preds = torch.cat([self.forward(s, w).reshape(-1) for s in samples])
        y = torch.tensor(np.concatenate([np.asarray(l, dtype=np.float64).reshape(-1) for l in labels]), dtype=self.dtype)
        mse = torch.mean((y - preds) ** 2)
        reg = torch.zeros((), dtype=self.dtype)
        for prefix, layers in self.np_oracle._regularized_layers():
            for l in layers:
                lam = float(l.get("kernel_regularizer", 0.0) or 0.0)
                if lam:
                    reg = reg + lam * (w[prefix + "/" + l["name"] + "/kernel"] ** 2).sum()
        total = mse + reg
        total.backward()
        grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(v.shape)) for k, v in w.items()}
        return float(mse.detach()), float(reg.detach()), preds.detach().numpy(), grads
