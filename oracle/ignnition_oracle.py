"""CPU ORACLE (test infrastructure, not product code).

NumPy restatement, op for op, of the model IGNNITION generates -- ``ComnetModel.call``
(reference ``code/utils/generate_model.py:384-658``) and the compute methods of its descriptor
classes (``code/utils/auxilary_classes.py``) -- INCLUDING the dense right-padded
``[num_dst, max_len, F]`` message tensor the reference builds with ``tf.scatter_nd``
(generate_model.py:479-490), so that it is visibly the same program.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the
product (``ignnition_b200``) never does.

PARITY PINNING: the reference ships no tests, golden vectors or dataset fixtures
and its float half needs tensorflow==2.1.0 / Keras==2.4.1, which cannot be installed here.
 * integer half (src/dst/seq/indices arrays): PINNED -- ``tests/golden/*.json`` were produced by
   the reference's own ``generator_std_to_framework.generator`` imported from /root/reference
   under a stub ``tensorflow`` module (``oracle/make_golden.py``), and this file's consumers are
   checked bit-for-bit against them.
 * float half: **parity unpinned** against TensorFlow.  The Keras formulas restated below
   (GRUCell v2 ``reset_after=True``, masked ``K.rnn``, Dense, SELU, l2, MSE, Adam,
   ExponentialDecay) are written from the published TF-2.1 semantics and cross-checked against
   ``torch.nn.GRUCell`` / ``torch.nn.functional.selu`` (``tests/test_oracle.py``) and an fp64
   shadow run of the same code.

The oracle interprets the raw ``model_description.json`` dict itself (it does not share the
product's parser) and takes weights as a ``{name: ndarray}`` dict in Keras layout:
  ``<dst>_update/kernel [in,3u]``, ``/recurrent_kernel [u,3u]``, ``/bias [2,3u]`` (gates z|r|h),
  ``<src>_to_<dst>_message_creation_<k>/<layer>/kernel [in,out]``, ``/bias [out]``,
  ``<dst>_ff_update/<layer>/...``, ``readout_model_<k>/<layer>/...``.
"""

from __future__ import annotations

import copy
import math
from typing import Dict, List, Optional

import numpy as np

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


# ----------------------------------------------------------------------------- activations
def sigmoid(x):
    e = np.exp(-np.abs(x))                     # overflow-free form of 1 / (1 + exp(-x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e)).astype(x.dtype)


# The user normalisation functions of the two examples, restated on NumPy
# (examples/Routenet/main.py:26-38, examples/Q-size/main.py:27-39).
def normalization_routenet(feature, feature_name):
    if feature_name == 'traffic':
        feature = (feature - 170) / 130
    if feature_name == 'link_capacity':
        feature = (feature - 25000) / 40000
    return feature


def normalization_queue_size(feature, feature_name):
    if feature_name == 'delay':
        feature = (np.log(feature) + 1.78) / 0.93
    if feature_name == 'traffic':
        feature = (feature - 0.28) / 0.15
    if feature_name == 'jitter':
        feature = (feature - 1.5) / 1.5
    if feature_name == 'link_capacity':
        feature = (feature - 27.0) / 14.86
    if feature_name == 'queue_sizes':
        feature = (feature - 16.5) / 15.5
    return feature


EXAMPLE_NORMALIZATIONS = {
    "normalization_routenet": normalization_routenet,
    "normalization_queue_size": normalization_queue_size,
    "log": lambda f, n: np.log(f),
    "exp": lambda f, n: np.exp(f),
}


def normalize_inputs(model_json: dict, tensors: dict, fns=None, dtype=np.float32) -> dict:
    """input_fn's ``normalization`` map (generate_model.py:46-86): features are float32 tensors."""
    fns = fns or EXAMPLE_NORMALIZATIONS
    out = dict(tensors)
    for e in model_json["entities"]:
        for f in e.get("features", []):
            x = np.asarray(out[f["name"]], dtype=dtype)
            norm = str(f.get("normalization", "None"))
            if norm != "None":
                x = np.asarray(fns[norm](x, f["name"]), dtype=dtype)
            out[f["name"]] = x
    return out


def activation(name: Optional[str], x):
    """tf.keras.activations.<name> for the names the examples / schema use."""
    if name is None or name in ("None", "linear"):
        return x
    if name == "relu":
        return np.maximum(x, 0)
    if name == "selu":
        return (SELU_SCALE * np.where(x > 0, x, SELU_ALPHA * (np.exp(np.minimum(x, 0)) - 1))).astype(x.dtype)
    if name == "sigmoid":
        return sigmoid(x)
    if name == "tanh":
        return np.tanh(x)
    if name == "elu":
        return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1).astype(x.dtype)
    if name == "softplus":
        return np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0)
    if name == "leaky_relu":
        return np.where(x > 0, x, 0.2 * x).astype(x.dtype)
    raise ValueError("oracle: unsupported activation " + str(name))


# ----------------------------------------------------------------------------- Keras cells
def gru_cell(x, h, kernel, recurrent_kernel, bias, reset_after=True):
    """One step of tf.keras.layers.GRUCell (TF 2.1 v2 defaults: implementation=2,
    reset_after=True, gates z|r|h, bias [2,3u]).  Called at auxilary_classes.py:764 and,
    through keras.layers.RNN, at :785-790.  [TF-2.1 semantics, restated]"""
    u = h.shape[1]
    if reset_after:
        mx = x @ kernel + bias[0]
        mh = h @ recurrent_kernel + bias[1]
        z = sigmoid(mx[:, :u] + mh[:, :u])
        r = sigmoid(mx[:, u:2 * u] + mh[:, u:2 * u])
        hh = np.tanh(mx[:, 2 * u:] + r * mh[:, 2 * u:])
    else:
        mx = x @ kernel + bias
        z = sigmoid(mx[:, :u] + h @ recurrent_kernel[:, :u])
        r = sigmoid(mx[:, u:2 * u] + h @ recurrent_kernel[:, u:2 * u])
        hh = np.tanh(mx[:, 2 * u:] + (r * h) @ recurrent_kernel[:, 2 * u:])
    return z * h + (1 - z) * hh


def masked_rnn_last(cell, inputs, initial_state, final_len):
    """keras.layers.RNN(cell, return_sequences=True)(inputs, initial_state,
    mask=sequence_mask(final_len)) followed by gather_nd(outputs, [i, final_len[i]-1])
    (auxilary_classes.py:785-796).  Masked steps carry state and repeat the previous output
    (K.rnn).  [TF-2.1 semantics, restated]"""
    n, max_len, _ = inputs.shape
    if np.any(final_len <= 0):
        raise ValueError("oracle: ordered aggregation with an empty destination "
                         "(gather_nd index -1 in the reference)")
    if np.any(final_len > max_len):
        raise ValueError("oracle: final_len exceeds the padded length")
    h = initial_state
    prev_out = np.zeros_like(initial_state)
    outputs = np.zeros((n, max_len, initial_state.shape[1]), dtype=initial_state.dtype)
    for t in range(max_len):
        m = (t < final_len)[:, None]
        new_h = cell(inputs[:, t, :], h)
        out = np.where(m, new_h, prev_out)
        h = np.where(m, new_h, h)
        outputs[:, t, :] = out
        prev_out = out
    return outputs[np.arange(n), final_len - 1]


def dense_stack(x, layers: List[dict], prefix: str, w: Dict[str, np.ndarray], last_units=None):
    """Keras functional Model of Dense layers (auxilary_classes.py:918-975)."""
    for j, l in enumerate(layers):
        if l["type_layer"] != "Dense":
            raise ValueError("oracle: only Dense layers are restated, got " + l["type_layer"])
        name = l.get("name")
        k = w[prefix + "/" + name + "/kernel"]
        x = x @ k
        if (prefix + "/" + name + "/bias") in w:
            x = x + w[prefix + "/" + name + "/bias"]
        act = l.get("activation")
        x = activation(None if act == "None" else act, x)
    return x


# ----------------------------------------------------------------------------- integer side
def csr_from_edges(src_idx, dst_idx, seq, num_dst):
    """What the device CSR builder must produce from the reference's flat arrays
    (generator_std_to_framework.py:145-181): edges grouped by destination, ``seq`` ascending
    inside each group.  rowptr = exclusive scan of unsorted_segment_sum(1, dst) (generate_model.py
    :481); col[rowptr[d]+seq] = src; perm = original edge position."""
    src_idx = np.asarray(src_idx, dtype=np.int64)
    dst_idx = np.asarray(dst_idx, dtype=np.int64)
    seq = np.asarray(seq, dtype=np.int64)
    lens = np.bincount(dst_idx, minlength=int(num_dst)).astype(np.int64)
    rowptr = np.zeros(int(num_dst) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    pos = rowptr[dst_idx] + seq
    col = np.full(len(src_idx), -1, dtype=np.int64)
    perm = np.full(len(src_idx), -1, dtype=np.int64)
    col[pos] = src_idx
    perm[pos] = np.arange(len(src_idx))
    return rowptr, col, perm


def stable_sort_csr(src_idx, dst_idx, num_dst):
    """CSR of a plain edge list with no seq: stable sort by destination (seq := rank in input order)."""
    dst_idx = np.asarray(dst_idx, dtype=np.int64)
    perm = np.argsort(dst_idx, kind="stable")
    lens = np.bincount(dst_idx, minlength=int(num_dst))
    rowptr = np.zeros(int(num_dst) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    return rowptr, np.asarray(src_idx, dtype=np.int64)[perm], perm


# ----------------------------------------------------------------------------- the model
class Oracle:
    """Interprets one model_description dict; ``forward`` restates ComnetModel.call."""

    def __init__(self, model_json: dict, dimensions: Optional[Dict[str, int]] = None,
                 dtype=np.float32):
        self.m = copy.deepcopy(model_json)
        self.dims = dict(dimensions or {})
        self.dtype = dtype
        self.nn = {n["nn_name"]: n for n in self.m["neural_networks"]}
        self.hs = {e["name"]: int(e["hidden_state_dimension"]) for e in self.m["entities"]}
        self.T = int(self.m["message_passing"]["num_iterations"])

    # --- weight shapes (ComnetModel.__init__, generate_model.py:235-382)
    def layer_names(self, nn_name, role):
        out = []
        for i, l in enumerate(self.nn[nn_name]["nn_architecture"]):
            l = dict(l)
            if l["type_layer"] == "Dropout":       # tf.keras.layers.Dropout: the identity outside training
                continue
            l.setdefault("name", "layer_%d_%s_%s" % (i, l["type_layer"], role))
            out.append(l)
        return out

    def message_dim(self, src):
        d = self.hs[src["name"]]
        for op in src["message"]:
            if op["type"] == "neural_network":
                d = int(self.layer_names(op["nn_name"], "x")[-1]["units"])
        return d

    def weight_shapes(self) -> Dict[str, tuple]:
        shapes: Dict[str, tuple] = {}
        for st in self.m["message_passing"]["stages"]:
            for mp in st["stage_mp"]:
                dst = mp["destination_entity"]
                fd = self.hs[dst]
                msg_dims = []
                for src in mp["source_entities"]:
                    for k, op in enumerate(src["message"]):
                        if op["type"] != "neural_network":
                            continue
                        din = 0
                        for i in op["input"]:
                            din += {"hs_source": self.hs[src["name"]], "hs_dest": fd,
                                    "edge_params": int(self.dims.get(src["adj_vector"], 0))}[i]
                        pre = "%s_to_%s_message_creation_%d" % (src["name"], dst, k)
                        for l in self.layer_names(op["nn_name"], "message_creation_%d" % k):
                            shapes[pre + "/" + l["name"] + "/kernel"] = (din, int(l["units"]))
                            shapes[pre + "/" + l["name"] + "/bias"] = (int(l["units"]),)
                            din = int(l["units"])
                    msg_dims.append(self.message_dim(src))
                up = mp["update"]
                agg = mp["aggregation"]
                fin = msg_dims[0]
                if agg["type"] == "concat" and int(agg.get("concat_axis", 1)) == 2:
                    fin = sum(msg_dims)
                if up["type"] == "recurrent_neural_network":
                    shapes[dst + "_update/kernel"] = (fin, 3 * fd)
                    shapes[dst + "_update/recurrent_kernel"] = (fd, 3 * fd)
                    ra = self.nn[up["nn_name"]].get("reset_after", True)       # GRUCell kwargs come from the JSON
                    shapes[dst + "_update/bias"] = (2, 3 * fd) if str(ra) in ("True", "true", "1") else (3 * fd,)
                else:
                    din = fin + fd
                    ls = self.layer_names(up["nn_name"], "update")
                    for j, l in enumerate(ls):
                        units = fd if j == len(ls) - 1 else int(l["units"])   # auxilary_classes.py:852-865
                        shapes[dst + "_ff_update/" + l["name"] + "/kernel"] = (din, units)
                        shapes[dst + "_ff_update/" + l["name"] + "/bias"] = (units,)
                        din = units
                if agg["type"] == "attention":
                    shapes[dst + "_attention/kernel1"] = (fin, fin)
                    shapes[dst + "_attention/kernel2"] = (fd, fin)
                    shapes[dst + "_attention/attn_kernel"] = (2 * fd, 1)
                if agg["type"] == "convolution":
                    shapes[dst + "_convolution/conv_kernel"] = (fd, fd)
        for k, op in enumerate(self.m["readout"]):
            if op["type"] in ("predict", "neural_network"):
                din = sum(int(self.hs.get(i, self.dims.get(i, 0))) for i in op["input"])
                for l in self.layer_names(op["nn_name"], "readout"):
                    shapes["readout_model_%d/%s/kernel" % (k, l["name"])] = (din, int(l["units"]))
                    shapes["readout_model_%d/%s/bias" % (k, l["name"])] = (int(l["units"]),)
                    din = int(l["units"])
                if op["type"] == "neural_network":
                    self.hs[op["output_name"]] = din
            elif op["type"] == "pooling":
                self.hs[op["output_name"]] = int(self.hs[op["input"][0]])
            elif op["type"] == "product" and op["type_product"] == "element_wise":
                self.hs[op["output_name"]] = int(self.hs[op["input"][0]])
            elif op["type"] == "product" and op["type_product"] == "dot_product":     # generate_model.py:375-376
                self.hs[op["output_name"]] = 1
            elif op["type"] == "extend_adjacencies":
                self.hs[op["output_name_src"]] = int(self.hs[op["input"][0]])
                self.hs[op["output_name_dst"]] = int(self.hs[op["input"][1]])
        return shapes

    def init_weights(self, seed=1234) -> Dict[str, np.ndarray]:
        """Seeded test weights: glorot-uniform kernels, U(-0.1, 0.1) biases."""
        rng = np.random.RandomState(seed)
        w = {}
        for name, shp in self.weight_shapes().items():
            if name.endswith("bias"):
                w[name] = rng.uniform(-0.1, 0.1, shp).astype(self.dtype)
            else:
                lim = math.sqrt(6.0 / (shp[0] + shp[-1]))
                w[name] = rng.uniform(-lim, lim, shp).astype(self.dtype)
        return w

    # --- Entity.calculate_hs (auxilary_classes.py:128-160)
    def initial_state(self, ent: dict, inp: dict):
        n = int(inp["num_" + ent["name"]])
        parts, total = [], 0
        for f in ent.get("features", []):
            size = int(self.dims.get(f["name"], f.get("size", 1)) or 1)
            parts.append(np.asarray(inp[f["name"]], dtype=self.dtype).reshape(n, size))
            total += size
        parts.append(np.zeros((n, int(ent["hidden_state_dimension"]) - total), dtype=self.dtype))
        return np.concatenate(parts, axis=1)

    # --- ComnetModel.call (generate_model.py:384-658)
    def forward(self, inp: dict, w: Dict[str, np.ndarray], return_states: bool = False,
                iterations: Optional[int] = None):
        dt = self.dtype
        w = {k: np.asarray(v, dtype=dt) for k, v in w.items()}
        state = {e["name"]: self.initial_state(e, inp) for e in self.m["entities"]}
        T = self.T if iterations is None else iterations
        for _ in range(T):
            for st in self.m["message_passing"]["stages"]:
                for mp in st["stage_mp"]:
                    dst = mp["destination_entity"]
                    state[dst] = self._message_passing(mp, dst, state, inp, w)
        result = None
        for k, op in enumerate(self.m["readout"]):
            if op["type"] in ("predict", "neural_network"):
                x = np.concatenate([state[i] if i in state else np.asarray(inp[i], dtype=dt)
                                    for i in op["input"]], axis=1)
                y = dense_stack(x, self.layer_names(op["nn_name"], "readout"),
                                "readout_model_%d" % k, w)
                if op["type"] == "predict":
                    result = y
                    break
                state[op["output_name"]] = y
            elif op["type"] == "pooling":           # auxilary_classes.py:1165-1185
                x = state[op["input"][0]]
                red = {"sum": np.sum, "mean": np.mean, "max": np.max}[op["type_pooling"]]
                state[op["output_name"]] = red(x, axis=0).reshape(1, -1)
            elif op["type"] == "product":           # auxilary_classes.py:1072-1088
                a, b = state[op["input"][0]], state[op["input"][1]]
                state[op["output_name"]] = (np.tensordot(a, b, axes=0)
                                            if op["type_product"] == "dot_product" else a * b)
            elif op["type"] == "extend_adjacencies":  # auxilary_classes.py:1236-1265
                state[op["output_name_src"]] = state[op["input"][0]][np.asarray(inp["src_" + op["adj_list"]])]
                state[op["output_name_dst"]] = state[op["input"][1]][np.asarray(inp["dst_" + op["adj_list"]])]
        if return_states:
            return result, state
        return result

    def _message_passing(self, mp, dst, state, inp, w):
        dt = self.dtype
        dst_states = state[dst]
        num_dst = int(inp["num_" + dst])
        agg = mp["aggregation"]
        first = True
        src_input = final_len = indices = None
        comb_src = comb_dst = comb_seq = None
        for src in mp["source_entities"]:
            sname = src["name"]
            src_idx = np.asarray(inp["src_" + src["adj_vector"]], dtype=np.int64)
            dst_idx = np.asarray(inp["dst_" + src["adj_vector"]], dtype=np.int64)
            seq = np.asarray(inp["seq_" + sname + "_" + dst], dtype=np.int64)
            src_messages = state[sname][src_idx]                     # tf.gather :432
            dst_messages = dst_states[dst_idx]                       # tf.gather :433
            final_messages = src_messages
            for k, op in enumerate(src["message"]):                  # :440-475
                if op["type"] != "neural_network":
                    continue
                parts = []
                for i in op["input"]:
                    if i == "hs_source":
                        parts.append(src_messages)
                    elif i == "hs_dest":
                        parts.append(dst_messages)
                    elif i == "edge_params":
                        # declared tf.int64 then cast (generate_model.py:149, :454-456): truncation
                        if len(src_idx) == 0:            # a sample without edges carries no parameter rows
                            parts.append(np.zeros((0, int(self.dims.get(src["adj_vector"], 0))), dtype=dt))
                            continue
                        p = np.asarray(inp["params_" + src["adj_vector"]])
                        parts.append(np.trunc(p).astype(dt).reshape(len(src_idx), -1))
                    else:
                        raise ValueError("oracle: named message inputs are broken in the reference (quirk 2)")
                x = np.concatenate(parts, axis=1)
                final_messages = dense_stack(
                    x, self.layer_names(op["nn_name"], "message_creation_%d" % k),
                    "%s_to_%s_message_creation_%d" % (sname, dst, k), w)
            lens = np.bincount(dst_idx, minlength=num_dst).astype(np.int64)     # :481
            max_len = int(seq.max()) + 1 if len(seq) else 0                       # :484
            s = np.zeros((num_dst, max_len, final_messages.shape[1]), dtype=dt)   # scatter_nd :490
            s[dst_idx, seq] = final_messages          # (dst, seq) pairs are unique: no duplicate-sum
            if agg["type"] == "concat":                                           # :496-505
                if first:
                    src_input, final_len, first = s, lens.copy(), False
                else:
                    ax = int(agg["concat_axis"])
                    src_input = np.concatenate([src_input, s], axis=ax)
                    if ax == 1:
                        final_len = final_len + lens
            elif agg["type"] == "interleave":                                     # :507-519
                idx_src = np.asarray(inp["indices_" + sname + "_to_" + dst], dtype=np.int64)
                if first:
                    src_input, indices, final_len, first = s, idx_src, lens.copy(), False
                else:
                    src_input = np.concatenate([src_input, s], axis=1)
                    indices = np.concatenate([indices.reshape(-1), idx_src])      # stack+reshape(-1)
                    final_len = final_len + lens
            else:                                                                 # :523-543
                if first:
                    src_input, final_len, first = s, lens.copy(), False
                    comb_src, comb_dst, comb_seq = final_messages, dst_idx, seq
                else:
                    src_input = np.concatenate([src_input, s], axis=1)
                    comb_src = np.concatenate([comb_src, final_messages], axis=0)
                    comb_dst = np.concatenate([comb_dst, dst_idx])
                    comb_seq = np.concatenate([comb_seq, seq + lens[dst_idx]])    # quirk 7
                    final_len = final_len + lens
        # aggregation :552-569
        t = agg["type"]
        if t == "sum":
            src_input = src_input.sum(axis=1)                                     # auxilary_classes.py:261
        elif t in ("mean", "max"):
            # north-star extensions, no reference counterpart: defined here (SURVEY section 8a note)
            valid = np.arange(src_input.shape[1])[None, :] < final_len[:, None]
            if t == "mean":
                src_input = src_input.sum(axis=1) / np.maximum(final_len, 1)[:, None].astype(dt)
            else:
                neg = np.where(valid[:, :, None], src_input, -np.inf)
                src_input = np.where(final_len[:, None] > 0, neg.max(axis=1), 0).astype(dt)
        elif t == "interleave":                                                   # auxilary_classes.py:421-440
            tr = np.transpose(src_input, (1, 0, 2))
            out = np.zeros_like(tr)
            out[indices.reshape(-1)] = tr             # scatter_nd with unique indices
            src_input = np.transpose(out, (1, 0, 2))
        elif t == "attention":                                                    # auxilary_classes.py:278-344
            k1, k2, ak = (w[dst + "_attention/kernel1"], w[dst + "_attention/kernel2"],
                          w[dst + "_attention/attn_kernel"])
            a_in = np.concatenate([comb_src @ k1, dst_states[comb_dst] @ k2], axis=1) @ ak
            a_in = np.where(a_in > 0, a_in, 0.2 * a_in)
            mx = int(comb_seq.max()) + 1
            aux = np.zeros((num_dst, mx, 1), dtype=dt)
            np.add.at(aux, (comb_dst, comb_seq), a_in)
            e = np.exp(aux - aux.max(axis=0, keepdims=True))                      # softmax over axis 0 (sic)
            coef = e / e.sum(axis=0, keepdims=True)
            weighted = comb_src * coef[comb_dst, comb_seq]
            src_input = np.zeros((num_dst, comb_src.shape[1]), dtype=dt)
            np.add.at(src_input, comb_dst, weighted)
        elif t == "convolution":                                                  # auxilary_classes.py:366-401
            ck = w[dst + "_convolution/conv_kernel"]
            nsum = np.zeros((num_dst, ck.shape[1]), dtype=dt)
            np.add.at(nsum, comb_dst, comb_src @ ck)
            deg = np.bincount(comb_dst, minlength=num_dst).astype(dt)
            with np.errstate(divide="ignore", invalid="ignore"):
                src_input = activation(agg.get("activation_function", "relu"),
                                       (nsum + dst_states) / deg[:, None])
        # update :573-600
        up = mp["update"]
        if up["type"] == "recurrent_neural_network":
            arch = self.nn[up["nn_name"]]
            if arch["recurrent_type"] != "GRU":
                raise ValueError("oracle: only GRU is restated (LSTM cannot run in the reference: "
                                 "single-tensor state, auxilary_classes.py:764)")
            ra = str(arch.get("reset_after", True)) in ("True", "true", "1")
            K, R, b = (w[dst + "_update/kernel"], w[dst + "_update/recurrent_kernel"],
                       w[dst + "_update/bias"])
            cell = lambda x, h: gru_cell(x, h, K, R, b, ra)
            if t in ("sum", "attention", "convolution", "mean", "max"):
                return cell(src_input, dst_states)                               # perform_unsorted_update
            return masked_rnn_last(cell, src_input, dst_states, final_len)       # perform_sorted_update
        # feed-forward update (semantics of call :594-600; __init__ crashes in the reference, quirk 1)
        ls = self.layer_names(up["nn_name"], "update")
        ls[-1]["units"] = self.hs[dst]
        return dense_stack(np.concatenate([src_input, dst_states], axis=1), ls, dst + "_ff_update", w)

    # --- model_fn (generate_model.py:697-830)
    def regularization(self, w):
        """sum(model.losses): l2(lambda) = lambda * sum(w^2) per Dense kernel  [TF-2.1]."""
        total = 0.0
        for nn_prefix, layers in self._regularized_layers():
            for l in layers:
                lam = float(l.get("kernel_regularizer", 0.0) or 0.0)
                if lam:
                    k = np.asarray(w[nn_prefix + "/" + l["name"] + "/kernel"], dtype=np.float64)
                    total += lam * float((k * k).sum())
        return total

    def _regularized_layers(self):
        out = []
        for st in self.m["message_passing"]["stages"]:
            for mp in st["stage_mp"]:
                dst = mp["destination_entity"]
                for src in mp["source_entities"]:
                    for k, op in enumerate(src["message"]):
                        if op["type"] == "neural_network":
                            out.append(("%s_to_%s_message_creation_%d" % (src["name"], dst, k),
                                        self.layer_names(op["nn_name"], "message_creation_%d" % k)))
                if mp["update"]["type"] == "neural_network":
                    out.append((dst + "_ff_update", self.layer_names(mp["update"]["nn_name"], "update")))
        for k, op in enumerate(self.m["readout"]):
            if op["type"] in ("predict", "neural_network"):
                out.append(("readout_model_%d" % k, self.layer_names(op["nn_name"], "readout")))
        return out

    def loss(self, samples: List[dict], labels: List[np.ndarray], w):
        """MeanSquaredError over the concatenated predictions of all samples + regularisation."""
        preds = np.concatenate([self.forward(s, w).reshape(-1) for s in samples])
        y = np.concatenate([np.asarray(l, dtype=self.dtype).reshape(-1) for l in labels])
        mse = float(np.mean((y.astype(np.float64) - preds.astype(np.float64)) ** 2))
        reg = self.regularization(w)
        return mse, reg, preds


def exponential_decay(step, initial_learning_rate, decay_steps, decay_rate, staircase=False):
    """tf.keras.optimizers.schedules.ExponentialDecay  [TF-2.1]; any truthy ``staircase`` floors."""
    p = step / float(decay_steps)
    if staircase:
        p = math.floor(p)
    return initial_learning_rate * decay_rate ** p


def adam_step(w, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam (non-amsgrad) [TF-2.1]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t * m / (sqrt(v) + eps).  ``step`` is 1-based."""
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    return w - lr_t * m / (np.sqrt(v) + eps), m, v


def eval_metrics(labels, preds):
    """label/prediction mean, MAE, MRE (normaliser |label|), R^2 (generate_model.py:770-787, 201-216)."""
    y = np.asarray(labels, dtype=np.float64).reshape(-1)
    p = np.asarray(preds, dtype=np.float64).reshape(-1)
    return {"label/mean": y.mean(), "prediction/mean": p.mean(),
            "mae": np.abs(y - p).mean(), "mre": (np.abs(y - p) / np.abs(y)).mean(),
            "r-squared": 1.0 - ((y - p) ** 2).sum() / ((y - y.mean()) ** 2).sum()}
