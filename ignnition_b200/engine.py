"""The B200 message-passing engine: drop-in for the model IGNNITION generates.

``Engine`` mirrors ``ComnetModel`` (reference ``code/utils/generate_model.py:219-695``):
 * ``Engine(model_description)`` creates the weights ``ComnetModel.__init__`` creates (:235-382),
   in Keras layout, inside ONE flat fp32 device buffer (a name -> (offset, shape) table maps TF
   checkpoint variables 1:1),
 * ``engine(input_dict, training=False)`` is ``ComnetModel.call`` (:384-658) for one sample,
 * ``engine.forward(graph)`` runs a block-diagonal batch of samples in one pass (what
   ``model_fn`` does with a Python loop, :712-724).

Host Python only decides WHICH kernels run; every tensor operation on the path is a hand-written
sm_100a kernel behind the C-ABI (``ops.py`` -> ``libignnition_b200.so``).  There is no CPU path.
"""

from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .batching import AdjacencySpec, Batch, SequenceSpec, assemble
from .model_description import FeedForward, MessagePassing, ModelDescription

_ALIGN = 64   # floats: every parameter starts on a 256-byte boundary (float4 / cp.async loads)


class DeviceGraph:
    """A batch resident in HBM: features, CSR per adjacency, step tables, per-entity counts."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None          # the packed upload
        self.t: Dict[str, torch.Tensor] = {}             # views into buf
        self.num: Dict[str, int] = {}
        self.n_samples = 0
        self.csr: Dict[str, Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]] = {}
        self.csr_t: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}   # transposed (by source), training
        self.steps: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.order: Dict[str, torch.Tensor] = {}
        self.meta: Dict[str, torch.Tensor] = {}
        self.max_seq: Dict[str, int] = {}                 # host-known longest list per adjacency
        self.partner: Dict[str, list] = {}               # concat axis 2: per-source row index per CSR position
        self.host_offsets: Dict[str, np.ndarray] = {}    # per-sample row offsets per entity, host copy
        self.small = False                               # built for the one-launch loop of small graphs (no walk order)
        self.small_rows = False                          # every entity small enough for ign_csr_build_small
        self.small_fit = False                           # ... and the model and row lengths fit the one-launch loop
        self.attn_comb: Dict[str, tuple] = {}            # attention over several sources: (rowptr, perm, slot_col, max_len)
        self.step_plan: Dict[str, tuple] = {}            # step-major plan of short ordered updates
        self.step_plan_bwd: Dict[str, tuple] = {}        # the same plan for the step-synchronous backward pass
        self.status: Dict[str, torch.Tensor] = {}
        self.h2d_bytes = 0


class _MPPlan:
    """One message passing (destination + sources) compiled to kernel choices."""

    def __init__(self, mp: MessagePassing, key: str):
        self.mp = mp
        self.key = key
        self.dst = mp.destination_entity
        self.adjs: List[AdjacencySpec] = []
        self.kind = ""         # 'agg_gru' | 'seq_gru' | 'agg_ff'
        self.op = ops.OP_SUM
        self.seq: Optional[SequenceSpec] = None
        self.msg_dim = 0
        self.concat2 = False   # concat along the feature axis: rows gathered per CSR position of source 0
        self.msg_rows = False  # ordered walk over message-MLP rows instead of source states
        self.v1 = False        # GRUCell(reset_after=False): generic composition only
        self.msg_src: List[bool] = []   # per source: its messages come from a message network (rows in edge order)
        self.attn = False      # attention aggregation: column softmax over one sample's destinations
        self.conv = False      # convolution aggregation: (sum . conv_kernel + self) / degree, activation


class Engine:
    def __init__(self, model: ModelDescription, device: Optional[torch.device] = None, seed: int = 0,
                 csr_mode: int = ops.CSR_RANK, sort_by_length: bool = True, max_step_launches: int = 0,
                 fuse_sum_gru: Optional[bool] = None):
        self.model = model
        # device="plan": compile the model description into kernel choices and the parameter table only -- no device, no
        # weights, nothing can run (host-side checks and tests of the plan logic)
        plan_only = isinstance(device, str) and device == "plan"
        self.device = None if plan_only else (
            torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()))
        if not plan_only and self.device.type != "cuda":
            raise RuntimeError("IGNNITION: ignnition_b200 runs on CUDA devices only (no CPU fallback)")
        # adjacency builder: IGN_CSR_RANK places every edge at rowptr[dst] + seq (the reference's scatter_nd((dst, seq)),
        # generate_model.py:490; histogram + scan + placement = 0.12 ms for 6.1 M edges), IGN_CSR_SORT ignores seq and
        # stable-sorts by destination (seq_* need not be uploaded; 0.27 ms when the list is not already in order)
        self.csr_mode = csr_mode
        self.sort_by_length = sort_by_length
        # ordered updates whose longest sequence is <= this run step-synchronously (one launch per step,
        # ign_gru_seq_step); 0 = always the sequence walk (ign_gru_seq), which is faster in round 1
        self.max_step_launches = max_step_launches
        # backward of ordered updates: sequences up to this long run as step-synchronous tensor-core launches
        # (two per step); longer ones walk tiles on the fp32 CUDA cores (ign_gru_seq_bwd)
        self.max_bwd_step_launches = 64
        # sum aggregation + GRU: one fused fp32 kernel, or segment_reduce + tensor-core cell.  None picks by
        # measurement (GEANT2 x 4096 links: 0.121 + 0.087 ms unfused on tensor cores vs 0.333 ms fused fp32)
        self.fuse_sum_gru = fuse_sum_gru
        # aggregation + GRU update as ONE tensor-core kernel (ign_agg_gru_cell_tc: gather, sum / mean / max, gate GEMMs,
        # TMA stores); False = ign_segment_reduce + ign_gru_cell (the round-1 pair, kept for ablation)
        # Measured (profiles/r2_agg_gru.md): a tie at 64-wide states / degree 20 (12.2 vs 12.1 ms on 10 M nodes, and
        # no [n, F] aggregate in HBM), slower at 32-wide (RouteNet links: 0.31 vs 0.19 ms) -> used for 64-wide states
        # unless IGN_FUSED_TC=1 / 0 forces it on / off.  Partitioned graphs always use it (parallel.py).
        self.fused_tc = {"1": True, "0": False}.get(os.environ.get("IGN_FUSED_TC", ""), None)
        self.dims = model.get_input_dimensions()
        self.entities = [e.name for e in model.get_entities()]
        self.hidden = {e.name: e.hidden_state_dimension for e in model.get_entities()}
        self.T = model.get_mp_iterations()
        self.features = [(f.name, e.name, f.size) for e in model.get_entities() for f in e.features]
        self.adjacencies = [AdjacencySpec(a[0], a[1], a[2], a[3] == "True") for a in model.get_adjecency_info()]
        self._adj_by_name = {a.name: a for a in self.adjacencies}
        self._msg_adjacencies = list(self.adjacencies)
        # inference of graphs whose largest entity has at most this many rows runs the whole T-iteration loop as ONE
        # persistent launch (csrc/small_graph.cu) when every stage is an ordered or sum update of 16 / 32-wide states;
        # 0 = never.  The reference's default batch (3 samples) is 546 paths: 16 dependent launches were all gaps
        self.small_graph_rows = int(os.environ.get("IGN_SMALL_GRAPH_ROWS", "8192"))
        self.max_graphs = 32            # captured CUDA graphs kept by forward_graphed (one per batch shape)
        self.bwd_steps_min_rows = ops.BWD_STEPS_MIN_ROWS   # below: the fp32 BPTT walk (one launch) instead of 2 launches per step
        self._needs_perm = set()
        self.has_dropout = False        # a network holds Dropout layers with a non-zero rate: inference only
        self.plans: List[List[_MPPlan]] = []
        self.sequences: List[SequenceSpec] = []
        self.param_table: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        self._param_init: Dict[str, str] = {}
        self._reg: Dict[str, float] = {}
        self._n_params = 0
        self._compile()
        if plan_only:
            return
        self.weights = torch.zeros(self._n_params, dtype=torch.float32, device=self.device)
        self.reset_parameters(seed)

    @property
    def gpu_launches(self) -> int:
        """Kernels launched by libignnition_b200.so in this process (counted inside the library)."""
        from . import _lib
        return int(_lib.load().ign_launch_count())

    # ------------------------------------------------------------------ model compilation
    def _add_param(self, name: str, shape: Tuple[int, ...], init: str, reg: float = 0.0):
        if name in self.param_table:
            return
        off = (self._n_params + _ALIGN - 1) // _ALIGN * _ALIGN
        self.param_table[name] = (off, tuple(shape))
        self._param_init[name] = init
        if reg:
            self._reg[name] = reg
        self._n_params = off + int(np.prod(shape))

    def _add_ff(self, prefix: str, ff: FeedForward, in_dim: int, last_units: Optional[int] = None) -> int:
        n = len(ff.layers)
        if getattr(ff, "dropout", False):
            self.has_dropout = True
        for j, l in enumerate(ff.layers):
            if l.type_layer != "Dense":
                raise RuntimeError("IGNNITION: layer type %s is not supported by the B200 engine "
                                   "(Dense, and Dropout as the identity at inference)" % l.type_layer)
            units = last_units if (j == n - 1 and last_units is not None) else l.units
            if units is None:
                raise RuntimeError("IGNNITION: Dense layer %s has no units" % l.name)
            if l.activation not in ops.ACTIVATIONS:
                raise RuntimeError("IGNNITION: activation %s is not supported" % l.activation)
            self._add_param("%s/%s/kernel" % (prefix, l.name), (in_dim, units), "glorot", l.kernel_regularizer)
            if l.use_bias:
                self._add_param("%s/%s/bias" % (prefix, l.name), (units,), "zeros")
            in_dim = units
        return in_dim

    def _compile(self):
        for si, (stage_name, mps) in enumerate(self.model.get_mp_instances()):
            stage_plans = []
            for mi, mp in enumerate(mps):
                p = _MPPlan(mp, "s%d_m%d" % (si, mi))
                fd = self.hidden[p.dst]
                msg_dims = []
                for src in mp.source_entities:
                    p.adjs.append(self._adj_by_name[src.adj_vector])
                    d = self.hidden[src.name]
                    for k, op in enumerate(src.message_formation):
                        if op.type != "feed_forward_nn":
                            continue
                        self._needs_perm.add(src.adj_vector)
                        din = 0
                        for i in op.input:
                            if i == "hs_source":
                                din += self.hidden[src.name]
                            elif i == "hs_dest":
                                din += fd
                            elif i == "edge_params":
                                din += int(src.extra_parameters)
                            else:
                                raise RuntimeError(
                                    "IGNNITION: named message inputs (%s) cannot run in the reference "
                                    "either (generate_model.py:458 vs :470)" % i)
                        d = self._add_ff("%s_to_%s_message_creation_%d" % (src.name, p.dst, k), op.model, din)
                    msg_dims.append(d)
                concat2 = mp.aggregation.type == "concat" and mp.aggregation.concat_axis == 2
                if len(set(msg_dims)) != 1 and not concat2:
                    raise RuntimeError("IGNNITION: all sources of a message passing must send messages of "
                                       "the same dimension, got %s" % msg_dims)
                p.msg_dim = msg_dims[0]
                agg = mp.aggregation.type
                p.msg_src = [any(op.type == "feed_forward_nn" for op in src.message_formation)
                             for src in mp.source_entities]
                has_msg_nn = any(p.msg_src)
                if concat2:
                    # padded blocks side by side along the FEATURE axis, lengths of the first source
                    # (generate_model.py:496-505); one GRU walk over rows gathered per step
                    if mp.update.type != "recurrent_nn":
                        raise RuntimeError("IGNNITION: concat aggregation needs a recurrent update")
                    p.msg_dim = sum(msg_dims)
                    p.kind = "seq_gru"
                    p.concat2 = True
                    for a in p.adjs:
                        self._needs_perm.add(a.name)
                elif agg in ("sum", "mean", "max"):
                    p.op = {"sum": ops.OP_SUM, "mean": ops.OP_MEAN, "max": ops.OP_MAX}[agg]
                    if p.op == ops.OP_MAX:      # the backward of a max writes per-edge gradients in input edge order
                        for a_ in p.adjs:
                            self._needs_perm.add(a_.name)
                    p.kind = "agg_gru" if mp.update.type == "recurrent_nn" else "agg_ff"
                elif agg in ("ordered", "interleave") or (agg == "concat" and mp.aggregation.concat_axis == 1):
                    # concat along axis 1 = the sources' padded blocks one after the other
                    # (generate_model.py:496-505): the multi-source ordered layout of the step table
                    if mp.update.type != "recurrent_nn":
                        raise RuntimeError("IGNNITION: %s aggregation needs a recurrent update" % agg)
                    p.kind = "seq_gru"
                    if len(p.adjs) > 1 or agg == "interleave":
                        # step entries name rows of the source states, or edge positions (perm) where a message
                        # network produced the source's messages
                        p.seq = SequenceSpec(p.key, p.dst, p.adjs, agg == "interleave")
                        self.sequences.append(p.seq)
                    p.msg_rows = has_msg_nn           # the walk reads message rows (edge order) through perm
                    for k_, a_ in enumerate(p.adjs):
                        if p.msg_src[k_]:
                            self._needs_perm.add(a_.name)
                elif agg == "attention":
                    # several sources: one combined edge list, colliding padded columns add up (generate_model.py:523-543,
                    # SURVEY quirk 7; csrc/attention.cu)
                    if p.msg_dim != fd:          # [m.kernel1 | h.kernel2] . attn_kernel[2 * fd, 1] must conform
                        raise RuntimeError("IGNNITION: attention needs messages as wide as the destination state "
                                           "(%d vs %d)" % (p.msg_dim, fd))
                    self._add_param(p.dst + "_attention/kernel1", (p.msg_dim, p.msg_dim), "glorot")
                    self._add_param(p.dst + "_attention/kernel2", (fd, p.msg_dim), "glorot")
                    self._add_param(p.dst + "_attention/attn_kernel", (2 * fd, 1), "glorot")
                    p.op = ops.OP_SUM
                    p.attn = True
                    for a_ in p.adjs:            # its backward writes per-edge gradients in input edge order
                        self._needs_perm.add(a_.name)
                    p.kind = "agg_gru" if mp.update.type == "recurrent_nn" else "agg_ff"
                elif agg == "convolution":
                    if p.msg_dim != fd:
                        raise RuntimeError(
                            "IGNNITION: When doing the a convolution, both the dimension of the messages sent and the "
                            "destination hidden states should match. In this case, however,the dimensions are %d and "
                            "%d of the source and destination respectively." % (p.msg_dim, fd))
                    if mp.aggregation.activation_function not in ops.ACTIVATIONS:
                        raise RuntimeError("IGNNITION: activation %s is not supported" % mp.aggregation.activation_function)
                    self._add_param(p.dst + "_convolution/conv_kernel", (fd, fd), "glorot")
                    p.op = ops.OP_SUM
                    p.conv = True
                    p.kind = "agg_gru" if mp.update.type == "recurrent_nn" else "agg_ff"
                else:
                    raise RuntimeError("IGNNITION: aggregation '%s' is not built yet in the B200 engine "
                                       "(SURVEY.md section 8f, next rows)" % agg)
                if mp.update.type == "recurrent_nn":
                    if mp.update.recurrent_type != "GRU":
                        raise RuntimeError("IGNNITION: only GRU cells are supported (LSTM cannot run in "
                                           "the reference: single-tensor state, auxilary_classes.py:764)")
                    params = mp.update.cell_parameters
                    # GRUCell(reset_after=False), the Keras v1 cell: one bias vector, the reset gate applied before
                    # the candidate's recurrent product -> three Dense products per step (ops.gru_cell_v1), none of
                    # the fused kernels
                    p.v1 = str(params.get("reset_after", True)) not in ("True", "true", "1")
                    # every other key of the JSON entry goes to tf.keras.layers.GRUCell(**parameters)
                    # (auxilary_classes.py:740-750): values that are Keras' defaults change nothing, anything else
                    # would change the arithmetic these kernels hard-wire
                    same = {"activation": ("tanh",), "recurrent_activation": ("sigmoid",), "use_bias": (True, "True"),
                            "dropout": (0, 0.0, "0", "0.0"), "recurrent_dropout": (0, 0.0, "0", "0.0"),
                            "implementation": (1, 2, "1", "2"), "kernel_initializer": None,
                            "recurrent_initializer": None, "bias_initializer": None, "reset_after": None,
                            "units": None, "name": None}
                    for key, val in params.items():
                        if key not in same or (same[key] is not None and val not in same[key]):
                            raise RuntimeError("IGNNITION: GRUCell parameter %s = %r is not built in the B200 engine "
                                               "(supported: reset_after, and Keras' defaults for the rest)" % (key, val))
                    self._add_param(p.dst + "_update/kernel", (p.msg_dim, 3 * fd), "glorot")
                    self._add_param(p.dst + "_update/recurrent_kernel", (fd, 3 * fd), "orthogonal")
                    self._add_param(p.dst + "_update/bias", (3 * fd,) if p.v1 else (2, 3 * fd), "zeros")
                else:
                    self._add_ff(p.dst + "_ff_update", mp.update.model, p.msg_dim + fd, last_units=fd)
                stage_plans.append(p)
            self.plans.append(stage_plans)
        self.readout = []
        dims = dict(self.hidden)
        for k, op in enumerate(self.model.get_readout_operations()):
            if op.type in ("predict", "neural_network"):
                din = sum(int(dims[i]) for i in op.input)
                dout = self._add_ff("readout_model_%d" % k, op.architecture, din)
                if op.type == "neural_network":
                    dims[op.output_name] = dout
                self.readout.append((k, op))
                if op.type == "predict":
                    break
            elif op.type == "pooling":                   # auxilary_classes.py:1165-1185, per sample
                if op.type_pooling not in ("sum", "mean", "max"):
                    raise RuntimeError("IGNNITION: pooling '%s' is not supported" % op.type_pooling)
                dims[op.output_name] = dims[op.input[0]]
                self.readout.append((k, op))
            elif op.type == "product":                   # auxilary_classes.py:1072-1094
                if op.type_product == "dot_product":
                    # tf.tensordot(a, b, axes=0): the outer product [n, Fa, m, Fb], registered with dimension 1
                    # (generate_model.py:375-376).  A network built on that dimension only accepts a last axis of 1
                    if dims[op.input[1]] != 1:
                        raise RuntimeError("IGNNITION: dot_product is tf.tensordot(axes=0) in the reference and its "
                                           "result is registered with dimension 1: the second input must be 1 wide "
                                           "(got %d), any other last axis fails in the networks built on it"
                                           % dims[op.input[1]])
                    dims[op.output_name] = 1
                elif op.type_product == "element_wise":
                    dims[op.output_name] = dims[op.input[0]]
                else:
                    raise RuntimeError("IGNNITION: product '%s' is not a product of the reference's schema"
                                       % op.type_product)
                self.readout.append((k, op))
            elif op.type == "extend_adjacencies":        # auxilary_classes.py:1236-1265
                dims[op.output_name[0]] = dims[op.input[0]]
                dims[op.output_name[1]] = dims[op.input[1]]
                self._needs_perm.add(op.adj_list)        # its backward reduces per-edge rows in input edge order
                self.readout.append((k, op))
            else:
                raise RuntimeError("IGNNITION: readout operation '%s' is not built yet in the B200 engine "
                                   "(SURVEY.md section 8f, next rows)" % op.type)

    # ------------------------------------------------------------------ weights
    def param(self, name: str) -> torch.Tensor:
        off, shape = self.param_table[name]
        return self.weights[off:off + int(np.prod(shape))].view(*shape)

    def _pview(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.param_table[name]
        return buf[off:off + int(np.prod(shape))].view(*shape)

    def reset_parameters(self, seed: int = 0):
        """Keras default initialisers: glorot_uniform kernels, orthogonal recurrent kernels, zero biases."""
        rng = np.random.RandomState(seed)
        host = np.zeros(self._n_params, dtype=np.float32)
        for name, (off, shape) in self.param_table.items():
            kind = self._param_init[name]
            if kind == "glorot":
                lim = math.sqrt(6.0 / (shape[0] + shape[-1]))
                v = rng.uniform(-lim, lim, shape)
            elif kind == "orthogonal":
                a = rng.normal(size=(shape[1], shape[0]))
                q, r = np.linalg.qr(a)
                q = q * np.sign(np.diag(r))
                v = q.T
            else:
                v = np.zeros(shape)
            host[off:off + v.size] = v.astype(np.float32).reshape(-1)
        self.weights.copy_(torch.from_numpy(host))

    def set_weights(self, w: Dict[str, np.ndarray]):
        host = self.weights.cpu().numpy()
        for name, v in w.items():
            if name not in self.param_table:
                raise RuntimeError("IGNNITION: unknown variable " + name)
            off, shape = self.param_table[name]
            v = np.asarray(v, dtype=np.float32)
            if tuple(v.shape) != shape:
                raise RuntimeError("IGNNITION: variable %s has shape %s, expected %s" % (name, v.shape, shape))
            host[off:off + v.size] = v.reshape(-1)
        self.weights.copy_(torch.from_numpy(host))

    def get_weights(self) -> Dict[str, np.ndarray]:
        host = self.weights.cpu().numpy()
        return {n: host[o:o + int(np.prod(s))].reshape(s).copy() for n, (o, s) in self.param_table.items()}

    # ------------------------------------------------------------------ batches
    def assemble(self, samples: Sequence[dict], labels=None) -> Batch:
        return assemble(samples, self.entities, self.features, self.adjacencies, self.sequences, labels)

    def upload_skip(self) -> tuple:
        """Host arrays the device does not need: ``seq_*`` when the CSR is built by the stable sort
        (seq is the rank in input order, generator_std_to_framework.py:153), ``sample_of_*`` when no
        multi-source sequence needs the sample of a destination."""
        skip = []
        if self.csr_mode == ops.CSR_SORT:
            skip.append("seq_")
        if not self.sequences:
            skip.append("sample_of_")
        return tuple(skip)

    def pack(self, batch: Batch, pin: bool = True, full: bool = False):
        skip = () if full else self.upload_skip()
        if not full and self.csr_mode == ops.CSR_RANK:
            # an edge list that arrives in destination order needs no seq on the device: the stable-sort builder
            # takes its pre-sorted fast path (and still verifies the order on the device)
            skip = skip + tuple("seq_" + n for n, srt in batch.dst_sorted.items() if srt)
        return batch.pack(pin=pin, skip=skip)

    # ------------------------------------------------------------------ small batches: captured CUDA graphs
    def forward_graphed(self, batch: Batch, pinned=None, copy: bool = True) -> torch.Tensor:
        """Inference of ``batch`` through a captured CUDA graph: adjacency build + T iterations + readout are ONE
        graph launch instead of ~50 kernel launches issued from Python, which is what bounds the reference's own batch
        sizes (train_options.ini: batch_size 3 ... 32; framework_operations.py:200-236).  One graph per batch SHAPE
        (array layout of the packed batch, row counts, sequence lengths); a batch of a known shape costs one
        host -> device copy into the graph's static staging buffer and one replay.  The returned predictions live in
        the graph's static output buffer: read them before the next call with the same shape.
        ``copy=False`` replays on whatever the staging buffer holds (bench: device-resident inputs)."""
        if pinned is None:
            pinned = self.pack(batch)
        buf, layout = pinned
        sig = (batch.n_samples, tuple(sorted(batch.num.items())), tuple(sorted(batch.max_seq.items())),
               tuple(sorted(batch.dst_sorted.items())),
               tuple((k, off, str(np.dtype(dt)), tuple(shape)) for k, (off, dt, shape) in layout.items()))
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        entry = self._graphs.get(sig)
        if entry is None and len(self._graphs) >= self.max_graphs:
            # a dataset whose samples rarely share a shape: every captured graph keeps its own memory pool, so the
            # cache is bounded; further new shapes run the same kernels launch by launch
            return self.forward(self.build_graph(self.upload(batch, pinned)))
        if entry is None:
            stage = torch.empty(buf.numel(), dtype=torch.uint8, device=self.device)
            dg = self.upload(batch, pinned, out=stage)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # eager pass first: one-time kernel attributes, workspace sizes
                self.build_graph(dg)
                self.forward(dg)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.load().ign_launch_count()
            with torch.cuda.graph(graph):
                self.build_graph(dg)
                pred = self.forward(dg)
            entry = (graph, stage, dg, pred, int(_lib.load().ign_launch_count() - l0))
            self._graphs[sig] = entry
            copy = True
        graph, stage, dg, pred, _ = entry
        if copy:
            stage.copy_(buf, non_blocking=True)
        graph.replay()
        return pred

    def graphed_kernels(self, batch: Batch, pinned=None) -> int:
        """Kernels inside the captured graph of this batch shape (0 if none was captured yet)."""
        if pinned is None:
            pinned = self.pack(batch)
        buf, layout = pinned
        for sig, entry in getattr(self, "_graphs", {}).items():
            if sig[0] == batch.n_samples and sig[1] == tuple(sorted(batch.num.items())):
                return entry[4]
        return 0

    def upload(self, batch: Batch, pinned=None, out: Optional[torch.Tensor] = None) -> DeviceGraph:
        """Host -> device copy of the packed batch (one cudaMemcpyAsync) + tensor views."""
        if pinned is None:
            pinned = self.pack(batch, full=True)
        buf, layout = pinned
        g = DeviceGraph()
        g.buf = out if out is not None else torch.empty(buf.numel(), dtype=torch.uint8, device=self.device)
        g.buf.copy_(buf, non_blocking=True)
        g.h2d_bytes = buf.numel()
        for k, (off, dtype, shape) in layout.items():
            n = int(np.prod(shape)) if len(shape) else 1
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
            g.t[k] = g.buf[off:off + n * 4].view(tdt).view(*shape) if n else torch.empty(shape, dtype=tdt, device=self.device)
        g.num = dict(batch.num)
        g.n_samples = batch.n_samples
        g.max_seq = dict(batch.max_seq)
        g.host_offsets = {e: batch.arrays["offsets_" + e] for e in self.entities if "offsets_" + e in batch.arrays}
        return g

    def build_graph(self, g: DeviceGraph, training: bool = False, check: bool = False) -> DeviceGraph:
        """Device adjacency builder: CSR per adjacency, length order, step tables."""
        small_rows = 0 < max(g.num.values()) <= self.small_graph_rows
        g.small_rows = small_rows                  # every adjacency fits the one-launch CSR builder
        g.small_fit = small_rows and self._small_program_ok() and self._small_fit(g.max_seq)
        g.small = not training and g.small_fit
        if small_rows and not check and len(self.adjacencies) <= 8:
            # one launch for every adjacency (ign_csr_build_small) instead of a dozen launches of ~3 us each
            specs = [(g.t["dst_" + a.name], g.t["src_" + a.name],
                      g.t.get("seq_" + a.name) if self.csr_mode == ops.CSR_RANK else None, g.num[a.dst],
                      training or a.name in self._needs_perm) for a in self.adjacencies]
            for a, built in zip(self.adjacencies, ops.csr_build_small(specs)):
                g.csr[a.name] = built
            per_adjacency = []
        else:
            per_adjacency = self.adjacencies
        for a in per_adjacency:
            dst, src, seq = g.t["dst_" + a.name], g.t["src_" + a.name], g.t.get("seq_" + a.name)
            # no seq on the device (the host saw the list in destination order): stable sort, pre-sorted fast path
            mode = self.csr_mode if seq is not None else ops.CSR_SORT
            rowptr, col, perm, status = ops.csr_build(dst, src, seq, g.num[a.dst], mode,
                                                      want_perm=training or check or a.name in self._needs_perm,
                                                      want_status=check)
            g.csr[a.name] = (rowptr, col, perm)
            if status is not None:
                g.status[a.name] = status
        for stage in self.plans:
            for p in stage:
                if p.attn and len(p.adjs) > 1:
                    # the first source's columns are its seq, the others' start at the destination's edge count in
                    # that source (generate_model.py:538-539): at most twice its longest list
                    ms = [g.max_seq.get(a.name, 0) for a in p.adjs]
                    max_len = max([ms[0], 1] + [2 * m for m in ms[1:]])
                    g.attn_comb[p.key] = ops.attention_combine(
                        [g.csr[a.name][0] for a in p.adjs], [g.csr[a.name][2] for a in p.adjs],
                        [int(g.t["dst_" + a.name].numel()) for a in p.adjs], max_len) + (max_len,)
                if p.kind != "seq_gru":
                    continue
                if p.concat2:
                    names = [a.name for a in p.adjs]
                    if len(set(g.max_seq.get(n, 0) for n in names)) != 1:
                        raise RuntimeError("IGNNITION: concat along axis 2 needs padded blocks of equal length, "
                                           "got %s (tf.concat fails in the reference)"
                                           % [g.max_seq.get(n, 0) for n in names])
                    rowptr0, col0, perm0 = g.csr[names[0]]
                    use_msgs = [any(op.type == "feed_forward_nn" for op in src.message_formation)
                                for src in p.mp.source_entities]
                    idx = [perm0 if use_msgs[0] else col0]
                    for k in range(1, len(names)):
                        rp, c, pm = g.csr[names[k]]
                        idx.append(ops.partner_index(rowptr0, rp, pm if use_msgs[k] else c, int(col0.numel())))
                    g.partner[p.key] = idx
                    g.steps[p.key] = (rowptr0, torch.arange(col0.numel(), dtype=torch.int32, device=self.device))
                elif p.seq is None:
                    rowptr, col, perm = g.csr[p.adjs[0].name]
                    g.steps[p.key] = (rowptr, perm if p.msg_rows else col)
                else:
                    rps = [g.csr[a.name][0] for a in p.adjs]
                    cols = [g.csr[a.name][2 if p.msg_src[k] else 1] for k, a in enumerate(p.adjs)]
                    total = sum(int(c.numel()) for c in cols)
                    multi = g.n_samples > 1
                    g.steps[p.key] = ops.steps_build(
                        rps, cols, g.t["sample_of_" + p.dst] if multi else None, g.t["pos_off_" + p.key],
                        g.t["pos_src_" + p.key], g.t["pos_col_" + p.key], g.num[p.dst], total)
                if g.small:          # one warp per destination in the one-launch loop: no length order, no tile table
                    continue
                if self.sort_by_length and g.num[p.dst] > 0:
                    g.order[p.key] = ops.length_order(g.steps[p.key][0])
                if g.num[p.dst] > 0:
                    g.meta[p.key] = ops.seq_meta(g.steps[p.key][0], g.steps[p.key][1], g.order.get(p.key))
                    # short sequences (RouteNet paths): one streaming launch per step instead of a walk
                    max_steps = sum(g.max_seq.get(a.name, 1 << 30) for a in p.adjs)
                    if (p.key in g.order and 1 <= max_steps <= self.max_step_launches and p.msg_dim == 32
                            and self.hidden[p.dst] == 32 and ops.tensor_cores_enabled() and not p.v1):
                        g.step_plan[p.key] = (ops.seq_step_plan(g.meta[p.key], g.steps[p.key][1], max_steps), max_steps)
                    # training: BPTT runs step-synchronously on the tensor cores (ign_gru_seq_bwd_steps)
                    # (below one tile per SM the 2 x max_steps launches cost more than the fp32 walk they replace)
                    if (training and p.key in g.order and 1 <= max_steps <= self.max_bwd_step_launches and not p.v1
                            and p.msg_dim == 32 and self.hidden[p.dst] == 32 and ops.tensor_cores_enabled()
                            and g.num[p.dst] >= self.bwd_steps_min_rows):
                        g.step_plan_bwd[p.key] = g.step_plan.get(p.key) or (
                            ops.seq_step_plan(g.meta[p.key], g.steps[p.key][1], max_steps), max_steps)
        return g

    def prepare(self, samples_or_batch, labels=None, training: bool = False, check: bool = False) -> DeviceGraph:
        batch = samples_or_batch if isinstance(samples_or_batch, Batch) else self.assemble(samples_or_batch, labels)
        return self.build_graph(self.upload(batch), training=training, check=check)

    # ------------------------------------------------------------------ forward
    def _act(self, name):
        return ops.ACTIVATIONS[name]

    def _run_ff(self, prefix: str, ff: FeedForward, x: torch.Tensor, saves: Optional[list] = None) -> torch.Tensor:
        layers = ff.layers
        # inference: a Dense layer followed by a single-output linear layer runs as one kernel
        # (the hidden activations never leave the SM) -- the readout head of RouteNet / Q-size
        if (saves is None and len(layers) >= 2 and layers[-1].activation in (None, "None", "linear")
                and self.param("%s/%s/kernel" % (prefix, layers[-1].name)).shape[1] == 1):
            wl = self.param("%s/%s/kernel" % (prefix, layers[-2].name))
            pb = lambda l: self.param("%s/%s/bias" % (prefix, l.name)) if l.use_bias else None
            if len(layers) == 3:
                w1 = self.param("%s/%s/kernel" % (prefix, layers[0].name))
                if ops.mlp_head_supported(x.shape[0], w1.shape[0], w1.shape[1], wl.shape[1]):
                    return ops.mlp_head(x, w1, pb(layers[0]), self._act(layers[0].activation), wl, pb(layers[1]),
                                        self._act(layers[1].activation),
                                        self.param("%s/%s/kernel" % (prefix, layers[2].name)).reshape(-1), pb(layers[2]))
            if ops.dense_head_supported(x.shape[0], wl.shape[0], wl.shape[1]):
                for l in layers[:-2]:
                    x = self._dense_layer(prefix, l, x, None)
                l2, l3 = layers[-2], layers[-1]
                return ops.dense_head(
                    x, wl, self.param("%s/%s/bias" % (prefix, l2.name)) if l2.use_bias else None,
                    self._act(l2.activation), self.param("%s/%s/kernel" % (prefix, l3.name)).reshape(-1),
                    self.param("%s/%s/bias" % (prefix, l3.name)) if l3.use_bias else None)
        for l in layers:
            x = self._dense_layer(prefix, l, x, saves)
        return x

    def _dense_layer(self, prefix: str, l, x: torch.Tensor, saves: Optional[list]) -> torch.Tensor:
        w = self.param("%s/%s/kernel" % (prefix, l.name))
        b = self.param("%s/%s/bias" % (prefix, l.name)) if l.use_bias else None
        y = ops.dense(x, w, b, self._act(l.activation))
        if saves is not None:
            # the backward pass takes act' from the layer's OUTPUT (every supported activation allows it), so the
            # pre-activation is never written: one [M, N] tensor per layer instead of two
            saves.append((prefix, l, x, y))
        return y

    def _messages(self, p: _MPPlan, k: int, g: DeviceGraph, state: Dict[str, torch.Tensor],
                  tape: Optional[list] = None):
        """Per-edge messages of source k in INPUT edge order, or None for direct_assignation."""
        src = p.mp.source_entities[k]
        a = p.adjs[k]
        msgs = None
        last = None
        for j, op in enumerate(src.message_formation):
            if op.type != "feed_forward_nn":
                continue
            parts, idx = [], []
            for i in op.input:
                if i == "hs_source":
                    parts.append(state[src.name]); idx.append(g.t["src_" + a.name])
                elif i == "hs_dest":
                    parts.append(state[p.dst]); idx.append(g.t["dst_" + a.name])
                else:
                    parts.append(g.t["params_" + a.name]); idx.append(None)
            n_edges = g.t["src_" + a.name].numel()
            prefix = "%s_to_%s_message_creation_%d" % (src.name, p.dst, j)
            first = op.model.layers[0]
            w0 = self.param("%s/%s/kernel" % (prefix, first.name))
            if tape is None and ops.gather_dense_supported([int(t.shape[1]) for t in parts], int(w0.shape[1]), n_edges):
                # inference: the gather and the concat run inside the first layer's GEMM (ign_gather_dense), the
                # [E, sum F] input is never written; the remaining layers follow on its output
                y0 = ops.gather_dense(parts, idx, n_edges, w0,
                                      self.param("%s/%s/bias" % (prefix, first.name)) if first.use_bias else None,
                                      self._act(first.activation))
                msgs = y0
                for l in op.model.layers[1:]:
                    msgs = self._dense_layer(prefix, l, msgs, None)
                last = (list(op.input), [int(t.shape[1]) for t in parts], None)
                continue
            x = ops.gather_concat(parts, idx, n_edges)
            saves = [] if tape is not None else None
            msgs = self._run_ff(prefix, op.model, x, saves)
            last = (list(op.input), [int(t.shape[1]) for t in parts], saves)
        if tape is not None and last is not None:       # only the last network's output is the message (:470-475)
            tape.append(("msg_ff", p, k) + last)
        return msgs

    def _mp_forward(self, p: _MPPlan, g: DeviceGraph, state: Dict[str, torch.Tensor], tape: Optional[list]):
        dst = p.dst
        h = state[dst]
        n_dst = g.num[dst]
        K = self.param(dst + "_update/kernel") if p.mp.update.type == "recurrent_nn" else None
        R = self.param(dst + "_update/recurrent_kernel") if K is not None else None
        B = self.param(dst + "_update/bias") if K is not None else None
        out = torch.empty_like(h)
        msgs = [self._messages(p, k, g, state, tape) for k in range(len(p.adjs))]
        has_msg = [m is not None for m in msgs]
        if p.kind == "seq_gru":
            rowptr_s, steps = g.steps[p.key]
            srcs = []
            for k, a in enumerate(p.adjs):
                srcs.append(msgs[k] if msgs[k] is not None else state[a.src])
            concat_widths = None
            if p.concat2:
                concat_widths = [int(s_.shape[1]) for s_ in srcs]
                srcs = [ops.gather_concat(srcs, g.partner[p.key], int(g.partner[p.key][0].numel()))]
            h_seq = None
            if tape is not None:
                h_seq = torch.empty(steps.numel(), h.shape[1], dtype=torch.float32, device=self.device)
            if p.key in g.step_plan and ops.tensor_cores_enabled():
                plan, max_steps = g.step_plan[p.key]
                ops.gru_seq_steps(plan, g.meta[p.key], srcs, h, K, R, B, max_steps, out=out, h_seq=h_seq)
            else:
                ops.gru_seq(rowptr_s, steps, g.order.get(p.key), srcs, h, K, R, B, out=out, h_seq=h_seq,
                            meta=g.meta.get(p.key))
            if tape is not None:          # the rows the walk read: source states, or the message network's rows
                tape.append(("seq_gru", p, list(srcs), h, h_seq, concat_widths))
            return out

        # aggregating kinds
        fused = (p.kind == "agg_gru" and p.op == ops.OP_SUM and len(p.adjs) == 1 and msgs[0] is None
                 and not p.conv and not p.attn and not p.v1 and self._fusable(p.msg_dim, h.shape[1])
                 and (self.fuse_sum_gru if self.fuse_sum_gru is not None
                      else (not ops.tensor_cores_enabled() or n_dst < ops.SMALL_ROWS)))   # small: 1 launch, not 3
        fused_tc = (p.kind == "agg_gru" and len(p.adjs) == 1 and msgs[0] is None and not p.conv and not p.attn
                    and not fused and n_dst > 0 and not p.v1
                    and (self.fused_tc if self.fused_tc is not None else h.shape[1] == 64)
                    and ops.agg_gru_cell_tc_supported(p.msg_dim, h.shape[1]))
        if fused_tc:
            rowptr, col, _ = g.csr[p.adjs[0].name]
            agg = torch.empty(n_dst, p.msg_dim, dtype=torch.float32, device=self.device) if tape is not None else None
            ops.agg_gru_cell_tc(p.op, rowptr, col, state[p.adjs[0].src], h, K, R, B, [out], agg_out=agg)
            if tape is not None:
                tape.append(("agg_gru_unfused", p, has_msg, h, agg, (state[p.adjs[0].src], col)))
            return out
        if fused:
            rowptr, col, _ = g.csr[p.adjs[0].name]
            agg = torch.empty(n_dst, p.msg_dim, dtype=torch.float32, device=self.device) if tape is not None else None
            ops.agg_gru_cell(rowptr, col, state[p.adjs[0].src], h, K, R, B, out=out, agg_out=agg)
            if tape is not None:
                tape.append(("agg_gru", p, has_msg, h, agg, None))
            return out
        agg = None
        max_src = None       # (rows, index per CSR slot) the aggregation read: what the backward of a max compares
        if p.attn:      # Attention_aggr (auxilary_classes.py:278-344)
            F = p.msg_dim
            slot_col = None
            if len(p.adjs) == 1:
                a = p.adjs[0]
                rowptr, col, perm = g.csr[a.name]
                rows, idx = (state[a.src], col) if msgs[0] is None else (msgs[0], perm)
                max_len = max(g.max_seq.get(a.name, 0), 1)
            else:       # comb_src_states: the per-edge messages of all sources, one list (generate_model.py:531-541)
                rowptr, idx, slot_col, max_len = g.attn_comb[p.key]
                counts = [int(g.t["dst_" + a.name].numel()) for a in p.adjs]
                rows = torch.empty(sum(counts), F, dtype=torch.float32, device=self.device)
                off = 0
                for k, a in enumerate(p.adjs):
                    if counts[k] and msgs[k] is None:
                        ops.gather_concat([state[a.src]], [g.t["src_" + a.name]], counts[k], out=rows[off:off + counts[k]])
                    elif counts[k]:
                        rows[off:off + counts[k]].copy_(msgs[k])
                    off += counts[k]
            ak = self.param(dst + "_attention/attn_kernel")
            v1 = ops.dense(self.param(dst + "_attention/kernel1"), ak[:F], None, 0)
            v2 = ops.dense(self.param(dst + "_attention/kernel2"), ak[F:], None, 0)
            agg = ops.attention_aggregate(rowptr, idx, rows, ops.dense(rows, v1, None, 0), ops.dense(h, v2, None, 0),
                                          g.t["offsets_" + dst], max_len, keep_ws=tape is not None, slot_col=slot_col)
            if tape is not None:
                agg, attn_ws = agg
                max_src = ("attn", rows, idx, v1, v2, attn_ws, h, max_len)
        for k, a in enumerate(p.adjs if not p.attn else []):
            rowptr, col, perm = g.csr[a.name]
            if msgs[k] is None:
                part = ops.segment_reduce(ops.OP_SUM if len(p.adjs) > 1 else p.op, rowptr, col, state[a.src])
                max_src = (state[a.src], col)
            else:
                part = ops.segment_reduce(ops.OP_SUM if len(p.adjs) > 1 else p.op, rowptr, perm, msgs[k])
                max_src = (msgs[k], perm)
            if agg is None:
                agg = part
            else:
                ops.axpy(1.0, part, agg)
        if len(p.adjs) > 1 and p.op != ops.OP_SUM:
            raise RuntimeError("IGNNITION: mean/max over several sources is not built")
        if p.conv:      # Conv_aggr (auxilary_classes.py:366-401); the kernel product commutes with the sum
            if len(p.adjs) != 1:
                raise RuntimeError("IGNNITION: convolution over several sources is not built")
            agg_sum = agg
            nsum = ops.dense(agg, self.param(dst + "_convolution/conv_kernel"), None, 0)
            agg = ops.conv_finish(nsum, h, g.csr[p.adjs[0].name][0], self._act(p.mp.aggregation.activation_function))
            max_src = ("conv", agg_sum)          # the backward needs the plain neighbour sum (train.py)
        if p.kind == "agg_gru":
            ops.gru_cell(agg, h, K, R, B, out=out)
            if tape is not None:
                tape.append(("agg_gru_unfused", p, has_msg, h, agg, max_src))
            return out
        x = ops.gather_concat([agg, h], [None, None], n_dst)          # FF update, generate_model.py:599
        ff = p.mp.update.model
        layers = FeedForward(list(ff.layers))
        saves = [] if tape is not None else None
        y = self._run_ff(dst + "_ff_update", layers, x, saves)
        if tape is not None:
            tape.append(("agg_ff", p, has_msg, int(agg.shape[1]), saves, agg, max_src))
        return y

    @staticmethod
    def _fusable(f_in: int, units: int) -> bool:
        # weights must stay resident in shared memory next to the tiles (csrc/gru.cu)
        return f_in in (16, 32) and units in (16, 32)

    def initial_states(self, g: DeviceGraph) -> Dict[str, torch.Tensor]:
        state = {}
        for e in self.model.get_entities():
            feats = [g.t["feat_" + f.name] for f in e.features]
            sizes = [f.size for f in e.features]
            state[e.name] = ops.init_state(feats, sizes, g.num[e.name], e.hidden_state_dimension,
                                           out=torch.empty(g.num[e.name], e.hidden_state_dimension,
                                                           dtype=torch.float32, device=self.device))
        return state

    def _small_program_ok(self) -> bool:
        """Every stage is an ordered / interleave / concat-axis-1 walk or a sum + GRU update over source STATES, all
        states and messages 16 or 32 wide: what csrc/small_graph.cu runs in one launch."""
        ok = getattr(self, "_small_ok", None)
        if ok is None:
            flat = [p for stage in self.plans for p in stage]
            widths = set(self.hidden.values())
            ok = (len(widths) == 1 and next(iter(widths)) in (16, 32)
                  and 1 <= len(flat) <= 8 and len(self.entities) <= 8
                  and all(p.msg_dim == self.hidden[p.dst] and not any(p.msg_src) and len(p.adjs) <= 4 and not p.v1
                          and ((p.kind == "seq_gru" and not p.concat2)
                               or (p.kind == "agg_gru" and p.op == ops.OP_SUM and not p.attn and not p.conv
                                   and len(p.adjs) == 1)) for p in flat))
            self._small_ok = ok
        return ok

    def _small_fit(self, max_seq: Dict[str, int]) -> bool:
        """Row lengths the one-launch loop has been verified on (tests/test_gpu_graphs.py): every walk at most one block
        of entries long (units per lane group), sums of any length at 32 units and up to one block at 16.  Longer
        rows take the per-stage kernels.  ``max_seq``: the batch's longest list per adjacency (host-known)."""
        units = next(iter(self.hidden.values()))
        for stage in self.plans:
            for p in stage:
                longest = sum(int(max_seq.get(a.name, 1 << 30)) for a in p.adjs)
                if p.kind == "seq_gru" and longest > units:
                    return False
                if p.kind != "seq_gru" and units < 32 and longest > units:
                    return False
        return True

    def _mp_small(self, g: DeviceGraph, state: Dict[str, torch.Tensor], T: int,
                  tape: Optional[list] = None) -> Dict[str, torch.Tensor]:
        """The T-iteration loop in one launch.  With a ``tape`` (training) every stage keeps what the backward pass
        reads -- its new states, the state after every step of a walk, the neighbour sum -- and the tape gets the same
        entries, in the same order, as the per-stage kernels would have written."""
        eid = {e: i for i, e in enumerate(self.entities)}
        buf0 = [state[e] for e in self.entities]
        buf1 = [torch.empty_like(t) if tape is None else t for t in buf0]
        flat = [p for stage in self.plans for p in stage]
        kinds, dsts, srcs, rowptrs, idxs, Ks, Rs, Bs = [], [], [], [], [], [], [], []
        for p in flat:
            seq = p.kind == "seq_gru"
            rowptr, idx = g.steps[p.key] if seq else g.csr[p.adjs[0].name][:2]
            kinds.append(0 if seq else 1)
            dsts.append(eid[p.dst])
            srcs.append([eid[a.src] for a in p.adjs])
            rowptrs.append(rowptr)
            idxs.append(idx)
            Ks.append(self.param(p.dst + "_update/kernel"))
            Rs.append(self.param(p.dst + "_update/recurrent_kernel"))
            Bs.append(self.param(p.dst + "_update/bias"))
        U = next(iter(self.hidden.values()))
        rows = [g.num[e] for e in self.entities]
        if tape is None:
            final = ops.small_graph_forward(U, rows, buf0, buf1, kinds, dsts, srcs, rowptrs, idxs, Ks, Rs, Bs, T)
            return {e: (buf1 if final[i] else buf0)[i] for i, e in enumerate(self.entities)}
        outs, hseqs, aggs = [], [], []
        cur = dict(state)
        for _ in range(T):
            for k, p in enumerate(flat):
                h = cur[p.dst]
                out = torch.empty_like(h)
                if kinds[k] == 0:
                    h_seq = torch.empty(max(int(idxs[k].numel()), 1), U, dtype=torch.float32, device=self.device)
                    tape.append(("seq_gru", p, [cur[a.src] for a in p.adjs], h, h_seq, None))
                    hseqs.append(h_seq)
                    aggs.append(None)
                else:
                    agg = torch.empty_like(h)
                    tape.append(("agg_gru", p, [False], h, agg, None))
                    hseqs.append(None)
                    aggs.append(agg)
                outs.append(out)
                cur[p.dst] = out
        ops.small_graph_forward(U, rows, buf0, buf1, kinds, dsts, srcs, rowptrs, idxs, Ks, Rs, Bs, T,
                                step_out=outs, step_hseq=hseqs, step_agg=aggs)
        return cur

    def message_passing(self, g: DeviceGraph, state: Dict[str, torch.Tensor], iterations: Optional[int] = None,
                        tape: Optional[list] = None) -> Dict[str, torch.Tensor]:
        T = self.T if iterations is None else iterations
        if tape is None and g.small and T > 0:
            return self._mp_small(g, state, T)
        n_stages = sum(len(stage) for stage in self.plans)
        if tape is not None and g.small_fit and T > 0 and T * n_stages <= 64:
            return self._mp_small(g, state, T, tape)
        for _ in range(T):
            for stage in self.plans:
                for p in stage:
                    state[p.dst] = self._mp_forward(p, g, state, tape)      # written back at once (:602)
        return state

    def readout_forward(self, state: Dict[str, torch.Tensor], tape: Optional[list] = None,
                        g: Optional[DeviceGraph] = None, return_states: bool = False):
        st = dict(state)
        owner = {e: e for e in self.entities}            # which entity's rows a derived state is laid out by
        result = None
        for k, op in self.readout:
            if op.type == "pooling":
                ent = owner.get(op.input[0])
                if ent is None or g is None:
                    raise RuntimeError("IGNNITION: pooling needs an entity-shaped input")
                red = {"sum": ops.OP_SUM, "mean": ops.OP_MEAN, "max": ops.OP_MAX}[op.type_pooling]
                st[op.output_name] = ops.segment_reduce(red, g.t["offsets_" + ent], None, st[op.input[0]])
                owner[op.output_name] = None              # one row per sample
                if tape is not None:
                    tape.append(("pool", op, red, g.t["offsets_" + ent], st[op.input[0]], st[op.output_name]))
                continue
            if op.type == "product" and op.type_product == "dot_product":
                # per sample [n, Fa] (x) [m, 1] -> [n, Fa, m, 1], laid out as (n Fa m) rows of width 1: vec(a) b^T is a
                # Dense product of inner width 1 (ign_dense); the samples' blocks follow one another
                a, b = st[op.input[0]], st[op.input[1]]
                ea, eb = owner.get(op.input[0]), owner.get(op.input[1])
                if ea is None or eb is None or g is None:
                    raise RuntimeError("IGNNITION: dot_product needs entity-shaped inputs")
                oa, ob = self._host_offsets(g, ea), self._host_offsets(g, eb)
                fa = int(a.shape[1])
                sizes = [int(oa[i + 1] - oa[i]) * fa * int(ob[i + 1] - ob[i]) for i in range(g.n_samples)]
                y = torch.empty(sum(sizes), 1, dtype=torch.float32, device=self.device)
                pos = 0
                for i, sz in enumerate(sizes):
                    if sz:
                        na, nb = int(oa[i + 1] - oa[i]), int(ob[i + 1] - ob[i])
                        ops.dense(a[oa[i]:oa[i + 1]].view(na * fa, 1), b[ob[i]:ob[i + 1]].view(1, nb), None, 0,
                                  out=y[pos:pos + sz].view(na * fa, nb), tensor_cores=False)
                    pos += sz
                st[op.output_name] = y
                owner[op.output_name] = None
                if tape is not None:
                    tape.append(("outer", op, a, b, oa, ob))
                continue
            if op.type == "product":
                a, b = st[op.input[0]], st[op.input[1]]
                if a.shape != b.shape:
                    raise RuntimeError("IGNNITION:  The product operation between %s and %s failed. Check that the "
                                       "dimensions are compatible." % (op.input[0], op.input[1]))
                st[op.output_name] = ops.mul(a, b)
                owner[op.output_name] = owner.get(op.input[0])
                if tape is not None:
                    tape.append(("product", op, a, b))
                continue
            if op.type == "extend_adjacencies":
                if g is None:
                    raise RuntimeError("IGNNITION: extend_adjacencies needs the graph tensors")
                src_idx, dst_idx = g.t["src_" + op.adj_list], g.t["dst_" + op.adj_list]
                n_e = src_idx.numel()
                st[op.output_name[0]] = ops.gather_concat([st[op.input[0]]], [src_idx], n_e)
                st[op.output_name[1]] = ops.gather_concat([st[op.input[1]]], [dst_idx], n_e)
                owner[op.output_name[0]] = owner[op.output_name[1]] = None
                if tape is not None:
                    tape.append(("extend", op))
                continue
            if len(op.input) == 1:
                x = st[op.input[0]]
            else:
                x = ops.gather_concat([st[i] for i in op.input], [None] * len(op.input), st[op.input[0]].shape[0])
            saves = [] if tape is not None else None
            y = self._run_ff("readout_model_%d" % k, op.architecture, x, saves)
            if tape is not None:
                tape.append(("readout", op, saves, [int(st[i].shape[1]) for i in op.input]))
            if op.type == "predict":
                result = y
                break
            st[op.output_name] = y
            owner[op.output_name] = owner.get(op.input[0])
        if return_states:
            return result, st
        return result

    @staticmethod
    def _host_offsets(g: DeviceGraph, entity: str):
        """Per-sample row offsets of an entity on the host (kept by ``upload``; read back otherwise)."""
        off = g.host_offsets.get(entity)
        if off is None:
            off = g.host_offsets[entity] = g.t["offsets_" + entity].cpu().numpy()
        return [int(v) for v in off]

    def forward(self, g: DeviceGraph, training: bool = False, return_states: bool = False,
                tape: Optional[list] = None):
        state = self.initial_states(g)
        state = self.message_passing(g, state, tape=tape)
        pred = self.readout_forward(state, tape=tape, g=g)
        if return_states:
            return pred, state
        return pred

    def __call__(self, input: dict, training: bool = False) -> torch.Tensor:
        """ComnetModel.call for ONE sample given as the reference's tensor dict (host arrays)."""
        g = self.prepare([input])
        return self.forward(g, training=training)
