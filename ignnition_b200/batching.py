"""Block-diagonal batch assembly on the host.

The reference keeps the samples of a batch separate and calls the model once per sample inside
one graph (``model_fn``, ``code/utils/generate_model.py:712-724``; ``batching_func`` :90-99).
Samples share nothing but weights, so here they are concatenated into one block-diagonal graph:
entity indices of sample s are shifted by the number of entities of samples < s.  The result is a
single packed host buffer (pinned when CUDA is present) that goes to the device in one copy; the
device then builds the CSR (``csrc/csr_build.cu``).

Index dtype: the reference contract is int64 (``generate_model.py:141-158``); every BASELINE config
fits int32, which is what the device uses.
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

BIG_COL = 1 << 30          # "no such column": always >= any in-degree -> zero message


class AdjacencySpec:
    """(adjacency name, source entity, destination entity, has edge params)."""

    def __init__(self, name: str, src: str, dst: str, uses_params: bool = False):
        self.name, self.src, self.dst, self.uses_params = name, src, dst, uses_params

    @property
    def seq_key(self):
        return "seq_%s_%s" % (self.src, self.dst)


class SequenceSpec:
    """A multi-source ordered / interleave message passing: destination + ordered adjacency list."""

    def __init__(self, key: str, dst: str, adjs: Sequence[AdjacencySpec], interleave: bool):
        self.key, self.dst, self.adjs, self.interleave = key, dst, list(adjs), interleave


class Batch:
    """Host-side batch: ``arrays`` (name -> ndarray) + entity counts / per-sample offsets."""

    def __init__(self):
        self.arrays: Dict[str, np.ndarray] = {}
        self.num: Dict[str, int] = {}
        self.offsets: Dict[str, np.ndarray] = {}
        self.n_samples = 0
        self.n_edges: Dict[str, int] = {}
        self.max_seq: Dict[str, int] = {}     # adjacency -> longest per-destination list (max(seq) + 1)
        self.dst_sorted: Dict[str, bool] = {}  # adjacency -> the edge list arrives in destination order

    # ---- a window of samples of a larger batch (a dataset file parsed in one go, cut into training batches)
    def take(self, lo: int, hi: int, adjacencies: Sequence["AdjacencySpec"], label_entity: Optional[str] = None) -> "Batch":
        """Samples [lo, hi) as a batch of their own: the same arrays ``assemble`` gives for those samples.  Edges,
        position tables and labels of a sample are contiguous in every array (samples are concatenated in order), so
        a window is a set of array slices with the row offsets of the first sample subtracted.  ``label_entity``: the
        entity whose rows carry the labels (one label per row)."""
        b = Batch()
        b.n_samples = hi - lo
        for e, off in self.offsets.items():
            r0, r1 = int(off[lo]), int(off[hi])
            b.offsets[e] = off[lo:hi + 1] - off[lo]
            b.num[e] = r1 - r0
            b.arrays["sample_of_" + e] = self.arrays["sample_of_" + e][r0:r1] - np.int32(lo)
            b.arrays["offsets_" + e] = _as_i32(b.offsets[e])
        if not hasattr(self, "_edge_bounds"):
            self._edge_bounds = {}
        for a in adjacencies:
            if a.name not in self._edge_bounds:         # edges per sample, found once per file
                es = self.arrays["sample_of_" + a.dst][self.arrays["dst_" + a.name]] if self.n_edges.get(a.name, 0) \
                    else np.zeros(0, np.int32)
                self._edge_bounds[a.name] = np.searchsorted(es, np.arange(self.n_samples + 1))
            eb = self._edge_bounds[a.name]
            e0, e1 = int(eb[lo]), int(eb[hi])
            b.arrays["src_" + a.name] = self.arrays["src_" + a.name][e0:e1] - np.int32(self.offsets[a.src][lo])
            b.arrays["dst_" + a.name] = self.arrays["dst_" + a.name][e0:e1] - np.int32(self.offsets[a.dst][lo])
            b.arrays["seq_" + a.name] = seq = self.arrays["seq_" + a.name][e0:e1]
            b.n_edges[a.name] = e1 - e0
            b.max_seq[a.name] = int(seq.max()) + 1 if e1 > e0 else 0
            if "params_" + a.name in self.arrays:
                b.arrays["params_" + a.name] = self.arrays["params_" + a.name][e0:e1]
            if a.name in self.dst_sorted:
                b.dst_sorted[a.name] = self.dst_sorted[a.name]
        for k, arr in self.arrays.items():
            if k.startswith("pos_off_"):
                key = k[len("pos_off_"):]
                p0, p1 = int(arr[lo]), int(arr[hi])
                b.arrays[k] = arr[lo:hi + 1] - arr[lo]
                b.arrays["pos_src_" + key] = self.arrays["pos_src_" + key][p0:p1]
                b.arrays["pos_col_" + key] = self.arrays["pos_col_" + key][p0:p1]
        return b

    def take_rows(self, b: "Batch", lo: int, hi: int, features: Sequence[Tuple[str, str, int]],
                  label_entity: Optional[str] = None) -> "Batch":
        """fills the per-row arrays (features, labels) of a window made by ``take``"""
        for name, ent, size in features:
            off = self.offsets[ent]
            b.arrays["feat_" + name] = self.arrays["feat_" + name][int(off[lo]) * size:int(off[hi]) * size]
        if "labels" in self.arrays:
            if label_entity is None:
                raise RuntimeError("IGNNITION: a window of a labelled batch needs the entity that carries the labels")
            off = self.offsets[label_entity]
            if self.arrays["labels"].size != int(off[-1]):
                raise RuntimeError("IGNNITION: %d labels for %d rows of %s: windows need one label per row"
                                   % (self.arrays["labels"].size, int(off[-1]), label_entity))
            b.arrays["labels"] = self.arrays["labels"][int(off[lo]):int(off[hi])]
        return b

    # ---- packing: one contiguous buffer, 256-byte aligned slices
    def pack(self, pin: bool = False, skip=()):
        """One contiguous host buffer (pinned when asked) + layout; arrays whose name starts with a
        prefix in ``skip`` are left out (e.g. ``seq_`` when the device builds the CSR by sorting)."""
        import torch
        layout = {}
        off = 0
        arrays = {k: a for k, a in self.arrays.items() if not any(k.startswith(p) for p in skip)}
        for k, a in arrays.items():
            off = (off + 255) // 256 * 256
            layout[k] = (off, a.dtype, a.shape)
            off += a.nbytes
        total = max(off, 256)
        buf = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
        view = buf.numpy()
        for k, a in arrays.items():
            o = layout[k][0]
            view[o:o + a.nbytes] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        return buf, layout


def _as_f32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _as_i32(x) -> np.ndarray:
    a = np.asarray(x)
    if a.size and (a.max(initial=0) >= 2 ** 31 or a.min(initial=0) < -2 ** 31):
        raise RuntimeError("IGNNITION: index does not fit int32")
    return np.ascontiguousarray(a.astype(np.int32))


def position_table(sample: dict, spec: SequenceSpec) -> Tuple[np.ndarray, np.ndarray]:
    """(pos_src, pos_col) of ONE sample: which (source, column) of the reference's concatenated
    padded tensor sits at sequence position p.

    default combine (generate_model.py:523-543): source k's block occupies columns
    [off_k, off_k + max_len_k).  interleave (auxilary_classes.py:421-440): column c of that
    concatenation moves to position indices[c] (tf.scatter_nd), indices = concat over sources of
    ``indices_<src>_to_<dst>`` (generate_model.py:509-518)."""
    max_lens = []
    for a in spec.adjs:
        seq = np.asarray(sample[a.seq_key])
        max_lens.append(int(seq.max()) + 1 if seq.size else 0)
    total = sum(max_lens)
    col_src = np.concatenate([np.full(m, k, dtype=np.int32) for k, m in enumerate(max_lens)]) \
        if total else np.zeros(0, np.int32)
    col_col = np.concatenate([np.arange(m, dtype=np.int32) for m in max_lens]) if total else np.zeros(0, np.int32)
    if not spec.interleave:
        return col_src, col_col
    idx = np.concatenate([np.asarray(sample["indices_%s_to_%s" % (a.src, spec.dst)], dtype=np.int64)
                          for a in spec.adjs])
    if idx.size != total:
        raise RuntimeError("IGNNITION: interleave indices (%d) do not match the padded length (%d)"
                           % (idx.size, total))
    pos_src = np.zeros(total, dtype=np.int32)
    pos_col = np.full(total, BIG_COL, dtype=np.int32)
    pos_src[idx] = col_src
    pos_col[idx] = col_col
    return pos_src, pos_col


def assemble(samples: Sequence[dict], entities: Sequence[str], features: Sequence[Tuple[str, str, int]],
             adjacencies: Sequence[AdjacencySpec], sequences: Sequence[SequenceSpec] = (),
             labels: Optional[Sequence] = None) -> Batch:
    """Concatenate per-sample tensor dicts (reference ``input_fn`` contract) into one Batch.

    features: (feature name, entity, size)."""
    b = Batch()
    b.n_samples = len(samples)
    for e in entities:
        counts = np.array([int(s["num_" + e]) for s in samples], dtype=np.int64)
        off = np.zeros(len(samples) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        b.offsets[e] = off
        b.num[e] = int(off[-1])
        b.arrays["sample_of_" + e] = np.repeat(np.arange(len(samples), dtype=np.int32), counts)
        b.arrays["offsets_" + e] = _as_i32(off)
    for name, ent, size in features:
        parts = []
        for s in samples:
            x = _as_f32(s[name]).reshape(-1)
            if x.size != int(s["num_" + ent]) * size:
                raise RuntimeError("IGNNITION: feature %s has %d values, expected %d x %d"
                                   % (name, x.size, int(s["num_" + ent]), size))
            parts.append(x)
        b.arrays["feat_" + name] = np.concatenate(parts) if parts else np.zeros(0, np.float32)
    for a in adjacencies:
        so, do = b.offsets[a.src], b.offsets[a.dst]
        src = [np.asarray(s["src_" + a.name], dtype=np.int64) + so[i] for i, s in enumerate(samples)]
        dst = [np.asarray(s["dst_" + a.name], dtype=np.int64) + do[i] for i, s in enumerate(samples)]
        seq = [np.asarray(s[a.seq_key], dtype=np.int64) for s in samples]
        b.arrays["src_" + a.name] = _as_i32(np.concatenate(src))
        b.arrays["dst_" + a.name] = _as_i32(np.concatenate(dst))
        b.arrays["seq_" + a.name] = _as_i32(np.concatenate(seq))
        b.n_edges[a.name] = int(b.arrays["src_" + a.name].size)
        b.max_seq[a.name] = int(b.arrays["seq_" + a.name].max()) + 1 if b.n_edges[a.name] else 0
        # per-sample lists in destination order stay in order after the block-diagonal shift
        b.dst_sorted[a.name] = all(bool(np.all(np.diff(np.asarray(s["dst_" + a.name], dtype=np.int64)) >= 0))
                                   for s in samples)
        if a.uses_params:
            # declared tf.int64 then cast to float32 (generate_model.py:149, :454-456): truncation
            # (a sample without any edge of this adjacency carries no parameter rows, nor the key)
            p = [np.trunc(np.asarray(s["params_" + a.name], dtype=np.float64)).astype(np.float32)
                 .reshape(len(s["src_" + a.name]), -1) for s in samples if len(s["src_" + a.name])]
            b.arrays["params_" + a.name] = (np.ascontiguousarray(np.concatenate(p, axis=0)) if p
                                            else np.zeros((0, 1), dtype=np.float32))
    for q in sequences:
        tabs = [position_table(s, q) for s in samples]
        off = np.zeros(len(samples) + 1, dtype=np.int64)
        np.cumsum([t[0].size for t in tabs], out=off[1:])
        b.arrays["pos_off_" + q.key] = _as_i32(off)
        b.arrays["pos_src_" + q.key] = np.concatenate([t[0] for t in tabs]).astype(np.int32)
        b.arrays["pos_col_" + q.key] = np.concatenate([t[1] for t in tabs]).astype(np.int32)
    if labels is not None:
        b.arrays["labels"] = np.concatenate([_as_f32(l).reshape(-1) for l in labels])
    return b


def assemble_tiled(base: dict, n_samples: int, entities: Sequence[str],
                   features: Sequence[Tuple[str, str, int]], adjacencies: Sequence[AdjacencySpec],
                   sequences: Sequence[SequenceSpec], feature_fns: Dict[str, Callable], seed: int = 0,
                   label_fn: Optional[Callable] = None, label_entity: Optional[str] = None) -> Batch:
    """``n_samples`` copies of ONE sample's graph with fresh features (synthetic benchmark batches).

    Same result as ``assemble([sample_k ...])`` where every sample_k has ``base``'s index arrays and
    features drawn by ``feature_fns[name](rng, count)``; built with vectorised numpy."""
    rng = np.random.RandomState(seed)
    b = Batch()
    b.n_samples = n_samples
    ar = np.arange(n_samples, dtype=np.int64)
    for e in entities:
        n = int(base["num_" + e])
        b.offsets[e] = ar_off = np.arange(n_samples + 1, dtype=np.int64) * n
        b.num[e] = int(ar_off[-1])
        b.arrays["sample_of_" + e] = np.repeat(ar.astype(np.int32), n)
        b.arrays["offsets_" + e] = _as_i32(ar_off)
    for name, ent, size in features:
        b.arrays["feat_" + name] = _as_f32(feature_fns[name](rng, b.num[ent] * size))
    for a in adjacencies:
        ns, nd = int(base["num_" + a.src]), int(base["num_" + a.dst])
        src = np.asarray(base["src_" + a.name], dtype=np.int64)
        dst = np.asarray(base["dst_" + a.name], dtype=np.int64)
        seq = np.asarray(base[a.seq_key], dtype=np.int64)
        b.arrays["src_" + a.name] = _as_i32((src[None, :] + (ar * ns)[:, None]).reshape(-1))
        b.arrays["dst_" + a.name] = _as_i32((dst[None, :] + (ar * nd)[:, None]).reshape(-1))
        b.arrays["seq_" + a.name] = _as_i32(np.tile(seq, n_samples))
        b.n_edges[a.name] = int(src.size) * n_samples
        b.max_seq[a.name] = int(seq.max()) + 1 if seq.size else 0
        b.dst_sorted[a.name] = bool(np.all(np.diff(dst) >= 0))
    for q in sequences:
        ps, pc = position_table(base, q)
        b.arrays["pos_off_" + q.key] = _as_i32(np.arange(n_samples + 1, dtype=np.int64) * ps.size)
        b.arrays["pos_src_" + q.key] = np.tile(ps, n_samples).astype(np.int32)
        b.arrays["pos_col_" + q.key] = np.tile(pc, n_samples).astype(np.int32)
    if label_fn is not None:
        b.arrays["labels"] = _as_f32(label_fn(rng, b.num[label_entity]))
    return b
