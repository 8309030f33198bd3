"""Native dataset ingest: ``data.json`` text -> :class:`Batch` in one pass of C++.

Same result, array for array, as ``generator.sample_to_tensors`` on every sample followed by
``batching.assemble`` (the reference's ``generator`` + per-sample tensors,
``code/utils/generator_std_to_framework.py:53-230``), without the per-edge Python loops: SURVEY.md
section 8f rank 1.  The on-disk format is unchanged (``*.tar.gz`` holding ``data.json``).

Multi-source ordered / concat / interleave message passings get their per-sample position tables
(``batching.position_table``) from the same pass; an interleave needs the name of the sample key that
holds its pattern (``interleave_names``: ``ModelDescription.get_interleave_tensors()``).
"""

from __future__ import annotations

import ctypes as C
import glob
import tarfile
from typing import Callable, Dict, Iterator, Optional, Sequence

import numpy as np

from . import _lib
from .batching import Batch


def _strings(names: Sequence[str]):
    arr = (C.c_char_p * max(len(names), 1))()
    for i, n in enumerate(names):
        arr[i] = n.encode("utf-8")
    return arr


def _ints(values: Sequence[int]):
    return (C.c_int32 * max(len(values), 1))(*[int(v) for v in values])


class NativeIngest:
    """Parses dataset text into the batch arrays of one :class:`Engine`."""

    def __init__(self, engine, label_name: Optional[str] = None, interleave_names: Optional[Sequence] = None):
        self.lib = _lib.load()
        if interleave_names is None and getattr(engine, "model", None) is not None:
            interleave_names = engine.model.get_interleave_tensors()
        pattern_of = {dst: name for name, dst in (interleave_names or [])}      # destination entity -> sample key
        self.entities = list(engine.entities)
        self.features = list(engine.features)            # (name, entity, size)
        self.adjacencies = list(engine.adjacencies)
        self.label_name = label_name
        e_index = {e: i for i, e in enumerate(self.entities)}
        self._keep = (_strings(self.entities), _strings([f[0] for f in self.features]),
                      _ints([e_index[f[1]] for f in self.features]), _strings([a.name for a in self.adjacencies]),
                      _ints([e_index[a.src] for a in self.adjacencies]), _ints([e_index[a.dst] for a in self.adjacencies]),
                      _ints([1 if a.uses_params else 0 for a in self.adjacencies]))
        k = self._keep
        self.handle = self.lib.ign_ingest_create(len(self.entities), k[0], len(self.features), k[1], k[2],
                                                 len(self.adjacencies), k[3], k[4], k[5], k[6],
                                                 label_name.encode("utf-8") if label_name else None)
        if not self.handle:
            raise RuntimeError(_lib.last_error() or "IGNNITION: ingest_create failed")
        self.sequences = list(engine.sequences)
        a_index = {a.name: i for i, a in enumerate(self.adjacencies)}
        for q in self.sequences:
            if q.interleave and q.dst not in pattern_of:
                raise RuntimeError("IGNNITION: no interleave definition was given for destination " + q.dst)
            rc = self.lib.ign_ingest_add_sequence(self.handle, len(q.adjs), _ints([a_index[a.name] for a in q.adjs]),
                                                  1 if q.interleave else 0,
                                                  pattern_of[q.dst].encode("utf-8") if q.interleave else None)
            if rc < 0:
                raise RuntimeError(_lib.last_error() or "IGNNITION: ingest_add_sequence failed")

    @staticmethod
    def supported(engine) -> bool:
        return True

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.ign_ingest_destroy(h)

    # ------------------------------------------------------------------ parsing
    def reset(self):
        self.lib.ign_ingest_reset(self.handle)

    def parse(self, text, max_samples: int = -1) -> int:
        """Append the samples of ``text`` (bytes / str: one sample object or a list of them)."""
        if isinstance(text, str):
            text = text.encode("utf-8")
        n = self.lib.ign_ingest_parse(self.handle, text, len(text), max_samples)
        if n < 0:
            raise RuntimeError(_lib.last_error() or "IGNNITION: ingest_parse failed")
        return int(n)

    # ------------------------------------------------------------------ results
    def _array(self, ptr, count: int, ctype, dtype):
        if count == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)

    def batch(self, feature_fns: Optional[Dict[str, Callable]] = None,
              label_fn: Optional[Callable] = None) -> Batch:
        """The batch parsed so far.  ``feature_fns[name](array)`` / ``label_fn(array)`` are the user's
        normalisation functions (elementwise: applied to the whole batch at once)."""
        lib, h = self.lib, self.handle
        b = Batch()
        b.n_samples = int(lib.ign_ingest_n_samples(h))
        for i, e in enumerate(self.entities):
            off = self._array(lib.ign_ingest_offsets(h, i), b.n_samples + 1, C.c_int64, np.int64)
            b.offsets[e] = off
            b.num[e] = int(off[-1])
            b.arrays["sample_of_" + e] = np.repeat(np.arange(b.n_samples, dtype=np.int32), np.diff(off))
            b.arrays["offsets_" + e] = off.astype(np.int32)
        for i, (name, ent, size) in enumerate(self.features):
            ptr = C.c_void_p()
            n = lib.ign_ingest_feature(h, i, C.byref(ptr))
            if n != b.num[ent] * size:
                raise RuntimeError("IGNNITION: feature %s has %d values, expected %d x %d" % (name, n, b.num[ent], size))
            x = self._array(ptr, n, C.c_float, np.float32)
            if feature_fns and name in feature_fns:
                x = np.asarray(feature_fns[name](x), dtype=np.float32)
            b.arrays["feat_" + name] = x
        for i, a in enumerate(self.adjacencies):
            ps, pd, pq, pp, w = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32()
            n = lib.ign_ingest_adjacency(h, i, C.byref(ps), C.byref(pd), C.byref(pq), C.byref(pp), C.byref(w))
            b.arrays["src_" + a.name] = self._array(ps, n, C.c_int32, np.int32)
            b.arrays["dst_" + a.name] = self._array(pd, n, C.c_int32, np.int32)
            b.arrays["seq_" + a.name] = seq = self._array(pq, n, C.c_int32, np.int32)
            b.n_edges[a.name] = int(n)
            b.max_seq[a.name] = int(seq.max()) + 1 if n else 0
            if a.uses_params:
                b.arrays["params_" + a.name] = self._array(pp, n * w.value, C.c_float, np.float32).reshape(n, -1)
        for i, q in enumerate(self.sequences):
            po, ps, pc = C.c_void_p(), C.c_void_p(), C.c_void_p()
            n = lib.ign_ingest_sequence(h, i, C.byref(po), C.byref(ps), C.byref(pc))
            b.arrays["pos_off_" + q.key] = self._array(po, b.n_samples + 1, C.c_int32, np.int32)
            b.arrays["pos_src_" + q.key] = self._array(ps, n, C.c_int32, np.int32)
            b.arrays["pos_col_" + q.key] = self._array(pc, n, C.c_int32, np.int32)
        if self.label_name:
            ptr = C.c_void_p()
            n = lib.ign_ingest_labels(h, C.byref(ptr))
            y = self._array(ptr, n, C.c_float, np.float32)
            if label_fn is not None:
                y = np.asarray(label_fn(y), dtype=np.float32)
            b.arrays["labels"] = y
        return b

    # ------------------------------------------------------------------ files
    @staticmethod
    def batches_parallel(engine, directory: str, workers: int = 8, label_name: Optional[str] = None,
                         interleave_names: Optional[Sequence] = None, **kw):
        """One batch per ``*.tar.gz`` of a directory, files parsed concurrently by ``workers`` threads (each
        with its own handle; zlib and the parser both run without the GIL).  Yields in file order."""
        from concurrent.futures import ThreadPoolExecutor
        paths = sorted(glob.glob(str(directory) + "/*.tar.gz"))

        def one(path):
            ing = NativeIngest(engine, label_name, kw.pop("interleave_names", None) if False else interleave_names)
            with tarfile.open(path, "r:gz") as tar:
                ing.parse(tar.extractfile("data.json").read())
            return ing.batch(**kw)

        with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
            for b in pool.map(one, paths):
                yield b

    def batches_from_directory(self, directory: str, batch_size: int, **kw) -> Iterator[Batch]:
        """Batches of ``batch_size`` samples from every ``*.tar.gz`` (``data.json`` inside) of a directory;
        a file's samples never straddle two batches unless the file holds more than ``batch_size``."""
        for path in sorted(glob.glob(str(directory) + "/*.tar.gz")):
            with tarfile.open(path, "r:gz") as tar:
                text = tar.extractfile("data.json").read()
            self.reset()
            # one call per file: the parser walks the top-level array itself
            self.parse(text)
            whole = self.batch(**kw)
            if whole.n_samples <= batch_size:
                yield whole
                continue
            # split by re-parsing windows (rare: files larger than a batch)
            import json
            samples = json.loads(text)
            for i in range(0, len(samples), batch_size):
                self.reset()
                self.parse(json.dumps(samples[i:i + batch_size]))
                yield self.batch(**kw)
