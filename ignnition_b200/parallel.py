"""Multi-GPU execution: one process per GPU.

Two ways the generated model shards:

* **Batches of independent samples** (RouteNet / Q-size): ``model_fn`` loops over samples
  (``code/utils/generate_model.py:712-724``), so contiguous slices of the sample list go to the ranks
  (``shard_samples``).  Inference needs no communication.  Training all-reduces the flat gradient
  buffer (NCCL over NVLink) and every rank scales its squared-error gradient by 1 / GLOBAL prediction
  count, which reproduces ``MeanSquaredError`` over all predictions of the global batch (:745-751).

* **One large graph, partitioned by destination row** (``PartitionedEngine``): rank r owns the rows
  ``[bounds[r], bounds[r + 1])`` of every entity -- their in-edges (CSR rows) and the authoritative
  state of those nodes -- and runs the message-passing loop of ``ComnetModel.call``
  (``generate_model.py:405-602``) on them.  Source states are needed from every rank, so every rank
  holds a full copy of each entity's state array; after an update the owners' new rows have to reach
  all copies.  Four exchanges are built:

  ``copy``      (default) the update kernel runs over the owned rows in a few row chunks, each chunk
                writes this rank's own array, and as soon as a chunk is done the COPY ENGINE pushes it
                into every peer's mapped array (``ign_peer_copy`` = cudaMemcpyAsync on a side stream,
                peers visited in rotated order so that no receiver has two senders at once) while the
                SMs are already gathering the next chunk.  Measured (tools/p2p_rate.cu,
                profiles/r2_p2p_rate.md): a copy-engine push runs at 735-764 GB/s next to a kernel that
                saturates HBM and does not slow it; stores issued by the busy SMs themselves (TMA or
                st.global alike) crawl at 30-70 GB/s.
  ``peer``      the update kernel itself (``ign_agg_gru_cell_tc``) stores every finished 128-row tile
                into the state array of EVERY rank: the arrays are cudaMalloc'ed by the library and
                mapped into all processes through CUDA IPC (``PeerBuffer``), the stores are TMA tensor
                stores over NVLink issued from the epilogue while the gather warps reduce the next
                tile.  No separate collective runs; one tiny NCCL all-reduce per message passing is
                the barrier that orders "all tiles landed" before the next gather.  Kept as the measured
                counter-example: the stores starve behind the SM's own gather traffic (12.8 ms per
                iteration at 2 GPUs against 8.3 ms of ``nccl``).
  ``boundary``  only the rows a peer's edges actually read are sent (lists built once): what a graph with
                locality needs.  The rows for all peers are packed into one contiguous block (one
                ``ign_gather_concat``), the copy engine moves each peer's slice into that peer's inbox
                (``ign_peer_copy``), and after the barrier the receiver scatters its inbox
                (``ign_rows_unpack``): two kernels and world - 1 copies per update, whatever the number of rows.
  ``nccl``      the baseline: the kernel writes the owner's rows locally, then
                ``ncclAllGather`` (``torch.distributed.all_gather_into_tensor``) fills the rest.

  The edge list may arrive in any sharding: ``route_edges`` sends every edge to the owner of its
  destination (stable, so the slot order of a destination -- and with it the order of the fp32 sum --
  is the order of the global list, whatever the number of ranks) and the owner builds its CSR with the
  device radix sort.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous slices of the sample list, balanced by cost (e.g. edges per sample), one per rank."""
    n = len(costs)
    total = float(sum(costs))
    bounds, start, acc = [], 0, 0.0
    for r in range(world):
        target = total * (r + 1) / world
        end = start
        while end < n and (acc + costs[end] <= target + 1e-9 or end == start) and (n - end) > (world - 1 - r):
            acc += costs[end]
            end += 1
        if r == world - 1:
            end = n
        bounds.append((start, end))
        start = end
    return bounds


def shard_samples(samples: Sequence, rank: int, world: int, costs: Sequence[float] = None) -> List:
    costs = costs if costs is not None else [1.0] * len(samples)
    lo, hi = shard_bounds(costs, world)[rank]
    return list(samples[lo:hi])


# --------------------------------------------------------------------------------- partitioned graph
def node_bounds(n: int, world: int) -> List[int]:
    """Row bounds of the destination partition: rank r owns [b[r], b[r + 1]), sizes differ by at most one row.
    (The update kernel tiles every rank's rows from its own first row and the TMA unit clips the last tile at
    the owner's last row, so the bounds need no alignment.)"""
    return [(r * n) // world for r in range(world + 1)]


def chunk_cuts(n_rows: int, chunks: int, growth: float, tile: int = 128) -> List[int]:
    """Row cuts [0, ..., n_rows] of the 'copy' exchange: at most ``chunks`` chunks on ``tile``-row boundaries whose sizes
    form a geometric series of ratio ``growth`` (> 1: small first chunk, the copy engine starts early; < 1: small last
    chunk, little exchange left after the last kernel)."""
    tiles = -(-n_rows // tile)
    k = max(1, min(int(chunks), tiles))
    weights = [float(growth) ** c for c in range(k)]
    acc, marks = 0.0, [0]
    for wgt in weights:
        acc += wgt
        marks.append(int(round(tiles * acc / sum(weights))))
    marks[-1] = tiles
    return sorted({min(n_rows, m * tile) for m in marks} | {n_rows})


def exchange_growth(world: int, row_bytes: int, edges_per_row: float, f_in: int,
                    peer_gbs: float = 745.0, kernel_gbs: float = 4700.0) -> float:
    """(push time per row) / (kernel time per row) from the measured rates (profiles/r2_p2p_rate.md, r2_agg_gru.md):
    copy engine 745 GB/s out per GPU, fused update 4.7 TB/s of algorithmic bytes; clamped to [0.25, 3]."""
    t_push = (world - 1) * row_bytes / (peer_gbs * 1e9)
    t_kern = (edges_per_row * (4 + 4 * f_in) + 2 * row_bytes + 4) / (kernel_gbs * 1e9)
    return min(3.0, max(0.25, t_push / t_kern))


def split_counts(owner_rowptr: Sequence[int]) -> List[int]:
    return [int(owner_rowptr[r + 1]) - int(owner_rowptr[r]) for r in range(len(owner_rowptr) - 1)]


def _dist():
    import torch.distributed as dist
    return dist


def _world(group) -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class _CudaArray:
    """A raw device allocation seen through ``__cuda_array_interface__`` (torch.as_tensor wraps it)."""

    def __init__(self, ptr: int, shape: Tuple[int, ...]):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerBuffer:
    """``[rows, width]`` fp32 array allocated by the library on this GPU and mapped into every rank.

    ``ptrs[r]`` is the address of rank r's array in THIS process (``ptrs[rank]`` is the local one);
    ``tensor`` is the local array as a torch tensor (a view, no copy)."""

    def __init__(self, rows: int, width: int, device, group=None):
        import torch
        from . import _lib
        lib = _lib.load()
        dist = _dist()
        self.rows, self.width, self.group = int(rows), int(width), group
        self.rank, self.world = _world(group)
        self.nbytes = max(self.rows * self.width * 4, 256)
        p = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.ign_peer_alloc(self.nbytes, C.byref(p)), "peer_alloc")
            self.ptr = int(p.value)
            self.ptrs: List[int] = [0] * self.world
            self.ptrs[self.rank] = self.ptr
            self._opened: List[int] = []
            if self.world > 1:
                h = C.create_string_buffer(64)
                _lib.check(lib.ign_peer_export(C.c_void_p(self.ptr), h), "peer_export")
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(h.raw), group=group)
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    q = C.c_void_p()
                    _lib.check(lib.ign_peer_open(C.create_string_buffer(handles[r], 64), C.byref(q)), "peer_open")
                    self.ptrs[r] = int(q.value)
                    self._opened.append(int(q.value))
            self.tensor = torch.as_tensor(_CudaArray(self.ptr, (self.rows, self.width)), device=device)
        self.device = device

    def close(self):
        """Unmap the peers' arrays and free the local one (collective: every rank must call it)."""
        import torch
        from . import _lib
        if self.ptr == 0:
            return
        lib = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for q in self._opened:
                lib.ign_peer_close(C.c_void_p(q))
            self._opened = []
            if self.world > 1:
                _dist().barrier(group=self.group)       # nobody frees while a peer still has it mapped
            self.tensor = None
            lib.ign_peer_free(C.c_void_p(self.ptr))
            self.ptr = 0


def route_edges(src, dst, bounds: Sequence[int], group=None):
    """Send every edge (global ``src`` / ``dst`` row ids, int32 device tensors: this rank's shard of the
    global list, any order) to the rank that owns its destination.  Stable: a destination's edges
    arrive ordered by (sending rank, position in that rank's shard), i.e. in the order of the global
    list when the shards are its contiguous pieces.  SURVEY.md section 8e, "CSR build"."""
    import torch
    from . import ops
    rank, world = _world(group)
    if world == 1:
        return src, dst
    dist = _dist()
    owner = ops.edge_owner(dst, bounds)
    # stable counting sort of the shard by owner rank: the radix-sort adjacency builder with W "rows"
    rowptr, src_sorted, perm, _ = ops.csr_build(owner, src, None, world, ops.CSR_SORT, want_perm=True)
    dst_sorted = ops.gather_int(dst, perm)
    send = split_counts(rowptr.cpu().tolist())
    send_t = torch.tensor(send, dtype=torch.int64, device=src.device)
    recv_t = torch.empty_like(send_t)
    dist.all_to_all_single(recv_t, send_t, group=group)
    recv = [int(v) for v in recv_t.cpu().tolist()]
    out_src = torch.empty(sum(recv), dtype=torch.int32, device=src.device)
    out_dst = torch.empty(sum(recv), dtype=torch.int32, device=src.device)
    dist.all_to_all_single(out_src, src_sorted, recv, send, group=group)
    dist.all_to_all_single(out_dst, dst_sorted, recv, send, group=group)
    return out_src, out_dst


class PartitionedEngine:
    """The message-passing loop of one large graph on the rows this rank owns (module docstring).

    Supports the message passings whose sources send their states (``direct_assignation``) into a
    sum / mean / max aggregation with a GRU update -- BASELINE config 5 and every model of that
    family; anything else raises.  ``engine`` supplies the model, the weights and the kernels."""

    EXCHANGES = ("copy", "peer", "boundary", "nccl")

    def __init__(self, engine, group=None, exchange: str = "copy", chunks: Optional[int] = None):
        from . import ops
        if exchange not in self.EXCHANGES:
            raise RuntimeError("IGNNITION: unknown exchange '%s' (one of %s)" % (exchange, ", ".join(self.EXCHANGES)))
        self.engine, self.group, self.exchange = engine, group, exchange
        self.rank, self.world = _world(group)
        if self.world > ops.MAX_PEERS:
            raise RuntimeError("IGNNITION: a partitioned graph spans at most %d GPUs of one node" % ops.MAX_PEERS)
        self.plans = []
        for stage in engine.plans:
            for p in stage:
                if (p.kind != "agg_gru" or len(p.adjs) != 1 or p.conv or p.attn or p.v1 or
                        any(op.type == "feed_forward_nn" for s in p.mp.source_entities for op in s.message_formation)):
                    raise RuntimeError("IGNNITION: partitioned graphs run sum / mean / max aggregations of source "
                                       "states with a recurrent update; message passing to '%s' is not of that "
                                       "kind" % p.dst)
                self.plans.append(p)
        self.bounds: Dict[str, List[int]] = {}
        self.num_global: Dict[str, int] = {}
        self.csr: Dict[str, tuple] = {}
        self.full: Dict[str, List[PeerBuffer]] = {}
        self.cur: Dict[str, int] = {}
        self.send_rows: Dict[str, List[Optional[object]]] = {}
        self.recv_rows: Dict[str, int] = {}         # boundary exchange: rows this rank receives per update
        self.bnd: Dict[str, dict] = {}              # boundary exchange: packed send / receive lists, inboxes, offsets
        self.n_edges: Dict[str, int] = {}
        self.exchanged_bytes = 0          # bytes this rank received over NVLink in message_passing()
        self._flag = None
        # 'copy' exchange: row chunks per update (each a kernel launch followed by world - 1 copy-engine pushes)
        self.chunks = int(chunks if chunks is not None else os.environ.get("IGN_EXCHANGE_CHUNKS", "4"))
        # one copy stream: pushes to two peers at once collide at the receivers (6.1 vs 4.1 ms at 8 GPUs)
        self.copy_streams = int(os.environ.get("IGN_EXCHANGE_STREAMS", "1"))
        # chunk sizes form a geometric series with ratio (push time per row) / (kernel time per row), estimated per
        # update in _mp_copy: where NVLink is the longer leg (8 GPUs: ratio 2) a small first chunk starts the copy engine
        # early and chunk c + 1 is ready just before the push of chunk c ends; where the kernel is (2 GPUs: ratio 0.3)
        # the LAST chunk is the small one and little of the exchange is left after the last kernel.  0 = estimate.
        self.growth = float(os.environ.get("IGN_EXCHANGE_GROWTH", "0"))
        self._copy_streams: list = []

    # ------------------------------------------------------------------ build
    def own(self, entity: str) -> Tuple[int, int]:
        b = self.bounds[entity]
        return b[self.rank], b[self.rank + 1]

    def build(self, num_global: Dict[str, int], edges: Dict[str, tuple], feats: Dict[str, object]):
        """``num_global``: rows per entity of the WHOLE graph.  ``edges[adjacency] = (src, dst)``: this
        rank's shard of the global edge list (int32 device tensors, global row ids).  ``feats[feature]``:
        the feature values of the rows this rank owns."""
        import torch
        from . import ops
        eng = self.engine
        dev = eng.device
        self.close()
        self.num_global = {e: int(num_global[e]) for e in eng.entities}
        self.bounds = {e: node_bounds(self.num_global[e], self.world) for e in eng.entities}
        for a in eng.adjacencies:
            src, dst = edges[a.name]
            src, dst = route_edges(src, dst, self.bounds[a.dst], self.group)
            lo, hi = self.own(a.dst)
            dst_local = ops.gather_int(dst, None, add=-lo)
            rowptr, col, _, _ = ops.csr_build(dst_local, src, None, hi - lo, ops.CSR_SORT)
            self.csr[a.name] = (rowptr, col)
            self.n_edges[a.name] = int(src.numel())
        n_buf = 2
        for e in eng.entities:
            self.full[e] = [PeerBuffer(self.num_global[e], eng.hidden[e], dev, self.group) for _ in range(n_buf)]
            self.cur[e] = 0
        # h0 of the owned rows (Entity.calculate_hs), then one all-gather so every copy starts complete
        for ent in eng.model.get_entities():
            lo, hi = self.own(ent.name)
            own_view = self.full[ent.name][0].tensor[lo:hi]
            ops.init_state([feats[f.name] for f in ent.features], [f.size for f in ent.features], hi - lo,
                           ent.hidden_state_dimension, out=own_view)
            self._all_gather(ent.name, 0)
        if self.exchange == "boundary":
            self._build_boundary_lists()
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self._barrier()
        return self

    def _all_gather(self, entity: str, buf: int):
        """ncclAllGather of the owners' rows into every rank's copy (in place)."""
        if self.world == 1:
            return
        dist = _dist()
        b = self.bounds[entity]
        t = self.full[entity][buf].tensor
        sizes = [b[r + 1] - b[r] for r in range(self.world)]
        lo, hi = self.own(entity)
        if len(set(sizes)) == 1:
            dist.all_gather_into_tensor(t, t[lo:hi], group=self.group)
        else:
            dist.all_gather([t[b[r]:b[r + 1]] for r in range(self.world)], t[lo:hi], group=self.group)

    def _build_boundary_lists(self):
        """send_rows[entity][p]: the rows this rank owns that rank p's edges read (global ids)."""
        import torch
        from . import ops
        dist = _dist()
        dev = self.engine.device
        for e in self.engine.entities:
            self.send_rows[e] = [None] * self.world
        if self.world == 1:
            return
        needed = {e: torch.zeros(self.num_global[e], dtype=torch.int32, device=dev) for e in self.engine.entities}
        for a in self.engine.adjacencies:
            ops.mark_rows(self.csr[a.name][1], needed[a.src])
        for e in self.engine.entities:
            b = self.bounds[e]
            lists, counts = [], []
            for r in range(self.world):
                if r == self.rank:
                    lists.append(torch.empty(0, dtype=torch.int32, device=dev))
                    counts.append(0)
                    continue
                rows, cnt = ops.flag_compact(needed[e][b[r]:b[r + 1]], add=b[r])
                c = int(cnt.item())
                lists.append(rows[:c])
                counts.append(c)
            self.recv_rows[e] = sum(counts)
            send_t = torch.tensor(counts, dtype=torch.int64, device=dev)
            recv_t = torch.empty_like(send_t)
            dist.all_to_all_single(recv_t, send_t, group=self.group)
            recv = [int(v) for v in recv_t.cpu().tolist()]
            want = torch.cat(lists) if sum(counts) else torch.empty(0, dtype=torch.int32, device=dev)
            got = torch.empty(sum(recv), dtype=torch.int32, device=dev)
            dist.all_to_all_single(got, want, recv, counts, group=self.group)
            off = 0
            for p in range(self.world):
                self.send_rows[e][p] = got[off:off + recv[p]].contiguous() if recv[p] else None
                off += recv[p]
            # packed exchange: `got` is the concatenation (by peer) of the rows this rank sends, `want` the concatenation
            # (by owner) of the rows it receives = the layout of its inbox; every sender needs the offset of its segment
            # in every receiver's inbox
            recv_off = [0]
            for c_ in counts:
                recv_off.append(recv_off[-1] + c_)
            table = torch.tensor(recv_off[:-1], dtype=torch.int64, device=dev)
            all_off = [torch.empty_like(table) for _ in range(self.world)]
            dist.all_gather(all_off, table, group=self.group)
            off_at = [int(all_off[p][self.rank].item()) for p in range(self.world)]
            send_off = [0]
            for c_ in recv:
                send_off.append(send_off[-1] + c_)
            width = self.engine.hidden[e]
            self.bnd[e] = {"send_all": got, "send_off": send_off, "send_cnt": recv, "recv_all": want, "off_at": off_at,
                           "inbox": [PeerBuffer(max(sum(counts), 1), width, dev, self.group) for _ in range(2)],
                           "turn": 0}

    def _barrier(self):
        """Orders "every rank's stores have landed" before the next reads: a 4-byte NCCL all-reduce on the
        compute stream (kernel completion makes the peer stores visible; NCCL orders the ranks)."""
        if self.world > 1:
            _dist().all_reduce(self._flag, group=self.group)

    # ------------------------------------------------------------------ run
    def state(self, entity: str):
        """The rows of ``entity`` this rank owns, current values."""
        lo, hi = self.own(entity)
        return self.full[entity][self.cur[entity]].tensor[lo:hi]

    def full_state(self, entity: str):
        return self.full[entity][self.cur[entity]].tensor

    def _mp(self, p):
        from . import ops
        eng = self.engine
        a = p.adjs[0]
        rowptr, col = self.csr[a.name]
        dst = p.dst
        lo, hi = self.own(dst)
        cur, nxt = self.cur[dst], self.cur[dst] ^ 1
        src_states = self.full[a.src][self.cur[a.src]].tensor
        h = self.full[dst][cur].tensor[lo:hi]
        K, R, B = eng.param(dst + "_update/kernel"), eng.param(dst + "_update/recurrent_kernel"), eng.param(dst + "_update/bias")
        bufs = self.full[dst][nxt]
        fused = ops.agg_gru_cell_tc_supported(p.msg_dim, eng.hidden[dst])
        if self.exchange == "copy" and self.world > 1:
            self._mp_copy(p, rowptr, col, src_states, h, K, R, B, bufs, lo, hi, fused)
            self.cur[dst] = nxt
            return
        if self.exchange == "peer" and not fused and self.world > 1:
            raise RuntimeError("IGNNITION: the peer exchange needs the fused update kernel (message width == units "
                               "in {32, 64}); use exchange='nccl' for %d -> %d" % (p.msg_dim, eng.hidden[dst]))
        if hi > lo:
            if fused:
                outs = ([bufs.ptrs[self.rank]] + [bufs.ptrs[r] for r in range(self.world) if r != self.rank]
                        if self.exchange == "peer" else [bufs.ptrs[self.rank]])
                ops.agg_gru_cell_tc(p.op, rowptr, col, src_states, h, K, R, B, outs, out_row0=lo)
            else:
                agg = ops.segment_reduce(p.op, rowptr, col, src_states)
                ops.gru_cell(agg, h, K, R, B, out=bufs.tensor[lo:hi])
        if self.world > 1:
            width = eng.hidden[dst]
            if self.exchange == "nccl":
                self._all_gather(dst, nxt)
                self.exchanged_bytes += (self.num_global[dst] - (hi - lo)) * width * 4
            elif self.exchange == "boundary":
                bd = self.bnd[dst]
                inbox = bd["inbox"][bd["turn"]]
                bd["turn"] ^= 1
                n_send = int(bd["send_all"].numel())
                if n_send:
                    packed = ops.gather_concat([bufs.tensor], [bd["send_all"]], n_send)
                    for j in range(1, self.world):
                        r = (self.rank + j) % self.world
                        cnt = bd["send_cnt"][r]
                        if cnt:
                            ops.peer_copy(inbox.ptrs[r] + bd["off_at"][r] * width * 4,
                                          packed.data_ptr() + bd["send_off"][r] * width * 4, cnt * width * 4)
                self.exchanged_bytes += self.recv_rows.get(dst, 0) * width * 4
                self._barrier()
                if int(bd["recv_all"].numel()):
                    ops.rows_unpack(inbox.tensor, bd["recv_all"], bufs.tensor)
            else:
                self.exchanged_bytes += (self.num_global[dst] - (hi - lo)) * width * 4
                self._barrier()
        self.cur[dst] = nxt

    def _mp_copy(self, p, rowptr, col, src_states, h, K, R, B, bufs, lo, hi, fused):
        """One update with the 'copy' exchange: kernel launches over row chunks on the compute stream; after each,
        the copy stream pushes the chunk's rows into every peer's array.  The next gather may start when every
        rank's pushes have landed: the compute stream waits for the copy stream, then the NCCL barrier."""
        import torch
        from . import ops
        width = self.engine.hidden[p.dst]
        n = hi - lo
        while len(self._copy_streams) < max(1, self.copy_streams):
            self._copy_streams.append(torch.cuda.Stream(device=self.engine.device))
        streams = self._copy_streams[:max(1, self.copy_streams)]
        row_bytes = width * 4
        growth = self.growth
        if growth <= 0.0:
            growth = exchange_growth(self.world, row_bytes, float(col.numel()) / max(n, 1), int(src_states.shape[1]))
        cuts = chunk_cuts(n, self.chunks, growth)
        trace = getattr(self, "trace", None)             # profiling: [(label, start event, end event)] of one update
        def mark(stream=None):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream) if stream is not None else ev.record()
            return ev
        for a, b in zip(cuts, cuts[1:]):
            if b <= a:
                continue
            k0 = mark() if trace is not None else None
            if fused:
                ops.agg_gru_cell_tc(p.op, rowptr[a:b + 1], col, src_states, h[a:b], K, R, B, [bufs.ptr], out_row0=lo + a)
            else:
                agg = ops.segment_reduce(p.op, rowptr[a:b + 1], col, src_states)
                ops.gru_cell(agg, h[a:b], K, R, B, out=bufs.tensor[lo + a:lo + b])
            done = mark() if trace is not None else torch.cuda.Event()
            if trace is None:
                done.record()
            else:
                trace.append(("kernel rows %d-%d" % (a, b), k0, done))
            off = (lo + a) * row_bytes
            for cs in streams:
                cs.wait_event(done)
            c0 = mark(streams[0]) if trace is not None else None
            for j in range(1, self.world):
                r = (self.rank + j) % self.world
                with torch.cuda.stream(streams[(j - 1) % len(streams)]):
                    ops.peer_copy(bufs.ptrs[r] + off, bufs.ptr + off, (b - a) * row_bytes)
            if trace is not None:
                trace.append(("push rows %d-%d" % (a, b), c0, mark(streams[0])))
        for cs in streams:
            pushed = torch.cuda.Event()
            pushed.record(cs)
            torch.cuda.current_stream().wait_event(pushed)
        self.exchanged_bytes += (self.num_global[p.dst] - n) * row_bytes
        b0 = mark() if trace is not None else None
        self._barrier()
        if trace is not None:
            trace.append(("barrier", b0, mark()))

    def exchange_only(self, entity: str):
        """Profiling aid: push this rank's rows of the current array to every peer (the values are already there) and
        run the barrier -- the exchange of one update without its kernel."""
        import torch
        from . import ops
        if self.world == 1:
            return
        lo, hi = self.own(entity)
        bufs = self.full[entity][self.cur[entity]]
        row_bytes = self.engine.hidden[entity] * 4
        for j in range(1, self.world):
            r = (self.rank + j) % self.world
            ops.peer_copy(bufs.ptrs[r] + lo * row_bytes, bufs.ptr + lo * row_bytes, (hi - lo) * row_bytes)
        self._barrier()

    def message_passing(self, iterations: Optional[int] = None):
        T = self.engine.T if iterations is None else iterations
        for _ in range(T):
            for p in self.plans:
                self._mp(p)                  # written back at once (generate_model.py:602)
        return {e: self.state(e) for e in self.engine.entities}

    def readout(self):
        """Predictions for the rows of the output entity this rank owns."""
        return self.engine.readout_forward({e: self.state(e) for e in self.engine.entities})

    def forward(self):
        self.message_passing()
        return self.readout()

    def checksum(self) -> Dict[str, float]:
        """Per-entity sum of |h| and a position-weighted sum over ALL rows of the graph (fp64, all-reduced):
        equal for every partitioning of the same graph when the states are."""
        import torch
        out = {}
        for e in self.engine.entities:
            lo, hi = self.own(e)
            h = self.state(e).double()
            w = (torch.arange(lo, hi, device=h.device, dtype=torch.float64) % 1009.0 + 1.0).unsqueeze(1)
            v = torch.stack([h.abs().sum(), (h * w).sum()])
            if self.world > 1:
                _dist().all_reduce(v, group=self.group)
            out[e] = [float(v[0].item()), float(v[1].item())]
        return out

    def close(self):
        for bufs in self.full.values():
            for b in bufs:
                b.close()
        for bd in self.bnd.values():
            for b in bd["inbox"]:
                b.close()
        self.full = {}
        self.bnd = {}
        self.csr = {}
