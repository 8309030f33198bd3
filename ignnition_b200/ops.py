"""Thin torch-tensor wrappers over the C-ABI (``include/ignnition_b200.h``).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; every function below
passes raw device pointers (``tensor.data_ptr()``) and the stream handle to the library and checks
the status.  No torch op computes anything on this path and nothing falls back to the CPU.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib

OP_SUM, OP_MEAN, OP_MAX, OP_SUM_ADD = 0, 1, 2, 3
CSR_SORT, CSR_RANK = 0, 1
STEP_SRC_SHIFT = 28
MAX_PEERS = 8               # IGN_MAX_PEERS: state arrays one fused update can write
ACTIVATIONS = {None: 0, "None": 0, "linear": 0, "relu": 1, "selu": 2, "sigmoid": 3, "tanh": 4,
               "elu": 5, "softplus": 6, "leaky_relu": 7}
ACT_FROM_OUTPUT = 0x100      # dense_bwd: the `pre_act` argument holds the layer's output (IGN_ACT_FROM_OUTPUT)


def set_tensor_cores(enable: bool) -> bool:
    """Process-wide switch between the tcgen05 (3xTF32) kernels and their fp32 CUDA-core twins."""
    return bool(_lib.load().ign_set_tensor_cores(1 if enable else 0))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None, name="tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("IGNNITION: %s must be a CUDA tensor (no CPU path exists)" % name)
    if not t.is_contiguous():
        raise RuntimeError("IGNNITION: %s must be contiguous" % name)
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's stream: an Engine on another GPU must be driven under
        # torch.cuda.device(engine.device) (peer-mapped buffers are passed as raw addresses, not tensors)
        raise RuntimeError("IGNNITION: %s lives on cuda:%d but the current device is cuda:%d; wrap the call in "
                           "torch.cuda.device(...)" % (name, t.device.index, torch.cuda.current_device()))
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("IGNNITION: %s must be %s, got %s" % (name, dtype, t.dtype))
    return t.data_ptr()


def _f(t, name="tensor"):
    return _ptr(t, torch.float32, name)


def _i(t, name="tensor"):
    return _ptr(t, torch.int32, name)


def _ptr_array(ts: Sequence[Optional[torch.Tensor]], dtype):
    arr = (C.c_void_p * max(len(ts), 1))()
    for k, t in enumerate(ts):
        arr[k] = _ptr(t, dtype, "tensor list element")
    return arr


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------- adjacency
def csr_build(dst: torch.Tensor, src: torch.Tensor, seq: Optional[torch.Tensor], num_dst: int,
              mode: int = CSR_SORT, want_perm: bool = False, want_status: bool = False):
    """(rowptr[num_dst+1], col[E], perm[E]|None, status[2]|None) -- see ign_csr_build."""
    lib = _lib.load()
    E = dst.numel()
    dev = dst.device
    rowptr = torch.empty(num_dst + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    perm = torch.empty(E, dtype=torch.int32, device=dev) if want_perm else None
    status = torch.empty(2, dtype=torch.int32, device=dev) if want_status else None
    nbytes = lib.ign_csr_build_ws_bytes(E, num_dst)
    ws = _workspace(nbytes, dev)
    _lib.check(lib.ign_csr_build(_i(dst, "dst"), _i(src, "src"), _i(seq, "seq"), E, num_dst, mode,
                                 _i(rowptr), _i(col), _i(perm), _i(status), ws.data_ptr(), ws.numel(),
                                 _stream()), "csr_build")
    return rowptr, col, perm, status


def csr_build_small(specs):
    """CSR by destination of every adjacency of a small graph in ONE launch (ign_csr_build_small).
    ``specs`` = [(dst, src, seq | None, num_dst, want_perm)]; returns [(rowptr, col, perm | None)]."""
    lib = _lib.load()
    dev = specs[0][0].device
    out = []
    for dst, src, seq, num_dst, want_perm in specs:
        E = dst.numel()
        out.append((torch.empty(num_dst + 1, dtype=torch.int32, device=dev),
                    torch.empty(E, dtype=torch.int32, device=dev),
                    torch.empty(E, dtype=torch.int32, device=dev) if want_perm else None))
    n = len(specs)
    nz = lambda t: t if t is not None and t.numel() else None
    _lib.check(lib.ign_csr_build_small(
        n, _ptr_array([nz(s_[0]) for s_ in specs], torch.int32), _ptr_array([nz(s_[1]) for s_ in specs], torch.int32),
        _ptr_array([nz(s_[2]) for s_ in specs], torch.int32), (C.c_int64 * n)(*[int(s_[0].numel()) for s_ in specs]),
        (C.c_int64 * n)(*[int(s_[3]) for s_ in specs]), _ptr_array([o[0] for o in out], torch.int32),
        _ptr_array([nz(o[1]) for o in out], torch.int32), _ptr_array([nz(o[2]) for o in out], torch.int32), _stream()),
        "csr_build_small")
    return out


def length_order(rowptr: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    n = rowptr.numel() - 1
    order = torch.empty(n, dtype=torch.int32, device=rowptr.device)
    ws = _workspace(lib.ign_length_order_ws_bytes(n), rowptr.device)
    _lib.check(lib.ign_length_order(_i(rowptr), n, _i(order), ws.data_ptr(), ws.numel(), _stream()),
               "length_order")
    return order


def steps_build(rowptrs: List[torch.Tensor], cols: List[torch.Tensor], dst_sample: Optional[torch.Tensor],
                pos_off: torch.Tensor, pos_src: torch.Tensor, pos_col: torch.Tensor, num_dst: int,
                total_steps: int):
    """Step table of a multi-source ordered / interleave aggregation (ign_steps_build).
    ``total_steps`` = sum of all sources' edge counts (known on the host)."""
    lib = _lib.load()
    dev = rowptrs[0].device
    steps_rowptr = torch.empty(num_dst + 1, dtype=torch.int32, device=dev)
    steps = torch.empty(max(total_steps, 1), dtype=torch.int32, device=dev)
    ws = _workspace(lib.ign_steps_build_ws_bytes(num_dst), dev)
    rp = _ptr_array(rowptrs, torch.int32)
    cp = _ptr_array(cols, torch.int32)
    _lib.check(lib.ign_steps_build(len(rowptrs), rp, cp, _i(dst_sample), _i(pos_off), _i(pos_src),
                                   _i(pos_col), num_dst, _i(steps_rowptr), _i(steps), ws.data_ptr(),
                                   ws.numel(), _stream()), "steps_build")
    return steps_rowptr, steps


def steps_keys(steps: torch.Tensor, src_id: int, n_rows: int) -> torch.Tensor:
    lib = _lib.load()
    keys = torch.empty_like(steps)
    _lib.check(lib.ign_steps_keys(_i(steps), steps.numel(), src_id, n_rows, _i(keys), _stream()), "steps_keys")
    return keys


# ------------------------------------------------------------------------------- forward
def init_state(feats: List[torch.Tensor], sizes: List[int], n: int, hidden: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    dev = feats[0].device if feats else (out.device if out is not None else torch.device("cuda"))
    if out is None:
        out = torch.empty(n, hidden, dtype=torch.float32, device=dev)
    fp = _ptr_array(feats, torch.float32)
    sz = (C.c_int32 * max(len(sizes), 1))(*sizes)
    _lib.check(lib.ign_init_state(len(feats), fp, sz, n, hidden, _f(out), _stream()), "init_state")
    return out


def segment_reduce(op: int, rowptr, col, src_states, out: Optional[torch.Tensor] = None):
    lib = _lib.load()
    n = rowptr.numel() - 1
    F = src_states.shape[1]
    if out is None:
        out = torch.empty(n, F, dtype=torch.float32, device=src_states.device)
    _lib.check(lib.ign_segment_reduce(op, _i(rowptr), _i(col), _f(src_states), F, n, _f(out), _stream()),
               "segment_reduce")
    return out


GRU_FUSED_SHAPES = {(16, 16), (32, 32), (64, 64), (16, 32), (64, 32), (32, 64), (32, 16)}     # IGN_GRU_DISPATCH


def gru_cell(x, h, kernel, rkernel, bias, out: Optional[torch.Tensor] = None, tensor_cores: bool = True):
    """One GRU step.  3xTF32 on tcgen05 when built for the shape (and out is not h), else fp32 FMA; shapes the
    fused kernels do not cover (hidden_state_dimension is free in the reference's schema) run as two Dense GEMMs and
    one element-wise kernel."""
    lib = _lib.load()
    n, units = h.shape
    if out is None:
        out = torch.empty_like(h)
    if bias.dim() == 1:          # GRUCell(reset_after=False): bias [3 units], see gru_cell_v1
        return gru_cell_v1(x, h, kernel, rkernel, bias, out)
    if (int(x.shape[1]), int(units)) not in GRU_FUSED_SHAPES:
        zx = dense(x, kernel, bias[0], 0)
        zh = dense(h, rkernel, bias[1], 0)
        _lib.check(lib.ign_gru_gates_fwd(_f(zx), _f(zh), _f(h), n, units, _f(out), _stream()), "gru_gates_fwd")
        return out
    ws, nbytes = None, 0
    if tensor_cores and out.data_ptr() != h.data_ptr():
        nbytes = lib.ign_gru_cell_ws_bytes(x.shape[1], units)
        if nbytes:
            ws = _workspace(nbytes, x.device)
    _lib.check(lib.ign_gru_cell(_f(x), _f(h), n, x.shape[1], units, _f(kernel), _f(rkernel), _f(bias),
                                _f(out), ws.data_ptr() if ws is not None else None, nbytes, _stream()),
               "gru_cell")
    return out


def gru_cell_v1(x, h, kernel, rkernel, bias, out=None):
    """One step of GRUCell(reset_after=False) (bias [3 units]): the reset gate multiplies h before the candidate's
    recurrent product, so the step is three Dense products around two element-wise kernels (ign_gru_v1_*)."""
    lib = _lib.load()
    n, u = h.shape
    if out is None:
        out = torch.empty_like(h)
    zx = dense(x, kernel, bias, 0)
    zh2 = dense(h, slice_cols(rkernel, 0, 2 * u), None, 0)
    rh = torch.empty_like(h)
    _lib.check(lib.ign_gru_v1_reset(_f(zx), _f(zh2), _f(h), n, u, _f(rh), _stream()), "gru_v1_reset")
    zhh = dense(rh, slice_cols(rkernel, 2 * u, u), None, 0)
    _lib.check(lib.ign_gru_v1_out(_f(zx), _f(zh2), _f(zhh), _f(h), n, u, _f(out), _stream()), "gru_v1_out")
    return out


def gru_cell_bwd_v1(x, h, kernel, rkernel, bias, d_out, dx, dh, dk, drk, db):
    """Backward of gru_cell_v1: the gates recomputed, ign_gru_v1_bwd_out / _reset for the gate gradients, ign_dense_bwd
    for the three products; dk / drk / db accumulate."""
    lib = _lib.load()
    n, u = h.shape
    r_zr, r_h = slice_cols(rkernel, 0, 2 * u), slice_cols(rkernel, 2 * u, u)
    zx = dense(x, kernel, bias, 0)
    zh2 = dense(h, r_zr, None, 0)
    rh = torch.empty_like(h)
    _lib.check(lib.ign_gru_v1_reset(_f(zx), _f(zh2), _f(h), n, u, _f(rh), _stream()), "gru_v1_reset")
    zhh = dense(rh, r_h, None, 0)
    direct = torch.empty_like(h)
    _lib.check(lib.ign_gru_v1_bwd_out(_f(zx), _f(zh2), _f(zhh), _f(h), _f(d_out), n, u, _f(direct), _stream()),
               "gru_v1_bwd_out")
    d_rh = torch.empty_like(h)
    d_r_h = torch.zeros_like(r_h)
    dense_bwd(rh, r_h, 0, None, zhh, d_rh, d_r_h, None)
    _lib.check(lib.ign_gru_v1_bwd_reset(_f(zx), _f(zh2), _f(h), _f(d_rh), n, u, _f(direct), _stream()),
               "gru_v1_bwd_reset")
    d_r_zr = torch.zeros_like(r_zr)
    dense_bwd(x, kernel, 0, None, zx, dx, dk, db)
    dense_bwd(h, r_zr, 0, None, zh2, dh, d_r_zr, None)
    axpy(1.0, direct, dh)
    axpy(1.0, gather_concat([d_r_zr, d_r_h], [None, None], u), drk)


def agg_gru_cell(rowptr, col, src_states, h_dst, kernel, rkernel, bias, out=None, agg_out=None):
    lib = _lib.load()
    n, units = h_dst.shape
    if out is None:
        out = torch.empty_like(h_dst)
    _lib.check(lib.ign_agg_gru_cell(_i(rowptr), _i(col), _f(src_states), src_states.shape[1], _f(h_dst), n,
                                    units, _f(kernel), _f(rkernel), _f(bias), _f(out), _f(agg_out),
                                    _stream()), "agg_gru_cell")
    return out


def agg_gru_cell_tc_supported(f_in: int, units: int) -> bool:
    return _lib.load().ign_agg_gru_cell_tc_ws_bytes(f_in, units) > 0 and tensor_cores_enabled()


def agg_gru_cell_tc(op: int, rowptr, col, src_states, h_dst, kernel, rkernel, bias, outs, out_row0: int = 0,
                    agg_out=None):
    """Fused gather + aggregation + GRU update (ign_agg_gru_cell_tc).  ``outs``: the state arrays that receive
    rows [out_row0, out_row0 + num_dst): torch tensors or raw device pointers (peer-mapped buffers)."""
    lib = _lib.load()
    n, units = h_dst.shape
    f_in = src_states.shape[1]
    nbytes = lib.ign_agg_gru_cell_tc_ws_bytes(f_in, units)
    if not nbytes:
        raise RuntimeError("IGNNITION: agg_gru_cell_tc is built for message width == units in {32, 64}, got %d, %d"
                           % (f_in, units))
    ws = _workspace(nbytes, h_dst.device)
    arr = (C.c_void_p * max(len(outs), 1))()
    for k, o in enumerate(outs):
        arr[k] = _f(o, "output state array") if torch.is_tensor(o) else int(o)
    _lib.check(lib.ign_agg_gru_cell_tc(op, _i(rowptr), _i(col), _f(src_states), f_in, _f(h_dst), n, units,
                                       _f(kernel), _f(rkernel), _f(bias), len(outs), arr, out_row0, _f(agg_out),
                                       ws.data_ptr(), ws.numel(), _stream()), "agg_gru_cell_tc")


def seq_meta(steps_rowptr, steps, order) -> torch.Tensor:
    lib = _lib.load()
    n = steps_rowptr.numel() - 1
    meta = torch.empty(max(n, 1), 4, dtype=torch.int32, device=steps_rowptr.device)
    _lib.check(lib.ign_seq_meta(_i(steps_rowptr), _i(steps), _i(order), n, _i(meta), _stream()), "seq_meta")
    return meta


def seq_step_plan(meta, steps, max_steps: int):
    """(nt[max_steps], off[max_steps+1], steps_T) of ign_seq_step_plan."""
    lib = _lib.load()
    n = meta.shape[0]
    dev = meta.device
    nt = torch.empty(max_steps, dtype=torch.int32, device=dev)
    off = torch.empty(max_steps + 1, dtype=torch.int32, device=dev)
    steps_t = torch.empty(max(int(steps.numel()), max_steps + 1), dtype=torch.int32, device=dev)
    _lib.check(lib.ign_seq_step_plan(_i(meta), _i(steps), n, max_steps, _i(nt), _i(off), _i(steps_t), _stream()),
               "seq_step_plan")
    return nt, off, steps_t


def gru_seq_steps(plan, meta, srcs: List[torch.Tensor], h0, kernel, rkernel, bias, max_steps: int, out=None,
                  h_seq=None, hs=None):
    """Ordered update as max_steps step-synchronous launches (ign_gru_seq_step)."""
    lib = _lib.load()
    nt, off, steps_t = plan
    n, units = h0.shape
    if out is None:
        out = torch.empty_like(h0)
    if hs is None:
        hs = torch.empty_like(h0)
    sp = _ptr_array(srcs, torch.float32)
    for t in range(max_steps):
        _lib.check(lib.ign_gru_seq_step(t, _i(nt), _i(off), _i(meta), _i(steps_t), len(srcs), sp, srcs[0].shape[1],
                                        _f(h0), _f(hs), n, units, _f(kernel), _f(rkernel), _f(bias), _f(out),
                                        _f(h_seq), _stream()), "gru_seq_step")
    return out


# fewer destination rows than one 128-row tile per SM: the update is bound by launches, not by the kernels
SMALL_ROWS = 148 * 128
# destinations from which the step-synchronous tensor-core BPTT (2 launches per step) beats the fp32 walk (one launch)
BWD_STEPS_MIN_ROWS = int(os.environ.get("IGN_BWD_STEPS_MIN_ROWS", "250000"))      # measured crossover: 186 k .. 373 k paths


def gru_seq_proj_pays(n_steps: int, srcs, units: int, meta, n_dst: int = 1 << 30) -> bool:
    """The hoisted-projection walk (ign_gru_seq_proj) is built for 32-wide states with a walk plan, and pays when the
    source rows are walked over at least twice on average (the projected table is 3 x the source rows)."""
    if meta is None or units != 32 or srcs[0].shape[1] != 32 or not tensor_cores_enabled():
        return False
    if os.environ.get("IGN_GRU_SEQ_PROJ", "1") == "0":
        return False
    if n_dst < SMALL_ROWS:           # launch-bound sizes: one kernel (the plain walk) instead of two
        return False
    return sum(int(s.shape[0]) for s in srcs) * 2 <= n_steps


def _seq_generic_plan(steps_rowptr, steps):
    """Step-major plan of an ordered update for the generic (any width) path: destinations by descending length, per
    step t the entries of the nt[t] destinations that have one (ign_length_order / ign_seq_meta / ign_seq_step_plan).
    The per-step counts come to the host once (this path is a fallback: a sync per update)."""
    order = length_order(steps_rowptr)
    meta = seq_meta(steps_rowptr, steps, order)
    n = steps_rowptr.numel() - 1
    max_steps = int(meta[:n, 2].max().item()) if n else 0
    if max_steps == 0:
        return meta, 0, [], [], None
    nt, off, steps_t = seq_step_plan(meta, steps, max_steps)
    return meta, max_steps, nt.cpu().tolist(), off.cpu().tolist(), steps_t


def _seq_generic_rows(srcs, entries):
    """messages of one step: rows of the source arrays named by step entries ((source << 28) | row, -1 = zero row)"""
    m = entries.numel()
    if len(srcs) == 1:
        return gather_concat([srcs[0]], [torch.where(entries >= 0, entries & 0x0FFFFFFF, entries)], m)
    x = None
    for k, s_k in enumerate(srcs):
        idx = torch.where((entries >= 0) & ((entries >> 28) == k), entries & 0x0FFFFFFF, torch.full_like(entries, -1))
        part = gather_concat([s_k], [idx], m)
        if x is None:
            x = part
        else:
            axpy(1.0, part, x)
    return x


def gru_seq_generic(steps_rowptr, steps, srcs, h0, kernel, rkernel, bias, out, h_seq=None):
    """Ordered update for shapes the walk kernels do not cover (hidden_state_dimension is free in the reference's
    schema; concat along the features makes messages wider than the state): step-synchronous, one gather + one GRU
    step (ign_gru_cell or its generic composition) per step over the destinations that still have one."""
    n = h0.shape[0]
    meta, max_steps, nt, off, steps_t = _seq_generic_plan(steps_rowptr, steps)
    dest, lo = meta[:n, 0].contiguous(), meta[:n, 1].contiguous()
    hs = gather_concat([h0], [dest], n)
    for t in range(max_steps):
        m = nt[t]
        if m == 0:
            break
        x = _seq_generic_rows(srcs, steps_t[off[t]:off[t] + m])
        hn = gru_cell(x, hs[:m], kernel, rkernel, bias)
        hs[:m].copy_(hn)
        if h_seq is not None:
            rows_unpack(hn, (lo[:m] + t).to(torch.int32), h_seq)
    if n:
        rows_unpack(hs, dest, out)
    return out


def gru_seq_bwd_generic(steps_rowptr, steps, srcs, h0, h_seq, kernel, rkernel, bias, d_out, d_steps, dh0, dk, drk, db):
    """BPTT of gru_seq_generic: the steps backwards, ign_gru_cell_bwd (fused or generic) per step."""
    n = h0.shape[0]
    meta, max_steps, nt, off, steps_t = _seq_generic_plan(steps_rowptr, steps)
    dest, lo = meta[:n, 0].contiguous(), meta[:n, 1].contiguous()
    dhs = gather_concat([d_out], [dest], n)
    for t in range(max_steps - 1, -1, -1):
        m = nt[t]
        if m == 0:
            continue
        x = _seq_generic_rows(srcs, steps_t[off[t]:off[t] + m])
        h_prev = (gather_concat([h0], [dest[:m]], m) if t == 0
                  else gather_concat([h_seq], [(lo[:m] + (t - 1)).to(torch.int32)], m))
        dx = torch.empty_like(x)
        dh = torch.empty_like(h_prev)
        gru_cell_bwd(x, h_prev, kernel, rkernel, bias, dhs[:m].contiguous(), dx, dh, dk, drk, db)
        dhs[:m].copy_(dh)
        rows_unpack(dx, (lo[:m] + t).to(torch.int32), d_steps)
    if n:
        rows_unpack(dhs, dest, dh0)


def gru_seq(steps_rowptr, steps, order, srcs: List[torch.Tensor], h0, kernel, rkernel, bias, out=None,
            h_seq=None, meta=None):
    lib = _lib.load()
    n, units = h0.shape
    if out is None:
        out = torch.empty_like(h0)
    if bias.dim() == 1 or (int(srcs[0].shape[1]), int(units)) not in GRU_FUSED_SHAPES:
        return gru_seq_generic(steps_rowptr, steps, srcs, h0, kernel, rkernel, bias, out, h_seq)
    sp = _ptr_array(srcs, torch.float32)
    if gru_seq_proj_pays(int(steps.numel()), srcs, units, meta, n):
        rows = (C.c_int64 * len(srcs))(*[int(s.shape[0]) for s in srcs])
        nbytes = lib.ign_gru_seq_proj_ws_bytes(len(srcs), rows, srcs[0].shape[1], units)
        ws = _workspace(nbytes, h0.device)
        _lib.check(lib.ign_gru_seq_proj(_i(steps_rowptr), _i(steps), _i(meta), len(srcs), sp, rows, srcs[0].shape[1],
                                        _f(h0), n, units, _f(kernel), _f(rkernel), _f(bias), _f(out), _f(h_seq),
                                        ws.data_ptr(), ws.numel(), _stream()), "gru_seq_proj")
        return out
    _lib.check(lib.ign_gru_seq(_i(steps_rowptr), _i(steps), _i(order), len(srcs), sp, srcs[0].shape[1],
                               _f(h0), n, units, _f(kernel), _f(rkernel), _f(bias), _f(out), _f(h_seq),
                               _i(meta), _stream()), "gru_seq")
    return out


def dense(x, w, bias, act: int, out=None, pre_act=None, tensor_cores: bool = True):
    """y = act(x w + b).  With ``tensor_cores`` the layer runs as 3xTF32 on tcgen05 when the shape is
    built for it (a workspace for the split / swizzled weight image is passed), else fp32 FMA."""
    lib = _lib.load()
    m, k = x.shape
    n = w.shape[1]
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=x.device)
    ws, nbytes = None, 0
    if tensor_cores:
        nbytes = lib.ign_dense_ws_bytes(k, n)
        if nbytes:
            ws = _workspace(nbytes, x.device)
    _lib.check(lib.ign_dense(_f(x), m, k, _f(w), _f(bias), n, act, _f(out), _f(pre_act),
                             ws.data_ptr() if ws is not None else None, nbytes, _stream()), "dense")
    return out


def tensor_cores_enabled() -> bool:
    lib = _lib.load()
    prev = lib.ign_set_tensor_cores(1)
    lib.ign_set_tensor_cores(prev)
    return bool(prev)


def dense_head_supported(m: int, k: int, n: int) -> bool:
    return m >= 128 and _lib.load().ign_dense_ws_bytes(k, n) > 0 and tensor_cores_enabled()


def dense_head(x, w, bias, act: int, head_w, head_b, out=None):
    """out[m, 1] = act(x w + b) . head_w + head_b on the tensor cores, hidden activations stay on chip."""
    lib = _lib.load()
    m, k = x.shape
    n = w.shape[1]
    if out is None:
        out = torch.empty(m, 1, dtype=torch.float32, device=x.device)
    nbytes = lib.ign_dense_ws_bytes(k, n)
    ws = _workspace(nbytes, x.device)
    _lib.check(lib.ign_dense_head(_f(x), m, k, _f(w), _f(bias), n, act, _f(head_w), _f(head_b),
                                  _f(out), ws.data_ptr(), nbytes, _stream()), "dense_head")
    return out


def mlp_head_supported(m: int, k1: int, n1: int, n2: int) -> bool:
    return m >= 128 and _lib.load().ign_mlp_head_ws_bytes(k1, n1, n2) > 0 and tensor_cores_enabled()


def mlp_head(x, w1, b1, act1: int, w2, b2, act2: int, w3, b3, out=None):
    """out[m, 1] = act2(act1(x w1 + b1) w2 + b2) . w3 + b3 in one tensor-core kernel."""
    lib = _lib.load()
    m, k1 = x.shape
    n1, n2 = w1.shape[1], w2.shape[1]
    if out is None:
        out = torch.empty(m, 1, dtype=torch.float32, device=x.device)
    nbytes = lib.ign_mlp_head_ws_bytes(k1, n1, n2)
    ws = _workspace(nbytes, x.device)
    _lib.check(lib.ign_mlp_head(_f(x), m, k1, _f(w1), _f(b1), n1, act1, _f(w2), _f(b2), n2, act2, _f(w3), _f(b3),
                                _f(out), ws.data_ptr(), nbytes, _stream()), "mlp_head")
    return out


def gather_concat(parts: List[torch.Tensor], idx: List[Optional[torch.Tensor]], rows: int, out=None):
    lib = _lib.load()
    widths = [p.shape[1] for p in parts]
    if out is None:
        out = torch.empty(rows, sum(widths), dtype=torch.float32, device=parts[0].device)
    pp = _ptr_array(parts, torch.float32)
    ip = _ptr_array(idx, torch.int32)
    wd = (C.c_int32 * len(widths))(*widths)
    _lib.check(lib.ign_gather_concat(len(parts), pp, ip, wd, rows, _f(out), _stream()), "gather_concat")
    return out


def gather_dense_supported(widths: Sequence[int], units: int, rows: int) -> bool:
    lib = _lib.load()
    wd = (C.c_int32 * len(widths))(*[int(v) for v in widths])
    return (rows >= 4096 and len(widths) <= 4 and tensor_cores_enabled()      # fewer rows: fp32 kernels (ign_dense)
            and os.environ.get("IGN_GATHER_DENSE", "1") != "0"
            and lib.ign_gather_dense_ws_bytes(len(widths), wd, int(units)) > 0)


def gather_dense(parts: List[torch.Tensor], idx: List[Optional[torch.Tensor]], rows: int, w, bias, act: int, out=None):
    """act(concat_k parts[k][idx[k]] w + b) with the gather and the concat fused into the GEMM's operand loaders
    (ign_gather_dense): the concatenated input is never materialised."""
    lib = _lib.load()
    widths = [int(p.shape[1]) for p in parts]
    n = int(w.shape[1])
    if out is None:
        out = torch.empty(rows, n, dtype=torch.float32, device=parts[0].device)
    wd = (C.c_int32 * len(widths))(*widths)
    nbytes = lib.ign_gather_dense_ws_bytes(len(widths), wd, n)
    ws = _workspace(nbytes, parts[0].device)
    _lib.check(lib.ign_gather_dense(len(parts), _ptr_array(parts, torch.float32), _ptr_array(idx, torch.int32), wd, rows,
                                    _f(w), _f(bias), n, act, _f(out), ws.data_ptr(), ws.numel(), _stream()),
               "gather_dense")
    return out


# ------------------------------------------------------------------------------- train step
def mse_loss(pred, label, grad_scale: float, d_pred, sse):
    lib = _lib.load()
    _lib.check(lib.ign_mse_loss(_f(pred), _f(label), pred.numel(), grad_scale, _f(d_pred),
                                _ptr(sse, torch.float64, "sse"), _stream()), "mse_loss")


LOSSES = {"MeanSquaredError": 0, "MeanAbsoluteError": 1, "MeanAbsolutePercentageError": 2,
          "MeanSquaredLogarithmicError": 3, "Huber": 4, "LogCosh": 5, "BinaryCrossentropy": 6}
OPTIMIZERS = {"SGD": 1, "RMSprop": 2, "Adagrad": 3, "Adamax": 4}


def loss(kind: int, pred, label, grad_scale: float, d_pred, acc, delta: float = 1.0):
    """acc += sum of the per-prediction losses (fp64); d_pred = dl/dpred * grad_scale (ign_loss)."""
    lib = _lib.load()
    _lib.check(lib.ign_loss(kind, _f(pred), _f(label), pred.numel(), grad_scale, delta, _f(d_pred),
                            _ptr(acc, torch.float64, "acc"), _stream()), "loss")


def optimizer_step(kind: int, w, g, s1, s2, lr: float, a: float, b: float, eps: float, flags: int = 0):
    lib = _lib.load()
    _lib.check(lib.ign_optimizer_step(kind, _f(w), _f(g), _f(s1), _f(s2), w.numel(), lr, a, b, eps, flags, _stream()),
               "optimizer_step")


def dense_bwd(x, w, act: int, pre_act, dy, dx, dw, db):
    lib = _lib.load()
    m, k = x.shape
    n = w.shape[1]
    ws, nbytes = None, 0
    if dx is not None and tensor_cores_enabled():
        nbytes = lib.ign_dense_bwd_ws_bytes(k, n)
        if nbytes:
            ws = _workspace(nbytes, x.device)
    _lib.check(lib.ign_dense_bwd(_f(x), m, k, _f(w), n, act, _f(pre_act), _f(dy), _f(dx), _f(dw), _f(db),
                                 ws.data_ptr() if ws is not None else None, nbytes if ws is not None else 0,
                                 _stream()), "dense_bwd")


def dense_head_bwd_chain_supported(m: int, k: int, n: int) -> bool:
    return n == 1 and k in (128, 256, 512) and m >= 1024


def dense_head_bwd_chain(x, w, dz, prev_act: int, dz_prev, dw, db_prev):
    """Backward of a linear k -> 1 head fused with act' of the layer below (ign_dense_head_bwd_chain)."""
    lib = _lib.load()
    m, k = x.shape
    _lib.check(lib.ign_dense_head_bwd_chain(_f(x), m, k, _f(w), _f(dz), prev_act, _f(dz_prev), _f(dw), _f(db_prev),
                                            _stream()), "dense_head_bwd_chain")


def gru_cell_bwd(x, h, kernel, rkernel, bias, d_out, dx, dh, dk, drk, db):
    lib = _lib.load()
    n, units = h.shape
    f_in = x.shape[1]
    if bias.dim() == 1:
        return gru_cell_bwd_v1(x, h, kernel, rkernel, bias, d_out, dx, dh, dk, drk, db)
    if not (f_in == units and units in (16, 32)):
        # generic shapes (config 5's 64-wide model, f_in != units): recompute the two gate GEMMs, one element-wise
        # kernel for the gate gradients, and the Dense backward for dx / dh and the weight gradients
        zx = dense(x, kernel, bias[0], 0)
        zh = dense(h, rkernel, bias[1], 0)
        direct = torch.empty_like(h)
        _lib.check(lib.ign_gru_gates_bwd(_f(zx), _f(zh), _f(h), _f(d_out), n, units, _f(direct), _stream()),
                   "gru_gates_bwd")
        dense_bwd(x, kernel, 0, None, zx, dx, dk, db[0])
        dense_bwd(h, rkernel, 0, None, zh, dh, drk, db[1])
        axpy(1.0, direct, dh)
        return
    _lib.check(lib.ign_gru_cell_bwd(_f(x), _f(h), n, x.shape[1], units, _f(kernel), _f(rkernel), _f(bias),
                                    _f(d_out), _f(dx), _f(dh), _f(dk), _f(drk), _f(db), _stream()),
               "gru_cell_bwd")


def gru_seq_bwd(steps_rowptr, steps, order, srcs, h0, h_seq, kernel, rkernel, bias, d_out, d_steps, dh0,
                dk, drk, db):
    lib = _lib.load()
    n, units = h0.shape
    if bias.dim() == 1 or not (int(srcs[0].shape[1]) == int(units) and int(units) in (16, 32)):
        return gru_seq_bwd_generic(steps_rowptr, steps, srcs, h0, h_seq, kernel, rkernel, bias, d_out, d_steps, dh0,
                                   dk, drk, db)
    sp = _ptr_array(srcs, torch.float32)
    _lib.check(lib.ign_gru_seq_bwd(_i(steps_rowptr), _i(steps), _i(order), len(srcs), sp, srcs[0].shape[1],
                                   _f(h0), _f(h_seq), n, units, _f(kernel), _f(rkernel), _f(bias),
                                   _f(d_out), _f(d_steps), _f(dh0), _f(dk), _f(drk), _f(db), _stream()),
               "gru_seq_bwd")


def gru_seq_bwd_steps(plan, meta, max_steps: int, srcs, h0, h_seq, kernel, rkernel, bias, d_out, d_steps, dh0,
                      dk, drk, db):
    """BPTT of an ordered update as step-synchronous tensor-core launches (ign_gru_seq_bwd_steps)."""
    lib = _lib.load()
    nt, off, steps_t = plan
    n, units = h0.shape
    sp = _ptr_array(srcs, torch.float32)
    nbytes = lib.ign_gru_seq_bwd_steps_ws_bytes(n)
    ws = _workspace(nbytes, h0.device)
    _lib.check(lib.ign_gru_seq_bwd_steps(max_steps, _i(nt), _i(off), _i(meta), _i(steps_t), len(srcs), sp,
                                         srcs[0].shape[1], _f(h0), _f(h_seq), n, units, _f(kernel), _f(rkernel),
                                         _f(bias), _f(d_out), _f(d_steps), _f(dh0), _f(dk), _f(drk), _f(db),
                                         ws.data_ptr(), ws.numel(), _stream()), "gru_seq_bwd_steps")


def l2_reg(w, lam: float, dw, reg):
    lib = _lib.load()
    _lib.check(lib.ign_l2_reg(_f(w), w.numel(), lam, _f(dw), _ptr(reg, torch.float64, "reg"), _stream()),
               "l2_reg")


def adam_step(w, g, m, v, lr: float, beta1: float, beta2: float, eps: float, step: int):
    lib = _lib.load()
    _lib.check(lib.ign_adam_step(_f(w), _f(g), _f(m), _f(v), w.numel(), lr, beta1, beta2, eps, step,
                                 _stream()), "adam_step")


def axpy(a: float, x, y):
    lib = _lib.load()
    _lib.check(lib.ign_axpy(x.numel(), a, _f(x), _f(y), _stream()), "axpy")


def slice_cols(x, col0: int, width: int, out=None):
    lib = _lib.load()
    rows, ld = x.shape
    if out is None:
        out = torch.empty(rows, width, dtype=torch.float32, device=x.device)
    _lib.check(lib.ign_slice_cols(_f(x), rows, ld, col0, width, _f(out), _stream()), "slice_cols")
    return out


def mul(a, b, out=None):
    lib = _lib.load()
    if out is None:
        out = torch.empty_like(a)
    _lib.check(lib.ign_mul(a.numel(), _f(a), _f(b), _f(out), _stream()), "mul")
    return out


def conv_finish(nsum, self_state, rowptr, act: int, out=None):
    lib = _lib.load()
    n, F = nsum.shape
    if out is None:
        out = torch.empty_like(nsum)
    _lib.check(lib.ign_conv_finish(_f(nsum), _f(self_state), _i(rowptr), F, n, act, _f(out), _stream()), "conv_finish")
    return out


def attention_combine(src_rowptrs, src_perms, edge_counts: Sequence[int], max_len: int):
    """(rowptr, perm, slot_col) of the CSR over the concatenated edge lists of several sources
    (ign_attention_combine; generate_model.py:523-543)."""
    lib = _lib.load()
    dev = src_rowptrs[0].device
    num_dst = src_rowptrs[0].numel() - 1
    total = int(sum(edge_counts))
    rowptr = torch.empty(num_dst + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(max(total, 1), dtype=torch.int32, device=dev)[:total]
    slot_col = torch.empty(max(total, 1), dtype=torch.int32, device=dev)[:total]
    counts = (C.c_int64 * len(edge_counts))(*[int(c) for c in edge_counts])
    _lib.check(lib.ign_attention_combine(len(edge_counts), _ptr_array(src_rowptrs, torch.int32),
                                         _ptr_array(src_perms, torch.int32), counts, num_dst, max_len, _i(rowptr),
                                         _i(perm) if total else None, _i(slot_col) if total else None, _stream()),
               "attention_combine")
    return rowptr, perm, slot_col


def attention_aggregate(rowptr, col, rows, src_score, dst_score, sample_offsets, max_len: int, out=None,
                        keep_ws: bool = False, slot_col=None):
    lib = _lib.load()
    num_dst = rowptr.numel() - 1
    n_edges = col.numel()
    n_samples = sample_offsets.numel() - 1
    F = rows.shape[1]
    if out is None:
        out = torch.empty(num_dst, F, dtype=torch.float32, device=rows.device)
    nbytes = lib.ign_attention_ws_bytes(n_edges, n_samples, max_len)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rows.device)
    _lib.check(lib.ign_attention_aggregate(_i(rowptr), _i(col), _i(slot_col), _f(rows), F, _f(src_score), _f(dst_score),
                                           _i(sample_offsets), n_samples, num_dst, n_edges, max_len, _f(out),
                                           ws.data_ptr(), nbytes, _stream()), "attention_aggregate")
    return (out, ws) if keep_ws else out


def attention_aggregate_bwd(rowptr, idx, perm, rows, g_out, sample_offsets, max_len: int, fwd_ws, slot_col=None):
    """(d_msg [E, F], d_pre4 [E, 4], d_ds [num_dst, 1]) of ign_attention_aggregate_bwd, per-edge arrays in input edge order"""
    lib = _lib.load()
    num_dst = rowptr.numel() - 1
    n_edges = idx.numel()
    n_samples = sample_offsets.numel() - 1
    F = rows.shape[1]
    dev = rows.device
    d_msg = torch.zeros(n_edges, F, dtype=torch.float32, device=dev)
    d_pre4 = torch.zeros(n_edges, 4, dtype=torch.float32, device=dev)
    d_ds = torch.zeros(num_dst, 1, dtype=torch.float32, device=dev)
    nbytes = lib.ign_attention_bwd_ws_bytes(n_edges, n_samples, max_len)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.ign_attention_aggregate_bwd(_i(rowptr), _i(idx), _i(perm), _i(slot_col), _f(rows), F, _f(g_out),
                                               _i(sample_offsets),
                                               n_samples, num_dst, n_edges, max_len, fwd_ws.data_ptr(), _f(d_msg),
                                               _f(d_pre4), _f(d_ds), ws.data_ptr(), nbytes, _stream()),
               "attention_aggregate_bwd")
    return d_msg, d_pre4, d_ds


def small_graph_forward(units: int, rows: Sequence[int], buf0, buf1, kinds: Sequence[int], dsts: Sequence[int],
                        srcs: Sequence[Sequence[int]], rowptrs, idxs, kernels, rkernels, biases, iterations: int,
                        step_out=None, step_hseq=None, step_agg=None):
    """All ``iterations`` of the message-passing loop of a small graph in one launch (ign_small_graph_forward).
    Returns, per entity, which of its two buffers (0 / 1) holds the final state.  ``step_out`` / ``step_hseq`` /
    ``step_agg`` (lists of iterations * stages tensors or None): keep every stage's outputs, for training."""
    lib = _lib.load()
    n_ent, n_ops = len(rows), len(kinds)
    dev = next(b.device for b in buf0 if b is not None)
    flat_src = []
    for s_ in srcs:
        flat_src += list(s_) + [-1] * (4 - len(s_))
    final = (C.c_int32 * n_ent)()
    nbytes = lib.ign_small_graph_ws_bytes()
    ws = _workspace(nbytes, dev)
    _lib.check(lib.ign_small_graph_forward(
        int(units), n_ent, (C.c_int64 * n_ent)(*[int(r) for r in rows]), _ptr_array(buf0, torch.float32),
        _ptr_array(buf1, torch.float32), n_ops, (C.c_int32 * n_ops)(*[int(k) for k in kinds]),
        (C.c_int32 * n_ops)(*[int(d) for d in dsts]), (C.c_int32 * (4 * n_ops))(*flat_src),
        _ptr_array(rowptrs, torch.int32), _ptr_array([i if i is not None and i.numel() else None for i in idxs], torch.int32),
        _ptr_array(kernels, torch.float32), _ptr_array(rkernels, torch.float32), _ptr_array(biases, torch.float32),
        int(iterations), _ptr_array(step_out, torch.float32) if step_out is not None else None,
        _ptr_array(step_hseq, torch.float32) if step_out is not None else None,
        _ptr_array(step_agg, torch.float32) if step_out is not None else None,
        final, ws.data_ptr(), nbytes, _stream()), "small_graph_forward")
    return [int(v) for v in final]


def partner_index(rowptr0, rowptr1, idx1, n_edges0: int):
    lib = _lib.load()
    out = torch.empty(n_edges0, dtype=torch.int32, device=rowptr0.device)
    _lib.check(lib.ign_partner_index(_i(rowptr0), _i(rowptr1), _i(idx1) if idx1.numel() else None,
                                     rowptr0.numel() - 1, _i(out) if n_edges0 else None, _stream()), "partner_index")
    return out


# ------------------------------------------------------------------------------- partitioned graphs
def edge_owner(dst, bounds: Sequence[int]) -> torch.Tensor:
    """Owner rank of every edge's destination; ``bounds`` = world + 1 ascending row bounds (host)."""
    lib = _lib.load()
    world = len(bounds) - 1
    owner = torch.empty_like(dst)
    b = (C.c_int32 * (world + 1))(*[int(v) for v in bounds])
    _lib.check(lib.ign_edge_owner(_i(dst), dst.numel(), b, world, _i(owner), _stream()), "edge_owner")
    return owner


def gather_int(values, perm=None, add: int = 0, out=None) -> torch.Tensor:
    lib = _lib.load()
    n = perm.numel() if perm is not None else values.numel()
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=values.device)
    _lib.check(lib.ign_gather_int(_i(values), _i(perm), n, add, _i(out), _stream()), "gather_int")
    return out


def mark_rows(col, flags):
    lib = _lib.load()
    _lib.check(lib.ign_mark_rows(_i(col), col.numel(), _i(flags), _stream()), "mark_rows")


def flag_compact(flags, add: int = 0):
    """(rows[n] int32 with the flagged positions + add first, count[1] device int32)."""
    lib = _lib.load()
    n = flags.numel()
    out = torch.empty(max(n, 1), dtype=torch.int32, device=flags.device)
    count = torch.empty(1, dtype=torch.int32, device=flags.device)
    ws = _workspace(lib.ign_flag_compact_ws_bytes(n), flags.device)
    _lib.check(lib.ign_flag_compact(_i(flags), n, add, _i(out), _i(count), ws.data_ptr(), ws.numel(), _stream()),
               "flag_compact")
    return out, count


def rows_put(src, rows, dst):
    """dst[rows] = src[rows]; ``dst``: a tensor or the raw device pointer of a peer-mapped array."""
    lib = _lib.load()
    d = _f(dst) if torch.is_tensor(dst) else int(dst)
    _lib.check(lib.ign_rows_put(_f(src), _i(rows), rows.numel(), src.shape[1], d, _stream()), "rows_put")


def scale_rows_inv_degree(d, rowptr):
    """d[r, :] /= max(in-degree of r, 1), in place: dL/d(mean) -> dL/d(sum)."""
    lib = _lib.load()
    _lib.check(lib.ign_scale_rows_inv_degree(_f(d), _i(rowptr), d.shape[0], d.shape[1], _stream()), "scale_rows_inv_degree")
    return d


def segment_broadcast(rowptr, rows, n_out: int):
    """out[r] = rows[s] for every row r of segment s: backward of a per-sample sum."""
    lib = _lib.load()
    out = torch.zeros(n_out, rows.shape[1], dtype=torch.float32, device=rows.device)
    _lib.check(lib.ign_segment_broadcast(_i(rowptr), _f(rows), rows.shape[0], rows.shape[1], _f(out), _stream()),
               "segment_broadcast")
    return out


def segment_max_bwd(rowptr, idx, perm, rows, agg, d_agg, n_edges: int):
    """Per-edge message gradients of a max aggregation (ties share evenly), [n_edges, width], at row perm[slot]."""
    lib = _lib.load()
    d_msg = torch.zeros(n_edges, agg.shape[1], dtype=torch.float32, device=agg.device)
    _lib.check(lib.ign_segment_max_bwd(_i(rowptr), _i(idx), _i(perm), _f(rows), _f(agg), _f(d_agg), agg.shape[0],
                                       agg.shape[1], _f(d_msg), _stream()), "segment_max_bwd")
    return d_msg


def rows_unpack(packed, rows, dst):
    """dst[rows[i]] = packed[i] (ign_rows_unpack)."""
    lib = _lib.load()
    _lib.check(lib.ign_rows_unpack(_f(packed), _i(rows), rows.numel(), dst.shape[1], _f(dst), _stream()), "rows_unpack")


def peer_copy(dst_ptr: int, src_ptr: int, nbytes: int):
    """``nbytes`` from ``src_ptr`` to ``dst_ptr`` (raw device addresses; dst usually a peer-mapped buffer) by the copy
    engine, on the current stream."""
    lib = _lib.load()
    _lib.check(lib.ign_peer_copy(int(dst_ptr), int(src_ptr), int(nbytes), _stream()), "peer_copy")


def index_range_check(idx, bound: int, bad):
    lib = _lib.load()
    _lib.check(lib.ign_index_range_check(_i(idx), idx.numel(), bound, _i(bad), _stream()), "index_range_check")
