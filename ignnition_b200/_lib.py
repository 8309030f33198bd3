"""ctypes binding of ``libignnition_b200.so`` (the C-ABI declared in ``include/ignnition_b200.h``).

There is no fallback: if the shared library has not been built the import of any compute entry
point raises, and every non-zero status is turned into ``RuntimeError("IGNNITION: ...")`` with the
library's own message (the reference logs ``IGNNITION: ...`` and exits, e.g.
``code/utils/generator_std_to_framework.py:229-230``).
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libignnition_b200.so")

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/ignnition_b200.h one to one
SIGNATURES = {
    "ign_version": (_int, []),
    "ign_last_error": (_int, [C.c_char_p, _sz]),
    "ign_launch_count": (_i64, []),
    "ign_set_tensor_cores": (_int, [_int]),
    "ign_csr_build_ws_bytes": (_sz, [_i64, _i64]),
    "ign_csr_build": (_int, [_p, _p, _p, _i64, _i64, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "ign_length_order_ws_bytes": (_sz, [_i64]),
    "ign_length_order": (_int, [_p, _i64, _p, _p, _sz, _p]),
    "ign_steps_build_ws_bytes": (_sz, [_i64]),
    "ign_steps_build": (_int, [_int, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "ign_steps_keys": (_int, [_p, _i64, _int, _i64, _p, _p]),
    "ign_init_state": (_int, [_int, _p, _p, _i64, _int, _p, _p]),
    "ign_segment_reduce": (_int, [_int, _p, _p, _p, _int, _i64, _p, _p]),
    "ign_gru_cell_ws_bytes": (_sz, [_int, _int]),
    "ign_gru_cell": (_int, [_p, _p, _i64, _int, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "ign_agg_gru_cell": (_int, [_p, _p, _p, _int, _p, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "ign_gru_seq": (_int, [_p, _p, _p, _int, _p, _int, _p, _i64, _int, _p, _p, _p, _p, _p, _p, _p]),
    "ign_gru_seq_proj_ws_bytes": (_sz, [_int, _p, _int, _int]),
    "ign_gru_seq_proj": (_int, [_p, _p, _p, _int, _p, _p, _int, _p, _i64, _int, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ign_seq_step_plan": (_int, [_p, _p, _i64, _int, _p, _p, _p, _p]),
    "ign_gru_seq_step": (_int, [_int, _p, _p, _p, _p, _int, _p, _int, _p, _p, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "ign_seq_meta": (_int, [_p, _p, _p, _i64, _p, _p]),
    "ign_dense_ws_bytes": (_sz, [_int, _int]),
    "ign_dense": (_int, [_p, _i64, _int, _p, _p, _int, _int, _p, _p, _p, _sz, _p]),
    "ign_dense_head": (_int, [_p, _i64, _int, _p, _p, _int, _int, _p, _p, _p, _p, _sz, _p]),
    "ign_mlp_head_ws_bytes": (_sz, [_int, _int, _int]),
    "ign_mlp_head": (_int, [_p, _i64, _int, _p, _p, _int, _int, _p, _p, _int, _int, _p, _p, _p, _p, _sz, _p]),
    "ign_gather_concat": (_int, [_int, _p, _p, _p, _i64, _p, _p]),
    "ign_gather_dense_ws_bytes": (_sz, [_int, _p, _int]),
    "ign_gather_dense": (_int, [_int, _p, _p, _p, _i64, _p, _p, _int, _int, _p, _p, _sz, _p]),
    "ign_mse_loss": (_int, [_p, _p, _i64, _f, _p, _p, _p]),
    "ign_dense_bwd_ws_bytes": (_sz, [_int, _int]),
    "ign_dense_bwd": (_int, [_p, _i64, _int, _p, _int, _int, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ign_gru_cell_bwd": (_int, [_p, _p, _i64, _int, _int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ign_gru_gates_bwd": (_int, [_p, _p, _p, _p, _i64, _int, _p, _p]),
    "ign_scale_rows_inv_degree": (_int, [_p, _p, _i64, _int, _p]),
    "ign_segment_broadcast": (_int, [_p, _p, _i64, _int, _p, _p]),
    "ign_segment_max_bwd": (_int, [_p, _p, _p, _p, _p, _p, _i64, _int, _p, _p]),
    "ign_gru_seq_bwd": (_int, [_p, _p, _p, _int, _p, _int, _p, _p, _i64, _int, _p, _p, _p, _p, _p, _p,
                               _p, _p, _p, _p]),
    "ign_gru_seq_bwd_steps_ws_bytes": (_sz, [_i64]),
    "ign_gru_seq_bwd_steps": (_int, [_int, _p, _p, _p, _p, _int, _p, _int, _p, _p, _i64, _int, _p, _p, _p, _p,
                                     _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ign_dense_head_bwd_chain": (_int, [_p, _i64, _int, _p, _p, _int, _p, _p, _p, _p]),
    "ign_l2_reg": (_int, [_p, _i64, _f, _p, _p, _p]),
    "ign_loss": (_int, [_int, _p, _p, _i64, C.c_float, C.c_float, _p, _p, _p]),
    "ign_optimizer_step": (_int, [_int, _p, _p, _p, _p, _i64, C.c_float, C.c_float, C.c_float, C.c_float, _int, _p]),
    "ign_gru_gates_fwd": (_int, [_p, _p, _p, _i64, _int, _p, _p]),
    "ign_adam_step": (_int, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i64, _p]),
    "ign_attention_ws_bytes": (_sz, [_i64, _i64, _int]),
    "ign_attention_aggregate": (_int, [_p, _p, _p, _p, _int, _p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "ign_gru_v1_reset": (_int, [_p, _p, _p, _i64, _int, _p, _p]),
    "ign_gru_v1_out": (_int, [_p, _p, _p, _p, _i64, _int, _p, _p]),
    "ign_gru_v1_bwd_out": (_int, [_p, _p, _p, _p, _p, _i64, _int, _p, _p]),
    "ign_gru_v1_bwd_reset": (_int, [_p, _p, _p, _p, _i64, _int, _p, _p]),
    "ign_csr_build_small": (_int, [_int, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ign_small_graph_ws_bytes": (_sz, []),
    "ign_small_graph_forward": (_int, [_int, _int, _p, _p, _p, _int, _p, _p, _p, _p, _p, _p, _p, _p, _int, _p, _p, _p, _p,
                                       _p, _sz, _p]),
    "ign_attention_combine": (_int, [_int, _p, _p, _p, _i64, _int, _p, _p, _p, _p]),
    "ign_attention_bwd_ws_bytes": (_sz, [_i64, _i64, _int]),
    "ign_attention_aggregate_bwd": (_int, [_p, _p, _p, _p, _p, _int, _p, _p, _i64, _i64, _i64, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "ign_partner_index": (_int, [_p, _p, _p, _i64, _p, _p]),
    "ign_ingest_create": (_p, [_int, _p, _int, _p, _p, _int, _p, _p, _p, _p, C.c_char_p]),
    "ign_ingest_add_sequence": (_int, [_p, _int, _p, _int, C.c_char_p]),
    "ign_ingest_sequence": (_i64, [_p, _int, _p, _p, _p]),
    "ign_ingest_destroy": (None, [_p]),
    "ign_ingest_reset": (None, [_p]),
    "ign_ingest_parse": (_i64, [_p, C.c_char_p, _sz, _i64]),
    "ign_ingest_n_samples": (_i64, [_p]),
    "ign_ingest_offsets": (_p, [_p, _int]),
    "ign_ingest_feature": (_i64, [_p, _int, _p]),
    "ign_ingest_adjacency": (_i64, [_p, _int, _p, _p, _p, _p, _p]),
    "ign_ingest_labels": (_i64, [_p, _p]),
    "ign_mul": (_int, [_i64, _p, _p, _p, _p]),
    "ign_slice_cols": (_int, [_p, _i64, _int, _int, _int, _p, _p]),
    "ign_conv_finish": (_int, [_p, _p, _p, _int, _i64, _int, _p, _p]),
    "ign_axpy": (_int, [_i64, _f, _p, _p, _p]),
    "ign_agg_gru_cell_tc_ws_bytes": (_sz, [_int, _int]),
    "ign_agg_gru_cell_tc": (_int, [_int, _p, _p, _p, _int, _p, _i64, _int, _p, _p, _p, _int, _p, _i64, _p, _p, _sz, _p]),
    "ign_peer_alloc": (_int, [_sz, C.POINTER(_p)]),
    "ign_peer_free": (_int, [_p]),
    "ign_peer_export": (_int, [_p, _p]),
    "ign_peer_open": (_int, [_p, C.POINTER(_p)]),
    "ign_peer_close": (_int, [_p]),
    "ign_edge_owner": (_int, [_p, _i64, _p, _int, _p, _p]),
    "ign_gather_int": (_int, [_p, _p, _i64, _int, _p, _p]),
    "ign_mark_rows": (_int, [_p, _i64, _p, _p]),
    "ign_flag_compact_ws_bytes": (_sz, [_i64]),
    "ign_flag_compact": (_int, [_p, _i64, _int, _p, _p, _p, _sz, _p]),
    "ign_rows_put": (_int, [_p, _p, _i64, _int, _p, _p]),
    "ign_peer_copy": (_int, [_p, _p, _sz, _p]),
    "ign_rows_unpack": (_int, [_p, _p, _i64, _int, _p, _p]),
    "ign_index_range_check": (_int, [_p, _i64, _i64, _p, _p]),
}

_lib = None


def load():
    """Load the library once; raise loudly when it is missing (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "IGNNITION: the CUDA extension %s is not built. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); "
            "ignnition_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().ign_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(status: int, what: str = ""):
    if status != 0:
        msg = last_error() or ("IGNNITION: %s failed with status %d" % (what, status))
        if not msg.startswith("IGNNITION"):
            msg = "IGNNITION: " + msg
        raise RuntimeError(msg)
