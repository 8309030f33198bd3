#!/usr/bin/env python
"""Where does the state deviation come from?  For one workload and the first K samples of a synthetic batch, the
max-norm relative deviation from the fp64 oracle of: the tcgen05 (3xTF32) path, the fp32 CUDA-core twins, and the
NumPy fp32 oracle (= what any fp32 implementation with another summation order shows), with the package's Keras
default initialiser (orthogonal recurrent kernels) and with the oracle's glorot test weights.

    python tools/parity_probe.py [--workload routenet_geant2_b4096] [--samples 16]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="routenet_geant2_b4096")
    ap.add_argument("--samples", type=int, default=16)
    args = ap.parse_args()
    import torch
    from ignnition_b200 import Engine, ModelDescription, ops
    from ignnition_b200.batching import assemble_tiled
    from oracle import ignnition_oracle as orc
    g, shape, qsize, _ = bench.load_case(args.workload)
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    base = g["reference_tensors"][0]
    K = args.samples
    for wname in ("keras_default(seed 0)", "oracle_glorot(1234)"):
        eng = Engine(md, device="cuda", seed=0)
        if wname.startswith("oracle"):
            eng.set_weights(orc.Oracle(g["model_json"], dims).init_weights(1234))
        w32 = eng.get_weights()
        batch = assemble_tiled(base, K, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                               bench.feature_fns(qsize), seed=0)
        o64 = orc.Oracle(g["model_json"], dims, dtype=np.float64)
        o32 = orc.Oracle(g["model_json"], dims, dtype=np.float32)
        want = []
        for k in range(K):
            t = dict(base)
            for name, ent, size in eng.features:
                n_e = int(base["num_" + ent])
                t[name] = batch.arrays["feat_" + name][k * n_e * size:(k + 1) * n_e * size]
            want.append((t, o64.forward(t, w32, return_states=True)))

        def err(pred, states):
            wp, ws = 0.0, {e: 0.0 for e in eng.entities}
            for k, (t, (p64, s64)) in enumerate(want):
                p64 = p64.reshape(-1)
                wp = max(wp, float(np.abs(pred[k] - p64).max() / np.abs(p64).max()))
                for e in eng.entities:
                    n_e = int(base["num_" + e])
                    ws[e] = max(ws[e], float(np.abs(states[e][k * n_e:(k + 1) * n_e] - s64[e]).max() / np.abs(s64[e]).max()))
            return {"pred": wp, **ws}

        rows = {}
        graph = eng.upload(batch, eng.pack(batch))
        real_cell, real_seq = ops.gru_cell, ops.gru_seq

        def forced(fn, tc):
            def call(*a, **kw):
                prev = ops.tensor_cores_enabled()
                ops.set_tensor_cores(tc)
                try:
                    return fn(*a, **kw)
                finally:
                    ops.set_tensor_cores(prev)
            return call

        for label, tc, cell_tc, seq_tc in (("tcgen05_3xtf32", True, None, None), ("fp32_cuda_cores", False, None, None),
                                           ("tc_walk+fp32_cell", True, False, None),
                                           ("fp32_walk+tc_cell", True, None, False)):
            ops.set_tensor_cores(tc)
            ops.gru_cell = forced(real_cell, cell_tc) if cell_tc is not None else real_cell
            ops.gru_seq = forced(real_seq, seq_tc) if seq_tc is not None else real_seq
            eng.build_graph(graph)
            p, st = eng.forward(graph, return_states=True)
            rows[label] = err(p.cpu().numpy().reshape(K, -1), {e: st[e].cpu().numpy() for e in eng.entities})
        ops.gru_cell, ops.gru_seq = real_cell, real_seq
        ops.set_tensor_cores(True)
        preds, sts = [], {e: [] for e in eng.entities}
        for t, _ in want:
            p, s = o32.forward(t, w32, return_states=True)
            preds.append(p.reshape(-1))
            for e in eng.entities:
                sts[e].append(s[e])
        rows["numpy_fp32_oracle"] = err(preds, {e: np.concatenate(v) for e, v in sts.items()})
        print(json.dumps({"workload": args.workload, "weights": wname, "samples": K,
                          "exact_math_env": os.environ.get("IGN_GRU_TC_EXACT_MATH"), "max_rel_dev_vs_fp64": rows}))


if __name__ == "__main__":
    main()
