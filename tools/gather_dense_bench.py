#!/usr/bin/env python
"""Message network input: gather + concat fused into the first layer's GEMM (ign_gather_dense) against
ign_gather_concat + ign_dense, config-5-like shapes (source and destination states 64 wide, 64 units), CUDA events.

    python tools/gather_dense_bench.py [--edges 20000000] [--nodes 1000000]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ignnition_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--edges", type=int, default=20_000_000)
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--width", type=int, default=64)
ap.add_argument("--units", type=int, default=64)
args = ap.parse_args()
E, N, F, U = args.edges, args.nodes, args.width, args.units
g = torch.Generator(device="cuda").manual_seed(0)
states = torch.randn(N, F, device="cuda", generator=g)
src = torch.randint(0, N, (E,), device="cuda", dtype=torch.int32, generator=g)
dst = torch.randint(0, N, (E,), device="cuda", dtype=torch.int32, generator=g).sort().values.to(torch.int32)
W = torch.randn(2 * F, U, device="cuda", generator=g) * 0.1
b = torch.randn(U, device="cuda", generator=g) * 0.1
out = torch.empty(E, U, device="cuda")
x = torch.empty(E, 2 * F, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_fused = timed(lambda: ops.gather_dense([states, states], [src, dst], E, W, b, 2, out=out))
ref = out.clone()
t_gc = timed(lambda: ops.gather_concat([states, states], [src, dst], E, out=x))
t_d = timed(lambda: ops.dense(x, W, b, 2, out=out))
same = bool(torch.equal(ref, out))
alg = E * (8 + 2 * F * 4 + U * 4)           # two indices, two gathered rows, one output row per edge
print(json.dumps({"edges": E, "nodes": N, "widths": [F, F], "units": U,
                  "fused_ms": t_fused, "fused_algorithmic_gbs": alg / t_fused / 1e6,
                  "gather_concat_ms": t_gc, "dense_ms": t_d, "unfused_ms": t_gc + t_d,
                  "bit_identical": same,
                  "input_tensor_bytes_not_written": E * 2 * F * 4,
                  "at_200M_edges_bytes_not_written_and_not_read": 2 * 200_000_000 * 2 * F * 4}))
