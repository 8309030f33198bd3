#!/usr/bin/env python
"""Per-stage timeline of the one-launch message-passing loop (csrc/small_graph.cu built with -DIGN_SG_PROFILE, see
tools/build_profile_lib.sh): for CTA 0, when its rows of a stage were done and when it left the grid barrier."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ignnition_b200 import Engine, ModelDescription, ops
from ignnition_b200.batching import assemble_tiled

g, shape, qsize, _ = bench.load_case("routenet_nsfnet_b4096")
md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
keep = {}
orig = ops._workspace


def spy(nbytes, device):
    t = orig(nbytes, device)
    if nbytes == 1024:
        keep["ws"] = t
    return t


ops._workspace = spy
for n in (3, 32):
    eng = Engine(md, device="cuda", seed=0)
    batch = assemble_tiled(g["reference_tensors"][0], n, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                           bench.feature_fns(qsize), seed=0)
    graph = eng.prepare(batch)
    st = eng.initial_states(graph)
    for _ in range(3):
        eng.message_passing(graph, dict(st))
    torch.cuda.synchronize()
    t = keep["ws"].cpu().numpy().view(np.uint64)[8:8 + 32].astype(np.int64).reshape(16, 2)
    t0 = t[0, 0]
    rows = [{"stage": i, "rows_done_us": (t[i, 0] - t0) / 1e3, "barrier_left_us": (t[i, 1] - t0) / 1e3} for i in range(16)]
    work = [(t[i, 0] - t[i - 1, 1]) / 1e3 for i in range(1, 16)]
    wait = [(t[i, 1] - t[i, 0]) / 1e3 for i in range(15)]
    print(json.dumps({"batch": n, "walk_stage_us": float(np.mean(work[1::2])), "sum_stage_us": float(np.mean(work[0::2])),
                      "barrier_after_walk_us": float(np.mean(wait[0::2])), "barrier_after_sum_us": float(np.mean(wait[1::2])),
                      "total_us": (t[15, 0] - t0) / 1e3}))
    print(json.dumps(rows))
