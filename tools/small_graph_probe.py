#!/usr/bin/env python
"""Config 1 sizes (RouteNet NSFNET, batch 3 / 32): where the time of one forward goes -- adjacency build, initial
states, the one-launch message-passing loop (csrc/small_graph.cu) at T = 8 and T = 1, readout -- each timed alone with
CUDA events over 200 repetitions (eager launches, so launch overhead is included in every part)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ignnition_b200 import Engine, ModelDescription
from ignnition_b200.batching import assemble_tiled

g, shape, qsize, _ = bench.load_case("routenet_nsfnet_b4096")
md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])


def timed(fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps


for n in (3, 32, 128):
    eng = Engine(md, device="cuda", seed=0)
    batch = assemble_tiled(g["reference_tensors"][0], n, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                           bench.feature_fns(qsize), seed=0)
    pinned = eng.pack(batch)
    dg = eng.upload(batch, pinned)
    out = {"batch": n, "rows": dict(dg.num)}
    out["build_us"] = timed(lambda: eng.build_graph(eng.upload(batch, pinned)))
    out["upload_us"] = timed(lambda: eng.upload(batch, pinned))
    graph = eng.build_graph(eng.upload(batch, pinned))
    out["small"] = graph.small
    out["init_us"] = timed(lambda: eng.initial_states(graph))
    st = eng.initial_states(graph)
    for T in (8, 1):
        out["loop_T%d_us" % T] = timed(lambda: eng.message_passing(graph, dict(st), iterations=T))
    fin = eng.message_passing(graph, dict(st))
    out["readout_us"] = timed(lambda: eng.readout_forward(fin, g=graph))
    eng.forward_graphed(batch, pinned)
    out["graph_replay_us"] = timed(lambda: eng.forward_graphed(batch, pinned, copy=False))
    out["kernels_in_graph"] = eng.graphed_kernels(batch, pinned)
    print(json.dumps(out))
