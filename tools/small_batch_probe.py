#!/usr/bin/env python
"""Config 1 sizes (RouteNet NSFNET, batch 3 / 32): graph replay time with the tensor-core kernels and with the fp32
twins, and the kernels inside the captured graph."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ignnition_b200 import Engine, ModelDescription, ops
from ignnition_b200.batching import assemble_tiled

g, shape, qsize, _ = bench.load_case("routenet_nsfnet_b4096")
dims = g["reference_meta"]["dimensions"]
md = ModelDescription(g["model_json"], dims)
for tc in (True, False):
    ops.set_tensor_cores(tc)
    for n in (3, 32):
        eng = Engine(md, device="cuda", seed=0)
        batch = assemble_tiled(g["reference_tensors"][0], n, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                               bench.feature_fns(qsize), seed=0)
        pinned = eng.pack(batch)
        eng.forward_graphed(batch, pinned)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            eng.forward_graphed(batch, pinned, copy=False)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"tensor_cores": tc, "batch": n, "replay_us": e0.elapsed_time(e1) * 5.0,
                          "kernels_in_graph": eng.graphed_kernels(batch, pinned)}))
ops.set_tensor_cores(True)
