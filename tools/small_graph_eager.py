#!/usr/bin/env python
"""A few eager forwards of the NSFNET batch (argv[1] samples) through the one-launch loop, for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ignnition_b200 import Engine, ModelDescription
from ignnition_b200.batching import assemble_tiled

g, shape, qsize, _ = bench.load_case("routenet_nsfnet_b4096")
md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = Engine(md, device="cuda", seed=0)
batch = assemble_tiled(g["reference_tensors"][0], n, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                       bench.feature_fns(qsize), seed=0)
graph = eng.prepare(batch)
for _ in range(3):
    eng.forward(graph)
torch.cuda.synchronize()
