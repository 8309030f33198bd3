#!/usr/bin/env python
"""A few replays of the captured NSFNET train step (batch 3 / 32) for an ncu launch list of the graph's kernel nodes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ignnition_b200 import Engine, ModelDescription
from ignnition_b200.batching import assemble_tiled
from ignnition_b200.train import Trainer

g, shape, qsize, _ = bench.load_case("routenet_nsfnet_b4096")
md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = Engine(md, device="cuda", seed=0)
out_entity = [o for o in md.get_readout_operations() if o.type == "predict"][0].input[0]
batch = assemble_tiled(g["reference_tensors"][0], n, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                       bench.feature_fns(qsize), seed=0, label_fn=lambda r, k: r.normal(-1.0, 0.5, k), label_entity=out_entity)
pinned = eng.pack(batch)
tr = Trainer(eng)
for _ in range(4):
    tr.train_step_graphed(batch, pinned)
torch.cuda.synchronize()
