#!/usr/bin/env python
"""Data-parallel training check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/dp_train_check.py

Every rank trains on its shard of a RouteNet batch (NCCL all-reduce of the flat gradient buffer,
gradients scaled by 1 / GLOBAL prediction count); rank 0 also trains a single-process copy on the
whole batch.  After a few Adam steps the two sets of weights must agree to fp32 round-off, which is
the reference's semantics: MeanSquaredError over all predictions of the batch
(code/utils/generate_model.py:745-751).  Also prints the step time."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ignnition_b200 import Engine, ModelDescription, synthetic                      # noqa: E402
from ignnition_b200.generator import sample_to_tensors                               # noqa: E402
from ignnition_b200.parallel import rank_world, shard_samples                        # noqa: E402
from ignnition_b200.train import Trainer                                             # noqa: E402


def main():
    rank, world, local = rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "routenet_nsfnet.json")))
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    feats = [f.name for f in md.get_all_features()]
    n_samples = 64
    samples, labels = [], []
    for k in range(n_samples):
        s = synthetic.routenet_sample("nsfnet" if k % 2 else "geant2", k % 5, k)
        t, y = sample_to_tensors(s, feats, "delay", md.get_adjecency_info(), [], [], True)
        t["traffic"] = (np.asarray(t["traffic"], np.float32) - 170) / 130
        t["link_capacity"] = (np.asarray(t["link_capacity"], np.float32) - 25000) / 40000
        samples.append(t)
        labels.append(np.log(np.asarray(y, np.float32)))
    costs = [len(t["src_adj_links_paths"]) for t in samples]
    idx = shard_samples(list(range(n_samples)), rank, world, costs)
    n_glob = sum(len(l) for l in labels)

    eng = Engine(md, device=dev, seed=7)
    tr = Trainer(eng, world_size=world)
    graph = eng.prepare([samples[i] for i in idx], labels=[labels[i] for i in idx], training=True)
    steps = 5
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(steps):
        if k == 1:
            e0.record()
        tr.train_step(graph, global_n=n_glob)
    e1.record()
    torch.cuda.synchronize()
    loss = tr.losses()
    if rank == 0:
        ref = Engine(md, device=dev, seed=7)
        rt = Trainer(ref, world_size=1)
        rg = ref.prepare(samples, labels=labels, training=True)
        for _ in range(steps):
            rt.train_step(rg)
        diff = float((ref.weights - eng.weights).abs().max())
        scale = float(ref.weights.abs().max())
        print(json.dumps({"world": world, "samples": n_samples, "steps": steps, "dp_loss": loss,
                          "single_loss": rt.losses(), "max_weight_diff": diff, "max_weight": scale,
                          "ms_per_train_step": e0.elapsed_time(e1) / (steps - 1),
                          "ok": bool(diff <= 2e-5 * max(scale, 1.0))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
