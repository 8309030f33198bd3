"""Data-parallel training of the PRODUCT path under NCCL (needs 2 GPUs; skipped otherwise).

Two ranks train on their shards of one RouteNet batch through ``Trainer.train_step`` (per-rank gradient of
SSE / GLOBAL prediction count, one NCCL all-reduce of the flat gradient buffer, replicated Adam); the weights
after a few steps must equal those of a single process trained on the whole batch, which is the reference's
MeanSquaredError over all predictions of the batch (``code/utils/generate_model.py:745-751``)."""

import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(md, n_samples):
    from ignnition_b200 import synthetic
    from ignnition_b200.generator import sample_to_tensors
    feats = [f.name for f in md.get_all_features()]
    samples, labels = [], []
    for k in range(n_samples):
        s = synthetic.routenet_sample("nsfnet" if k % 2 else "geant2", k % 5, k)
        t, y = sample_to_tensors(s, feats, "delay", md.get_adjecency_info(), [], [], True)
        t["traffic"] = (np.asarray(t["traffic"], np.float32) - 170) / 130
        t["link_capacity"] = (np.asarray(t["link_capacity"], np.float32) - 25000) / 40000
        samples.append(t)
        labels.append(np.log(np.asarray(y, np.float32)))
    return samples, labels


def _rank_main(rank, world, port, steps, n_samples, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ignnition_b200 import Engine, ModelDescription, ops
    from ignnition_b200.parallel import shard_samples
    from ignnition_b200.train import Trainer
    ops.set_tensor_cores(False)          # see the test: the same arithmetic at every batch size
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "routenet_nsfnet.json")))
        md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
        samples, labels = _batch(md, n_samples)
        costs = [len(t["src_adj_links_paths"]) for t in samples]
        idx = shard_samples(list(range(n_samples)), rank, world, costs)
        n_glob = sum(len(l) for l in labels)
        eng = Engine(md, device=dev, seed=7)
        tr = Trainer(eng, world_size=world)
        graph = eng.prepare([samples[i] for i in idx], labels=[labels[i] for i in idx], training=True)
        for _ in range(steps):
            tr.train_step(graph, global_n=n_glob)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "w_%d.npy" % rank), eng.weights.cpu().numpy())
        if rank == 0:
            json.dump(tr.losses(), open(os.path.join(out_dir, "loss.json"), "w"))
    finally:
        dist.destroy_process_group()


def test_trainer_nccl_world2_matches_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ignnition_b200 import Engine, ModelDescription
    from ignnition_b200.train import Trainer
    steps, n_samples, world = 4, 32, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, steps, n_samples, str(tmp_path)), nprocs=world, join=True)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "routenet_nsfnet.json")))
    md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
    samples, labels = _batch(md, n_samples)
    # fp32 kernels in the workers and here: the engine picks kernels by batch size (3xTF32 Dense layers from 4096 rows:
    # the 32-sample reference batch has 5824 paths, a rank's half 2912), and four Adam steps turn a 1e-6 difference of a
    # near-zero gradient into 1e-4 of a weight; with one arithmetic only the summation order differs
    from ignnition_b200 import ops
    prev = ops.set_tensor_cores(False)
    try:
        ref = Engine(md, device="cuda", seed=7)
        rt = Trainer(ref, world_size=1)
        rg = ref.prepare(samples, labels=labels, training=True)
        for _ in range(steps):
            rt.train_step(rg)
    finally:
        ops.set_tensor_cores(prev)
    want = ref.weights.cpu().numpy()
    w0 = np.load(os.path.join(str(tmp_path), "w_0.npy"))
    w1 = np.load(os.path.join(str(tmp_path), "w_1.npy"))
    assert np.array_equal(w0, w1)                        # replicated Adam on the same all-reduced gradient
    scale = max(float(np.abs(want).max()), 1.0)
    assert float(np.abs(w0 - want).max()) <= 2e-5 * scale
    loss = json.load(open(os.path.join(str(tmp_path), "loss.json")))
    assert abs(loss["loss"] - rt.losses()["loss"]) <= 1e-4 * max(abs(rt.losses()["loss"]), 1e-6)
