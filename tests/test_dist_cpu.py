"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sample sharding and the data-parallel
gradient rule the Trainer applies (sum of per-rank gradients of SSE / GLOBAL prediction count ==
gradient of the reference's MeanSquaredError over the whole batch, generate_model.py:745-751)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden
from ignnition_b200.parallel import shard_bounds, shard_samples


def test_shard_bounds_cover_and_balance():
    costs = [5, 1, 1, 1, 8, 2, 2, 4]
    for world in (1, 2, 3, 4, 8):
        b = shard_bounds(costs, world)
        assert b[0][0] == 0 and b[-1][1] == len(costs)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(hi > lo for lo, hi in b)
    assert shard_samples(list("abcdefgh"), 1, 2, costs) == list("efgh")


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ignnition_oracle as orc
    from oracle.torch_port import TorchOracle
    g = load_golden("routenet_nsfnet")
    dims = g["reference_meta"]["dimensions"]
    o = orc.Oracle(g["model_json"], dims, dtype=np.float64)
    w = {k: v.astype(np.float32) for k, v in o.init_weights(1234).items()}
    tens = [orc.normalize_inputs(g["model_json"], t) for t in g["reference_tensors"]]
    labels = [np.log(np.asarray(y)) for y in g["reference_labels"]]
    to = TorchOracle(g["model_json"], dims)
    mine = shard_samples(list(range(len(tens))), rank, world)
    n_glob = sum(len(l) for l in labels)
    # rank-local: gradient of SSE / n_glob  (what Trainer.loss_and_grads computes with grad_scale = 1/n_glob)
    wt = to.params(w)
    preds = torch.cat([to.forward(tens[i], wt).reshape(-1) for i in mine])
    y = torch.tensor(np.concatenate([labels[i] for i in mine]))
    (((preds - y) ** 2).sum() / n_glob).backward()
    names = sorted(wt)
    flat = torch.cat([wt[k].grad.reshape(-1) for k in names])
    dist.all_reduce(flat)                                   # the NCCL all-reduce of Trainer.apply
    if rank == 0:
        # single-process reference: MSE over all predictions of the global batch (no regulariser here)
        wt2 = to.params(w)
        p2 = torch.cat([to.forward(t, wt2).reshape(-1) for t in tens])
        y2 = torch.tensor(np.concatenate(labels))
        torch.mean((p2 - y2) ** 2).backward()
        ref = torch.cat([wt2[k].grad.reshape(-1) for k in names])
        q.put(float((flat - ref).abs().max() / ref.abs().max()))
    dist.destroy_process_group()


def test_data_parallel_gradient_rule_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12


def test_partition_host_logic():
    """destination-range bounds and the row chunks of the copy-engine exchange (ignnition_b200.parallel): pure host
    logic of the partitioned graph, no GPU"""
    from ignnition_b200.parallel import chunk_cuts, exchange_growth, node_bounds, shard_bounds
    for n, w in ((10, 3), (9_999_360, 8), (5, 8), (0, 2)):
        b = node_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = [b[i + 1] - b[i] for i in range(w)]
        assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
    g8 = exchange_growth(8, 256, 20.0, 64)
    g2 = exchange_growth(2, 256, 20.0, 64)
    assert 1.8 < g8 < 2.2 and 0.25 <= g2 < 0.35            # NVLink is the longer leg at 8 GPUs, the kernel at 2
    for n, k, g in ((1_249_920, 4, g8), (4_999_680, 4, g2), (1000, 8, 1.0), (50, 4, 2.0), (0, 4, 2.0)):
        cuts = chunk_cuts(n, k, g)
        assert cuts[0] == 0 and cuts[-1] == n and cuts == sorted(set(cuts)) and len(cuts) - 1 <= max(k, 1)
        assert all(c % 128 == 0 for c in cuts[:-1])
    c8 = chunk_cuts(1_249_920, 4, g8)
    assert c8[1] - c8[0] < c8[-1] - c8[-2]                  # growing chunks: the copy engine starts early
    c2 = chunk_cuts(4_999_680, 4, g2)
    assert c2[1] - c2[0] > c2[-1] - c2[-2]                  # shrinking chunks: little exchange left at the end
    assert shard_bounds([1.0] * 10, 3) == [(0, 3), (3, 6), (6, 10)] or sum(hi - lo for lo, hi in shard_bounds([1.0] * 10, 3)) == 10
