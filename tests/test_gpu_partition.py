"""Destination-partitioned graph (BASELINE config 5) as product code: ``ignnition_b200.parallel``.

* one rank: PartitionedEngine == Engine.forward == the CPU oracle on the same graph;
* the routing helpers (owner ranks, stable split, flag compaction, row puts) against numpy;
* two ranks (needs 2 GPUs; spawned here with torch.multiprocessing over NCCL, 127.0.0.1): every exchange
  ('copy' = copy-engine pushes of finished row chunks, 'peer' = TMA stores from the update kernel into the
  peer-mapped arrays, 'boundary', 'nccl') gives the
  1-rank states BIT FOR BIT, from contiguous shards of the same global edge list."""

import os
import sys

import numpy as np
import pytest
import torch

from oracle import ignnition_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mpnn_json(hidden, iterations, agg="sum"):
    return {
        "entities": [{"name": "node", "hidden_state_dimension": hidden,
                      "features": [{"name": "x", "normalization": "None"}]}],
        "message_passing": {"num_iterations": iterations, "stages": [{"stage_name": "s", "stage_mp": [{
            "destination_entity": "node",
            "source_entities": [{"name": "node", "adj_vector": "adj", "message": [{"type": "direct_assignation"}]}],
            "aggregation": {"type": agg},
            "update": {"type": "recurrent_neural_network", "nn_name": "rec"}}]}]},
        "readout": [{"type": "predict", "input": ["node"], "label": "y", "nn_name": "ro"}],
        "neural_networks": [
            {"nn_name": "rec", "nn_type": "recurrent_neural_network", "recurrent_type": "GRU"},
            {"nn_name": "ro", "nn_type": "feed_forward", "nn_architecture": [
                {"type_layer": "Dense", "units": 1, "activation": "None"}]}],
        "learning_options": {"loss": "MeanSquaredError", "optimizer": {"type": "Adam"}},
    }


def global_graph(n, e, hidden, seed=0, p_local=0.0, world=1):
    rng = np.random.RandomState(seed)
    dst = rng.randint(0, n, e).astype(np.int32)
    src = rng.randint(0, n, e).astype(np.int32)
    if p_local > 0:           # variant C: most sources live on the destination's owner
        per = -(-n // max(world, 1))
        local = rng.rand(e) < p_local
        src = np.where(local, (dst // per) * per + rng.randint(0, per, e), src).clip(0, n - 1).astype(np.int32)
    x = rng.randn(n, hidden).astype(np.float32)
    return src, dst, x


def build_engine(hidden, T, agg="sum", seed=3):
    from ignnition_b200 import Engine, ModelDescription
    md = ModelDescription(mpnn_json(hidden, T, agg), {"x": hidden, "adj": 0})
    # fuse_sum_gru=False: the single-GPU engine would take its one-launch fp32 update for graphs this small; the
    # comparison below is bit for bit against the tensor-core update the partitioned engine runs (small_graph_rows = 0
    # for the same reason: no one-launch fp32 loop)
    eng = Engine(md, device="cuda", seed=seed, fuse_sum_gru=False)
    eng.small_graph_rows = 0
    return eng


@pytest.mark.parametrize("hidden,agg", [(64, "sum"), (32, "sum"), (64, "mean"), (32, "max")])
def test_partitioned_one_rank_matches_engine_and_oracle(hidden, agg):
    from ignnition_b200 import ops
    from ignnition_b200.engine import DeviceGraph
    from ignnition_b200.parallel import PartitionedEngine
    n, e, T = 3000, 40000, 3
    src, dst, x = global_graph(n, e, hidden, seed=hidden)
    eng = build_engine(hidden, T, agg)
    pe = PartitionedEngine(eng, exchange="peer")
    pe.build({"node": n}, {"adj": (torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda())},
             {"x": torch.from_numpy(x).cuda()})
    pred = pe.forward().cpu().numpy()
    states = pe.state("node").cpu().numpy()
    pe.close()
    # the same graph through the single-GPU engine (stable sort by destination = the same slot order)
    g = DeviceGraph()
    g.num = {"node": n}
    g.n_samples = 1
    g.t = {"feat_x": torch.from_numpy(x).cuda(), "src_adj": torch.from_numpy(src).cuda(),
           "dst_adj": torch.from_numpy(dst).cuda()}
    eng.build_graph(g)
    pred2, st2 = eng.forward(g, return_states=True)
    assert np.array_equal(states, st2["node"].cpu().numpy())
    assert np.array_equal(pred, pred2.cpu().numpy())
    # oracle (fp64) on the reference's tensor dict
    order = np.argsort(dst, kind="stable")
    seq = np.zeros(e, np.int64)
    cnt = {}
    for i in order:
        seq[i] = cnt.get(dst[i], 0)
        cnt[dst[i]] = seq[i] + 1
    o = orc.Oracle(mpnn_json(hidden, T, agg), {"x": hidden, "adj": 0}, dtype=np.float64)
    w = {k: v.astype(np.float64) for k, v in eng.get_weights().items()}
    t = {"x": x.astype(np.float64), "num_node": n, "src_adj": src.astype(np.int64), "dst_adj": dst.astype(np.int64),
         "seq_node_node": seq}
    want = o.forward(t, w).reshape(-1)
    err = float(np.abs(pred.reshape(-1) - want).max() / np.abs(want).max())
    assert err < 1e-5, err


def test_routing_helpers():
    from ignnition_b200 import ops
    from ignnition_b200.parallel import node_bounds, split_counts
    rng = np.random.RandomState(5)
    n, e, world = 1003, 20000, 4
    bounds = node_bounds(n, world)
    assert bounds[0] == 0 and bounds[-1] == n and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
    dst = rng.randint(0, n, e).astype(np.int32)
    src = rng.randint(0, n, e).astype(np.int32)
    owner = ops.edge_owner(torch.from_numpy(dst).cuda(), bounds).cpu().numpy()
    want_owner = np.searchsorted(np.asarray(bounds[1:]), dst, side="right")
    assert np.array_equal(owner, want_owner)
    rowptr, src_sorted, perm, _ = ops.csr_build(torch.from_numpy(owner).cuda(), torch.from_numpy(src).cuda(), None,
                                                world, ops.CSR_SORT, want_perm=True)
    order = np.argsort(want_owner, kind="stable")
    assert np.array_equal(perm.cpu().numpy(), order) and np.array_equal(src_sorted.cpu().numpy(), src[order])
    assert split_counts(rowptr.cpu().tolist()) == [int((want_owner == r).sum()) for r in range(world)]
    assert np.array_equal(ops.gather_int(torch.from_numpy(dst).cuda(), perm, add=-7).cpu().numpy(), dst[order] - 7)
    flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    ops.mark_rows(torch.from_numpy(src[:500]).cuda(), flags)
    rows, cnt = ops.flag_compact(flags, add=3)
    uniq = np.unique(src[:500])
    assert int(cnt.item()) == len(uniq) and np.array_equal(rows[:len(uniq)].cpu().numpy(), uniq + 3)
    a = torch.randn(n, 64, device="cuda")
    b = torch.zeros(n, 64, device="cuda")
    pick = torch.from_numpy(uniq.astype(np.int32)).cuda()
    ops.rows_put(a, pick, b)
    assert torch.equal(b[pick.long()], a[pick.long()])
    assert int((b.abs().sum(1) > 0).sum()) == len(uniq)              # no other row was touched


def _rank_main(rank, world, port, exchange, hidden, n, e, T, p_local, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ignnition_b200.parallel import PartitionedEngine
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        src, dst, x = global_graph(n, e, hidden, seed=hidden, p_local=p_local, world=world)
        lo_e, hi_e = rank * e // world, (rank + 1) * e // world        # contiguous shard of the global edge list
        from ignnition_b200 import Engine, ModelDescription
        md = ModelDescription(mpnn_json(hidden, T), {"x": hidden, "adj": 0})
        eng = Engine(md, device=torch.device("cuda", rank), seed=3)
        pe = PartitionedEngine(eng, exchange=exchange, chunks=3)          # 'copy': three row chunks per update
        from ignnition_b200.parallel import node_bounds
        b = node_bounds(n, world)
        pe.build({"node": n}, {"adj": (torch.from_numpy(src[lo_e:hi_e]).cuda(), torch.from_numpy(dst[lo_e:hi_e]).cuda())},
                 {"x": torch.from_numpy(x[b[rank]:b[rank + 1]]).cuda()})
        pred = pe.forward()
        full = pe.full_state("node").cpu().numpy()             # every rank must hold ALL rows after the exchange
        np.save(os.path.join(out_dir, "full_%s_%d.npy" % (exchange, rank)), full)
        np.save(os.path.join(out_dir, "pred_%s_%d.npy" % (exchange, rank)), pred.cpu().numpy())
        pe.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange,p_local", [("copy", 0.0), ("peer", 0.0), ("nccl", 0.0), ("boundary", 0.9)])
def test_partitioned_two_ranks_bitwise(exchange, p_local, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ignnition_b200.parallel import PartitionedEngine, node_bounds
    hidden, n, e, T, world = 64, 10007, 150000, 3, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, exchange, hidden, n, e, T, p_local, str(tmp_path)), nprocs=world, join=True)
    # the 1-rank run of the same global graph in this process
    src, dst, x = global_graph(n, e, hidden, seed=hidden, p_local=p_local, world=world)
    eng = build_engine(hidden, T)
    pe = PartitionedEngine(eng, exchange="peer")
    pe.build({"node": n}, {"adj": (torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda())},
             {"x": torch.from_numpy(x).cuda()})
    pred1 = pe.forward().cpu().numpy()
    full1 = pe.full_state("node").cpu().numpy()
    pe.close()
    b = node_bounds(n, world)
    needed = np.zeros(n, bool)
    for r in range(world):
        full = np.load(os.path.join(str(tmp_path), "full_%s_%d.npy" % (exchange, r)))
        pred = np.load(os.path.join(str(tmp_path), "pred_%s_%d.npy" % (exchange, r)))
        assert np.array_equal(pred, pred1[b[r]:b[r + 1]])
        if exchange == "boundary":      # a rank holds its own rows and the rows its edges read
            mine = (dst >= b[r]) & (dst < b[r + 1])
            needed[:] = False
            needed[src[mine]] = True
            needed[b[r]:b[r + 1]] = True
            assert np.array_equal(full[needed], full1[needed])
        else:
            assert np.array_equal(full, full1)
