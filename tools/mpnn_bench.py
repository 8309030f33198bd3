#!/usr/bin/env python
"""Config-5 leg alone (generic sum MPNN on one large graph, destination-partitioned over the ranks).

    python tools/mpnn_bench.py [--nodes N] [--edges E] [--variant uniform|skew|local] [--exchange peer|nccl|boundary]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/mpnn_bench.py ...
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def sweep(args, torch, dist, dev, rank, world):
    from ignnition_b200 import Engine, ModelDescription
    from ignnition_b200.parallel import PartitionedEngine
    n_nodes = args.nodes // 1024 * 1024
    n_edges = args.edges // 8 * 8
    md = ModelDescription(bench.mpnn_model_json(args.hidden), {"x": args.hidden, "adj": 0})
    eng = Engine(md, device=dev, seed=0)
    src, dst, x = bench.mpnn_shard(n_nodes, n_edges, args.hidden, args.variant, rank, world, torch, dev)
    pe = PartitionedEngine(eng, exchange="copy")
    pe.build({"node": n_nodes}, {"adj": (src, dst)}, {"x": x})
    del src, dst, x
    def timed(fn):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps

    ms = timed(lambda: pe.exchange_only("node"))
    if rank == 0:
        recv = (n_nodes - (pe.own("node")[1] - pe.own("node")[0])) * args.hidden * 4
        print(json.dumps({"sweep": "exchange only (copy engine, all ranks at once)", "n_gpus": world, "ms": ms,
                          "received_gbs_per_gpu": recv / ms / 1e6}), flush=True)
    for item in args.sweep.split(","):
        parts = item.split("x")
        ch, ns = int(parts[0]), int(parts[1])
        pe.chunks, pe.copy_streams = ch, ns
        pe.growth = float(parts[2]) if len(parts) > 2 else 1.0
        for _ in range(3):
            pe.message_passing(iterations=1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        pe.message_passing(iterations=args.steps)
        t1.record()
        torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t.item()) / args.steps
            print(json.dumps({"sweep": "copy", "n_gpus": world, "chunks": ch, "copy_streams": ns,
                              "growth": pe.growth, "ms_per_iteration": ms,
                              "g_edges_per_s": n_edges / ms / 1e6}), flush=True)
    # timeline of one update with the last configuration (rank 0), ms from the first kernel's start
    pe.message_passing(iterations=2)
    if world > 1:
        dist.barrier()
    pe.trace = []
    pe.message_passing(iterations=1)
    torch.cuda.synchronize()
    if rank == 0:
        t0 = pe.trace[0][1]
        print(json.dumps({"timeline_ms": [(lab, round(t0.elapsed_time(a), 3), round(t0.elapsed_time(b), 3))
                                          for lab, a, b in pe.trace]}), flush=True)
    pe.trace = None
    pe.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--edges", type=int, default=200_000_000)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--variant", default="uniform")
    ap.add_argument("--exchange", default="copy")
    ap.add_argument("--sweep", default="", help="copy exchange: 'chunks x streams' pairs, e.g. 2x1,4x2 (one graph build)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.sweep:
        sweep(args, torch, dist, dev, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return
    for variant in args.variant.split(","):
        for exchange in args.exchange.split(","):
            out = bench.run_mpnn(args.nodes, args.edges, args.hidden, args.steps, 3, torch, dev, variant, rank, world, exchange)
            if rank == 0:
                print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
