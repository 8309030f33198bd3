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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--edges", type=int, default=200_000_000)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--variant", default="uniform")
    ap.add_argument("--exchange", default="copy")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for variant in args.variant.split(","):
        for exchange in args.exchange.split(","):
            out = bench.run_mpnn(args.nodes, args.edges, args.hidden, args.steps, 3, torch, dev, variant, rank, world, exchange)
            if rank == 0:
                print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
