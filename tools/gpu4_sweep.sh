#!/bin/bash
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 tools/mpnn_bench.py --steps 8 --sweep 4x1x0,8x1x0,8x1x1.0,6x1x0,12x1x1.0 2>/dev/null | grep sweep
