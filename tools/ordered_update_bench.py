#!/usr/bin/env python
"""Times one ordered update (RouteNet stage 1 shape: 2.26 M paths, 303 k links, ~6.1 M incidences, 32-wide)
with the sequence walk (ign_gru_seq) and with the step-synchronous launches (ign_gru_seq_step), CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ignnition_b200 import ops  # noqa: E402

rng = np.random.RandomState(0)
n_dst, n_src, u, max_len = 2_260_992, 303_104, 32, 6
lens = rng.choice([1, 2, 3, 4, 5, 6], n_dst, p=[0.16, 0.31, 0.29, 0.15, 0.07, 0.02])
r = np.zeros(n_dst + 1, np.int64)
np.cumsum(lens, out=r[1:])
c = rng.randint(0, n_src, int(r[-1]))
rp = torch.from_numpy(r).to(torch.int32).cuda()
cc = torch.from_numpy(c).to(torch.int32).cuda()
order = ops.length_order(rp)
meta = ops.seq_meta(rp, cc, order)
plan = ops.seq_step_plan(meta, cc, max_len)
states = torch.randn(n_src, u, device="cuda") * 0.5
h0 = torch.randn(n_dst, u, device="cuda")
K = torch.randn(u, 3 * u, device="cuda") * 0.2
R = torch.randn(u, 3 * u, device="cuda") * 0.2
b = torch.randn(2, 3 * u, device="cuda") * 0.1
out1, out2, hs = torch.empty_like(h0), torch.empty_like(h0), torch.empty_like(h0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


alg = int(r[-1]) * (4 + 4 * u) + n_dst * (8 * u + 8)
out3 = torch.empty_like(h0)
t_proj = timed(lambda: ops.gru_seq(rp, cc, order, [states], h0, K, R, b, out=out3, meta=meta))   # hoisted projection
os.environ["IGN_GRU_SEQ_PROJ"] = "0"
t_walk = timed(lambda: ops.gru_seq(rp, cc, order, [states], h0, K, R, b, out=out1, meta=meta))
t_step = timed(lambda: ops.gru_seq_steps(plan, meta, [states], h0, K, R, b, max_len, out=out2, hs=hs))
torch.cuda.synchronize()
alg = int(r[-1]) * (4 + 4 * u) + n_dst * (8 * u + 8)
print("hoisted projection + TMEM state operand, 3 walkers: %.3f ms (%.0f GB/s algorithmic), max |diff| vs walk %.2e"
      % (t_proj, alg / t_proj / 1e6, float((out3 - out1).abs().max())))
print("steps %d | walk %.3f ms (%.0f GB/s algorithmic) | step-synchronous %.3f ms (%.0f GB/s) | max |diff| %.2e"
      % (int(r[-1]), t_walk, alg / t_walk / 1e6, t_step, alg / t_step / 1e6, float((out1 - out2).abs().max())))
