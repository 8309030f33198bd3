#!/usr/bin/env python
"""Times the backward pass of one ordered update (RouteNet stage 1 shape: 2.26 M paths, 303 k links, ~6.1 M
incidences, 32-wide): fp32 tile walk (ign_gru_seq_bwd) vs step-synchronous tcgen05 launches
(ign_gru_seq_bwd_steps), CUDA events.

With the profiling build (tools/build_profile_lib.sh, copied over ignnition_b200/libignnition_b200.so on the GPU box):
  * extra arguments are ablation flags of the step kernel (results invalid, timing only): 1 no G stores, 2 no dL/dh
    loads in the epilogue, 4 no output rows, 8 no L2-prefetch warps, 16 skip the weight-gradient kernel, 32 all producer
    rows from one 128 KB region;
  * the per-phase cycle counts of one warp per role are printed (IGN_PROF_FLAGS=<flags> profiles under an ablation)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ignnition_b200 import ops, _lib  # noqa: E402

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
d_out = torch.randn(n_dst, u, device="cuda")
h_seq = torch.empty(int(r[-1]), u, device="cuda")
ops.gru_seq(rp, cc, order, [states], h0, K, R, b, h_seq=h_seq, meta=meta)
d_steps = torch.empty(int(r[-1]), u, device="cuda")
dh0 = torch.empty_like(h0)
dk, dr, db = torch.zeros_like(K), torch.zeros_like(R), torch.zeros_like(b)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


walk = lambda: ops.gru_seq_bwd(rp, cc, order, [states], h0, h_seq, K, R, b, d_out, d_steps, dh0, dk, dr, db)
step = lambda: ops.gru_seq_bwd_steps(plan, meta, max_len, [states], h0, h_seq, K, R, b, d_out, d_steps, dh0, dk, dr, db)
print("steps %d | fp32 walk %.3f ms | step-synchronous tcgen05 %.3f ms" % (int(r[-1]), timed(walk, 3), timed(step)))
dbg = getattr(_lib.load(), "ign_debug_bwd", None)
if dbg is not None and len(sys.argv) > 1:
    for v in [int(a) for a in sys.argv[1:]]:
        dbg(v)
        print("  debug flags %2d: %.3f ms" % (v, timed(step)))
    dbg(0)

prof = getattr(_lib.load(), "ign_debug_bwd_prof", None)
if prof is not None:
    import ctypes as C
    buf = (C.c_ulonglong * 16)()
    prof(buf, 1)
    if dbg is not None:
        dbg(int(os.environ.get("IGN_PROF_FLAGS", "0")))      # phase profile under an ablation
    step()
    torch.cuda.synchronize()
    prof(buf, 0)
    v = [x / 148.0 for x in buf]       # cycles per CTA (one warp of each role), summed over the launches of one call
    tiles = int(r[-1]) / 128 / 148
    names = ["prod: loads landed", "prod: wait GEMM1(j-1)", "prod: split+store", "epi: prologue/fetch", "epi: wait acc1",
             "epi: tmem ld + gates", "epi: wait gdone", "epi: split + tmem st", "epi: G staging + bulk", "epi: wait acc2",
             "epi: out stores"]
    print("cycles per tile-step (%.0f tile-steps per CTA):" % tiles)
    for n, x in zip(names, v):
        print("  %-24s %8.0f" % (n, x / tiles))
