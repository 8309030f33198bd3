"""Probe of the tensor-core weight-gradient kernel (csrc/dw_tc.cu) with patterned inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ignnition_b200 import ops, _lib

lib = _lib.load()
m, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.zeros(m, k, device="cuda"); dz = torch.zeros(m, n, device="cuda")
x[:] = torch.arange(1, k + 1, device="cuda").float()[None, :]
dz[:] = 1.0
dz[:, 1] = 2.0
w = torch.zeros(k, n, device="cuda"); dw = torch.zeros(k, n, device="cuda")
c0 = lib.ign_launch_count()
ops.dense_bwd(x, w, 0, None, dz, None, dw, None)
torch.cuda.synchronize()
print("launches", lib.ign_launch_count() - c0)
print(dw[:4, :4].cpu().numpy() / m)
print(dw[-2:, -3:].cpu().numpy() / m)
want = (x.double().T @ dz.double())
print("rel err", float((dw.double() - want).abs().max() / want.abs().max()))
