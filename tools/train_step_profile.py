#!/usr/bin/env python
"""Development tool: wall time and CUDA-kernel count of one PolicyValueNet.train_step (batch 128, 15x15) alone on the GPU."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet  # noqa: E402

net = PolicyValueNet(15, 15, batch_size=128, seed=0)
rs = np.random.RandomState(0)
st = torch.tensor((rs.rand(128, 9, 15, 15) < 0.2).astype(np.float32), device="cuda")
pi = torch.tensor(rs.dirichlet(np.ones(225), size=128).astype(np.float32), device="cuda")
z = torch.tensor(rs.choice([-1.0, 1.0], size=128).astype(np.float32), device="cuda")
for _ in range(3):
    net.train_step(st, pi, z, 1e-3, sync=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    net.train_step(st, pi, z, 1e-3, sync=False)
torch.cuda.synchronize()
print("train_step alone: %.2f ms" % (100 * (time.perf_counter() - t0)))
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net.train_step(st, pi, z, 1e-3, sync=False)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
print("CUDA kernels / memcpys in one train_step:", len(ev), " total device time %.2f ms" % (sum(e.device_time_total for e in ev) / 1e3))
t0 = time.perf_counter()
sth = st.cpu().numpy()
for _ in range(10):
    net.policy_value(sth)
print("policy_value(128 states) alone: %.2f ms" % (100 * (time.perf_counter() - t0)))
