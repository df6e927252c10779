#!/usr/bin/env python
"""Development tool: what does a second stream of small kernels see while ap_search_run saturates the GPU?
Times, with the 4096-game search running in a background thread: one tiny kernel + sync, a chain of 100 tiny kernels,
one train_step, one policy_value - on a default-priority and on a highest-priority torch stream."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet  # noqa: E402

net = PolicyValueNet(15, 15, batch_size=128, seed=0)
eng = net.search_engine(n_games=4096, n_playout=400, node_capacity=400 * 225 + 2)
bench.synthetic_positions(eng, 4096)
rs = np.random.RandomState(0)
st = torch.tensor((rs.rand(128, 9, 15, 15) < 0.2).astype(np.float32), device="cuda")
pi = torch.tensor(rs.dirichlet(np.ones(225), size=128).astype(np.float32), device="cuda")
z = torch.tensor(rs.choice([-1.0, 1.0], size=128).astype(np.float32), device="cuda")
sth = st.cpu().numpy()
for _ in range(2):
    net.train_step(st, pi, z, 1e-3, sync=False)
    net.policy_value(sth)
stop = False


def searcher():
    while not stop:
        eng.search_advance(-1)
        eng.search_run(400)


def measure(tag, stream):
    x = torch.zeros(1024, device="cuda")
    with torch.cuda.stream(stream):
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            x.add_(1.0)
            stream.synchronize()
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(100):
            x.add_(1.0)
        stream.synchronize()
        chain = time.perf_counter() - t0
        t0 = time.perf_counter()
        net.train_step(st, pi, z, 1e-3, sync=False)
        stream.synchronize()
        tr = time.perf_counter() - t0
    t0 = time.perf_counter()
    net.policy_value(sth)
    pv = time.perf_counter() - t0
    print("%-34s 1 kernel+sync %.3f ms (max %.3f) | 100-kernel chain %.2f ms | train_step %.1f ms | policy_value %.1f ms"
          % (tag, 1e3 * np.median(ts), 1e3 * max(ts), 1e3 * chain, 1e3 * tr, 1e3 * pv))


lo = torch.cuda.Stream()
hi = torch.cuda.Stream(priority=-1)
measure("idle GPU, default priority", lo)
measure("idle GPU, high priority", hi)
th = threading.Thread(target=searcher)
th.start()
time.sleep(1.0)
for _ in range(2):
    measure("search running, default priority", lo)
    measure("search running, high priority", hi)
stop = True
th.join()
