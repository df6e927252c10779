"""Self-play + train loop over all GPUs of one box (BASELINE.json configs[4]).

Every rank (one process per GPU) plays ``n_games`` concurrent self-play games on its own engine - no
cross-GPU traffic during search.  Once per iteration the two exchanges of SURVEY 8(e) happen:
finished-game records are all-gathered as fixed-size packed rows (``dist.gather_replay``, NCCL), rank 0
pushes them into its device replay ring and runs ``epochs`` train steps on minibatches the gather kernel
writes straight into device tensors, then the flat fp32 weight buffer is broadcast to every rank
(``dist.broadcast_weights``, one ncclBroadcast) and each rank rebuilds its fp16 operand images locally.

The single-GPU form of the same loop is the reference's ``TrainPipeline.run`` (train_mxnet.py:265-283:
collect one game -> policy_update -> repeat).
"""
import time

import numpy as np

from . import dist as apdist
from .replay import ReplayBuffer
from .selfplay import BatchedSelfPlay


def selfplay_train_loop(net, n_games, n_iters, plies_per_iter=4, n_playout=400, c_puct=5, temp=1.0, batch_size=128,
                        epochs=8, learn_rate=4e-4, buffer_size=2198800, n_in_row=5, seed=0, log=None, device_pick=False):
    """Returns a dict of counters / timings (per rank; wall clock).  device_pick: moves sampled on the device with
    the next ply's search overlapped with the host bookkeeping (``BatchedSelfPlay(device_pick=True)``); the search in
    flight is joined before the weights change, so it is one ply stale at most."""
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    world = dist.get_world_size() if multi else 1
    S = net.board_width * net.board_height
    sp = BatchedSelfPlay(net, n_games, n_playout=n_playout, c_puct=c_puct, temp=temp, n_in_row=n_in_row,
                         seed=seed + 1000 * rank, device_pick=device_pick)
    ring = ReplayBuffer(net._eng, buffer_size) if rank == 0 else None
    if multi:
        apdist.broadcast_weights(net, src=0)  # identical start
    out = dict(plies=0, playouts=0, games=0, train_steps=0, records=0, t_selfplay=0.0, t_exchange=0.0, t_train=0.0,
               losses=[])
    dev = "cuda:%d" % net._device
    for it in range(n_iters):
        t0 = time.perf_counter()
        packed = []
        for _ in range(plies_per_iter):
            for winner, states, pis, zs in sp.step():
                out["games"] += 1
                if states is not None:
                    packed.append(apdist.pack_records(states, pis, zs, S))
        sp.drain()  # nothing may search while the weights are trained / broadcast
        out["plies"] += plies_per_iter * n_games
        out["playouts"] += plies_per_iter * n_games * n_playout
        mine = np.concatenate(packed, axis=0) if packed else np.zeros((0, apdist.record_width(S)), np.uint8)
        t1 = time.perf_counter()
        allrec = apdist.gather_replay(mine) if multi else mine
        t2 = time.perf_counter()
        if rank == 0:
            if allrec.shape[0]:
                sb = (9 * S + 7) // 8
                w = apdist.record_width(S)
                off = w - 4 * S - 4
                pis = np.ascontiguousarray(allrec[:, off:off + 4 * S]).view(np.float32).reshape(-1, S)
                zs = np.ascontiguousarray(allrec[:, off + 4 * S:]).view(np.float32).reshape(-1)
                ring.extend_positions(np.ascontiguousarray(allrec[:, :sb]), pis, zs)
                out["records"] += allrec.shape[0]
            if len(ring) > batch_size:
                for _ in range(epochs):
                    st, pi, z = ring.sample_torch(batch_size, dev)
                    loss, _ = net.train_step(st, pi, z, learn_rate)
                    out["train_steps"] += 1
                out["losses"].append(float(loss[0]))
        if multi:
            apdist.broadcast_weights(net, src=0)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        out["t_selfplay"] += t1 - t0
        out["t_exchange"] += t2 - t1
        out["t_train"] += t3 - t2
        if log and rank == 0:
            log("iter %d: %d games finished, ring %d, %.2fs self-play %.3fs exchange %.2fs train+broadcast"
                % (it, out["games"], len(ring), t1 - t0, t2 - t1, t3 - t2))
    out["world"] = world
    return out
