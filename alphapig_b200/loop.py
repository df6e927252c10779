"""Self-play + train loop over all GPUs of one box (BASELINE.json configs[4]).

Every rank (one process per GPU) plays ``n_games`` concurrent self-play games on its own engine - no cross-GPU traffic
during search.  The two exchanges of SURVEY 8(e) happen once per iteration (= ``plies_per_iter`` plies), both as NCCL
collectives on device memory:

* finished-game records: ``BatchedSelfPlay(device_records=True)`` keeps every game's (state bits, pi, z) records in
  HBM and moves finished games to a device outbox (csrc/traj.cu); the outbox tensor is gathered to the trainer rank
  as is (``dist.gather_records_device``: NCCL send/recv of device memory) and pushed into its device replay ring
  (``ap_replay_push_packed``) - no record ever visits the host;
* weights: after the trainer's ``policy_update`` one broadcast of the flat fp32 master buffer, then each rank rebuilds
  its fp16 operand images locally at its next ply boundary.

``overlap=True`` (default): nothing waits for the trainer.  The search of the next ply always runs in the self-play
object's background thread; the collectives are issued from the main thread while it runs; ``policy_update`` runs in
a trainer thread on rank 0 (its own CUDA stream, the batch engine's handle) while all ranks - rank 0 included - keep
searching on the previous weights.  Pipeline per iteration i (boundary = the point inside ``step()`` where no search
is in flight):

    boundary   : drain the outbox (device copy); if a broadcast completed since the last boundary, swap the new
                 weights into the search engine (operand-image rebuild, < 1 ms)
    main thread: gather records(i) to the trainer rank  ->  if the previous policy_update has finished (rank 0 says so
                 in a 4-byte broadcast): broadcast its weights and start the next policy_update on everything gathered
                 since; otherwise the records queue up and nobody waits

so the weights a ply searches with are a few iterations old at most (the trainer shares rank 0's GPU with its search); the reference (train_mxnet.py:265-283: collect one
game -> policy_update -> repeat, everything sequential) has no staleness and no overlap.  ``overlap=False`` is the
synchronous form of the same loop (every engine drained while rank 0 trains), kept for A/B.

The trainer job is the reference's ``TrainPipeline.policy_update`` (train_mxnet.py:194-237): ONE minibatch drawn with
``random.sample`` semantics from the ring, up to ``epochs`` steps on it, early stop at KL > 4 kl_targ, adaptive
``lr_multiplier`` (``train_mxnet.kl_and_lr_rule``).
"""
import time

import numpy as np

from . import dist as apdist
from .replay import ReplayBuffer
from .selfplay import BatchedSelfPlay


class _Trainer(object):
    """``policy_update`` on the trainer rank: owns the replay ring (all ring calls come from its thread)."""

    def __init__(self, net, buffer_size, batch_size, epochs, learn_rate, kl_targ):
        import torch
        self.net, self.batch_size, self.epochs = net, batch_size, epochs
        self.learn_rate, self.kl_targ, self.lr_multiplier = learn_rate, kl_targ, 1.0
        self.ring = ReplayBuffer(net._eng, buffer_size)
        self.dev = torch.device("cuda", net._device)
        # highest priority: the trainer's few small kernels slip in whenever an SM frees up between two of the search's
        # persistent kernels (at default priority each of them waited ~1 ms: 2.5 s per policy_update instead of 0.03 s alone)
        self.stream = torch.cuda.Stream(self.dev, priority=-1)
        self.steps = self.records = self.early_stops = 0
        self.losses, self.kls = [], []
        self.seconds = 0.0

    def warm(self):
        """One throw-away train step on copies of the parameters (same shapes, same kernels): cuDNN / cuBLAS handles,
        algorithm selection and lazy module loading cost seconds the first time and would otherwise land in the first
        policy_update of the loop."""
        import torch
        from collections import OrderedDict
        from . import train as T
        net = self.net
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            _, views = net._views()
            arg = OrderedDict((k, views[k].clone()) for k in net._arg_names)
            aux = OrderedDict((k, views[k].clone()) for k in net._aux_names)
            S = net.board_width * net.board_height
            x = torch.zeros((self.batch_size, net.channelnum, net.board_height, net.board_width), device=self.dev)
            pi = torch.full((self.batch_size, S), 1.0 / S, device=self.dev)
            z = torch.zeros((self.batch_size, 1), device=self.dev)
            for _ in range(2):
                T.train_step(arg, aux, T.AdamState(), x, pi, z, 1e-3, net.arch,
                             net._n_blocks if net.arch != "simple" else 0, wd=net.l2_const)
            net.policy_value(np.zeros((self.batch_size, net.channelnum, net.board_height, net.board_width), np.float32))
            self.stream.synchronize()

    def job(self, gathered):
        """gathered: list of uint8 CUDA tensors of packed records (one per rank)"""
        import torch
        from .train_mxnet import kl_and_lr_rule
        t0 = time.perf_counter()
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            for recs in gathered:
                if recs.shape[0]:
                    self.ring.extend_packed(recs.contiguous())
                    self.records += recs.shape[0]
            if len(self.ring) > self.batch_size:
                st, pi, z = self.ring.sample_torch(self.batch_size, self.dev)
                st_host = st.cpu().numpy()
                old_probs, _ = self.net.policy_value(st_host)
                lr = self.learn_rate * self.lr_multiplier
                for i in range(self.epochs):
                    loss, _ = self.net.train_step(st, pi, z, lr, sync=False)
                    self.steps += 1
                    new_probs, _ = self.net.policy_value(st_host)
                    kl = np.mean(np.sum(old_probs * (np.log(old_probs + 1e-10) - np.log(new_probs + 1e-10)), axis=1))
                    if kl > self.kl_targ * 4:
                        self.early_stops += 1
                        break
                kl, self.lr_multiplier = kl_and_lr_rule(old_probs, new_probs, self.kl_targ, self.lr_multiplier)
                self.losses.append(float(loss[0]))
                self.kls.append(float(kl))
            self.stream.synchronize()
        self.seconds += time.perf_counter() - t0


def selfplay_train_loop(net, n_games, n_iters, plies_per_iter=4, n_playout=400, c_puct=5, temp=1.0, batch_size=128,
                        epochs=8, learn_rate=4e-4, buffer_size=2198800, n_in_row=5, seed=0, log=None, device_pick=True,
                        overlap=True, kl_targ=0.02, warmup_iters=0, start_positions=None, prefill=None, trainer_share=0.0):
    """Returns a dict of counters / timings (per rank; wall clock).  ``warmup_iters`` iterations run first and are left
    out of every counter (the timed region then starts at a ply boundary with the pipeline full).
    start_positions: (cells, meta) every slot's first game starts from (default: empty boards).
    trainer_share (multi-GPU, overlap): the trainer rank plays this fraction FEWER games than the others, so that its
    plies - slowed down by the policy_update kernels sharing its GPU - take as long as everybody else's and the job is
    not paced by rank 0 (the job's value is the sum of what all ranks played over the slowest rank's time).
    prefill: callable(ring) run once on the trainer rank before the first iteration - the reference fills its buffer
    from SGF records for the first 4000 batches before any self-play game is trained on (train_mxnet.py:270-273)."""
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    world = dist.get_world_size() if multi else 1
    if not overlap:
        return _synchronous_loop(net, n_games, n_iters, plies_per_iter, n_playout, c_puct, temp, batch_size, epochs,
                                 learn_rate, buffer_size, n_in_row, seed, log, device_pick, multi, rank, world)
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.device("cuda", net._device)
    S = net.board_width * net.board_height
    if multi and rank == 0 and trainer_share > 0:
        n_games = max(1, int(round(n_games * (1.0 - trainer_share))))
        if start_positions is not None:
            start_positions = (start_positions[0][:n_games], start_positions[1][:n_games])
    sp = BatchedSelfPlay(net, n_games, n_playout=n_playout, c_puct=c_puct, temp=temp, n_in_row=n_in_row,
                         seed=seed + 1000 * rank, device_pick=True, device_records=True)
    if start_positions is not None:
        sp.load_positions(*start_positions)
    trainer = _Trainer(net, buffer_size, batch_size, epochs, learn_rate, kl_targ) if rank == 0 else None
    pool = ThreadPoolExecutor(1) if rank == 0 else None
    if trainer is not None:
        pool.submit(trainer.warm).result()
        if prefill is not None:
            prefill(trainer.ring)
    # host-side rendezvous group (gloo): ranks meet on the CPU before any NCCL kernel is launched, see
    # dist.gather_records_device
    cpu_group = dist.new_group(backend="gloo") if multi and dist.get_backend() == "nccl" else None
    if multi:
        apdist.broadcast_weights(net, src=0)  # identical start
    from .nets import _DevView
    flat, _ = net._views()
    staged = flat.clone()  # the weights every rank agreed on last (the trainer keeps writing `flat` in place)
    wptr, wnum = sp.eng.net_weights()
    search_w = torch.as_tensor(_DevView(wptr, wnum), device=dev)  # the search engine's own fp32 master copy
    state = {"take": False, "outbox": None, "new_weights": False, "swaps": 0}

    def boundary(sp_):
        if state["new_weights"]:
            # no search in flight: staged fp32 weights -> the search engine's replica, operand images rebuilt there
            search_w.copy_(staged)
            torch.cuda.current_stream(dev).synchronize()
            sp_.eng.net_refresh()
            state["new_weights"] = False
            state["swaps"] += 1
        if state["take"]:
            state["outbox"] = sp_.take_outbox()
            state["take"] = False
    sp.boundary_hook = boundary

    out = dict(plies=0, playouts=0, games=0, records=0, bytes_gathered=0, bytes_broadcast=0, broadcasts=0, t_total=0.0,
               t_collectives=0.0, iters=0)
    fut = None
    pending = []
    ready = torch.zeros(1, dtype=torch.int32, device=dev if (multi and cpu_group is None) else "cpu")
    if multi:
        # first use of every collective of the loop, off the clock: NCCL sets up its send/recv and broadcast connections
        # lazily, which costs ~1.5 s on 8 GPUs the first time (the gather to the trainer is grouped point-to-point traffic)
        apdist.gather_records_device(torch.zeros((1, apdist.record_width(S)), dtype=torch.uint8, device=dev), dst=0,
                                     cpu_group=cpu_group)
        dist.broadcast(ready, src=0, group=cpu_group)
        dist.broadcast(flat, src=0)
        torch.cuda.synchronize(dev)
    # The timed region starts and ends right behind the launch of a ply's search (no device synchronisation: that would
    # let the search in flight finish off the clock), so it covers n_iters * plies_per_iter whole periods per rank.
    t_start = time.perf_counter()
    t_end = t_start
    for it in range(warmup_iters + n_iters):
        timed = it >= warmup_iters
        detail = []
        for p in range(plies_per_iter):
            if p == plies_per_iter - 1:
                state["take"] = True
            ts = time.perf_counter()
            done = sp.step()
            detail.append(round(time.perf_counter() - ts, 4))
            if timed:
                out["games"] += len(done)
        # both ends of the clock sit right behind a step() (= right behind the launch of the next ply's search): the region
        # holds n_iters x plies_per_iter whole ply periods and n_iters exchanges
        if it == warmup_iters - 1:
            t_start = time.perf_counter()
        t_end = time.perf_counter()
        recs = state["outbox"]
        state["outbox"] = None
        # ---- exchanges, issued while the next ply's search runs in sp's background thread --------------------------
        t0 = time.perf_counter()
        gathered = apdist.gather_records_device(recs, dst=0, cpu_group=cpu_group) if multi else [recs]
        tg = time.perf_counter()
        if rank == 0:
            pending.extend(g for g in gathered if g.shape[0])
        # The trainer shares its GPU with rank 0's search and is slowed down by it; nobody waits for it.  Rank 0 tells
        # the others whether a finished policy_update is there to be broadcast; if not, the records just queue up.
        ready[0] = 1 if (rank != 0 or fut is None or fut.done()) else 0
        if multi:
            dist.broadcast(ready, src=0, group=cpu_group)
        t1 = time.perf_counter()
        if int(ready.item()):
            if rank == 0 and fut is not None:
                fut.result()
                fut = None
            if multi:
                dist.broadcast(flat, src=0)
            staged.copy_(flat)  # the trainer is idle here: a consistent snapshot, swapped in at the next boundary
            torch.cuda.current_stream(dev).synchronize()  # (not the device: the next ply's search is in flight)
            state["new_weights"] = True
            if timed:
                out["broadcasts"] += 1
                out["bytes_broadcast"] += flat.numel() * 4 if multi else 0
            if rank == 0:
                fut = pool.submit(trainer.job, pending)
                pending = []
        t3 = time.perf_counter()
        if timed:
            out.setdefault("detail", []).append(detail + [round(tg - t0, 4), round(t1 - tg, 4), round(t3 - t1, 4)])
        if timed:
            out["iters"] += 1
            out["plies"] += plies_per_iter * n_games
            out["playouts"] += plies_per_iter * n_games * n_playout
            out["records"] += int(recs.shape[0])
            out["bytes_gathered"] += sum(int(g.shape[0]) for g in gathered) * int(recs.shape[1])
            out["t_collectives"] += t3 - t0
        if log and rank == 0:
            log("iter %d: %d games finished so far, ring %d, exchanges %.1f ms on the main thread"
                % (it, out["games"], trainer.records, 1e3 * (t3 - t0)))
    out["t_total"] = t_end - t_start
    sp.drain()
    torch.cuda.synchronize(dev)
    if multi:
        dist.barrier()
    if cpu_group is not None:
        dist.destroy_process_group(cpu_group)
    if fut is not None:
        fut.result()
    if rank == 0 and pending:
        trainer.job(pending)  # records gathered after the last policy_update started (off the clock)
    if trainer is not None:
        out.update(train_steps=trainer.steps, ring_records=trainer.records, losses=trainer.losses, kls=trainer.kls,
                   lr_multiplier=trainer.lr_multiplier, early_stops=trainer.early_stops, t_trainer=trainer.seconds)
    out.update(world=world, weight_swaps=state["swaps"], forced_openings=sp.forced_openings, overlap=True, n_games=n_games)
    sp.boundary_hook = None
    return out


def _synchronous_loop(net, n_games, n_iters, plies_per_iter, n_playout, c_puct, temp, batch_size, epochs, learn_rate,
                      buffer_size, n_in_row, seed, log, device_pick, multi, rank, world):
    """Everything in sequence per iteration: self-play plies (host-staged records), all-gather, train on rank 0 with every
    engine idle, broadcast.  A throughput harness for A/B against the overlapped loop: it draws a fresh minibatch for each
    of the ``epochs`` steps at a fixed learning rate, NOT the reference's ``policy_update`` schedule."""
    import torch
    S = net.board_width * net.board_height
    sp = BatchedSelfPlay(net, n_games, n_playout=n_playout, c_puct=c_puct, temp=temp, n_in_row=n_in_row,
                         seed=seed + 1000 * rank, device_pick=device_pick)
    ring = ReplayBuffer(net._eng, buffer_size) if rank == 0 else None
    if multi:
        apdist.broadcast_weights(net, src=0)
    out = dict(plies=0, playouts=0, games=0, train_steps=0, records=0, t_selfplay=0.0, t_exchange=0.0, t_train=0.0,
               losses=[], overlap=False)
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
                ring.extend_packed(allrec)
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
    out["t_total"] = out["t_selfplay"] + out["t_exchange"] + out["t_train"]
    return out
