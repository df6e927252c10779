#!/usr/bin/env python
"""Self-play + train loop on N GPUs of one box (BASELINE.json configs[4]):
    python tools/train_loop.py --games 1024 --iters 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_loop.py --games 1024 --iters 3
Prints one JSON line from rank 0 (whole-job moves/s and playouts/s INCLUDING the exchange and the train steps)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=1024)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--plies", type=int, default=4)
ap.add_argument("--playouts", type=int, default=400)
ap.add_argument("--board", type=int, default=15)
ap.add_argument("--arch", default="simple")
ap.add_argument("--blocks", type=int, default=10)
ap.add_argument("--epochs", type=int, default=8)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--device-pick", action="store_true", help="synchronous loop: sample the self-play moves on the device")
ap.add_argument("--no-overlap", action="store_true", help="the synchronous loop (host-staged records, engines drained while training)")
ap.add_argument("--warmup-iters", type=int, default=1)
ap.add_argument("--detail", action="store_true", help="print every rank's per-iteration timings to stderr")
a = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from alphapig_b200.loop import selfplay_train_loop  # noqa: E402
if a.arch == "simple":
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    net = PolicyValueNet(a.board, a.board, batch_size=a.batch, device=local, seed=0)
else:
    from alphapig_b200.policy_value_net_mxnet import PolicyValueNet
    net = PolicyValueNet(a.board, a.board, batch_size=a.batch, n_blocks=a.blocks, device=local, seed=0)
res = selfplay_train_loop(net, a.games, a.iters, plies_per_iter=a.plies, n_playout=a.playouts, batch_size=a.batch,
                          epochs=a.epochs, log=lambda s: print(s, file=sys.stderr), device_pick=a.device_pick,
                          overlap=not a.no_overlap, warmup_iters=a.warmup_iters)
t = torch.tensor([res["t_total"]], dtype=torch.float64, device="cuda")
cnt = torch.tensor([res["plies"], res["playouts"], res["games"]], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
if a.detail:
    print("rank %s detail [step s ..., gather, flag, bcast+submit]: %s host_s %.3f" % (os.environ.get("RANK", "0"), res.get("detail"), 0.0),
          file=sys.stderr)
if int(os.environ.get("RANK", "0")) == 0:
    keep = {k: v for k, v in res.items() if k not in ("losses", "kls", "detail")}
    print(json.dumps({"workload": "self-play + train loop, %s net, %d games/GPU, n_playout %d" % (a.arch, a.games, a.playouts),
                      "n_gpus": world, "moves_per_s": float(cnt[0] / t[0]), "playouts_per_s": float(cnt[1] / t[0]),
                      "games_finished": int(cnt[2]), "seconds_total_max": float(t[0]), "rank0": keep,
                      "last_losses": res.get("losses", [])[-3:]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
