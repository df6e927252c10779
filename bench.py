#!/usr/bin/env python
"""Benchmark of the self-play hot path (BASELINE.json metric: MCTS playouts/s, 15x15, 400 playouts).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one full MCTS move search (n_playout lock-steps of select -> features -> net ->
expand/backup) for every one of the G concurrent games of a rank, from the synthetic positions of
SURVEY 8(d), on a fresh tree.  Prints ONE JSON line (see the prompt's contract / DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 15
N_IN_ROW = 5
N_PLAYOUT = 400
C_PUCT = 5
G_PER_GPU = 4096
ARCH = "simple"
METRIC = "mcts_playouts_per_s"
UNIT = "playouts/s"


NETS = {  # --net: BASELINE configs[1] (default), the residual net train_mxnet.py:79-91 trains, configs[3]
    "simple": ("policy_value_net_mxnet_simple", "policy_value_net_mxnet_simple (6-conv net)", 0),
    "resnet": ("policy_value_net_mxnet", "policy_value_net_mxnet (residual net, 10 blocks x 128, as trained by the reference)", 10),
    "inception": ("policy_value_net_inception", "builder-defined Inception-ResNet variant (3x3 stem + 10 x block35, BASELINE configs[3])", 10),
}


def workload_name(G, net="simple"):
    return ("15x15 five-in-a-row batched MCTS move search, %s, "
            "%d concurrent games per GPU, n_playout=%d, c_puct=%d" % (NETS[net][1], G, N_PLAYOUT, C_PUCT))


# ------------------------------------------------------------------------------------------
# synthetic inputs: k = 2*randint(0,31) uniformly random legal plies per game
# ------------------------------------------------------------------------------------------
def draw_position(rs):
    S = W * H
    k = 2 * rs.randint(0, 31)
    perm = rs.permutation(S)[:k]
    cells = np.zeros(S, np.int8)
    cells[perm[0::2]] = 1
    cells[perm[1::2]] = 2
    hist = [int(m) for m in perm[::-1][:4]] + [-1] * max(0, 4 - k)
    meta = np.array([1, int(perm[-1]) if k else -1, k] + hist[:4] + [0], np.int32)
    return cells, meta


def synthetic_positions(eng, G, seed0=1234):
    """A position whose random play already ended the game is redrawn (checked on the device)."""
    S = W * H
    rss = [np.random.RandomState(seed0 + g) for g in range(G)]
    cells = np.zeros((G, S), np.int8)
    meta = np.zeros((G, 8), np.int32)
    todo = np.arange(G)
    while len(todo):
        for g in todo:
            cells[g], meta[g] = draw_position(rss[g])
        eng.boards_import(cells[todo], meta[todo], todo.astype(np.int32))
        end, _ = eng.boards_status(todo.astype(np.int32))
        todo = todo[end]
    return cells, meta


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: oracle MCTS + oracle net (PyTorch-CPU fp32, 1 thread), n_moves move searches."""
    seed, n_moves, n_playout, params = args
    import torch
    torch.set_num_threads(1)
    from oracle.board import OBoard
    from oracle.mcts import OMCTSPlayer
    from oracle.net import ONet
    net = ONet(W, H, arch=ARCH, params=params)
    rs = np.random.RandomState(seed)
    b = OBoard(W, H, N_IN_ROW)
    b.init_board(0)
    for m in rs.permutation(W * H)[:2 * rs.randint(0, 8)]:
        b.do_move(int(m))
    player = OMCTSPlayer(net.policy_value_fn, c_puct=C_PUCT, n_playout=n_playout, is_selfplay=1)
    np.random.seed(seed)
    t0 = time.perf_counter()
    done = 0
    for _ in range(n_moves):
        if b.game_end()[0]:
            break
        mv = player.get_action(b, temp=1.0)
        b.do_move(int(mv))
        done += 1
    return done * n_playout, time.perf_counter() - t0


def cpu_baseline_single(params, n_moves=2):
    """Faithful single-process reference-style path: Python tree + batch-1 net forward with all
    host threads given to the net (what the reference's MXNet-CPU engine would use)."""
    import torch
    from oracle.board import OBoard
    from oracle.mcts import OMCTSPlayer
    from oracle.net import ONet
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = ONet(W, H, arch=ARCH, params=params)
    b = OBoard(W, H, N_IN_ROW)
    b.init_board(0)
    player = OMCTSPlayer(net.policy_value_fn, c_puct=C_PUCT, n_playout=N_PLAYOUT, is_selfplay=1)
    np.random.seed(0)
    t0 = time.perf_counter()
    for _ in range(n_moves):
        b.do_move(int(player.get_action(b, temp=1.0)))
    dt = time.perf_counter() - t0
    return {"value": n_moves * N_PLAYOUT / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle port (Python MCTS + PyTorch-CPU fp32 stand-in for MXNet, batch 1): 1 game, %d moves x %d "
                      "playouts on 15x15, %.1f s" % (n_moves, N_PLAYOUT, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: MXNet absent, reference Python cannot
    travel to the GPU box) saturating the host: one process per core, one game each."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from alphapig_b200.params import init_params
    arg, aux = init_params(ARCH, W, H, seed=0, synthetic_stats=True)
    params = (arg, aux)
    procs = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    per_step = []
    with ctx.Pool(procs) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 * step + p, 1, N_PLAYOUT, params) for p in range(procs)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                per_step.append((sum(r[0] for r in res), dt))
    playouts = sum(p for p, _ in per_step)
    secs = sum(t for _, t in per_step)
    value = playouts / secs
    sample = ("%d processes x 1 game x 1 move x %d playouts per step (host-saturated, 1 torch thread each)"
              % (procs, N_PLAYOUT))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, len(per_step)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(G_PER_GPU), "l2": "n/a (CPU)"},
           "moves_per_s": value / N_PLAYOUT,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def ncu_conv_traffic(G):
    """dram bytes (read + write) per conv launch, averaged over the trunk launches of the committed ncu launch
    list of this workload (profiles/, `ncu --metrics ...,dram__bytes_read.sum,dram__bytes_write.sum` on
    tools/profile_step.py --games 4096).  None when the list is missing or the batch differs."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_launches_v7_final.csv")
    if G != 4096 or not os.path.exists(path):
        return None, None
    per = {}
    with open(path) as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    h = {k: i for i, k in enumerate(rows[0])}
    for r in rows[1:]:
        if "k_conv3x3_tc" in r[h["Kernel Name"]] and r[h["Metric Name"]].startswith("dram__bytes"):
            per.setdefault(r[h["ID"]], 0.0)
            per[r[h["ID"]]] += float(r[h["Metric Value"]])
    if not per:
        return None, None
    return sum(per.values()) / len(per), os.path.relpath(path, ROOT)


def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import importlib
    from alphapig_b200.params import flop_per_leaf, init_params
    from alphapig_b200 import dist as apdist

    G = args.games
    arch = args.net
    n_blocks = NETS[arch][2]
    PolicyValueNet = importlib.import_module("alphapig_b200." + NETS[arch][0]).PolicyValueNet
    arg, aux = init_params(arch, W, H, n_blocks=n_blocks, seed=0, synthetic_stats=True)
    kw = {} if arch == "simple" else {"n_blocks": n_blocks}
    net = PolicyValueNet(W, H, batch_size=128, model_params=(arg, aux), device=local, **kw)
    cap = N_PLAYOUT * W * H + 2
    eng = net.search_engine(n_in_row=N_IN_ROW, c_puct=C_PUCT, n_playout=N_PLAYOUT, n_games=G, node_capacity=cap)
    if world > 1:
        # the one collective of the path: post-train weight broadcast from the trainer rank (NCCL over NVLink)
        apdist.broadcast_weights(net, src=0)
    cells, meta = synthetic_positions(eng, G, seed0=1234 + rank * G)
    pin_cells = torch.from_numpy(cells).pin_memory().numpy()
    pin_meta = torch.from_numpy(meta).pin_memory().numpy()
    eng.boards_import(pin_cells, pin_meta)
    eng.search_profile(True)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        eng.search_advance(-1)
        eng.search_run(N_PLAYOUT)
        return eng.search_timing()[0]

    for _ in range(args.warmup):
        resident_step()
    eng.search_stats()
    clocks = ClockSampler(local)
    sync_all()
    clocks.start()
    l0 = eng.launch_count()
    dev_ms, phase_ms = 0.0, None
    for _ in range(args.steps):
        dev_ms += resident_step()
        ph = eng.search_profile(True)
        phase_ms = ph if phase_ms is None else phase_ms + ph
    sync_all()
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    stats = eng.search_stats()

    # end to end through the public API with host buffers: H2D positions, search, D2H visit counts
    e2e_s = 0.0
    sync_all()
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        eng.boards_import(pin_cells, pin_meta)
        eng.search_advance(-1)
        eng.search_run(N_PLAYOUT)
        count, acts, visits, _, rootn = eng.search_root()
        if i >= args.warmup:
            e2e_s += time.perf_counter() - t0
    assert int(rootn.min()) == N_PLAYOUT and int(visits.sum()) == G * (N_PLAYOUT - 1)
    sync_all()
    h2d = pin_cells.nbytes + pin_meta.nbytes + 4 * G * 3
    d2h = count.nbytes + acts.nbytes + visits.nbytes + rootn.nbytes

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_max = float(t[0]), float(t[1])
    total_playouts = world * G * N_PLAYOUT * args.steps
    value = total_playouts / (dev_ms_max / 1000.0)
    e2e_value = total_playouts / e2e_max

    if rank == 0:
        peaks = {}
        pk_src = "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            pk_src = "measured (sustained)"
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        n_conv = len(phase_ms) - 4
        conv_ms = float(phase_ms[2:2 + n_conv].sum())
        cfin = 256 if arch == "simple" else 128
        head_flop = 2 * (cfin * 6 * W * H + 4 * (W * H) ** 2 + 2 * W * H)
        conv_flop_leaf = flop_per_leaf(arch, W, H, n_blocks=n_blocks) - head_flop  # trunk convs only
        lockstep = args.steps * N_PLAYOUT
        conv_launches = lockstep * n_conv
        # terminal leaves never reach the net (compacted out on the device): only evaluated leaves count as work
        evaluated = stats["playouts"] - stats["terminal_leaves"]
        achieved = conv_flop_leaf * evaluated / (conv_ms / 1000.0) / 1e12
        traffic, traffic_src = ncu_conv_traffic(G) if arch == "simple" else (None, None)
        roof = {"bound": "tensor", "kernel": "k_conv3x3_tc (%d launches per lock-step, all trunk layers)" % n_conv,
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": pk_src, "traffic": traffic, "traffic_source": traffic_src,
                "avg_launch_ms": conv_ms / conv_launches,
                "algorithmic_flop_per_launch_avg": conv_flop_leaf * evaluated / conv_launches,
                "evaluated_leaves": int(evaluated), "terminal_leaves_skipped": int(stats["terminal_leaves"]),
                "phase_ms_per_lockstep": {"select": float(phase_ms[0]) / lockstep, "features": float(phase_ms[1]) / lockstep,
                                          "trunk_convs": [float(x) / lockstep for x in phase_ms[2:2 + n_conv]],
                                          "heads": float(phase_ms[2 + n_conv]) / lockstep,
                                          "expand_backup": float(phase_ms[3 + n_conv]) / lockstep}}
        # tree kernels against the HBM roofline (they are latency bound; reported for honesty)
        tree_bytes = 20 * stats["children_scanned"] + 20 * stats["children_written"] + 24 * stats["path_nodes"] + \
            128 * stats["playouts"]
        tree_ms = float(phase_ms[0] + phase_ms[3 + n_conv])
        roof["tree_kernels"] = {"bound": "hbm", "achieved_gbs": tree_bytes / (tree_ms / 1000.0) / 1e9,
                                "peak_gbs": float(peaks.get("hbm_gbs", 6650.0)),
                                "frac": tree_bytes / (tree_ms / 1000.0) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
                                "bytes_per_playout": tree_bytes / max(1, stats["playouts"]),
                                "terminal_leaf_frac": stats["terminal_leaves"] / max(1, stats["playouts"])}
        cpu = None
        if world == 1 and not args.no_cpu and arch == "simple":
            cpu = cpu_baseline_single((arg, aux), n_moves=2)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
               "config": {"workload": workload_name(G, arch), "games_per_gpu": G, "n_playout": N_PLAYOUT,
                          "net": NETS[arch][0], "flop_per_leaf": flop_per_leaf(arch, W, H, n_blocks=n_blocks),
                          "arithmetic": ("fp16 operands, fp32 TMEM accumulate (net); fp64 (tree)" if arch != "resnet" else
                                         "hi + lo fp16 activations x error-diffusion-rounded fp16 weights, two products per K "
                                         "step, fp32 TMEM accumulate (net, precision=%s); fp64 (tree)" % eng.net_precision),
                          "l2": "working set (node pools + activation planes, >10 GB) is larger than L2; no flush needed",
                          "timing": "CUDA events on the engine stream around each ap_search_run, summed over steps, max over ranks"},
               "moves_per_s": value / N_PLAYOUT,
               "clocks": clk,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "timing": "wall clock around boards_import + search_advance + search_run + search_root"},
               "gpu_launches": int(launches),
               "roofline": roof}
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# C3: mcts_pure batched random-rollout player (BASELINE.json configs[2]); `--workload pure`
# ------------------------------------------------------------------------------------------
PURE_PLAYOUT = 1000
PURE_GAMES = 8192


def cpu_pure_baseline(n_playout=60):
    """oracle port of mcts_pure.MCTSPlayer.get_action on one synthetic 15x15 position, 1 core"""
    from oracle.board import OBoard
    from oracle.mcts import OPureMCTSPlayer
    rs = np.random.RandomState(1234)
    b = OBoard(W, H, N_IN_ROW)
    b.init_board(0)
    for m in rs.permutation(W * H)[:2 * rs.randint(0, 8)]:
        b.do_move(int(m))
    np.random.seed(0)
    t0 = time.perf_counter()
    OPureMCTSPlayer(c_puct=5, n_playout=n_playout).get_action(b)
    dt = time.perf_counter() - t0
    return {"value": n_playout / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle port of mcts_pure (Python tree + rollouts): 1 game, 1 move x %d playouts on 15x15, %.1f s"
                      % (n_playout, dt)}


def ncu_pure_launch(G):
    """dram bytes (read + write) and issue-slot utilisation of one k_pure_run launch at the bench size, from the
    committed ncu launch list (profiles/, tools/profile_step.py --pure 0 --games 8192 --playouts 1000)."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_pure_launches_v4.csv")
    if G != PURE_GAMES or not os.path.exists(path):
        return None, None, None
    with open(path) as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    h = {k: i for i, k in enumerate(rows[0])}
    m = {r[h["Metric Name"]]: float(r[h["Metric Value"]].replace(",", "")) for r in rows[1:]}
    return (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0),
            m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), os.path.relpath(path, ROOT))


def run_gpu_pure(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from alphapig_b200.engine import Engine
    G = args.games if args.games != G_PER_GPU else PURE_GAMES
    eng = Engine(width=W, height=H, n_in_row=N_IN_ROW, n_games=G, c_puct=5, n_playout=PURE_PLAYOUT,
                 node_capacity=PURE_PLAYOUT * W * H + 2, device=local)
    cells, meta = synthetic_positions(eng, G, seed0=1234 + rank * G)
    pin_cells = torch.from_numpy(cells).pin_memory().numpy()
    pin_meta = torch.from_numpy(meta).pin_memory().numpy()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        eng.pure_run(PURE_PLAYOUT, seed=i, rollout_mode=args.rollout_mode)
    eng.search_stats()
    clocks = ClockSampler(local)
    sync_all()
    clocks.start()
    l0 = eng.launch_count()
    dev_ms = 0.0
    for i in range(args.steps):
        eng.pure_run(PURE_PLAYOUT, seed=100 + i, rollout_mode=args.rollout_mode)
        dev_ms += eng.search_timing()[0]
    sync_all()
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    stats = eng.search_stats()
    e2e_s = 0.0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        eng.boards_import(pin_cells, pin_meta)
        mv = eng.pure_run(PURE_PLAYOUT, seed=200 + i, rollout_mode=args.rollout_mode)
        if i >= args.warmup:
            e2e_s += time.perf_counter() - t0
    sync_all()
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total = world * G * PURE_PLAYOUT * args.steps
    value = total / (float(t[0]) / 1000.0)
    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            src = "measured (burst)"
        except Exception:
            peaks, src = {}, "fallback"
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        tree_bytes = 20 * stats["children_scanned"] + 20 * stats["children_written"] + 24 * stats["path_nodes"] + \
            64 * stats["playouts"]
        ach = tree_bytes / (dev_ms / 1000.0) / 1e9
        traffic, issue_pct, traffic_src = ncu_pure_launch(G)
        out = {"metric": "mcts_pure_playouts_per_s", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": float(t[0]) / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64 tree / u32 bitboards", "data": "synthetic",
               "config": {"workload": "15x15 mcts_pure batched random-rollout player, %d playouts/move, %d games per GPU"
                                      % (PURE_PLAYOUT, G),
                          "l2": "node pools (%d games x %d nodes x 32 B) are larger than L2; no flush needed"
                                % (G, PURE_PLAYOUT * W * H + 2),
                          "timing": "CUDA events on the engine stream around the fused k_pure_run launch"},
               "moves_per_s": value / PURE_PLAYOUT,
               "rollout_plies_per_s": stats["rollout_plies"] / (dev_ms / 1000.0),
               "rollout": ("permutation (one sorted random-key permutation of the empty cells + 8-step bit descent to "
                           "the first line; plies = length of the random game it decides)" if args.rollout_mode == 0
                           else "ply by ply"),
               "clocks": clk,
               "e2e": {"value": total / float(t[1]), "unit": UNIT, "h2d_bytes_per_step": int(pin_cells.nbytes + pin_meta.nbytes),
                       "d2h_bytes_per_step": int(mv.nbytes), "timing": "wall clock around boards_import + pure_run"},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": "k_pure_run (one launch = one move search for every game)",
                            "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_source": src,
                            "traffic": traffic, "traffic_source": traffic_src,
                            "issue_active_pct_ncu": issue_pct, "avg_launch_ms": dev_ms / args.steps,
                            "algorithmic_bytes_per_launch": tree_bytes / args.steps,
                            "bytes_per_playout": tree_bytes / max(1, stats["playouts"]),
                            "note": "issue bound (73 % issue-active under ncu): register-resident rollouts and warp-level select.  "
                                    "achieved = SURVEY 8(d)'s algorithmic bytes (what the reference's tree touches: every "
                                    "child of every scanned / expanded node) over the launch time; the kernel itself keeps "
                                    "children lazy and re-reads hot blocks from L2, so its DRAM traffic is ~2 % of that"}}
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_pure_baseline()
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# real self-play plies (BatchedSelfPlay.step: search + host-side sampling, recording, re-rooting); `--workload selfplay`
# ------------------------------------------------------------------------------------------
def run_gpu_selfplay(args):
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from alphapig_b200.params import init_params
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.selfplay import BatchedSelfPlay, PipelinedSelfPlay
    G = args.games
    arg, aux = init_params(ARCH, W, H, seed=0, synthetic_stats=True)
    net = PolicyValueNet(W, H, batch_size=128, model_params=(arg, aux), device=local)
    kw = dict(n_playout=N_PLAYOUT, c_puct=C_PUCT, temp=1.0, n_in_row=N_IN_ROW, seed=0, node_capacity=2 * N_PLAYOUT * W * H + 2)
    if args.device_pick:
        args.groups = 1
        kw["device_pick"] = True
    sp = PipelinedSelfPlay(net, G, n_groups=args.groups, **kw) if args.groups > 1 else BatchedSelfPlay(net, G, **kw)
    parts = sp.groups if args.groups > 1 else [sp]
    cm, k = [], 0
    for part in parts:
        cm.append(synthetic_positions(part.eng, part.G, seed0=1234 + k))
        k += part.G
    cells = np.concatenate([c for c, _ in cm])
    meta = np.concatenate([m for _, m in cm])
    sp.load_positions(cells, meta)
    calls = args.groups if args.groups > 1 else 1   # one pipelined step() advances one group
    for _ in range(args.warmup * calls):
        sp.step()
    torch.cuda.synchronize()
    host0 = sp.host_seconds
    t0 = time.perf_counter()
    games = moves = 0
    for _ in range(args.steps * calls):
        games += len(sp.step())
        moves += sp.last_moves if args.groups > 1 else G
    dt = time.perf_counter() - t0
    if args.groups > 1 or args.device_pick:
        sp.drain()
    print(json.dumps({"metric": "selfplay_moves_per_s", "value": moves / dt, "unit": "moves/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
                      "higher_is_better": True, "data": "synthetic start positions, then real self-play with tree reuse",
                      "playouts_per_s": moves * N_PLAYOUT / dt,
                      "host_ms_per_step": 1000 * (sp.host_seconds - host0) / args.steps, "groups": args.groups,
                      "move_sampling": "device (ap_selfplay_pick), next search overlapped with the host bookkeeping"
                                       if args.device_pick else "host (numpy)",
                      "games_finished": games,
                      "config": {"workload": "BatchedSelfPlay.step: %d games, n_playout=%d, temp=1.0, Dirichlet noise, records kept"
                                             % (G, N_PLAYOUT)}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--games", type=int, default=G_PER_GPU, help="concurrent games per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--groups", type=int, default=2, help="selfplay workload: pipelined game groups (1 = none)")
    ap.add_argument("--device-pick", action="store_true",
                    help="selfplay workload: one group, moves sampled on the device, next search overlapped with the host phase")
    ap.add_argument("--rollout-mode", type=int, default=0, choices=[0, 2],
                    help="pure workload: 0 = permutation rollouts (default), 2 = ply-by-ply rollouts (A/B)")
    ap.add_argument("--net", default="simple", choices=sorted(NETS),
                    help="az workload: simple = BASELINE configs[1] (default, the headline); resnet = the 10-block net the "
                         "reference trains; inception = configs[3] (builder-defined variant)")
    ap.add_argument("--workload", default="az", choices=["az", "pure", "selfplay"],
                    help="az = BASELINE configs[1] (default, the headline metric); pure = configs[2] (mcts_pure, 1000 playouts)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "pure":
        run_gpu_pure(args)
    elif args.workload == "selfplay":
        run_gpu_selfplay(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
