#!/usr/bin/env python
"""Benchmark of the self-play hot path (BASELINE.json metric: MCTS playouts/s and self-play moves/s, 15x15, 400 playouts).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline (the JSON line's `value`): one step = one full MCTS move search (n_playout lock-steps of select ->
features -> net -> expand/backup) for every one of the G = 4096 concurrent games of a rank, from the synthetic positions
of SURVEY 8(d), on a fresh tree (BASELINE configs[1]).  The same line carries, under their own keys, driver-timed
sub-measurements of the other configurations (each with its own `roofline`):

    selfplay        real self-play plies (BatchedSelfPlay, device-side move sampling, records kept, tree reuse);
                    the line's `moves_per_s` is THIS played figure (all ranks), `moves_per_s_search_only` = value / 400
    loop            configs[4]: the self-play + train loop with the record gather and the weight broadcast (NCCL)
                    INSIDE the timed region, training overlapped with search
    pure            configs[2]: mcts_pure, 8192 games x 1000 playouts                       (N = 1 only)
    resnet10        the 10-block residual net the reference trains, same move-search workload  (N = 1 only)
    inception       configs[3]: the builder-defined Inception-ResNet variant                   (N = 1 only)
    single_game_ms  configs[0]'s GPU side: ONE game through MCTSPlayer.get_action, 8x8 and 15x15 (N = 1 only)

`--legs none` prints the headline alone; `--workload pure|selfplay|loop` and `--net resnet|inception` print one leg as
its own line.  Prints ONE JSON line (see the prompt's contract / DESIGN.md 6).
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
def oracle_board_from_position(cells, meta):
    """(cells, meta) of the C-ABI board format -> oracle board (CPU legs only)"""
    from oracle.board import OBoard
    b = OBoard(W, H, N_IN_ROW)
    b.init_board(0)
    b.states = {int(m): int(cells[m]) for m in np.nonzero(cells)[0]}
    b.availables = [m for m in range(W * H) if m not in b.states]
    b.current_player, b.last_move = int(meta[0]), int(meta[1])
    recent = [int(h) for h in meta[3:7] if h >= 0]  # most recent first; only the last four plies are ordered
    b.history = [(m, p) for m, p in b.states.items() if m not in recent] + [(m, b.states[m]) for m in reversed(recent)]
    return b


def synthetic_position_cpu(index):
    """The SAME position game `index` of the GPU arm starts from (draw_position with RandomState(1234 + index),
    redrawn while the random play already ended the game), as an oracle board."""
    rs = np.random.RandomState(1234 + index)
    while True:
        b = oracle_board_from_position(*draw_position(rs))
        if not b.game_end()[0]:
            return b


def _cpu_worker(args):
    """One process: oracle MCTS + oracle net (PyTorch-CPU fp32, 1 thread), n_moves move searches from the synthetic
    position of game `seed` (0 - 60 stones, the GPU arm's own positions)."""
    seed, n_moves, n_playout, params = args
    import torch
    torch.set_num_threads(1)
    from oracle.mcts import OMCTSPlayer
    from oracle.net import ONet
    net = ONet(W, H, arch=ARCH, params=params)
    b = synthetic_position_cpu(seed)
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
    b = synthetic_position_cpu(0)
    player = OMCTSPlayer(net.policy_value_fn, c_puct=C_PUCT, n_playout=N_PLAYOUT, is_selfplay=1)
    np.random.seed(0)
    t0 = time.perf_counter()
    for _ in range(n_moves):
        b.do_move(int(player.get_action(b, temp=1.0)))
    dt = time.perf_counter() - t0
    return {"value": n_moves * N_PLAYOUT / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle port (Python MCTS + PyTorch-CPU fp32 stand-in for MXNet, batch 1): 1 game (synthetic position 0 "
                      "of the GPU arm), %d moves x %d playouts on 15x15, all %d host threads to the net, %.1f s"
                      % (n_moves, N_PLAYOUT, threads, dt)}


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
            res = pool.map(_cpu_worker, [((step * procs + p) % G_PER_GPU, 1, N_PLAYOUT, params) for p in range(procs)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                per_step.append((sum(r[0] for r in res), dt))
    playouts = sum(p for p, _ in per_step)
    secs = sum(t for _, t in per_step)
    value = playouts / secs
    sample = ("%d processes (= host cores) x 1 game x 1 move x %d playouts per step, each from one of the GPU arm's synthetic "
              "positions (0 - 60 stones), host-saturated, 1 torch thread each" % (procs, N_PLAYOUT))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, len(per_step)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(G_PER_GPU), "games_per_gpu": G_PER_GPU, "n_playout": N_PLAYOUT,
                      "net": NETS[ARCH][0], "l2": "n/a (CPU)", "host_cores": procs},
           "moves_per_s": value / N_PLAYOUT,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Ctx(object):
    """process-wide state of one bench run: rank / world / device and the NCCL group (one process per GPU)"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            import datetime
            # a rank that dies must not leave the others in a collective for NCCL's default 10 minutes
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local),
                                    timeout=datetime.timedelta(seconds=180))
            self.dist = dist
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            self.peak_src = "measured"
        except Exception:
            self.peaks, self.peak_src = {}, "fallback (B200_PROFILING.md)"

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="max"):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def ncu_conv_traffic(G):
    """dram bytes (read + write) per conv launch, averaged over the trunk launches of the committed ncu launch
    list of this workload (profiles/, `ncu --metrics ...,dram__bytes_read.sum,dram__bytes_write.sum` on
    tools/profile_step.py --games 4096).  None when the list is missing or the batch differs."""
    import csv
    for name in ("r2_launches_final.csv", "r1_launches_v7_final.csv"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            break
    else:
        return None, None
    if G != 4096:
        return None, None
    per = {}
    with open(path) as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    h = {k: i for i, k in enumerate(rows[0])}
    for r in rows[1:]:
        if ("k_conv" in r[h["Kernel Name"]] or "k_front_tc" in r[h["Kernel Name"]]) and r[h["Metric Name"]].startswith("dram__bytes"):
            per.setdefault(r[h["ID"]], 0.0)
            per[r[h["ID"]]] += float(r[h["Metric Value"]])
    if not per:
        return None, None
    return sum(per.values()) / len(per), os.path.relpath(path, ROOT)


def az_leg(ctx, args, arch, steps, warmup, G, e2e=True, cpu=False):
    """BASELINE configs[1] shape with net `arch`: `steps` move searches of G games x N_PLAYOUT playouts per rank.
    Returns the JSON line's dict on rank 0 (None elsewhere)."""
    import importlib
    torch = ctx.torch
    from alphapig_b200.params import flop_per_leaf, init_params
    from alphapig_b200 import dist as apdist
    n_blocks = NETS[arch][2]
    PolicyValueNet = importlib.import_module("alphapig_b200." + NETS[arch][0]).PolicyValueNet
    arg, aux = init_params(arch, W, H, n_blocks=n_blocks, seed=0, synthetic_stats=True)
    kw = {} if arch == "simple" else {"n_blocks": n_blocks}
    net = PolicyValueNet(W, H, batch_size=128, model_params=(arg, aux), device=ctx.local, **kw)
    cap = N_PLAYOUT * W * H + 2
    eng = net.search_engine(n_in_row=N_IN_ROW, c_puct=C_PUCT, n_playout=N_PLAYOUT, n_games=G, node_capacity=cap)
    if ctx.world > 1:
        # the collective of THIS leg sits outside its timed region (the search itself has no cross-GPU traffic); the
        # `loop` leg times the exchanges
        apdist.broadcast_weights(net, src=0)
    cells, meta = synthetic_positions(eng, G, seed0=1234 + ctx.rank * G)
    pin_cells = torch.from_numpy(cells).pin_memory().numpy()
    pin_meta = torch.from_numpy(meta).pin_memory().numpy()
    eng.boards_import(pin_cells, pin_meta)
    eng.search_profile(True)

    def resident_step():
        eng.search_advance(-1)
        eng.search_run(N_PLAYOUT)
        return eng.search_timing()[0]

    for _ in range(warmup):
        resident_step()
    eng.search_stats()
    clocks = ClockSampler(ctx.local)
    ctx.sync_all()
    clocks.start()
    l0 = eng.launch_count()
    dev_ms, phase_ms = 0.0, None
    for _ in range(steps):
        dev_ms += resident_step()
        ph = eng.search_profile(True)
        phase_ms = ph if phase_ms is None else phase_ms + ph
    ctx.sync_all()
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    stats = eng.search_stats()

    # end to end through the public API with host buffers: H2D positions, search, D2H visit counts
    e2e_s = h2d = d2h = 0.0
    if e2e:
        ctx.sync_all()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            eng.boards_import(pin_cells, pin_meta)
            eng.search_advance(-1)
            eng.search_run(N_PLAYOUT)
            count, acts, visits, _, rootn = eng.search_root()
            if i >= warmup:
                e2e_s += time.perf_counter() - t0
        assert int(rootn.min()) == N_PLAYOUT and int(visits.sum()) == G * (N_PLAYOUT - 1)
        ctx.sync_all()
        h2d = pin_cells.nbytes + pin_meta.nbytes + 4 * G * 3
        d2h = count.nbytes + acts.nbytes + visits.nbytes + rootn.nbytes
    dev_ms_max, e2e_max = ctx.reduce([dev_ms, e2e_s])
    total_playouts = ctx.world * G * N_PLAYOUT * steps
    value = total_playouts / (dev_ms_max / 1000.0)
    precision = eng.net_precision
    net.close()
    if ctx.rank != 0:
        return None
    peak_tf = float(ctx.peaks.get("bf16_tflops_sustained", 1400.0))
    n_conv = len(phase_ms) - 4
    conv_ms = float(phase_ms[2:2 + n_conv].sum())
    cfin = 256 if arch == "simple" else 128
    head_flop = 2 * (cfin * 6 * W * H + 4 * (W * H) ** 2 + 2 * W * H)
    conv_flop_leaf = flop_per_leaf(arch, W, H, n_blocks=n_blocks) - head_flop  # trunk convs only
    lockstep = steps * N_PLAYOUT
    # conv launches per lock-step: the 6-conv net runs conv1 + conv2 as ONE kernel (front_tc.cu), i.e. 5 launches for 6
    # layers; counted from the engine's own launch counter (select, FC and expand/backup are the other three)
    n_conv_launches = n_conv
    if arch == "simple":
        per_step = int(round(launches / float(lockstep)))
        if 3 < per_step - 3 <= n_conv:
            n_conv_launches = per_step - 3
    conv_launches = lockstep * n_conv_launches
    # terminal leaves never reach the net (compacted out on the device): only evaluated leaves count as work
    evaluated = stats["playouts"] - stats["terminal_leaves"]
    achieved = conv_flop_leaf * evaluated / (conv_ms / 1000.0) / 1e12
    traffic, traffic_src = ncu_conv_traffic(G) if arch == "simple" else (None, None)
    roof = {"bound": "tensor", "kernel": "conv trunk kernels (%d layers in %d launches per lock-step)" % (n_conv, n_conv_launches),
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "peak_source": ctx.peak_src + " (bf16 sustained: the kernels run inside a seconds-long power-capped step)",
            "traffic": traffic, "traffic_source": traffic_src,
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
    hbm = float(ctx.peaks.get("hbm_gbs", 6650.0))
    roof["tree_kernels"] = {"bound": "hbm", "achieved_gbs": tree_bytes / (tree_ms / 1000.0) / 1e9, "peak_gbs": hbm,
                            "frac": tree_bytes / (tree_ms / 1000.0) / 1e9 / hbm,
                            "bytes_per_playout": tree_bytes / max(1, stats["playouts"]),
                            "mean_select_depth": stats["path_nodes"] / max(1, stats["playouts"]) - 1.0,
                            "terminal_leaf_frac": stats["terminal_leaves"] / max(1, stats["playouts"])}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": steps,
           "warmup": warmup, "ms_per_step": dev_ms_max / steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
           "config": {"workload": workload_name(G, arch), "games_per_gpu": G, "n_playout": N_PLAYOUT,
                      "net": NETS[arch][0], "flop_per_leaf": flop_per_leaf(arch, W, H, n_blocks=n_blocks),
                      "arithmetic": ("fp16 operands, fp32 TMEM accumulate (net); fp64 (tree)" if arch != "resnet" else
                                     "hi + lo fp16 activations x error-diffusion-rounded fp16 weights, two products per K "
                                     "step, fp32 TMEM accumulate (net, precision=%s); fp64 (tree)" % precision),
                      "l2": "working set (node pools + activation planes, >10 GB) is larger than L2; no flush needed",
                      "timing": "CUDA events on the engine stream around each ap_search_run, summed over steps, max over ranks",
                      "host_cores": os.cpu_count()},
           "moves_per_s": value / N_PLAYOUT,
           "clocks": clk,
           "gpu_launches": int(launches),
           "roofline": roof}
    if e2e:
        out["e2e"] = {"value": total_playouts / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h),
                      "timing": "wall clock around boards_import + search_advance + search_run + search_root"}
    if cpu:
        out["cpu_baseline"] = cpu_baseline_single((arg, aux), n_moves=2)
    return out


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


def pure_leg(ctx, args, steps, warmup, G, cpu=False):
    torch = ctx.torch
    from alphapig_b200.engine import Engine
    eng = Engine(width=W, height=H, n_in_row=N_IN_ROW, n_games=G, c_puct=5, n_playout=PURE_PLAYOUT,
                 node_capacity=PURE_PLAYOUT * W * H + 2, device=ctx.local)
    cells, meta = synthetic_positions(eng, G, seed0=1234 + ctx.rank * G)
    pin_cells = torch.from_numpy(cells).pin_memory().numpy()
    pin_meta = torch.from_numpy(meta).pin_memory().numpy()
    for i in range(warmup):
        eng.pure_run(PURE_PLAYOUT, seed=i, rollout_mode=args.rollout_mode)
    eng.search_stats()
    clocks = ClockSampler(ctx.local)
    ctx.sync_all()
    clocks.start()
    l0 = eng.launch_count()
    dev_ms = 0.0
    for i in range(steps):
        eng.pure_run(PURE_PLAYOUT, seed=100 + i, rollout_mode=args.rollout_mode)
        dev_ms += eng.search_timing()[0]
    ctx.sync_all()
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    stats = eng.search_stats()
    e2e_s = 0.0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        eng.boards_import(pin_cells, pin_meta)
        mv = eng.pure_run(PURE_PLAYOUT, seed=200 + i, rollout_mode=args.rollout_mode)
        if i >= warmup:
            e2e_s += time.perf_counter() - t0
    ctx.sync_all()
    dev_max, e2e_max = ctx.reduce([dev_ms, e2e_s])
    total = ctx.world * G * PURE_PLAYOUT * steps
    value = total / (dev_max / 1000.0)
    eng.close()
    if ctx.rank != 0:
        return None
    hbm = float(ctx.peaks.get("hbm_gbs", 6650.0))
    tree_bytes = 20 * stats["children_scanned"] + 20 * stats["children_written"] + 24 * stats["path_nodes"] + \
        64 * stats["playouts"]
    ach = tree_bytes / (dev_ms / 1000.0) / 1e9
    traffic, issue_pct, traffic_src = ncu_pure_launch(G)
    out = {"metric": "mcts_pure_playouts_per_s", "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": steps,
           "warmup": warmup, "ms_per_step": dev_max / steps, "higher_is_better": True, "scaling": "weak",
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
           "e2e": {"value": total / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(pin_cells.nbytes + pin_meta.nbytes),
                   "d2h_bytes_per_step": int(mv.nbytes), "timing": "wall clock around boards_import + pure_run"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "k_pure_run (one launch = one move search for every game)",
                        "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                        "peak_source": ctx.peak_src + " (hbm copy bandwidth)",
                        "traffic": traffic, "traffic_source": traffic_src,
                        "issue_active_pct_ncu": issue_pct, "avg_launch_ms": dev_ms / steps,
                        "algorithmic_bytes_per_launch": tree_bytes / steps,
                        "bytes_per_playout": tree_bytes / max(1, stats["playouts"]),
                        "note": "issue bound (73 % issue-active under ncu): register-resident rollouts and warp-level select.  "
                                "achieved = SURVEY 8(d)'s algorithmic bytes (what the reference's tree touches: every "
                                "child of every scanned / expanded node) over the launch time; the kernel itself keeps "
                                "children lazy and re-reads hot blocks from L2, so its DRAM traffic is ~2 % of that"}}
    if cpu:
        out["cpu_baseline"] = cpu_pure_baseline()
    return out


def whole_period_roofline(ctx, playouts_per_s, terminal_frac, what):
    """Tensor roofline of a leg that is timed by wall clock over whole ply periods (search + tree kernels + sampling +
    records + exchanges): algorithmic trunk FLOPs of the evaluated leaves per second against the sustained bf16 peak.  The
    conv-kernel-only fraction is the headline's `roofline`; this one also pays for everything else in the period."""
    from alphapig_b200.params import flop_per_leaf
    head_flop = 2 * (256 * 6 * W * H + 4 * (W * H) ** 2 + 2 * W * H)
    conv_flop_leaf = flop_per_leaf("simple", W, H) - head_flop
    peak_tf = float(ctx.peaks.get("bf16_tflops_sustained", 1400.0)) * ctx.world
    ach = playouts_per_s * (1.0 - terminal_frac) * conv_flop_leaf / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            "peak_source": ctx.peak_src + " (bf16 sustained x n_gpus)", "traffic": None,
            "terminal_leaf_frac": terminal_frac, "what": what}


# ------------------------------------------------------------------------------------------
# real self-play plies (BatchedSelfPlay.step: search + move sampling, recording, re-rooting); `--workload selfplay`
# ------------------------------------------------------------------------------------------
def selfplay_leg(ctx, args, steps, warmup, G):
    torch = ctx.torch
    from alphapig_b200.params import init_params
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.selfplay import BatchedSelfPlay, PipelinedSelfPlay
    arg, aux = init_params(ARCH, W, H, seed=0, synthetic_stats=True)
    net = PolicyValueNet(W, H, batch_size=128, model_params=(arg, aux), device=ctx.local)
    kw = dict(n_playout=N_PLAYOUT, c_puct=C_PUCT, temp=1.0, n_in_row=N_IN_ROW, seed=ctx.rank)
    groups = 1 if args.device_pick else args.groups
    if args.device_pick:
        kw["device_pick"] = True
        kw["device_records"] = bool(args.device_records)
    sp = PipelinedSelfPlay(net, G, n_groups=groups, **kw) if groups > 1 else BatchedSelfPlay(net, G, **kw)
    parts = sp.groups if groups > 1 else [sp]
    cm, k = [], 0
    for part in parts:
        cm.append(synthetic_positions(part.eng, part.G, seed0=1234 + ctx.rank * G + k))
        k += part.G
    sp.load_positions(np.concatenate([c for c, _ in cm]), np.concatenate([m for _, m in cm]))
    calls = groups if groups > 1 else 1   # one pipelined step() advances one group
    for _ in range(warmup * calls):
        sp.step()
    ctx.sync_all()
    # The pipelined forms keep the NEXT ply's search in flight when step() returns.  One more untimed step after the
    # barrier puts every rank right behind such a launch; the timed region then is exactly `steps` periods of
    # [join search t, pick / record / re-root t, launch search t + 1] and ends right behind a launch again.
    for _ in range(calls):
        sp.step()
    for part in parts:
        if not args.device_pick:
            part.eng.search_stats()
    clocks = ClockSampler(ctx.local)
    clocks.start()
    host0 = sp.host_seconds
    t0 = time.perf_counter()
    games = moves = 0
    for _ in range(steps * calls):
        games += len(sp.step())
        moves += sp.last_moves if groups > 1 else G
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    sp.drain()
    dt_max, = ctx.reduce([dt])
    moves_all, games_all = ctx.reduce([moves, games], op="sum")
    # share of terminal leaves (they never reach the net) over all searches since the last counter reset
    term_frac = 0.0
    try:
        st = [part.eng.search_stats() for part in parts]
        term_frac = sum(x["terminal_leaves"] for x in st) / float(max(1, sum(x["playouts"] for x in st)))
    except Exception:
        pass
    host_ms = 1000 * (sp.host_seconds - host0) / steps
    forced = sum(getattr(part, "forced_openings", 0) for part in parts)
    cap = [part.eng.node_capacity() for part in parts]
    net.close()
    if ctx.rank != 0:
        return None
    return {"metric": "selfplay_moves_per_s", "value": moves_all / dt_max, "unit": "moves/s", "n_gpus": ctx.world,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000 * dt_max / steps,
            "higher_is_better": True, "data": "synthetic start positions, then real self-play with tree reuse",
            "playouts_per_s": moves_all * N_PLAYOUT / dt_max,
            "host_ms_per_step": host_ms, "groups": groups,
            "move_sampling": ("device (ap_selfplay_pick), next search overlapped with the host bookkeeping"
                              if args.device_pick else "host (numpy)"),
            "records": ("device trajectories + outbox (ap_traj_*)" if args.device_pick and args.device_records
                        else "host (features + pi copied per ply)"),
            "games_finished": int(games_all), "forced_openings": int(forced),
            "node_capacity": cap, "clocks": clk,
            "roofline": whole_period_roofline(ctx, moves_all * N_PLAYOUT / dt_max, term_frac,
                                              "trunk FLOPs of the evaluated leaves over the wall clock of whole self-play plies"),
            "timing": "wall clock over `steps` whole periods of every game's ply (search + sampling + re-root + records; the "
                      "region starts and ends right behind the launch of the next ply's search), max over ranks",
            "config": {"workload": "BatchedSelfPlay.step: %d games per GPU, n_playout=%d, temp=1.0, Dirichlet noise, records kept, "
                                   "tree reuse" % (G, N_PLAYOUT)}}


# ------------------------------------------------------------------------------------------
# configs[4]: self-play + train loop, collectives inside the timed region; `--workload loop`
# ------------------------------------------------------------------------------------------
def loop_leg(ctx, args, iters, warmup_iters, G, plies=2):
    from alphapig_b200.loop import selfplay_train_loop
    from alphapig_b200.params import init_params
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    arg, aux = init_params(ARCH, W, H, seed=0, synthetic_stats=True)
    net = PolicyValueNet(W, H, batch_size=128, model_params=(arg, aux), device=ctx.local)
    # every slot starts from a synthetic mid-game position (games then end within the few timed plies and their records
    # travel); the trainer's ring starts with an SGF-style bootstrap (train_mxnet.py:270-273: the reference trains on
    # replayed human games before self-play), so policy_update runs in every timed iteration
    probe = net.search_engine(n_in_row=N_IN_ROW, c_puct=C_PUCT, n_playout=1, n_games=G, node_capacity=8, tag="positions")
    cells, meta = synthetic_positions(probe, G, seed0=1234 + ctx.rank * G)
    net.release_engine(probe)

    def prefill(ring):
        rs = np.random.RandomState(99)
        seqs = [[int(m) for m in rs.permutation(W * H)[:40 + (g % 30)]] for g in range(64)]
        ring.eng.replay_push_sgf(seqs, [1 + (g % 2) for g in range(64)])

    clocks = ClockSampler(ctx.local)
    clocks.start()
    res = selfplay_train_loop(net, G, iters, plies_per_iter=plies, n_playout=N_PLAYOUT, c_puct=C_PUCT, temp=1.0,
                              batch_size=128, epochs=8, seed=7, warmup_iters=warmup_iters, overlap=not args.no_overlap,
                              start_positions=(cells, meta), prefill=prefill, trainer_share=args.trainer_share)
    clk = clocks.stop()
    if os.environ.get("AP_LOOP_DETAIL") == "1":  # development: every rank's per-iteration timings
        sys.stderr.write("rank %d [step s ..., gather, flag, bcast+submit] %s t_total %.3f\n" % (ctx.rank, res.get("detail"), res["t_total"]))
    t_max, coll_max = ctx.reduce([res["t_total"], res.get("t_collectives", 0.0)])
    playouts, plies_all, games, recs = ctx.reduce([res["playouts"], res["plies"], res["games"], res["records"]], op="sum")
    net.close()
    if ctx.rank != 0:
        return None
    return {"metric": "loop_playouts_per_s", "value": playouts / t_max, "unit": UNIT, "n_gpus": ctx.world,
            "moves_per_s": plies_all / t_max, "iters": iters, "warmup_iters": warmup_iters, "plies_per_iter": plies,
            "seconds": t_max, "higher_is_better": True,
            "games_finished": int(games), "records_to_trainer": int(recs),
            "train_steps": res.get("train_steps"), "trainer_seconds": res.get("t_trainer"),
            "last_losses": (res.get("losses") or [])[-3:], "last_kls": (res.get("kls") or [])[-3:],
            "lr_multiplier": res.get("lr_multiplier"), "weight_swaps": res.get("weight_swaps"),
            "collectives": {"in_timed_region": True,
                            "what": "per iteration: counts all-gather + record gather to the trainer rank (device memory), a "
                                    "4-byte ready flag and - when a policy_update has finished - one broadcast of the flat fp32 "
                                    "weights, NCCL" if ctx.world > 1 else "none (1 GPU)",
                            "seconds_main_thread_max": coll_max, "weight_broadcasts": res.get("broadcasts"),
                            "bytes_gathered": res.get("bytes_gathered"), "bytes_broadcast": res.get("bytes_broadcast")},
            "overlap": bool(res.get("overlap")), "trainer_share": args.trainer_share, "clocks": clk,
            "timing": "wall clock from the ply boundary after the warm-up iterations to the end of the last iteration's "
                      "last search (barrier + synchronize both sides), max over ranks",
            "config": {"workload": "self-play + train loop (configs[4]): %d games per GPU, n_playout=%d, simple net, %d plies per "
                                   "iteration, policy_update (batch 128, <= 8 epochs, KL rule) on rank 0 overlapped with search"
                                   % (G, N_PLAYOUT, plies)}}


# ------------------------------------------------------------------------------------------
# configs[0]'s GPU side: ONE game through the reference-named shims (what human_play / ChessClient pay per move)
# ------------------------------------------------------------------------------------------
def single_game_leg(ctx, n_moves=6):
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_alphaZero import MCTSPlayer
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    out = {"n_playout": N_PLAYOUT, "what": "MCTSPlayer.get_action(board, temp=1.0) wall clock per move, one game, "
                                            "is_selfplay=1 (tree reuse); first two moves are warm-up"}
    for Wb in (8, 15):
        net = PolicyValueNet(Wb, Wb, batch_size=128, seed=0)
        player = MCTSPlayer(net.policy_value_fn, c_puct=C_PUCT, n_playout=N_PLAYOUT, is_selfplay=1)
        b = Board(width=Wb, height=Wb, n_in_row=N_IN_ROW)
        b.init_board(0)
        np.random.seed(0)
        ts = []
        for _ in range(n_moves + 2):
            t0 = time.perf_counter()
            mv = player.get_action(b, temp=1.0)
            ts.append(time.perf_counter() - t0)
            b.do_move(mv)
        ts = np.array(ts[2:])
        out["%dx%d" % (Wb, Wb)] = {"ms_per_move": 1e3 * float(ts.mean()), "min_ms": 1e3 * float(ts.min()),
                                  "playouts_per_s": N_PLAYOUT / float(ts.mean())}
        # opt-in multi-leaf mode (virtual loss, K playouts in flight per lock-step): NOT the reference's sequential search
        K = 16
        player = MCTSPlayer(net.policy_value_fn, c_puct=C_PUCT, n_playout=N_PLAYOUT, is_selfplay=1, leaves_per_step=K)
        b.init_board(0)
        ts = []
        for _ in range(n_moves + 2):
            t0 = time.perf_counter()
            mv = player.get_action(b, temp=1.0)
            ts.append(time.perf_counter() - t0)
            b.do_move(mv)
        ts = np.array(ts[2:])
        out["%dx%d" % (Wb, Wb)]["virtual_loss_k%d_ms_per_move" % K] = 1e3 * float(ts.mean())
        del player
        net.close()
    return out


def run_gpu(args):
    ctx = Ctx()
    G = args.games
    if args.workload == "pure":
        out = pure_leg(ctx, args, args.steps, args.warmup, G if G != G_PER_GPU else PURE_GAMES,
                       cpu=ctx.world == 1 and not args.no_cpu)
    elif args.workload == "selfplay":
        out = selfplay_leg(ctx, args, args.steps, args.warmup, G)
    elif args.workload == "loop":
        out = loop_leg(ctx, args, args.steps, 1, G)
    else:
        headline = args.net == "simple"
        out = az_leg(ctx, args, args.net, args.steps, args.warmup, G, e2e=True,
                     cpu=ctx.world == 1 and not args.no_cpu and headline)
        if headline and args.legs != "none":
            legs = {}
            args.device_pick, args.device_records = True, True
            legs["selfplay"] = selfplay_leg(ctx, args, 3, 1, G)
            legs["loop"] = loop_leg(ctx, args, 2, 1, G)
            if ctx.world == 1:
                legs["pure"] = pure_leg(ctx, args, 5, 3, PURE_GAMES, cpu=False)
                legs["resnet10"] = az_leg(ctx, args, "resnet", 2, 1, G, e2e=False)
                legs["inception"] = az_leg(ctx, args, "inception", 3, 2, G, e2e=False)
                legs["single_game_ms"] = single_game_leg(ctx)
            if ctx.rank == 0:
                try:  # rooflines of the wall-clock legs (pure arithmetic on numbers already measured)
                    tf = legs["selfplay"]["roofline"]["terminal_leaf_frac"]
                    legs["loop"]["roofline"] = whole_period_roofline(
                        ctx, legs["loop"]["value"], tf, "trunk FLOPs of the evaluated leaves over the wall clock of the self-play + "
                        "train loop (terminal-leaf share taken from the selfplay leg)")
                    sg = legs.get("single_game_ms")
                    if sg:
                        from alphapig_b200.params import flop_per_leaf
                        peak = float(ctx.peaks.get("bf16_tflops_sustained", 1400.0))
                        for key, wb in (("8x8", 8), ("15x15", 15)):
                            ach = sg[key]["playouts_per_s"] * flop_per_leaf("simple", wb, wb) / 1e12
                            sg[key]["roofline"] = {"bound": "launch latency (one leaf in flight: ~9 dependent kernels per playout)",
                                                   "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None}
                except Exception as err:  # never lose the line over a derived number
                    out["legs_roofline_error"] = repr(err)
                out["moves_per_s_search_only"] = out["moves_per_s"]
                out["moves_per_s"] = legs["selfplay"]["value"]
                out["moves_per_s_note"] = ("played self-play moves per second (the `selfplay` leg: every ply decided by a full "
                                           "400-playout search, sampled, recorded and re-rooted); moves_per_s_search_only = "
                                           "value / n_playout")
                for k, v in legs.items():
                    if k in ("resnet10", "inception") and v is not None:
                        v = {kk: v[kk] for kk in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config",
                                                   "clocks", "gpu_launches", "roofline")}
                    out[k] = v
    if ctx.rank == 0:
        print(json.dumps(out))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--games", type=int, default=G_PER_GPU, help="concurrent games per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--legs", default="all", choices=["all", "none"],
                    help="default az run: all = append the selfplay / loop / pure / resnet10 / inception / single-game "
                         "sub-measurements to the line; none = headline only")
    ap.add_argument("--groups", type=int, default=2, help="selfplay workload: pipelined game groups (1 = none)")
    ap.add_argument("--device-pick", action="store_true",
                    help="selfplay workload: one group, moves sampled on the device, next search overlapped with the host phase")
    ap.add_argument("--device-records", action="store_true",
                    help="selfplay workload with --device-pick: records stay on the device (trajectories + outbox)")
    ap.add_argument("--no-overlap", action="store_true", help="loop workload: the synchronous loop (A/B)")
    ap.add_argument("--trainer-share", type=float, default=0.0,
                    help="loop workload, N > 1: the trainer rank plays this fraction fewer games (its GPU also trains)")
    ap.add_argument("--rollout-mode", type=int, default=0, choices=[0, 2],
                    help="pure workload: 0 = permutation rollouts (default), 2 = ply-by-ply rollouts (A/B)")
    ap.add_argument("--net", default="simple", choices=sorted(NETS),
                    help="az workload: simple = BASELINE configs[1] (default, the headline); resnet = the 10-block net the "
                         "reference trains; inception = configs[3] (builder-defined variant)")
    ap.add_argument("--workload", default="az", choices=["az", "pure", "selfplay", "loop"],
                    help="az = BASELINE configs[1] (default, the headline metric); pure = configs[2] (mcts_pure, 1000 playouts); "
                         "selfplay = real self-play plies; loop = configs[4] (self-play + train, collectives timed)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
