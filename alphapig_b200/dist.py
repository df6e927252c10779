"""Multi-GPU plumbing: one process per GPU, games sharded per rank, ZERO traffic during search.

Only two exchanges exist, both outside the search loop, both ``torch.distributed``
(NCCL over NVLink on GPUs; gloo in the CPU tests):

* ``broadcast_weights``  - after ``train_step`` on the trainer rank, ONE broadcast of the flat fp32
  master-weight buffer (5.4 MB simple / 12.7 MB res-10) straight from/into the engine-owned device
  memory, then a local rebuild of the fp16 operand images.  The reference's analogue is the
  in-process train->predict parameter copy (policy_value_net_mxnet.py:295-297).
* ``gather_replay``      - finished-game records as fixed-size packed rows
  (bit-packed 9xS planes | pi fp32[S] | z fp32) gathered to every rank (or used on the trainer);
  the reference's analogue is ``data_buffer.extend`` (train_mxnet.py:153,180).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_games(n_total, rank, world):
    """Contiguous block of games for this rank: (first, count)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def broadcast_weights(net, src=0):
    flat, _ = net._views()
    dist.broadcast(flat, src=src)
    net.sync_replicas()


def record_width(S):
    """bytes per packed record: ceil(9S/8) state bits padded to 4, S fp32 pi, 1 fp32 z"""
    return ((9 * S + 7) // 8 + 3) // 4 * 4 + 4 * S + 4


def pack_records(states_bits, pis, zs, S):
    """states_bits uint8[n][ceil(9S/8)], pis float[n][S], zs float[n] -> uint8[n][record_width]"""
    n = len(zs)
    w = record_width(S)
    sb = (9 * S + 7) // 8
    out = np.zeros((n, w), np.uint8)
    if n:
        out[:, :sb] = states_bits
        off = w - 4 * S - 4
        out[:, off:off + 4 * S] = np.ascontiguousarray(pis, dtype=np.float32).view(np.uint8).reshape(n, 4 * S)
        out[:, off + 4 * S:] = np.ascontiguousarray(zs, dtype=np.float32).reshape(n, 1).view(np.uint8)
    return out


def unpack_records(packed, S, width, height):
    """-> (states float32[n][9][width][height], pis float32[n][S], zs float32[n])"""
    n = packed.shape[0]
    w = record_width(S)
    sb = (9 * S + 7) // 8
    off = w - 4 * S - 4
    bits = np.unpackbits(np.ascontiguousarray(packed[:, :sb]), axis=1)[:, :9 * S]
    states = bits.reshape(n, 9, width, height).astype(np.float32)
    pis = np.ascontiguousarray(packed[:, off:off + 4 * S]).view(np.float32).reshape(n, S)
    zs = np.ascontiguousarray(packed[:, off + 4 * S:]).view(np.float32).reshape(n)
    return states, pis, zs


def split_records(packed, S):
    """packed rows -> (state bits uint8[n][ceil(9S/8)], pis float32[n][S], zs float32[n]) - the arguments of
    ``ReplayBuffer.extend_positions`` / ``ap_replay_push``"""
    n = packed.shape[0]
    w = record_width(S)
    sb = (9 * S + 7) // 8
    off = w - 4 * S - 4
    pis = np.ascontiguousarray(packed[:, off:off + 4 * S]).view(np.float32).reshape(n, S)
    zs = np.ascontiguousarray(packed[:, off + 4 * S:]).view(np.float32).reshape(n)
    return np.ascontiguousarray(packed[:, :sb]), pis, zs


def gather_replay(packed, device=None):
    """All-gather variable numbers of packed records from every rank.  Returns uint8[N_total][width]."""
    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    n = torch.tensor([packed.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = packed.shape[1]
    mx = max(counts + [1])
    buf = torch.zeros((mx, width), dtype=torch.uint8, device=dev)
    if packed.shape[0]:
        buf[:packed.shape[0]] = torch.from_numpy(packed).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)], axis=0)


def gather_records_device(recs, dst=None, group=None, cpu_group=None):
    """Gather packed records that live on the GPU (a uint8 CUDA tensor [n][width], n differs per rank) straight over
    NCCL: an all-gather of the counts, then the padded record blocks - to rank ``dst`` only (``dist.gather``,
    NCCL send/recv underneath: the trainer is the only consumer) or to every rank (dst=None, all-gather).  Returns
    the list of per-rank tensors trimmed to their counts (empty list on the ranks that receive nothing).

    cpu_group (a gloo group over the same ranks): the counts travel over it as CPU tensors.  That exchange is also the
    rendezvous: a rank that arrives early waits on the HOST while its GPU keeps searching, and the NCCL kernels that
    follow start within microseconds of each other on all GPUs.  Without it an early rank's NCCL kernel sits on its
    SMs spinning for the slowest peer - and the search's persistent conv kernels, which need every SM, stall behind
    it (measured on 8 x B200: 0.2 s per ply)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    if cpu_group is not None:
        n = torch.tensor([recs.shape[0]], dtype=torch.int64)
        clist = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(clist, n, group=cpu_group)
    else:
        n = torch.tensor([recs.shape[0]], dtype=torch.int64, device=recs.device)
        clist = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(clist, n, group=group)
    counts = [int(c.item()) for c in clist]
    mx = max(counts + [1])
    width = recs.shape[1]
    send = recs
    if recs.shape[0] != mx:
        send = torch.zeros((mx, width), dtype=torch.uint8, device=recs.device)
        send[:recs.shape[0]] = recs
    send = send.contiguous()
    if dst is None:
        if nccl:
            out = torch.empty((world * mx, width), dtype=torch.uint8, device=recs.device)
            dist.all_gather_into_tensor(out, send, group=group)
            parts = [out[r * mx:(r + 1) * mx] for r in range(world)]
        else:
            parts = [torch.empty_like(send) for _ in range(world)]
            dist.all_gather(parts, send, group=group)
        return [p[:c] for p, c in zip(parts, counts)]
    parts = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst, group=group)
    return [p[:c] for p, c in zip(parts, counts)] if rank == dst else []
