"""``data_buffer`` of the reference ``TrainPipeline`` on the device.

The reference keeps ``deque(maxlen=buffer_size)`` of augmented ``(state, mcts_prob, winner)`` tuples
(train_mxnet.py:57), fills it with ``data_buffer.extend(get_equi_data(play_data))`` (:115-135,153,180)
and draws minibatches with ``random.sample(data_buffer, batch_size)`` (:196).  ``ReplayBuffer`` keeps the
same logical sequence - same length, same eviction, same sample for the same Python ``random`` state -
but stores one packed record per position in HBM and lets the gather kernel (csrc/replay.cu) apply
the rotation / flip when the minibatch is drawn.
"""
import random

import numpy as np


class ReplayBuffer(object):
    def __init__(self, engine, maxlen):
        self.eng = engine
        self.maxlen = int(maxlen)
        engine.replay_create(self.maxlen)

    def __len__(self):
        return self.eng.replay_size()[0]

    def extend_positions(self, states, pis, zs):
        """Append 8 augmented samples per position.  states [n][9][H][W] 0/1 (any dtype) or already
        bit-packed uint8 [n][ceil(9S/8)]; pis [n][S]; zs [n]."""
        states = np.asarray(states)
        if states.dtype != np.uint8 or states.ndim != 2:
            n = states.shape[0]
            states = np.packbits(states.reshape(n, -1).astype(np.uint8), axis=1)
        self.eng.replay_push(states, pis, zs)

    def extend(self, play_data):
        """``data_buffer.extend(get_equi_data(play_data))`` given the UN-augmented play_data
        [(state, mcts_prob, winner_z), ...] of one game."""
        play_data = list(play_data)
        if not play_data:
            return
        self.extend_positions(np.stack([np.asarray(s) for s, _, _ in play_data]),
                              np.stack([np.asarray(p) for _, p, _ in play_data]),
                              np.array([z for _, _, z in play_data], dtype=np.float32))

    def extend_packed(self, records):
        """Append positions given as packed records (``alphapig_b200.dist.pack_records`` / the self-play outbox):
        a uint8 ndarray [n][record width] or a CUDA uint8 tensor of that shape on the ring's GPU (no host copy)."""
        if hasattr(records, "data_ptr"):
            if records.shape[0]:
                assert records.is_cuda and records.is_contiguous() and records.dtype.itemsize == 1
                import torch
                torch.cuda.current_stream(records.device).synchronize()
                self.eng.replay_push_packed(None, n=records.shape[0], device_ptr=records.data_ptr())
        elif len(records):
            self.eng.replay_push_packed(records)

    def sample_indices(self, batch_size):
        """The positions ``random.sample(data_buffer, batch_size)`` would pick (same ``random`` state,
        same population size => same indices)."""
        return random.sample(range(len(self)), batch_size)

    def sample(self, batch_size):
        """-> (state_batch [B][9][H][W], mcts_probs_batch [B][S], winner_batch [B]) numpy float32"""
        return self.eng.replay_gather(self.sample_indices(batch_size))

    def sample_torch(self, batch_size, device):
        """Same minibatch as torch tensors on ``device`` written by the gather kernel (no host copy)."""
        import torch
        idx = self.sample_indices(batch_size)
        e = self.eng
        st = torch.empty((batch_size, 9, e.height, e.width), dtype=torch.float32, device=device)
        pi = torch.empty((batch_size, e.S), dtype=torch.float32, device=device)
        z = torch.empty((batch_size,), dtype=torch.float32, device=device)
        e.replay_gather_device(idx, st.data_ptr(), pi.data_ptr(), z.data_ptr())
        return st, pi, z

    def __getitem__(self, j):
        n = len(self)
        if j < 0:
            j += n
        if not 0 <= j < n:
            raise IndexError(j)
        st, pi, z = self.eng.replay_gather([j])
        return st[0], pi[0], float(z[0])
