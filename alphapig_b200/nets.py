"""Shared implementation of the two ``PolicyValueNet`` shims (policy_value_net_mxnet{,_simple}.py).

Inference runs in the engine's tcgen05 kernels (csrc/conv_tc.cu + net.cu); the training step is
PyTorch on the same GPU (alphapig_b200/train.py).  The fp32 master weights live once on the device as
a flat buffer owned by the engine; the PyTorch parameters are zero-copy views of that buffer, so
``train_step`` updates it in place and ``ap_net_refresh`` rebuilds the fp16 operand images --
no host round trip (the reference copies every parameter train -> predict modules through the host
after each step, policy_value_net_mxnet.py:295-297).
"""
import pickle
from collections import OrderedDict

import numpy as np

from . import params as P
from .engine import Engine


def _to_numpy(x):
    if hasattr(x, "asnumpy"):  # an MXNet NDArray, should anyone still have one
        x = x.asnumpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


class _DevView(object):
    """__cuda_array_interface__ wrapper so torch can alias engine-owned device memory."""

    def __init__(self, ptr, numel, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PolicyValueNetBase(object):
    _is_alphapig_b200_net = True
    arch = "simple"

    def __init__(self, board_width, board_height, batch_size=512, n_blocks=8, n_filter=128, model_params=None,
                 device=0, n_in_row=5, seed=None, precision="auto"):
        self.board_width = board_width
        self.board_height = board_height
        self.batchsize = batch_size
        self.channelnum = 9
        self.l2_const = 1e-4
        self._n_blocks = n_blocks
        self._n_filter = n_filter
        self._device = device
        self._precision = precision  # see Engine.net_load: "auto" | "fp16" | "split" | "split_act"
        if model_params:
            arg, aux = model_params
            arg = OrderedDict((k, _to_numpy(v)) for k, v in arg.items())
            aux = OrderedDict((k, _to_numpy(v)) for k, v in aux.items())
        else:
            arg, aux = P.init_params(self.arch, board_width, board_height, n_blocks, n_filter, seed=seed)
        self._arg_names, self._aux_names = list(arg.keys()), list(aux.keys())
        self._engines = {}
        # the batch engine: forward for policy_value / policy_value_fn, owner of the master weights
        # (its stream has the highest priority: policy_value / the replay ring next to a search on the same GPU)
        self._eng = self._make_engine(1, n_in_row, 5.0, 1, arg, aux, high_priority=True)
        self._torch = None
        self._opt = None

    # -- engines ------------------------------------------------------------
    def _merged_host(self):
        arg, aux = self.get_policy_param()
        d = OrderedDict(arg)
        d.update(aux)
        return d

    def _make_engine(self, n_games, n_in_row, c_puct, n_playout, arg=None, aux=None, node_capacity=0, high_priority=False):
        eng = Engine(width=self.board_width, height=self.board_height, n_in_row=n_in_row, n_games=n_games,
                     c_puct=c_puct, n_playout=n_playout, node_capacity=node_capacity, device=self._device,
                     high_priority=high_priority)
        if arg is None:
            merged = self._merged_host()
        else:
            merged = OrderedDict(arg)
            merged.update(aux)
        eng.net_load(self.arch, merged, n_blocks=self._n_blocks if self.arch != "simple" else 0,
                     n_filter=self._n_filter, precision=self._precision)
        return eng

    def search_engine(self, n_in_row=5, c_puct=5.0, n_playout=400, n_games=1, node_capacity=0, tag=0):
        """An engine holding ``n_games`` trees and a replica of the current weights (kept in sync by
        ``train_step``); what ``MCTS`` uses when this net's ``policy_value_fn`` is the evaluator.
        Engines are cached per configuration; ``tag`` distinguishes several engines of the same shape."""
        key = (n_games, n_in_row, float(c_puct), int(n_playout), int(node_capacity), tag)
        eng = self._engines.get(key)
        if eng is None:
            eng = self._make_engine(n_games, n_in_row, c_puct, n_playout, node_capacity=node_capacity)
            self._engines[key] = eng
        return eng

    def close(self):
        """Free every engine of this net (device memory returns at once instead of at garbage collection)."""
        for eng in list(self._engines.values()):
            eng.close()
        self._engines.clear()
        self._torch = None
        if getattr(self, "_eng", None) is not None:
            self._eng.close()
            self._eng = None

    def release_engine(self, eng):
        """Drop a search engine from the replica cache and free its device memory (an ``MCTS`` going away)."""
        for key, e in list(self._engines.items()):
            if e is eng:
                del self._engines[key]
        eng.close()

    # -- inference (policy_value_net_mxnet_simple.py:178-226) ------------------
    def policy_value(self, state_batch):
        states = np.asarray(state_batch, dtype=np.float32)
        return self._eng.net_forward(states)

    def policy_value_fn(self, board):
        legal_positions = board.availables
        state = np.ascontiguousarray(board.current_state(), dtype=np.float32).reshape(
            1, self.channelnum, self.board_height, self.board_width)
        acts_probs, values = self._eng.net_forward(state)
        return zip(legal_positions, acts_probs[0][legal_positions]), values[0]

    # -- weights --------------------------------------------------------------
    def _views(self):
        """torch views (name -> tensor) of the engine's flat fp32 master buffer."""
        if self._torch is None:
            import torch
            ptr, numel = self._eng.net_weights()
            flat = torch.as_tensor(_DevView(ptr, numel), device="cuda:%d" % self._device)
            shapes = OrderedDict()
            a, x = P.param_shapes(self.arch, self.board_width, self.board_height, self._n_blocks, self._n_filter)
            shapes.update(a)
            shapes.update(x)
            views = OrderedDict()
            for name, off, n in self._eng.net_layout():
                views[name] = flat[off:off + n].view(shapes[name])
            self._torch = (flat, views)
        return self._torch

    def get_policy_param(self):
        """(arg_params, aux_params) as name -> float32 ndarray (reference: MXNet NDArrays)."""
        if getattr(self, "_eng", None) is None or self._eng.net_names is None:
            raise RuntimeError("net not loaded")
        _, views = self._views()
        arg = OrderedDict((k, views[k].detach().cpu().numpy().copy()) for k in self._arg_names)
        aux = OrderedDict((k, views[k].detach().cpu().numpy().copy()) for k in self._aux_names)
        return arg, aux

    def save_model(self, model_file):
        """pickle protocol 2 of (arg_params, aux_params) (policy_value_net_mxnet_simple.py:250-254)"""
        net_params = self.get_policy_param()
        print('>>>>>>>>>> saved into', model_file)
        with open(model_file, 'wb') as f:
            pickle.dump(net_params, f, protocol=2)

    def sync_replicas(self):
        """Push the master weights to every search engine replica (device to device) and rebuild
        all operand images."""
        import torch
        flat, _ = self._views()
        torch.cuda.synchronize(flat.device)
        self._eng.net_refresh()
        for eng in self._engines.values():
            ptr, numel = eng.net_weights()
            torch.as_tensor(_DevView(ptr, numel), device=flat.device).copy_(flat)
            torch.cuda.synchronize(flat.device)
            eng.net_refresh()

    # -- training (policy_value_net_mxnet_simple.py:228-244) -------------------
    def train_step(self, state_batch, mcts_probs, winner_batch, learning_rate, sync=True):
        """One Adam step on (states, pis, zs); returns (loss (1,), entropy (1,)) as the reference does.
        sync=False leaves the search-engine replicas on the OLD weights (a search may be running on them): the caller
        pushes the new ones with ``sync_replicas()`` when no search is in flight (alphapig_b200/loop.py)."""
        import torch
        from . import train as T
        flat, views = self._views()
        dev = flat.device
        if self._opt is None:
            self._opt = T.AdamState()
        S = self.board_width * self.board_height

        def dev_tensor(a, shape):
            # device tensors (e.g. ReplayBuffer.sample_torch) are used in place; anything else goes through numpy
            if isinstance(a, torch.Tensor):
                return a.to(device=dev, dtype=torch.float32).reshape(shape)
            return torch.as_tensor(np.asarray(a, dtype=np.float32).reshape(shape), device=dev)
        x = dev_tensor(state_batch, (-1, self.channelnum, self.board_height, self.board_width))
        pi = dev_tensor(mcts_probs, (-1, S))
        z = dev_tensor(winner_batch, (-1, 1))
        arg = OrderedDict((k, views[k]) for k in self._arg_names)
        aux = OrderedDict((k, views[k]) for k in self._aux_names)
        loss, entropy = T.train_step(arg, aux, self._opt, x, pi, z, float(learning_rate), self.arch,
                                     self._n_blocks if self.arch != "simple" else 0, wd=self.l2_const)
        if sync:
            self.sync_replicas()
        else:
            self._eng.net_refresh()  # the batch engine (policy_value / policy_value_fn) follows at once
        return loss.reshape(1).cpu().numpy(), entropy.reshape(1).cpu().numpy()
