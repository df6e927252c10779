"""Weight interchange.  The reference checkpoints are ``pickle.dump((arg_params, aux_params),
protocol=2)`` of name -> MXNet NDArray dicts (policy_value_net_mxnet.py:301-309, train_mxnet.py:286-293).
The same two name-keyed dicts are kept here, as float32 numpy arrays:

* ``save_npz`` / ``load_npz``  - portable ``.npz`` with ``arg:`` / ``aux:`` key prefixes
* ``load_model``               - a pickle written by ``PolicyValueNet.save_model`` of this package (numpy
  values) or by the reference (MXNet NDArrays: needs ``mxnet`` importable, values go through ``asnumpy``)
* ``python -m alphapig_b200.checkpoint in.model out.npz`` - offline converter for whoever has MXNet.
"""
import pickle
import sys
from collections import OrderedDict

import numpy as np


def _np(v):
    if hasattr(v, "asnumpy"):
        v = v.asnumpy()
    return np.ascontiguousarray(np.asarray(v), dtype=np.float32)


def save_npz(path, model_params):
    arg, aux = model_params
    flat = OrderedDict()
    for k, v in arg.items():
        flat["arg:" + k] = _np(v)
    for k, v in aux.items():
        flat["aux:" + k] = _np(v)
    np.savez(path, **flat)


def load_npz(path):
    arg, aux = OrderedDict(), OrderedDict()
    with np.load(path) as z:
        for k in z.files:
            kind, name = k.split(":", 1)
            (arg if kind == "arg" else aux)[name] = z[k].astype(np.float32)
    return arg, aux


def load_model(path):
    """(arg_params, aux_params) from a reference-style pickle (train_mxnet.py:286-293 tries both encodings)."""
    with open(path, "rb") as f:
        try:
            arg, aux = pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            arg, aux = pickle.load(f, encoding="bytes")
    dec = lambda k: k.decode() if isinstance(k, bytes) else k
    return (OrderedDict((dec(k), _np(v)) for k, v in arg.items()),
            OrderedDict((dec(k), _np(v)) for k, v in aux.items()))


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit("usage: python -m alphapig_b200.checkpoint <in.model (pickle)> <out.npz>")
    save_npz(sys.argv[2], load_model(sys.argv[1]))
    print("wrote", sys.argv[2])
