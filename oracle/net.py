"""Oracle restatement of the reference policy/value nets (PyTorch-CPU, fp32/fp64).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  STRUCTURE PINNED (``symbol_ops``
equals the reference's committed symbol file ``policy_value_loss.json`` node for
node, tests/test_oracle_golden.py); arithmetic **PARITY UNPINNED** at the
MXNet boundary: the arithmetic of the reference lives in MXNet
(``requirements.txt:8`` -> mxnet==1.6.0; graph file written by 1.5.1), which is
not present in ``/root/reference`` nor installable here.  This module restates
the documented MXNet-1.x operator semantics the reference call sites rely on:

* ``Convolution`` : NCHW cross-correlation, weight (O,I,kh,kw), bias on,
  pad = k//2 (``policy_value_net_mxnet_simple.py:39-46``)
* ``BatchNorm``   : eps=1e-3, inference uses moving mean/var,
  ``fix_gamma=True`` by default (gamma treated as 1) -- the ``conv_act`` BNs
  (``..._simple.py:47-53``); ``fix_gamma=False`` for ``bnA*/bnB*``
  (``policy_value_net_mxnet.py:77-81``)
* ``Flatten`` -> (N, C*H*W); ``Dropout`` identity at inference;
  ``FullyConnected`` y = x W^T + b with W (out,in);
  ``SoftmaxActivation`` = softmax over the S logits (probabilities).

Graphs: simple = ``policy_value_net_mxnet_simple.py:68-92``;
residual = ``policy_value_net_mxnet.py:70-102``.
Parameter names/shapes are the reference's (``policy_value_net_mxnet.py:125-138``).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
SIMPLE_TRUNK = (("conv1", 64), ("conv2", 64), ("conv3", 128), ("conv4", 128),
                ("conv5", 256), ("conv_final", 256))


def trunk_spec(arch, n_blocks=10, n_filter=128):
    """[(kind, names..., cin, cout)] for the 3x3 trunk."""
    if arch == "simple":
        spec, cin = [], 9
        for name, cout in SIMPLE_TRUNK:
            spec.append(("conv_act", name, cin, cout))
            cin = cout
        return spec, cin
    if arch == "resnet":
        spec = [("conv_act", "res_conv1", 9, 128)]
        cin = 128
        for i in range(1, n_blocks + 1):
            spec.append(("res_block", i, cin, n_filter))
            cin = n_filter
        return spec, cin
    if arch == "inception":  # builder-defined variant, see INCEPTION note below
        spec = [("conv_act", "incep_conv1", 9, n_filter)]
        for i in range(1, n_blocks + 1):
            spec.append(("block35", i, n_filter, n_filter))
        return spec, n_filter
    raise ValueError(arch)


# INCEPTION: the reference's inception-resnet-v2.py is an unwired ImageNet symbol; the board-sized variant
# restated here is DEFINED BY THIS REPO (alphapig_b200/params.py): 3x3 stem + n_blocks x block35
# (inception-resnet-v2.py:41-58) with ConvFactory = conv + bias, BatchNorm(fix_gamma=True), ReLU, scale 0.17.
BLOCK35 = (("t0", None, 32, 1), ("t1a", None, 32, 1), ("t1b", 32, 32, 3), ("t2a", None, 32, 1), ("t2b", 32, 48, 3),
           ("t2c", 48, 64, 3), ("up", 128, None, 1))


def param_shapes(arch, width, height, n_blocks=10, n_filter=128):
    """OrderedDict name -> shape, arg params then aux params (is_aux flag)."""
    S = width * height
    arg, aux = OrderedDict(), OrderedDict()

    def conv_act(name, cin, cout, k):
        arg[name + "_weight"] = (cout, cin, k, k)
        arg[name + "_bias"] = (cout,)
        arg[name + "_gamma"] = (cout,)
        arg[name + "_beta"] = (cout,)
        aux[name + "_mean"] = (cout,)
        aux[name + "_var"] = (cout,)

    spec, cfin = trunk_spec(arch, n_blocks, n_filter)
    for item in spec:
        if item[0] == "conv_act":
            conv_act(item[1], item[2], item[3], 3)
        elif item[0] == "block35":
            for name, cin, cout, k in BLOCK35:
                conv_act("b35_%d_%s" % (item[1], name), cin or n_filter, cout or n_filter, k)
        else:
            _, i, cin, cout = item
            for tag, ci in (("A", cin), ("B", cout)):
                arg["conv%s%d_weight" % (tag, i)] = (cout, ci, 3, 3)
                arg["conv%s%d_bias" % (tag, i)] = (cout,)
                arg["bn%s%d_gamma" % (tag, i)] = (cout,)
                arg["bn%s%d_beta" % (tag, i)] = (cout,)
                aux["bn%s%d_moving_mean" % (tag, i)] = (cout,)
                aux["bn%s%d_moving_var" % (tag, i)] = (cout,)
    conv_act("conv3_1_1", cfin, 4, 1)
    arg["fc_3_1_1_weight"] = (S, 4 * S)
    arg["fc_3_1_1_bias"] = (S,)
    conv_act("conv3_2_1", cfin, 2, 1)
    arg["fc_3_2_1_weight"] = (1, 2 * S)
    arg["fc_3_2_1_bias"] = (1,)
    return arg, aux


def init_params(arch, width, height, seed=0, n_blocks=10, n_filter=128, synthetic_stats=True):
    """Xavier(uniform, avg, magnitude 3) weights, zero biases, gamma 1, beta 0
    (MXNet ``mx.init.Xavier()`` defaults, ``..._simple.py:148``).  With
    ``synthetic_stats`` the moving stats are mu~N(0,0.1), var~U(0.5,1.5) and
    beta~N(0,0.1), biases~N(0,0.05) so BN folding is actually exercised
    (SURVEY 8(d) synthetic weights)."""
    g = torch.Generator().manual_seed(seed)
    arg_s, aux_s = param_shapes(arch, width, height, n_blocks, n_filter)
    arg, aux = OrderedDict(), OrderedDict()
    for name, shp in arg_s.items():
        if name.endswith("_weight"):
            hw = int(np.prod(shp[2:])) if len(shp) > 2 else 1
            fan_in, fan_out = shp[1] * hw, shp[0] * hw
            scale = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
            t = (torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) * scale
        elif name.endswith("_gamma"):
            t = torch.ones(shp)
            if synthetic_stats and name.startswith("bn"):
                t = 0.5 + torch.rand(shp, generator=g)
        elif name.endswith("_beta"):
            t = torch.randn(shp, generator=g) * 0.1 if synthetic_stats else torch.zeros(shp)
        else:  # bias
            t = torch.randn(shp, generator=g) * 0.05 if synthetic_stats else torch.zeros(shp)
        arg[name] = t.numpy().astype(np.float32)
    for name, shp in aux_s.items():
        if name.endswith("mean"):
            t = torch.randn(shp, generator=g) * 0.1 if synthetic_stats else torch.zeros(shp)
        else:
            t = 0.5 + torch.rand(shp, generator=g) if synthetic_stats else torch.ones(shp)
        aux[name] = t.numpy().astype(np.float32)
    return arg, aux


def _bn(x, gamma, beta, mean, var, fix_gamma):
    inv = 1.0 / torch.sqrt(var + BN_EPS)
    y = (x - mean[None, :, None, None]) * inv[None, :, None, None]
    if not fix_gamma:
        y = y * gamma[None, :, None, None]
    return y + beta[None, :, None, None]


def forward(arg, aux, states, arch, n_blocks=10, n_filter=128, dtype=torch.float32,
            return_logits=False):
    """states: (B, 9, H, W) array-like -> (probs (B,S), values (B,1)) numpy."""
    P = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in list(arg.items()) + list(aux.items())}
    x = torch.as_tensor(np.ascontiguousarray(states)).to(dtype)

    def conv_act(x, name, k, act=True):
        y = F.conv2d(x, P[name + "_weight"], P[name + "_bias"], padding=k // 2)
        y = _bn(y, P[name + "_gamma"], P[name + "_beta"], P[name + "_mean"], P[name + "_var"], True)
        return F.relu(y) if act else y

    spec, _ = trunk_spec(arch, n_blocks, n_filter)
    for item in spec:
        if item[0] == "conv_act":
            x = conv_act(x, item[1], 3)
        elif item[0] == "block35":
            pre = "b35_%d_" % item[1]
            t0 = conv_act(x, pre + "t0", 1)
            t1 = conv_act(conv_act(x, pre + "t1a", 1), pre + "t1b", 3)
            t2 = conv_act(conv_act(conv_act(x, pre + "t2a", 1), pre + "t2b", 3), pre + "t2c", 3)
            up = conv_act(torch.cat([t0, t1, t2], dim=1), pre + "up", 1, act=False)
            x = F.relu(x + 0.17 * up)
        else:
            i = item[1]
            idn = x
            y = F.conv2d(x, P["convA%d_weight" % i], P["convA%d_bias" % i], padding=1)
            y = F.relu(_bn(y, P["bnA%d_gamma" % i], P["bnA%d_beta" % i],
                           P["bnA%d_moving_mean" % i], P["bnA%d_moving_var" % i], False))
            y = F.conv2d(y, P["convB%d_weight" % i], P["convB%d_bias" % i], padding=1)
            y = _bn(y, P["bnB%d_gamma" % i], P["bnB%d_beta" % i],
                    P["bnB%d_moving_mean" % i], P["bnB%d_moving_var" % i], False)
            x = F.relu(y + idn)
    B = x.shape[0]
    p = conv_act(x, "conv3_1_1", 1).reshape(B, -1)
    logits = p @ P["fc_3_1_1_weight"].t() + P["fc_3_1_1_bias"]
    probs = torch.softmax(logits, dim=1)
    v = conv_act(x, "conv3_2_1", 1).reshape(B, -1)
    val = torch.tanh(v @ P["fc_3_2_1_weight"].t() + P["fc_3_2_1_bias"])
    if return_logits:
        return probs.numpy(), val.numpy(), logits.numpy()
    return probs.numpy(), val.numpy()


class ONet(object):
    """Oracle-side ``PolicyValueNet`` inference face (``..._simple.py:178-226``)."""

    def __init__(self, board_width, board_height, arch="simple", params=None, seed=0,
                 n_blocks=10, n_filter=128, dtype=torch.float32):
        self.board_width, self.board_height = board_width, board_height
        self.arch, self.n_blocks, self.n_filter, self.dtype = arch, n_blocks, n_filter, dtype
        self.arg, self.aux = params if params is not None else init_params(
            arch, board_width, board_height, seed, n_blocks, n_filter)

    def policy_value(self, state_batch):
        with torch.no_grad():
            return forward(self.arg, self.aux, np.asarray(state_batch), self.arch,
                           self.n_blocks, self.n_filter, self.dtype)

    def policy_value_fn(self, board):
        legal = board.availables
        st = np.ascontiguousarray(board.current_state()).reshape(
            1, 9, self.board_height, self.board_width)
        probs, values = self.policy_value(st)
        return zip(legal, probs[0][legal]), values[0]


FLOP_PER_LEAF = {  # 2*MAC of convs + FCs
    ("simple", 15): 517682700, ("simple", 8): 147169536, ("resnet", 15): 1332521100,
}


def flop_per_leaf(arch, width, height, n_blocks=10, n_filter=128):
    S = width * height
    spec, cfin = trunk_spec(arch, n_blocks, n_filter)
    mac = 0
    for item in spec:
        if item[0] == "conv_act":
            mac += 9 * item[2] * item[3] * S
        elif item[0] == "block35":
            mac += sum(k * k * (cin or n_filter) * (cout or n_filter) for _, cin, cout, k in BLOCK35) * S
        else:
            mac += 9 * item[2] * item[3] * S + 9 * item[3] * item[3] * S
    mac += cfin * 6 * S + 4 * S * S + 2 * S
    return 2 * mac


def symbol_ops(arch="resnet", n_blocks=10, n_filter=128, width=15, height=15):
    """The op list of the reference's TRAIN symbol (create_policy_value_train, policy_value_net_mxnet.py:173-212 /
    ..._simple.py:121-159) as THIS restatement understands it - derived from the same ``trunk_spec`` / head
    constants ``forward`` and ``alphapig_b200.train`` use - in MXNet's node order and naming:
    [{"op", "name", "attrs", "inputs"}].  ``tests/test_oracle_golden.py`` compares it with the op list of the
    reference's committed symbol file (``policy_value_loss.json`` -> ``tests/golden/res10_symbol_ops.json``): the
    structural pin of this module (layer order, kernel / pad / filters, which BatchNorms train gamma, residual adds,
    heads, Dropout, tanh / SoftmaxActivation, loss and entropy outputs)."""
    S = width * height
    ops = []

    def add(op, name, inputs, **attrs):
        ops.append({"op": op, "name": name, "attrs": {k: str(v) for k, v in attrs.items()}, "inputs": list(inputs)})
        return name

    def conv_act(x, name, nf, k, act="relu"):
        pad = k // 2
        c = add("Convolution", name, [x, name + "_weight", name + "_bias"], kernel="(%d, %d)" % (k, k), num_filter=nf,
                pad="(%d, %d)" % (pad, pad))
        # conv_act BatchNorms keep MXNet's default fix_gamma=True (no attr in the symbol file)
        b = add("BatchNorm", name + "_bn", [c, name + "_gamma", name + "_beta", name + "_mean", name + "_var"])
        return add("Activation", name + "_act", [b], act_type=act)

    spec, _ = trunk_spec(arch, n_blocks, n_filter)
    x = "input_states"
    plus = 0
    for item in spec:
        if item[0] == "conv_act":
            x = conv_act(x, item[1], item[3], 3)
        elif item[0] == "res_block":
            i = item[1]
            idn = x
            for half in ("A", "B"):
                cn, bn = "conv%s%d" % (half, i), "bn%s%d" % (half, i)
                c = add("Convolution", cn, [x, cn + "_weight", cn + "_bias"], kernel="(3, 3)", num_filter=item[3], pad="(1, 1)")
                x = add("BatchNorm", bn, [c, bn + "_gamma", bn + "_beta", bn + "_moving_mean", bn + "_moving_var"],
                        fix_gamma="False")
                if half == "A":
                    x = add("Activation", "actA%d" % i, [x], act_type="relu")
            x = add("elemwise_add", "_plus%d" % plus, [x, idn])
            plus += 1
            x = add("Activation", "actB%d" % i, [x], act_type="relu")
        else:
            raise ValueError("no reference symbol for " + item[0])
    trunk = x
    # value head + value loss (policy_value_net_mxnet.py:93-97,190-193)
    v = conv_act(trunk, "conv3_2_1", 2, 1)
    v = add("Flatten", "flatten1", [v])
    v = add("Dropout", "dropout1", [v], p=0.5)
    v = add("FullyConnected", "fc_3_2_1", [v, "fc_3_2_1_weight", "fc_3_2_1_bias"], num_hidden=1)
    v = add("Activation", "activation0", [v], act_type="tanh")
    d = add("elemwise_sub", "_minus0", ["input_labels", v])
    d = add("square", "square0", [d])
    vloss = add("mean", "mean1", [d])
    # policy head + policy loss (:85-91,194-199)
    p = conv_act(trunk, "conv3_1_1", 4, 1)
    p = add("Flatten", "flatten0", [p])
    p = add("Dropout", "dropout0", [p], p=0.5)
    p = add("FullyConnected", "fc_3_1_1", [p, "fc_3_1_1_weight", "fc_3_1_1_bias"], num_hidden=S)
    p = add("SoftmaxActivation", "Act_SILER", [p])
    lp = add("log", "log0", [p])
    m = add("elemwise_mul", "_mul0", [lp, "mcts_probs"])
    m = add("sum", "sum0", [m], axis=1)
    m = add("_mul_scalar", "_mulscalar0", [m], scalar=-1.0)
    ploss = add("mean", "mean0", [m])
    tot = add("elemwise_add", "_plus%d" % plus, [vloss, ploss])
    add("MakeLoss", "makeloss0", [tot])
    # entropy output (monitoring only, :200-204)
    n1 = add("_mul_scalar", "_mulscalar1", [p], scalar=-1.0)
    l1 = add("log", "log1", [p])
    e = add("elemwise_mul", "_mul1", [n1, l1])
    e = add("sum", "sum1", [e], axis=1)
    e = add("mean", "mean2", [e])
    e = add("BlockGrad", "blockgrad0", [e])
    add("MakeLoss", "makeloss1", [e])
    return ops
