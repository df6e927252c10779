"""Parameter tables of the reference nets: names, shapes and initialisation
(policy_value_net_mxnet.py:125-138 shape dump; Xavier defaults of ``mx.init.Xavier()``:
uniform, factor_type 'avg', magnitude 3 -- policy_value_net_mxnet_simple.py:148)."""
import math
from collections import OrderedDict

import numpy as np

# builder-defined board-sized Inception-ResNet variant (BASELINE configs[3]): the reference's inception-resnet-v2.py is an
# unwired ImageNet symbol (strides / pooling collapse a 15x15 plane), so the variant keeps what transfers: a 3x3 stem and
# n_blocks x block35 (inception-resnet-v2.py:41-58: towers 1x1(32) | 1x1(32)-3x3(32) | 1x1(32)-3x3(48)-3x3(64), concat 128,
# 1x1 up-projection with BN and no activation, net += 0.17 * up, ReLU), every conv = ConvFactory (conv + bias, BatchNorm with
# the default fix_gamma=True, ReLU), then the reference's policy / value heads.
INCEPTION_SCALE = 0.17
BLOCK35 = (("t0", None, 32, 1), ("t1a", None, 32, 1), ("t1b", 32, 32, 3), ("t2a", None, 32, 1), ("t2b", 32, 48, 3),
           ("t2c", 48, 64, 3), ("up", 128, None, 1))  # (name, cin (None = block input), cout (None = block input), k)

SIMPLE_TRUNK = (("conv1", 9, 64), ("conv2", 64, 64), ("conv3", 64, 128), ("conv4", 128, 128),
                ("conv5", 128, 256), ("conv_final", 256, 256))


def param_shapes(arch, width, height, n_blocks=8, n_filter=128):
    """-> (arg_shapes, aux_shapes) OrderedDicts with the reference's names."""
    S = width * height
    arg, aux = OrderedDict(), OrderedDict()

    def conv_act(name, cin, cout, k):
        arg[name + "_weight"] = (cout, cin, k, k)
        for suffix in ("_bias", "_gamma", "_beta"):
            arg[name + suffix] = (cout,)
        aux[name + "_mean"] = (cout,)
        aux[name + "_var"] = (cout,)

    if arch == "simple":
        for name, cin, cout in SIMPLE_TRUNK:
            conv_act(name, cin, cout, 3)
        cfin = 256
    elif arch == "resnet":
        conv_act("res_conv1", 9, 128, 3)
        cin = 128
        for i in range(1, n_blocks + 1):
            for tag, ci in (("A", cin), ("B", n_filter)):
                arg["conv%s%d_weight" % (tag, i)] = (n_filter, ci, 3, 3)
                arg["conv%s%d_bias" % (tag, i)] = (n_filter,)
                arg["bn%s%d_gamma" % (tag, i)] = (n_filter,)
                arg["bn%s%d_beta" % (tag, i)] = (n_filter,)
                aux["bn%s%d_moving_mean" % (tag, i)] = (n_filter,)
                aux["bn%s%d_moving_var" % (tag, i)] = (n_filter,)
            cin = n_filter
        cfin = n_filter
    elif arch == "inception":
        conv_act("incep_conv1", 9, n_filter, 3)
        for i in range(1, n_blocks + 1):
            for name, cin, cout, k in BLOCK35:
                conv_act("b35_%d_%s" % (i, name), cin or n_filter, cout or n_filter, k)
        cfin = n_filter
    else:
        raise ValueError("arch must be 'simple', 'resnet' or 'inception'")
    conv_act("conv3_1_1", cfin, 4, 1)
    arg["fc_3_1_1_weight"] = (S, 4 * S)
    arg["fc_3_1_1_bias"] = (S,)
    conv_act("conv3_2_1", cfin, 2, 1)
    arg["fc_3_2_1_weight"] = (1, 2 * S)
    arg["fc_3_2_1_bias"] = (1,)
    return arg, aux


def init_params(arch, width, height, n_blocks=8, n_filter=128, seed=None, synthetic_stats=False):
    """Xavier weights, zero biases, gamma 1, beta 0, moving mean 0 / var 1.  ``synthetic_stats``
    draws non-trivial BN statistics / affine terms instead (benchmark weights)."""
    rs = np.random.RandomState(seed)
    arg_s, aux_s = param_shapes(arch, width, height, n_blocks, n_filter)
    arg, aux = OrderedDict(), OrderedDict()
    for name, shp in arg_s.items():
        if name.endswith("_weight"):
            hw = int(np.prod(shp[2:])) if len(shp) > 2 else 1
            scale = math.sqrt(3.0 / ((shp[1] * hw + shp[0] * hw) / 2.0))
            a = rs.uniform(-scale, scale, size=shp)
        elif name.endswith("_gamma"):
            a = rs.uniform(0.5, 1.5, size=shp) if (synthetic_stats and name.startswith("bn")) else np.ones(shp)
        elif name.endswith("_beta"):
            a = rs.normal(0, 0.1, size=shp) if synthetic_stats else np.zeros(shp)
        else:
            a = rs.normal(0, 0.05, size=shp) if synthetic_stats else np.zeros(shp)
        arg[name] = a.astype(np.float32)
    for name, shp in aux_s.items():
        if name.endswith("mean"):
            a = rs.normal(0, 0.1, size=shp) if synthetic_stats else np.zeros(shp)
        else:
            a = rs.uniform(0.5, 1.5, size=shp) if synthetic_stats else np.ones(shp)
        aux[name] = a.astype(np.float32)
    return arg, aux


def flop_per_leaf(arch, width, height, n_blocks=8, n_filter=128):
    """Algorithmic FLOPs (2*MAC of convs + FCs) of one forward pass."""
    S = width * height
    arg, _ = param_shapes(arch, width, height, n_blocks, n_filter)
    mac = 0
    for name, shp in arg.items():
        if name.endswith("_weight"):
            mac += int(np.prod(shp)) * (S if len(shp) == 4 else 1)
    return 2 * mac
