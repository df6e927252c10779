"""The training step of the reference nets in PyTorch (north_star: "PyTorch allowed for ... the
training step").  Restates ``create_policy_value_train`` + ``train_step``
(policy_value_net_mxnet_simple.py:121-159,228-244; policy_value_net_mxnet.py:173-212,282-299)
with the MXNet-1.x semantics listed in SURVEY 8(c):

* loss = mean((z - v)^2) + mean(-sum(pi * log p)); entropy = mean(-sum(p log p)) (reported only)
* BatchNorm in training mode, momentum 0.9, eps 1e-3, ``fix_gamma=True`` for the ``conv_act`` BNs
  (gamma pinned to 1), trainable gamma for ``bnA*/bnB*``
* Dropout(0.5) before both FC layers
* Adam(beta 0.9/0.999, eps 1e-8), lr given per call, wd=1e-4 added to the gradient of ``*_weight``
  and ``*_gamma`` only, ``rescale_grad = 1/batch_size`` (Module.init_optimizer with a string optimizer)

Device-agnostic: runs on cuda in the product and on CPU (fp32/fp64) in the host-logic tests.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.9


def trunk_plan(arch, n_blocks):
    if arch == "simple":
        return [("conv_act", n) for n in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv_final")]
    if arch == "inception":
        return [("conv_act", "incep_conv1")] + [("b35", i) for i in range(1, n_blocks + 1)]
    return [("conv_act", "res_conv1")] + [("res", i) for i in range(1, n_blocks + 1)]


def _bn_train(x, P, gamma, beta, mean_name, var_name, fix_gamma, new_stats):
    """BatchNorm in training mode (batch statistics, biased variance, eps 1e-3) through the fused native op -
    one forward and one backward kernel instead of a dozen element-wise ones - plus the MXNet-style moving statistics
    (momentum 0.9 on the biased batch variance)."""
    # fix_gamma: gamma is the constant 1 (an explicit ones vector: cuDNN's backward wants a defined weight tensor)
    y = F.batch_norm(x, None, None, weight=x.new_ones(x.shape[1]) if fix_gamma else P[gamma], bias=P[beta], training=True,
                     momentum=0.0, eps=BN_EPS)
    with torch.no_grad():
        var, mu = torch.var_mean(x, dim=(0, 2, 3), unbiased=False)
        new_stats[mean_name] = P[mean_name] * BN_MOMENTUM + mu * (1 - BN_MOMENTUM)
        new_stats[var_name] = P[var_name] * BN_MOMENTUM + var * (1 - BN_MOMENTUM)
    return y


def forward_train(P, x, arch, n_blocks, dropout_gen=None, dropout=True):
    """P: dict name -> tensor (arg params require grad).  Returns probs, value, new moving stats."""
    new_stats = {}

    def conv_act(x, name, k, act=True):
        y = F.conv2d(x, P[name + "_weight"], P[name + "_bias"], padding=k // 2)
        y = _bn_train(y, P, name + "_gamma", name + "_beta", name + "_mean", name + "_var", True, new_stats)
        return F.relu(y) if act else y

    for kind, key in trunk_plan(arch, n_blocks):
        if kind == "conv_act":
            x = conv_act(x, key, 3)
        elif kind == "b35":  # block35 of the Inception-ResNet variant (params.py)
            pre = "b35_%d_" % key
            t0 = conv_act(x, pre + "t0", 1)
            t1 = conv_act(conv_act(x, pre + "t1a", 1), pre + "t1b", 3)
            t2 = conv_act(conv_act(conv_act(x, pre + "t2a", 1), pre + "t2b", 3), pre + "t2c", 3)
            up = conv_act(torch.cat([t0, t1, t2], dim=1), pre + "up", 1, act=False)
            x = F.relu(x + 0.17 * up)
        else:
            idn = x
            y = F.conv2d(x, P["convA%d_weight" % key], P["convA%d_bias" % key], padding=1)
            y = F.relu(_bn_train(y, P, "bnA%d_gamma" % key, "bnA%d_beta" % key, "bnA%d_moving_mean" % key,
                                 "bnA%d_moving_var" % key, False, new_stats))
            y = F.conv2d(y, P["convB%d_weight" % key], P["convB%d_bias" % key], padding=1)
            y = _bn_train(y, P, "bnB%d_gamma" % key, "bnB%d_beta" % key, "bnB%d_moving_mean" % key,
                          "bnB%d_moving_var" % key, False, new_stats)
            x = F.relu(y + idn)
    B = x.shape[0]

    def drop(t):
        if not dropout:
            return t
        keep = (torch.rand(t.shape, generator=dropout_gen, device=t.device, dtype=t.dtype) >= 0.5).to(t.dtype)
        return t * keep * 2.0

    p = drop(conv_act(x, "conv3_1_1", 1).reshape(B, -1))
    probs = torch.softmax(p @ P["fc_3_1_1_weight"].t() + P["fc_3_1_1_bias"], dim=1)
    v = drop(conv_act(x, "conv3_2_1", 1).reshape(B, -1))
    value = torch.tanh(v @ P["fc_3_2_1_weight"].t() + P["fc_3_2_1_bias"])
    return probs, value, new_stats


class AdamState(object):
    def __init__(self):
        self.t = 0
        self.m = {}
        self.v = {}


def train_step(arg, aux, opt, states, mcts_probs, winners, lr, arch, n_blocks=0, wd=1e-4,
               dropout=True, dropout_gen=None):
    """In-place update of ``arg`` (trainable) and ``aux`` (moving stats) tensors.
    Returns (loss, entropy) as 0-dim tensors."""
    B = states.shape[0]
    names = [k for k in arg if not (k.endswith("_gamma") and not k.startswith("bn"))]  # fix_gamma BNs are constants
    P = dict(aux)
    for k, t in arg.items():
        P[k] = t.detach().requires_grad_(k in names)
    probs, value, new_stats = forward_train(P, states, arch, n_blocks, dropout_gen, dropout)
    logp = torch.log(probs)
    policy_loss = (-(logp * mcts_probs).sum(dim=1)).mean()
    value_loss = ((winners.reshape(B, 1) - value) ** 2).mean()
    loss = value_loss + policy_loss
    entropy = (-(probs * logp).sum(dim=1)).mean().detach()
    grads = torch.autograd.grad(loss, [P[k] for k in names])
    opt.t += 1
    b1, b2, eps = 0.9, 0.999, 1e-8
    lr_t = lr * math.sqrt(1.0 - b2 ** opt.t) / (1.0 - b1 ** opt.t)
    with torch.no_grad():
        # multi-tensor (foreach) form of the per-parameter update: the same element-wise operations in the same
        # order, ~12 kernel launches instead of ~10 per parameter tensor (on a GPU that is also running the search
        # every trainer launch is a chance to delay a persistent conv kernel, alphapig_b200/loop.py)
        ws = [arg[k] for k in names]
        for k, w in zip(names, ws):
            if k not in opt.m:
                opt.m[k] = torch.zeros_like(w)
                opt.v[k] = torch.zeros_like(w)
        ms = [opt.m[k] for k in names]
        vs = [opt.v[k] for k in names]
        gs = torch._foreach_mul(list(grads), 1.0 / B)
        decayed = [i for i, k in enumerate(names) if k.endswith("_weight") or k.endswith("_gamma")]
        if decayed:
            torch._foreach_add_([gs[i] for i in decayed], [ws[i] for i in decayed], alpha=wd)
        torch._foreach_mul_(ms, b1)
        torch._foreach_add_(ms, gs, alpha=1 - b1)
        torch._foreach_mul_(vs, b2)
        torch._foreach_addcmul_(vs, gs, gs, value=1 - b2)
        den = torch._foreach_sqrt(vs)
        torch._foreach_add_(den, eps)
        upd = torch._foreach_mul(ms, lr_t)
        torch._foreach_div_(upd, den)
        torch._foreach_sub_(ws, upd)
        stat_names = list(new_stats)
        if stat_names:
            torch._foreach_copy_([aux[k] for k in stat_names], [new_stats[k] for k in stat_names])
    return loss.detach(), entropy
