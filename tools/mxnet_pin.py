#!/usr/bin/env python
"""Lift the "parity unpinned" cap of the policy/value nets the day an MXNet wheel is at hand.

The reference's net arithmetic lives in MXNet (requirements.txt:8 pins mxnet==1.6.0), which is neither under
/root/reference nor installable offline, so oracle/net.py restates the documented operator semantics and says
"arithmetic parity unpinned".  ONE command on any machine that has MXNet 1.x and the reference checkout:

    python tools/mxnet_pin.py --reference /path/to/AlphaPig            # writes tests/golden/mxnet_pin.npz

It builds the reference's OWN ``PolicyValueNet`` classes (policy_value_net_mxnet_simple.py:19-254 and
policy_value_net_mxnet.py:19-309, imported unmodified), loads this repo's seeded parameters into them
(``model_params=(arg_params, aux_params)`` as mx.nd arrays), runs ``policy_value`` on the SURVEY 8(d) synthetic
positions and ONE ``train_step``, and stores inputs, parameters-after-step, probabilities, values, loss and entropy.
Commit the file: tests/test_oracle_golden.py::test_mxnet_pin then checks oracle/net.py (forward, fp32) and
alphapig_b200/train.py (Adam / wd / rescale_grad semantics) against it on every CPU run, and tests/test_gpu_net.py
checks the tensor-core path against the same tensors on the GPU box.  Nothing else in the repo changes.

Without MXNet the script refuses to run (it never fabricates a fixture)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of anxingle/AlphaPig")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "mxnet_pin.npz"))
    ap.add_argument("--boards", type=int, default=24)
    a = ap.parse_args()
    try:
        import mxnet as mx
    except ImportError:
        sys.exit("mxnet is not importable here: install mxnet 1.x (the reference pins 1.6.0) and re-run")
    sys.path.insert(0, a.reference)
    sys.dont_write_bytecode = True
    import policy_value_net_mxnet as ref_res  # noqa: E402  (the reference's files, unmodified)
    import policy_value_net_mxnet_simple as ref_simple  # noqa: E402
    from alphapig_b200.params import init_params  # noqa: E402
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oboard_from, synth_position  # noqa: E402

    W = 15
    boards = [oboard_from(W, W, 5, synth_position(W, W, 5, 1234 + g)) for g in range(a.boards)]
    states = np.stack([np.ascontiguousarray(b.current_state()) for b in boards]).astype(np.float32)
    rs = np.random.RandomState(0)
    pis = rs.dirichlet(np.ones(W * W), size=a.boards).astype(np.float32)
    zs = rs.choice([-1.0, 1.0], size=a.boards).astype(np.float32)
    out = {"states": states, "pis": pis, "zs": zs, "mxnet_version": np.array(mx.__version__)}
    for tag, mod, arch, kw in (("simple", ref_simple, "simple", {}), ("res3", ref_res, "resnet", {"n_blocks": 3, "n_filter": 128})):
        arg, aux = init_params(arch, W, W, n_blocks=kw.get("n_blocks", 0), seed=0, synthetic_stats=True)
        params = ({k: mx.nd.array(v) for k, v in arg.items()}, {k: mx.nd.array(v) for k, v in aux.items()})
        net = mod.PolicyValueNet(W, W, batch_size=a.boards, model_params=params, **kw)
        probs, values = net.policy_value(states)
        out[tag + "_probs"], out[tag + "_values"] = np.asarray(probs), np.asarray(values)
        # policy_value_fn contract on one board (availables order, values[0])
        ap_, v_ = net.policy_value_fn(boards[0])
        out[tag + "_fn_acts"] = np.array([k for k, _ in ap_] if not isinstance(ap_, zip) else [], np.int64)
        loss, entropy = net.train_step(states, pis, zs, 2e-3)
        out[tag + "_loss"], out[tag + "_entropy"] = np.asarray(loss), np.asarray(entropy)
        arg2, aux2 = net.get_policy_param()
        for k, v in list(arg2.items()) + list(aux2.items()):
            out["%s_after/%s" % (tag, k)] = v.asnumpy()
        probs2, values2 = net.policy_value(states)
        out[tag + "_probs_after"], out[tag + "_values_after"] = np.asarray(probs2), np.asarray(values2)
    np.savez_compressed(a.out, **out)
    print("wrote", a.out, "- commit it; tests/test_oracle_golden.py::test_mxnet_pin picks it up")


if __name__ == "__main__":
    main()
