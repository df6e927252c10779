"""Drop-in for reference ``policy_value_net_mxnet.PolicyValueNet`` (residual net, stem 128 +
n_blocks x [conv-BN-ReLU-conv-BN-add-ReLU], policy_value_net_mxnet.py:19-309)."""
from .nets import PolicyValueNetBase


class PolicyValueNet(PolicyValueNetBase):
    arch = "resnet"

    def __init__(self, board_width, board_height, batch_size=512, n_blocks=8, n_filter=128, model_params=None, **kw):
        PolicyValueNetBase.__init__(self, board_width, board_height, batch_size=batch_size, n_blocks=n_blocks,
                                    n_filter=n_filter, model_params=model_params, **kw)
