"""Drop-in for reference ``policy_value_net_mxnet_simple.PolicyValueNet`` (6-conv net,
policy_value_net_mxnet_simple.py:19-254): same constructor, ``policy_value``, ``policy_value_fn``,
``train_step``, ``get_policy_param``, ``save_model``."""
from .nets import PolicyValueNetBase


class PolicyValueNet(PolicyValueNetBase):
    arch = "simple"

    def __init__(self, board_width, board_height, batch_size=512, model_params=None, **kw):
        PolicyValueNetBase.__init__(self, board_width, board_height, batch_size=batch_size,
                                    model_params=model_params, **kw)
