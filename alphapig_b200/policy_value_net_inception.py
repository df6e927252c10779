"""``PolicyValueNet`` for BASELINE configs[3]: a board-sized Inception-ResNet policy/value net.

The reference ships ``inception-resnet-v2.py`` (the MXNet ImageNet example, never wired into the game code and
not applicable to a 15x15 plane: its stride-2 convs and poolings collapse the board), so this variant is
DEFINED BY THIS REPO and pinned only by the repo's own fp32 restatement (``oracle/net.py``): a 3x3 stem
(9 -> 128), ``n_blocks`` x ``block35`` (inception-resnet-v2.py:41-58, scale 0.17) and the reference's policy /
value heads.  Same constructor and methods as the other ``PolicyValueNet`` shims."""
from .nets import PolicyValueNetBase


class PolicyValueNet(PolicyValueNetBase):
    arch = "inception"

    def __init__(self, board_width, board_height, batch_size=512, n_blocks=10, n_filter=128, model_params=None, **kw):
        PolicyValueNetBase.__init__(self, board_width, board_height, batch_size=batch_size, n_blocks=n_blocks,
                                    n_filter=n_filter, model_params=model_params, **kw)
