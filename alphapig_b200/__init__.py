"""alphapig_b200: B200-native (sm_100a) engine behind the Python API of anxingle/AlphaPig.

The package carries the reference's modules under their own names (``game``, ``game_ai``, ``mcts_alphaZero``,
``mcts_pure``, ``policy_value_net_mxnet``, ``policy_value_net_mxnet_simple``, ``train_mxnet``, ``utils``).
``install()`` registers them in ``sys.modules`` under the reference's TOP-LEVEL names so that the reference's own
scripts run unmodified on top of the engine:

    import alphapig_b200; alphapig_b200.install()
    from game import Board, Game                       # train_mxnet.py:18, human_play_mxnet.py:11
    from game_ai import Game_AI                        # train_mxnet.py:19
    from mcts_pure import MCTSPlayer as MCTS_Pure      # train_mxnet.py:20
    from mcts_alphaZero import MCTSPlayer              # train_mxnet.py:21
    from utils import config_loader, send_email        # train_mxnet.py:22
    from policy_value_net_mxnet import PolicyValueNet  # train_mxnet.py:29
    import AlphaPig as gomoku_zm                       # evaluate/ChessClient.py:20

Nothing here touches the GPU: the shared library is loaded by the first ``Engine()``.
"""
import importlib
import sys

__version__ = "0.2"

# reference top-level module name -> module of this package
REFERENCE_MODULES = {
    "game": "game",                                              # game.py
    "game_ai": "game_ai",                                        # game_ai.py
    "mcts_alphaZero": "mcts_alphaZero",                          # mcts_alphaZero.py
    "mcts_pure": "mcts_pure",                                    # mcts_pure.py
    "policy_value_net_mxnet": "policy_value_net_mxnet",          # policy_value_net_mxnet.py (residual net)
    "policy_value_net_mxnet_simple": "policy_value_net_mxnet_simple",  # policy_value_net_mxnet_simple.py
    "utils": "utils",                                            # utils/__init__.py
    "utils.sgf_dataIter": "utils.sgf_dataIter",                  # utils/sgf_dataIter.py
    "utils.config_loader": "utils.config_loader",                # utils/config_loader.py
    "utils.send_email": "utils.send_email",                      # utils/send_email.py (inert: e-mail is out of scope)
}
# modules the reference package __init__ exposes (``import AlphaPig as gomoku_zm``, evaluate/ChessClient.py:17-20,189-199)
PACKAGE_ALIAS = "AlphaPig"
_installed = {}


def install(package_alias=PACKAGE_ALIAS, train_pipeline=False, force=False):
    """Register this package's modules under the reference's import names.

    package_alias:  name under which the whole package is importable (``AlphaPig`` for evaluate/ChessClient.py);
                    None skips it.
    train_pipeline: also alias ``train_mxnet`` to this package's device-resident ``TrainPipeline``.  Off by default:
                    the reference's own ``train_mxnet.py`` is a CALLER of the aliased modules and runs unmodified.
    force:          replace foreign modules already imported under these names (default: raise ImportError).

    Returns the dict name -> module that was registered; ``uninstall()`` reverts it."""
    names = dict(REFERENCE_MODULES)
    if train_pipeline:
        names["train_mxnet"] = "train_mxnet"
    done = {}
    for ref_name, ours in names.items():
        mod = importlib.import_module("%s.%s" % (__name__, ours))
        have = sys.modules.get(ref_name)
        if have is not None and have is not mod and not force:
            raise ImportError("alphapig_b200.install(): a different module %r is already imported from %r "
                              "(pass force=True to replace it)" % (ref_name, getattr(have, "__file__", "?")))
        sys.modules[ref_name] = mod
        done[ref_name] = mod
    if package_alias:
        pkg = sys.modules[__name__]
        have = sys.modules.get(package_alias)
        if have is not None and have is not pkg and not force:
            raise ImportError("alphapig_b200.install(): %r is already imported" % package_alias)
        sys.modules[package_alias] = pkg
        done[package_alias] = pkg
        for ref_name, ours in names.items():
            sys.modules["%s.%s" % (package_alias, ref_name)] = sys.modules["%s.%s" % (__name__, ours)]
            done["%s.%s" % (package_alias, ref_name)] = sys.modules["%s.%s" % (__name__, ours)]
    _installed.update(done)
    return done


def uninstall():
    """Remove the aliases ``install()`` registered (modules imported under them stay alive where referenced)."""
    for name, mod in list(_installed.items()):
        if sys.modules.get(name) is mod:
            del sys.modules[name]
        del _installed[name]
