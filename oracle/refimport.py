"""Import the UNMODIFIED reference modules from /root/reference (never copied).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Only usable in the build
container; ``/root/reference`` does not exist on the GPU box, so everything
that runs there uses the committed ``tests/golden`` fixtures instead.

Recipe: ``game.py:9-16`` imports ``policy_value_net_mxnet``
(needs MXNet) and ``utils`` (py2-only), and configures logging from
``config_loader.config_`` at import time -- pre-seed ``sys.modules`` with stubs.
"""
import os
import sys
import types

REF = os.environ.get("ALPHAPIG_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "mcts_alphaZero.py"))


_cache = {}


def load():
    """-> namespace with Board, Game, Game_AI, mcts_alphaZero, mcts_pure, train_mxnet / TrainPipeline."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF)
    sys.dont_write_bytecode = True  # reference tree is read-only
    if "policy_value_net_mxnet" not in sys.modules:
        m = types.ModuleType("policy_value_net_mxnet")
        m.PolicyValueNet = type("PolicyValueNet", (), {})
        sys.modules["policy_value_net_mxnet"] = m
    if "utils" not in sys.modules:
        u = types.ModuleType("utils")
        sg = types.ModuleType("utils.sgf_dataIter")
        sg.get_data_from_files = lambda file_name, sgf_home: _cache["sgf"][file_name]
        cl = types.ModuleType("utils.config_loader")
        cl.config_ = {"train_logging": {"version": 1}}
        se = types.ModuleType("utils.send_email")
        se.send_mail = lambda *a, **k: None
        u.sgf_dataIter, u.config_loader, u.send_email = sg, cl, se
        sys.modules["utils.send_email"] = se
        sys.modules["utils"] = u
        sys.modules["utils.sgf_dataIter"] = sg
        sys.modules["utils.config_loader"] = cl
    _cache["sgf"] = {}
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import game
        import game_ai
        import mcts_alphaZero
        import mcts_pure
        # train_mxnet.py:22 overwrites CUDA_VISIBLE_DEVICES at import time: put the caller's value back
        cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
        import train_mxnet
        if cvd is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    ns = types.SimpleNamespace(
        Board=game.Board, Game=game.Game, Game_AI=game_ai.Game_AI, game=game, game_ai=game_ai,
        mcts_alphaZero=mcts_alphaZero, mcts_pure=mcts_pure, train_mxnet=train_mxnet,
        TrainPipeline=train_mxnet.TrainPipeline, sgf_records=_cache["sgf"])
    _cache["ns"] = ns
    return ns
