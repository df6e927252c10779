"""CPU (-m "not gpu"): the Python face drops in under the reference's own module names.

``alphapig_b200.install()`` registers ``game``, ``game_ai``, ``mcts_alphaZero``, ``mcts_pure``,
``policy_value_net_mxnet``, ``policy_value_net_mxnet_simple``, ``utils`` (+ ``AlphaPig`` for evaluate/ChessClient.py)
in ``sys.modules``.  The tests below execute the UNMODIFIED reference callers from /root/reference on top of those
aliases (each in its own interpreter: the oracle's ``refimport`` seeds the same module names with the reference's
own files).  Without a GPU the chain must run up to the first ``Engine()`` and fail loudly there - there is no CPU
fallback to fall into."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.environ.get("ALPHAPIG_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train_mxnet.py")), reason="reference tree not present")


def _run(code, cwd=None):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=ROOT, ALPHAPIG_REFERENCE=REF)
    p = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd or ROOT, env=env, capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stdout + "\n" + p.stderr
    return p.stdout


def test_install_registers_and_uninstall_reverts():
    out = _run("""
        import sys
        import alphapig_b200
        done = alphapig_b200.install()
        import game, game_ai, mcts_alphaZero, mcts_pure, policy_value_net_mxnet, policy_value_net_mxnet_simple
        from utils import sgf_dataIter, config_loader, send_email
        import AlphaPig as gomoku_zm
        assert game is sys.modules['alphapig_b200.game'] and game.Board.__module__ == 'alphapig_b200.game'
        assert game_ai.Game_AI.__mro__[1] is game.Game
        assert policy_value_net_mxnet.PolicyValueNet.arch == 'resnet'
        assert policy_value_net_mxnet_simple.PolicyValueNet.arch == 'simple'
        # what evaluate/ChessClient.py:189-199 dereferences
        for name in ('game', 'mcts_alphaZero', 'mcts_pure', 'policy_value_net_mxnet', 'policy_value_net_mxnet_simple'):
            assert getattr(gomoku_zm, name) is sys.modules[name], name
        assert gomoku_zm.mcts_alphaZero.MCTSPlayer is mcts_alphaZero.MCTSPlayer
        assert 'train_logging' in config_loader.config_ and config_loader.config_['n_playout'] == 400
        assert send_email.send_mail('t', 'm', 'x') is False
        assert 'train_mxnet' not in done
        assert 'train_mxnet' in alphapig_b200.install(train_pipeline=True)
        import train_mxnet
        assert train_mxnet.TrainPipeline.__module__ == 'alphapig_b200.train_mxnet'
        alphapig_b200.uninstall()
        assert 'game' not in sys.modules and 'AlphaPig' not in sys.modules and 'utils' not in sys.modules
        # a foreign module under one of the names is not silently replaced
        import types
        sys.modules['game'] = types.ModuleType('game')
        try:
            alphapig_b200.install()
            raise SystemExit('expected ImportError')
        except ImportError:
            pass
        alphapig_b200.install(force=True)
        assert sys.modules['game'] is sys.modules['alphapig_b200.game']
        print('ok')
    """)
    assert out.strip().endswith("ok")


@needs_ref
def test_reference_import_blocks_run_unmodified_on_the_aliases():
    """train_mxnet.py:18-29 and human_play_mxnet.py:10-14, read from the reference tree and executed verbatim."""
    out = _run("""
        import os, re, sys
        import alphapig_b200
        alphapig_b200.install()
        ref = os.environ['ALPHAPIG_REFERENCE']
        for fname, lo, hi in (('train_mxnet.py', 18, 29), ('human_play_mxnet.py', 10, 14)):
            lines = open(os.path.join(ref, fname), encoding='utf-8').read().split('\\n')[lo - 1:hi]
            block = '\\n'.join(l for l in lines if re.match(r'(from|import) ', l))
            assert 'from game import Board, Game' in block and 'PolicyValueNet' in block, block
            ns = {}
            exec(compile(block, fname, 'exec'), ns)
            for name in ('Board', 'Game', 'MCTS_Pure', 'MCTSPlayer', 'PolicyValueNet'):
                assert ns[name].__module__.startswith('alphapig_b200.'), (fname, name, ns[name].__module__)
            if fname == 'train_mxnet.py':
                assert ns['Game_AI'].__module__ == 'alphapig_b200.game_ai'
                assert hasattr(ns['config_loader'], 'config_') and hasattr(ns['send_email'], 'send_mail')
        print('ok')
    """)
    assert out.strip().endswith("ok")


@needs_ref
def test_unmodified_reference_train_script_drives_the_shims(tmp_path):
    """The reference's own train_mxnet.py (module body + TrainPipeline.__init__, train_mxnet.py:37-95) imported from
    /root/reference on top of the aliases: on a box without a GPU it must get as far as ``PolicyValueNet(...)`` (every
    import, the logging setup, Board / Game / Game_AI construction, the SGF directory scan) and stop at the engine's
    loud no-CPU-fallback error; human_play_mxnet.py:51-81 likewise up to its ``PolicyValueNet``."""
    out = _run("""
        import importlib.util, os, pickle, sys
        import alphapig_b200
        alphapig_b200.install()
        from alphapig_b200._lib import EngineError
        ref = os.environ['ALPHAPIG_REFERENCE']
        cvd = os.environ.get('CUDA_VISIBLE_DEVICES')

        def load(name):
            spec = importlib.util.spec_from_file_location('ref_' + name, os.path.join(ref, name + '.py'))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod

        tm = load('train_mxnet')
        assert tm.Board is sys.modules['alphapig_b200.game'].Board
        assert tm.PolicyValueNet is sys.modules['alphapig_b200.policy_value_net_mxnet'].PolicyValueNet
        conf = tm.config_loader.load_config(os.path.join(ref, 'conf', 'train_config.yaml'))
        assert conf['n_playout'] == 400 and conf['board_width'] == 15
        import torch
        if not torch.cuda.is_available():
            try:
                tm.TrainPipeline(conf)
                raise SystemExit('TrainPipeline() must not succeed without a GPU')
            except EngineError as e:
                assert 'no CPU fallback' in str(e), e
        hp = load('human_play_mxnet')
        assert hp.MCTSPlayer is sys.modules['alphapig_b200.mcts_alphaZero'].MCTSPlayer
        h = hp.Human(); h.set_player_ind(2); assert str(h) == 'Human 2'
        # run() loads ./logs/current_policy.model (a pickled (arg_params, aux_params)) and builds the net from it
        from alphapig_b200.params import init_params
        os.makedirs('logs', exist_ok=True)
        arg, aux = init_params('resnet', 15, 15, 8, 128, seed=0)
        pickle.dump((dict(arg), dict(aux)), open('logs/current_policy.model', 'wb'), protocol=2)
        if not torch.cuda.is_available():
            try:
                hp.run()
                raise SystemExit('run() must not succeed without a GPU')
            except EngineError as e:
                assert 'no CPU fallback' in str(e), e
        print('ok')
    """, cwd=str(tmp_path))
    assert out.strip().endswith("ok")


@needs_ref
def test_reference_train_script_runs_as_main_on_the_aliases(tmp_path):
    """``python train_mxnet.py`` itself - the reference's ``__main__`` block (train_mxnet.py:286-309: load
    ./conf/train_config.yaml, build TrainPipeline, run(), send the closing e-mail) - executed with runpy on the aliases
    from a working directory laid out like the reference's.  Without a GPU it gets through every import, the logging
    configuration of the reference's own YAML, Board / Game / Game_AI and the SGF directory scan, logs the engine's
    "no CPU fallback" error through ITS OWN ``except Exception`` handler and ends with the (inert) ``send_mail``."""
    import shutil
    (tmp_path / "conf").mkdir()
    (tmp_path / "sgf_data").mkdir()
    shutil.copy(os.path.join(REF, "conf", "train_config.yaml"), str(tmp_path / "conf" / "train_config.yaml"))
    out = _run("""
        import logging, os, runpy, sys
        import alphapig_b200
        alphapig_b200.install()
        import torch
        sent = []
        import utils.send_email as se
        real = se.send_mail
        se.send_mail = lambda *a: (sent.append(a), real(*a))[1]
        runpy.run_path(os.path.join(os.environ['ALPHAPIG_REFERENCE'], 'train_mxnet.py'), run_name='__main__')
        assert len(sent) == 1, sent                      # the `finally:` of the reference's main block ran
        assert os.path.isdir('logs')                      # its YAML's file handlers got their directory
        if not torch.cuda.is_available():
            log = open(os.path.join('logs', 'error.log')).read()
            assert 'no CPU fallback' in log, log[-400:]
        print('ok')
    """, cwd=str(tmp_path))
    assert out.strip().endswith("ok")
