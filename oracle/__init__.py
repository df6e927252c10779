"""CPU oracle for the AlphaPig self-play hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy / pure-Python
restatement of the reference algorithm (anxingle/AlphaPig: ``game.py``,
``game_ai.py``, ``mcts_alphaZero.py``, ``mcts_pure.py``,
``policy_value_net_mxnet{,_simple}.py``).  It exists so that the CUDA engine in
``alphapig_b200`` can be checked bit-for-bit (boards, visit counts, moves) and
to 1e-3 (net outputs) on a box where ``/root/reference`` does not exist.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs (``cpu_baseline`` / ``--impl reference``) may import it.  Nothing under
``alphapig_b200/`` imports it; the product path has no CPU fallback.

Pinning status
--------------
* ``oracle.board`` / ``oracle.mcts`` / ``oracle.selfplay``: PINNED.  They are
  checked against the *imported, unmodified* reference classes in this
  container (``tests/golden/make_golden.py``; ``tests/test_oracle_vs_reference.py``
  re-runs the comparison live whenever ``/root/reference`` is present) and
  against the committed golden vectors in ``tests/golden/*.npz`` that the same
  script wrote from the reference.
* ``oracle.pipeline`` (``get_equi_data``, deque + ``random.sample``, ``policy_update``,
  ``policy_evaluate``): PINNED live against ``train_mxnet.TrainPipeline`` driven with a fake net
  (``tests/test_oracle_vs_reference.py``) and by ``tests/golden/pipeline_cases.npz``
  (``tests/golden/make_golden_pipeline.py``).
* ``oracle.net``: PARITY UNPINNED at the MXNet boundary.  The reference's net
  arithmetic lives in third-party MXNet (``requirements.txt:8`` pins
  ``mxnet==1.6.0``) which is not installed and not installable here (no
  network); the reference ships no weights, golden tensors or tests for it.
  ``oracle.net`` restates the published MXNet operator semantics the reference
  call sites rely on (see the module docstring) and is cross-checked fp32 vs
  fp64 only.
"""
