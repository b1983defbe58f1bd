"""B200-native batched Othello self-play engine behind the call surface of
AfoninAndrei/alphaZero-Othello: ``envs.game.Game`` / ``envs.othello.OthelloGameNew``,
``MCTS_model.MCTS``, ``self_play_worker.one_self_play`` (+ the batched
``collect_self_play_games``).  CUDA kernels in ``csrc/`` behind the C ABI of
``include/othello_b200.h``; there is no CPU fallback."""
__version__ = "0.1.0"
