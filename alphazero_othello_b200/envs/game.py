"""The environment contract that MCTS, the self-play worker and the arena program against.

It is the drop-in seam of this package: the same seven members as the reference's abstract
``Game`` (``envs/game.py:5-57``) with the same names and argument order, so code written against
the reference runs unchanged against :class:`alphazero_othello_b200.envs.othello.OthelloGameNew`.

Conventions shared by every implementation here
  * a *state* is a numpy array seen from absolute colours (+1 / -1 discs, 0 empty);
  * *player* is +1 or -1, the side to move;
  * an *action* is an integer in ``range(action_size)``; the last index is "pass" for Othello;
  * states are never mutated: ``get_next_state`` returns a new array.
"""
import abc


class Game(abc.ABC):
    """Two-player, deterministic, perfect-information board game."""

    # -- sizes ---------------------------------------------------------------------------------
    @property
    @abc.abstractmethod
    def action_size(self):
        """Number of distinct actions (policy vector length)."""

    @property
    @abc.abstractmethod
    def state_size(self):
        """Number of board cells."""

    # -- rules ---------------------------------------------------------------------------------
    @abc.abstractmethod
    def get_initial_state(self):
        """The starting position."""

    @abc.abstractmethod
    def get_valid_moves(self, state, player):
        """0/1 vector of length ``action_size``: what ``player`` may play in ``state``."""

    @abc.abstractmethod
    def get_next_state(self, state, action, player):
        """The position after ``player`` plays ``action`` (raises ``ValueError`` if it is illegal)."""

    @abc.abstractmethod
    def get_value_and_terminated(self, state, action, player):
        """``(value, terminated)``; once the game is over ``value`` is +1 / -1 / 0 from ``player``'s side,
        before that it is 0.  ``action`` is the move that led to ``state`` (implementations may ignore it)."""

    @abc.abstractmethod
    def get_opponent(self, player):
        """The other side."""

    # -- conveniences built on the contract (not part of the reference's ABC) --------------------
    def is_terminal(self, state, player):
        return self.get_value_and_terminated(state, None, player)[1]

    def legal_actions(self, state, player):
        return [a for a, ok in enumerate(self.get_valid_moves(state, player)) if ok]
