"""The environment contract MCTS and the self-play worker program against.

Same surface as the reference's ``envs/game.py:5-57`` (an ABC with
get_initial_state / get_valid_moves / action_size / state_size /
get_next_state / get_value_and_terminated / get_opponent), kept as the drop-in
seam: anything written against the reference's ``Game`` runs against this one.
"""
import abc


class Game(abc.ABC):
    """Two-player, perfect-information board game seen through numpy arrays."""

    @abc.abstractmethod
    def get_initial_state(self):
        """Starting board."""

    @abc.abstractmethod
    def get_valid_moves(self, state, player):
        """0/1 vector over ``action_size`` actions for ``player`` to move."""

    @property
    @abc.abstractmethod
    def action_size(self):
        """Number of actions."""

    @property
    @abc.abstractmethod
    def state_size(self):
        """Number of board cells."""

    @abc.abstractmethod
    def get_next_state(self, state, action, player):
        """Board after ``player`` plays ``action`` (a new array)."""

    @abc.abstractmethod
    def get_value_and_terminated(self, state, action, player):
        """``(value, terminated)``: +1 / -1 / 0 from ``player``'s side once the game is over."""

    @abc.abstractmethod
    def get_opponent(self, player):
        """The other player."""
