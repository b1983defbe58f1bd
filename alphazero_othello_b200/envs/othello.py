"""Othello 8x8 behind the reference's Game API, computed on the B200.

``OthelloGameNew`` mirrors ``envs/othello.py:309-460`` of the reference (same
method names, argument meaning, dtypes and error behaviour); every rule
evaluation is a CUDA kernel call through the C ABI (``oth_host_*``,
include/othello_b200.h).  ``BatchedOthello`` is the same thing for device
tensors of packed bitboards, which is what the self-play engine uses.

There is no CPU implementation in this package.
"""
import ctypes as C
import sys

import numpy as np

from .. import _lib
from .game import Game


def _i8(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int8)
    return a if shape is None else a.reshape(shape)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class OthelloGameNew(Game):
    """Drop-in for the reference's ``OthelloGameNew`` (imported there as ``OthelloGame``)."""

    square_content = {-1: "X", 0: "-", 1: "O"}

    @staticmethod
    def get_square_piece(piece):
        return OthelloGameNew.square_content[piece]

    def __init__(self, n):
        assert n == 8, "Bitboard engine supports only standard 8×8 Othello"  # envs/othello.py:322
        _lib.require_device()
        self.n = n
        self._state_size = n * n
        self._action_size = n * n + 1

    @property
    def action_size(self):
        return self._action_size

    @property
    def state_size(self):
        return self._state_size

    def get_initial_state(self):
        s = np.zeros((8, 8), np.int8)  # envs/othello.py:390-392
        s[3, 4] = s[4, 3] = 1
        s[3, 3] = s[4, 4] = -1
        return s

    # -- batched forms (n states per call) -----------------------------------
    def valid_moves_batch(self, states, players):
        states = _i8(states, (-1, 64))
        players = _i8(players, (-1,))
        out = np.empty((len(states), 65), np.uint8)
        _lib.check(_lib.lib().oth_host_valid_moves(_ptr(states), _ptr(players), _ptr(out), len(states)), "get_valid_moves")
        return out

    def next_state_batch(self, states, actions, players):
        states = _i8(states, (-1, 64))
        players = _i8(players, (-1,))
        actions = np.ascontiguousarray(actions, dtype=np.int32).reshape(-1)
        out = np.empty_like(states)
        flags = np.empty(len(states), np.uint8)
        rc = _lib.lib().oth_host_next_state(_ptr(states), _ptr(actions), _ptr(players), _ptr(out), _ptr(flags), len(states))
        if rc == _lib.OTH_E_ILLEGAL:
            bad = int(np.nonzero(flags & _lib.F_ILLEGAL)[0][0])
            raise ValueError(f"Illegal move: {int(actions[bad])}")
        _lib.check(rc, "get_next_state")
        return out.reshape(-1, 8, 8)

    def value_and_terminated_batch(self, states, players):
        states = _i8(states, (-1, 64))
        players = _i8(players, (-1,))
        v = np.empty(len(states), np.int8)
        t = np.empty(len(states), np.uint8)
        _lib.check(_lib.lib().oth_host_value_terminated(_ptr(states), _ptr(players), _ptr(v), _ptr(t), len(states)),
                   "get_value_and_terminated")
        return v, t.astype(bool)

    # -- the reference's single-state API ------------------------------------
    def get_valid_moves(self, state, player):
        return self.valid_moves_batch(np.asarray(state)[None], [player])[0]  # envs/othello.py:394-411

    def get_next_state(self, state, action, player):
        return self.next_state_batch(np.asarray(state)[None], [action], [player])[0]  # envs/othello.py:413-433

    def get_value_and_terminated(self, state, action, player):
        v, t = self.value_and_terminated_batch(np.asarray(state)[None], [player])  # `action` ignored, :435-454
        return int(v[0]), bool(t[0])

    def get_score(self, state, player):
        state = np.asarray(state)
        return int(np.sum(state == player) - np.sum(state == -player))  # envs/othello.py:456-457

    def get_opponent(self, player):
        return -player

    def get_symmetries(self, state, pi):
        """The 8 dihedral images in the order of envs/othello.py:286-298."""
        assert len(pi) == self.n ** 2 + 1
        ks = np.repeat(np.arange(1, 5, dtype=np.int32), 2)
        fl = np.tile(np.array([1, 0], np.uint8), 4)
        s, p = symmetry_batch(np.repeat(_i8(state, (1, 64)), 8, 0), np.repeat(np.asarray(pi, np.float32)[None], 8, 0), ks, fl)
        dt = np.asarray(state).dtype
        return [(s[i].reshape(8, 8).astype(dt), list(p[i])) for i in range(8)]

    def print_board(self, state, player, ply=None):
        rows = ["  a b c d e f g h"]
        for r in range(8):
            rows.append(" ".join([str(r + 1)] + [self.square_content[-int(state[r, c])] for c in range(8)]))
        print("\n".join(rows))


OthelloGame = OthelloGameNew  # the name every runtime path of the reference imports


def symmetry_batch(states, pis, ks, flips):
    """Dihedral images (np.rot90 by ks[i], then np.fliplr if flips[i]) of int8 boards
    [n,64] and float32 policies [n,65]; returns float32 arrays."""
    states = _i8(states, (-1, 64))
    pis = np.ascontiguousarray(pis, dtype=np.float32).reshape(-1, 65)
    ks = np.ascontiguousarray(ks, dtype=np.int32).reshape(-1)
    flips = np.ascontiguousarray(flips, dtype=np.uint8).reshape(-1)
    n = len(states)
    assert len(pis) == n and len(ks) == n and len(flips) == n
    os_ = np.empty((n, 64), np.float32)
    op = np.empty((n, 65), np.float32)
    _lib.check(_lib.lib().oth_host_symmetry(_ptr(states), _ptr(pis), _ptr(ks), _ptr(flips), _ptr(os_), _ptr(op), n),
               "symmetry")
    return os_, op


def get_random_symmetry(state, pi):
    """Drop-in for envs/othello.py:501-526: same np.random draws (randint(4), then
    rand() < 0.5), same output shapes/dtypes ((1, n, n) float32, (n*n+1,) float32)."""
    tud = sys.modules.get("torch.utils.data")
    if tud is not None and tud.get_worker_info() is not None:
        # train.py:26-42 calls this per sample inside DataLoader workers (forked children of a process that already
        # holds a CUDA context): CUDA cannot be used there.  Augment the collated batch in the training process
        # instead -- replay.augment_batch / BatchedOthello.random_symmetry.
        raise RuntimeError("get_random_symmetry runs on the GPU and cannot be called from a DataLoader worker; "
                           "use alphazero_othello_b200.replay.augment_batch on the collated batch")
    k = np.random.randint(4)
    flip = np.random.rand() < 0.5
    state = np.asarray(state)
    assert state.shape[-2:] == (8, 8) and state.size == 64, "2-D 8x8 states only"
    s, p = symmetry_batch(state.reshape(1, 64), np.asarray(pi, np.float32)[None], [k], [flip])
    return s.reshape(1, 8, 8), p[0]


class BatchedOthello:
    """Device-tensor API over packed bitboards (bit i = row*8+col; ``own`` = side to move)."""

    def __init__(self, device="cuda:0"):
        import torch
        _lib.require_device()
        self.torch = torch
        self.device = torch.device(device)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def initial(self, n):
        t = self.torch
        own = t.full((n,), (1 << 28) | (1 << 35), dtype=t.int64, device=self.device)
        opp = t.full((n,), (1 << 27) | (1 << 36), dtype=t.int64, device=self.device)
        return own, opp

    def legal_moves(self, own, opp):
        own, opp = own.contiguous(), opp.contiguous()
        out = self.torch.empty_like(own)
        with self.torch.cuda.device(self.device):
            _lib.check(_lib.lib().oth_legal_moves(own.data_ptr(), opp.data_ptr(), out.data_ptr(), own.numel(), self._stream()))
        return out

    def step(self, own, opp, action):
        t = self.torch
        own, opp, action = own.contiguous(), opp.contiguous(), action.contiguous()
        n = own.numel()
        no, np_, nm = t.empty_like(own), t.empty_like(own), t.empty_like(own)
        fl = t.empty(n, dtype=t.uint8, device=self.device)
        with t.cuda.device(self.device):
            _lib.check(_lib.lib().oth_step(own.data_ptr(), opp.data_ptr(), action.data_ptr(), no.data_ptr(), np_.data_ptr(),
                                           nm.data_ptr(), fl.data_ptr(), n, self._stream()))
        return no, np_, nm, fl

    def rollout(self, n_games, seed=0, game_id_base=0, n_trace=0):
        t = self.torch
        score = t.empty(n_games, dtype=t.int32, device=self.device)
        plies = t.empty(n_games, dtype=t.int32, device=self.device)
        final = t.empty((n_games, 2), dtype=t.int64, device=self.device)
        ta = t.empty((max(n_trace, 1), _lib.OTH_MAX_PLIES), dtype=t.uint8, device=self.device)
        tm = t.empty((max(n_trace, 1), _lib.OTH_MAX_PLIES), dtype=t.int64, device=self.device)
        cnt = t.zeros(8, dtype=t.int64, device=self.device)
        with t.cuda.device(self.device):
            _lib.check(_lib.lib().oth_rollout(seed, game_id_base, n_games, score.data_ptr(), plies.data_ptr(), final.data_ptr(),
                                              n_trace, ta.data_ptr(), tm.data_ptr(), cnt.data_ptr(), self._stream()))
        return dict(score=score, plies=plies, final=final, trace_actions=ta[:n_trace], trace_moves=tm[:n_trace], counters=cnt)

    def symmetry(self, states, pis, ks, flips):
        """Training-time augmentation on device (RandomSymmetryDataset.__getitem__, train.py:36-42,
        batched): states int8[n,8,8], pis f32[n,65], ks int32[n] quarter turns, flips uint8[n].
        Returns (float32 [n,1,8,8], float32 [n,65]) exactly as get_random_symmetry would per sample."""
        t = self.torch
        n = states.shape[0]
        states, pis = states.contiguous(), pis.contiguous()
        ks, flips = ks.to(t.int32).contiguous(), flips.to(t.uint8).contiguous()
        os_ = t.empty((n, 1, 8, 8), dtype=t.float32, device=self.device)
        op = t.empty((n, 65), dtype=t.float32, device=self.device)
        with t.cuda.device(self.device):
            _lib.check(_lib.lib().oth_symmetry(states.data_ptr(), pis.data_ptr(), ks.data_ptr(), flips.data_ptr(), os_.data_ptr(),
                                               op.data_ptr(), n, self._stream()))
        return os_, op

    def random_symmetry(self, states, pis, generator=None):
        """One random dihedral image per sample (k ~ U{0..3}, flip ~ Bernoulli(1/2)), drawn on the device."""
        t = self.torch
        n = states.shape[0]
        ks = t.randint(0, 4, (n,), device=self.device, dtype=t.int32, generator=generator)
        flips = (t.rand(n, device=self.device, generator=generator) < 0.5).to(t.uint8)
        return self.symmetry(states, pis, ks, flips)

    def pack(self, states, players):
        t = self.torch
        n = states.shape[0]
        own = t.empty(n, dtype=t.int64, device=self.device)
        opp = t.empty(n, dtype=t.int64, device=self.device)
        with t.cuda.device(self.device):
            _lib.check(_lib.lib().oth_pack_states(states.data_ptr(), players.data_ptr(), own.data_ptr(), opp.data_ptr(), n,
                                                  self._stream()))
        return own, opp

    def unpack(self, own, opp, players):
        t = self.torch
        n = own.numel()
        out = t.empty((n, 8, 8), dtype=t.int8, device=self.device)
        with t.cuda.device(self.device):
            _lib.check(_lib.lib().oth_unpack_states(own.data_ptr(), opp.data_ptr(), players.data_ptr(), out.data_ptr(), n,
                                                    self._stream()))
        return out
