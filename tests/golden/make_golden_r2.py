"""Generate tests/golden/ref_golden_r2.npz by RUNNING THE REFERENCE ITSELF (round-2 additions).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r2.py        (dev container only)

Two parts, both results of the unmodified reference driven through its public entry points:

* ``ro_*``  -- ``MCTS(env, args, policy=None)`` (MCTS_model.py:276-303, 332-335: uniform priors + one random
  playout per leaf).  ``np.random.choice`` is wrapped so that the index every playout pick returned is logged;
  the oracle consumes that stream instead of its Philox draws and must reproduce visit counts, child values
  and root values bit for bit.  A forced pass (``possible_actions == [64]``) is checked not to advance MT19937
  and is not part of the stream.
* ``ar_*``  -- the arena: ``eval._run_one_match`` / ``eval.play_match`` (eval.py:86-178) on stub policies that are
  rebuilt from ``(class, config, state_dict)`` exactly as the workers do.  Stored per match: the action of every
  ply, the tie pick of every ply (``np.random.choice(best_actions)``, MCTS_model.py:249-255), and the result string
  the reference returned -- including odd match indices (roles swapped, result inverted) and a drawn game.

Nothing from the reference is copied; the hash stub below is the numpy twin of oracle/othello_oracle.c orc_stub_h.
"""
import os
import sys

import numpy as np

REF = os.environ.get("OTHELLO_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from envs.othello import OthelloGameNew  # noqa: E402
import MCTS_model  # noqa: E402
from MCTS_model import MCTS  # noqa: E402
import eval as ref_eval  # noqa: E402
from make_golden import StubH  # noqa: E402  (numpy twin of the oracle's / the engine's hash stub)

OUT = {}


# ------------------------------------------------------------- M13 rollouts
class PickTap:
    """Logs the index returned by every np.random.choice(possible_actions) call (no ``p``)."""

    def __init__(self):
        self.picks, self.forced_pass_calls = [], 0

    def __enter__(self):
        self.o_choice = np.random.choice
        tap = self

        def choice(a, size=None, replace=True, p=None):
            assert p is None and size is None
            arr = [int(x) for x in np.asarray(a)]
            st = np.random.get_state()
            r = tap.o_choice(a)
            if arr == [64]:
                st2 = np.random.get_state()
                assert st[2] == st2[2] and np.array_equal(st[1], st2[1]), "a forced pass must not advance MT19937"
                tap.forced_pass_calls += 1
            else:
                tap.picks.append(arr.index(int(r)))
            return r

        np.random.choice = choice
        return self

    def __exit__(self, *exc):
        np.random.choice = self.o_choice


def gen_rollout():
    g = OthelloGameNew(8)
    cases = [("ro0", 60, 1.4, 6, 11), ("ro1", 25, 2.0, 40, 12), ("ro2", 120, 0.8, 4, 13)]  # name, sims, c_puct, moves, seed
    for name, sims, c, n_moves, seed in cases:
        np.random.seed(seed)
        m = MCTS(g, {"c_puct": c, "num_simulations": sims, "num_threads": 1}, None)
        assert m.use_rollout
        s, pl = g.get_initial_state(), 1
        rec = dict(counts=[], cval=[], root_value=[], root_n=[], action=[], player=[], state=[], picks_after=[])
        with PickTap() as tap:
            for mv in range(n_moves):
                probs = m.policy_improve_step(s, pl, temp=1.0)
                counts = np.zeros(65, np.int32)
                cval = np.zeros(65)
                for a, ch in m.root.children.items():
                    counts[a] = ch.visit_count
                    cval[a] = ch.value
                rec["state"].append(s.copy()); rec["player"].append(pl)
                rec["counts"].append(counts); rec["cval"].append(cval)
                rec["root_value"].append(m.root.value); rec["root_n"].append(m.root.visit_count)
                rec["picks_after"].append(len(tap.picks))
                a = int(np.argmax(probs))
                rec["action"].append(a)
                m.make_move(a)
                s = g.get_next_state(s, a, pl)
                _, done = g.get_value_and_terminated(s, a, pl)
                if done:
                    break
                pl = -pl
        OUT[f"{name}_cfg"] = np.array([sims, c], np.float64)
        OUT[f"{name}_picks"] = np.array(tap.picks, np.int32)
        for k, v in rec.items():
            OUT[f"{name}_{k}"] = np.array(v)
        print(name, "moves", len(rec["action"]), "picks", len(tap.picks), "forced-pass calls", tap.forced_pass_calls,
              "root values", [round(x, 3) for x in rec["root_value"][:4]])


# -------------------------------------------------------------------- arena
class ArenaStub(StubH):
    """A policy the reference's workers can rebuild: ``policy_class(**policy_config)`` +
    ``load_state_dict`` + ``eval`` (eval.py:97-109), evaluating with the hash stub of salt ``salt``."""

    def __init__(self, salt):
        super().__init__(salt)

    def load_state_dict(self, sd):
        assert sd == {}

    def eval(self):
        return self


class MatchTap:
    """Logs, for the match being played, the action of every ply (arg-max of what policy_improve_step returned)
    and the tie pick np.random.choice(best_actions) made inside it."""

    def __init__(self):
        self.actions, self.u_tie = [], []

    def __enter__(self):
        self.o_pis, self.o_choice = MCTS.policy_improve_step, np.random.choice
        tap = self

        def choice(a, size=None, replace=True, p=None):
            assert p is None
            r = tap.o_choice(a)
            arr = [int(x) for x in np.asarray(a)]
            tap._tie = (arr.index(int(r)) + 0.5) / len(arr)
            return r

        def pis(self_, init_state, init_player, temp=1):
            tap._tie = 0.0
            probs = tap.o_pis(self_, init_state=init_state, init_player=init_player, temp=temp)
            tap.actions.append(int(np.argmax(probs)))
            tap.u_tie.append(tap._tie)
            return probs

        MCTS.policy_improve_step, np.random.choice = pis, choice
        return self

    def __exit__(self, *exc):
        MCTS.policy_improve_step, np.random.choice = self.o_pis, self.o_choice


def run_match(idx, args, salt_a, salt_b, seed):
    np.random.seed(seed)
    with MatchTap() as tap:
        res = ref_eval._run_one_match((idx, 8, args, (ArenaStub, {"salt": salt_a}, {}), (ArenaStub, {"salt": salt_b}, {})))
    return res, tap.actions, tap.u_tie


def gen_arena():
    args = {"c_puct": 2.0, "num_simulations": 12, "num_threads": 1}
    matches = []
    # ten matches with index = position (even: candidate plays +1, odd: roles swapped, result inverted) ...
    for i in range(10):
        matches.append((i, 100 + i, 200 + i, 5000 + i))
    # ... plus the first drawn game found by scanning salts (both parities are tried)
    found = None
    for k in range(400):
        idx = k % 2
        res, acts, _ = run_match(idx, args, 1000 + k, 3000 + k, 7000 + k)
        if res == "Draw":
            found = (idx, 1000 + k, 3000 + k, 7000 + k)
            break
    assert found is not None, "no drawn game found"
    # the list position is the match index the batched arena will use (only its parity matters, eval.py:115-126)
    if len(matches) % 2 != found[0] % 2:
        matches.append((len(matches), 150, 250, 5050))
    matches.append((len(matches),) + found[1:])
    T = 128
    n = len(matches)
    acts = np.full((n, T), -1, np.int32)
    uts = np.zeros((n, T))
    res_l, meta = [], []
    for j, (idx, sa, sb, seed) in enumerate(matches):
        res, a, u = run_match(idx, args, sa, sb, seed)
        acts[j, :len(a)] = a
        uts[j, :len(u)] = u
        res_l.append(res)
        meta.append((idx, sa, sb, seed, len(a)))
        print("match", j, "index", idx, "salts", sa, sb, "plies", len(a), "->", res)
    assert "Draw" in res_l and "A" in res_l and "B" in res_l
    # play_match's own mapping (first tree = +1), without the worker wrapper
    env = OthelloGameNew(8)
    np.random.seed(99)
    first = MCTS(env, args, ArenaStub(31))
    second = MCTS(env, args, ArenaStub(32))
    with MatchTap() as tap:
        direct = ref_eval.play_match(env, first, second)
    OUT.update(ar_cfg=np.array([args["num_simulations"], args["c_puct"]], np.float64), ar_actions=acts, ar_u_tie=uts,
               ar_results=np.array(res_l), ar_meta=np.array(meta, np.int64),
               ar_direct_result=np.array([direct]), ar_direct_actions=np.array(tap.actions, np.int32),
               ar_direct_u_tie=np.array(tap.u_tie), ar_direct_salts=np.array([31, 32], np.int64))
    print("direct play_match ->", direct, len(tap.actions), "plies")


if __name__ == "__main__":
    gen_rollout()
    gen_arena()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_r2.npz")
    np.savez_compressed(path, **OUT)
    print("wrote", path, os.path.getsize(path), "bytes")
