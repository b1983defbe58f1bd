"""Differential pinning of the oracle against the LIVE reference (dev container only).

tests/golden/ holds results of the reference that travel to the GPU box; here, where
/root/reference exists, the oracle is additionally driven side by side with the reference's own
classes on fresh random inputs (other seeds, stubs and hyper-parameters than the goldens).  Skipped
wherever the reference is absent.  Nothing is copied: the reference is imported and called.
"""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("OTHELLO_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "envs", "othello.py")),
                                reason="reference checkout not present (GPU box)")

M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    added = REF not in sys.path
    if added:
        sys.path.insert(0, REF)
    mine = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "envs" or k.startswith("envs.")}  # not this repo's envs
    try:
        from envs.othello import OthelloGameNew
        from MCTS_model import MCTS
        yield {"Game": OthelloGameNew, "MCTS": MCTS}
    finally:
        for k in [k for k in sys.modules if k == "envs" or k.startswith("envs.") or k == "MCTS_model"]:
            sys.modules.pop(k)
        sys.modules.update(mine)
        if added:
            sys.path.remove(REF)


def _mix(x):
    x &= M64
    x ^= x >> 31
    x = (x * 0x9E3779B97F4A7C15) & M64
    x ^= x >> 29
    return x


class HashStub:
    """Deterministic policy: float32 priors and a value in [-1, 1] from a hash of (player*state, salt)."""

    def __init__(self, salt):
        self.salt = salt

    def inference(self, state, player):
        h = _mix(int.from_bytes((np.asarray(state) * int(player)).astype(np.int8).tobytes(), "little") % M64 ^ self.salt)
        raw = np.array([1 + (_mix(h + 977 * a) % 97) for a in range(65)], dtype=np.float32)
        return raw / raw.sum(), float(np.float32(int(_mix(h ^ 0xABCDEF) % 401) - 200) / np.float32(200))


@pytest.mark.parametrize("seed", range(6))
def test_env_random_games_side_by_side(ref, seed):
    """Ten random games per seed: masks, next states, terminal tests (both players), scores, illegal moves."""
    import oracle as O
    g, o = ref["Game"](8), O.OracleGame()
    rs = np.random.RandomState(1000 + seed)
    for _ in range(10):
        s, player = g.get_initial_state(), 1
        assert np.array_equal(s, o.get_initial_state())
        for ply in range(130):
            m = g.get_valid_moves(s, player)
            assert np.array_equal(m, o.get_valid_moves(s, player))
            for p in (player, -player):
                assert g.get_value_and_terminated(s, None, p) == o.get_value_and_terminated(s, None, p)
                assert g.get_score(s, p) == o.get_score(s, p)
            if ply % 7 == 0:  # illegal actions raise in both
                for a in rs.choice(np.nonzero(m[:64] == 0)[0], 3):
                    with pytest.raises(ValueError):
                        g.get_next_state(s, int(a), player)
                    with pytest.raises(ValueError):
                        o.get_next_state(s, int(a), player)
            a = int(rs.choice(np.nonzero(m)[0]))
            ns = g.get_next_state(s, a, player)
            assert np.array_equal(ns, o.get_next_state(s, a, player)) and ns.dtype == np.int8
            s = ns
            if g.get_value_and_terminated(s, a, player)[1]:
                break
            player = -player
        else:
            raise AssertionError("game did not end")


CASES = [  # salt, sims, c_puct, eps, alpha, temp, moves
    (1, 30, 2.0, 0.0, 1.0, 1.0, 6),
    (2, 57, 1.0, 0.3, 1.0, 1.0, 5),
    (3, 25, 3.5, 0.25, 0.3, 0.5, 8),
    (4, 80, 2.0, 0.3, 0.03, 1.0, 4),
    (5, 16, 0.7, 0.5, 1.0, 2.0, 10),
    (6, 40, 2.0, 0.3, 1.0, 0.0, 8),
]


@pytest.mark.parametrize("salt,sims,c_puct,eps,alpha,temp,moves", CASES)
def test_mcts_side_by_side(ref, monkeypatch, salt, sims, c_puct, eps, alpha, temp, moves):
    """Reference MCTS (num_threads=1) and the oracle on the same stub, the same Dirichlet draws and the same
    moves, with tree re-use: child visit counts, child values / priors, root value / N bit for bit; policy
    targets bit for bit at temp 1 and within 1e-6 otherwise."""
    import oracle as O
    stub = HashStub(salt)
    g = ref["Game"](8)
    drawn, ties = [], []
    real_dirichlet, real_choice = np.random.dirichlet, np.random.choice

    def dirichlet(a, *k, **kw):
        drawn.append(real_dirichlet(a, *k, **kw))
        return drawn[-1]

    def choice(a, *k, **kw):  # tie pick at temp ~ 0 (MCTS_model.py:249-255): take the first, tell the oracle u = 0
        if "p" not in kw and not k and np.ndim(a) == 1:
            ties.append(len(a))
            return a[0]
        return real_choice(a, *k, **kw)

    monkeypatch.setattr(np.random, "dirichlet", dirichlet)
    monkeypatch.setattr(np.random, "choice", choice)
    np.random.seed(salt)
    m = ref["MCTS"](g, {"c_puct": c_puct, "num_simulations": sims, "num_threads": 1}, stub, dirichlet_alpha=alpha,
                    dirichlet_epsilon=eps)
    om = O.OracleMCTS(c_puct, sims, O.Evaluator(fn=stub.inference), dirichlet_epsilon=eps)
    s, player = g.get_initial_state(), 1
    rs = np.random.RandomState(77 + salt)
    for mv in range(moves):
        n0 = len(drawn)
        probs = m.policy_improve_step(s, player, temp=temp)
        noise = drawn[-1] if len(drawn) > n0 else None
        oprobs = om.policy_improve_step(s, player, temp=temp, noise=noise, u_tie=0.0)
        st = om.root_stats()
        counts = np.zeros(65, np.int32)
        cval, cpri = np.zeros(65), np.zeros(65)
        for a, ch in m.root.children.items():
            counts[a], cval[a], cpri[a] = ch.visit_count, ch.value, float(ch.prior)
        assert np.array_equal(counts, st["counts"]), mv
        assert np.array_equal(cval, st["child_value"]) and np.array_equal(cpri, st["child_prior"]), mv
        assert m.root.value == st["root_value"] and m.root.visit_count == st["root_n"], mv
        assert probs.dtype == np.float32
        if temp == 1.0 or temp == 0.0:
            assert np.array_equal(probs, oprobs), mv
        else:
            assert np.abs(probs - oprobs).max() <= 1e-6, mv
        a = int(rs.choice(65, p=probs.astype(np.float64) / probs.astype(np.float64).sum()))
        m.make_move(a)
        om.make_move(a)
        s = g.get_next_state(s, a, player)
        if g.get_value_and_terminated(s, a, player)[1]:
            break
        player = -player


class _RngTap:
    """Logs what the reference draws from np.random (dirichlet noise, the uniform behind choice(65, p), tie picks)."""

    def __init__(self, monkeypatch):
        self.noise, self.u_move, self.tie = [], [], []
        real_dirichlet, real_choice = np.random.dirichlet, np.random.choice
        tap = self

        def dirichlet(alpha, size=None):
            tap.noise.append(np.array(real_dirichlet(alpha, size), np.float64))
            return tap.noise[-1]

        def choice(a, size=None, replace=True, p=None):
            if p is None:
                r = real_choice(a)
                arr = list(np.asarray(a))
                tap.tie.append((arr.index(r), len(arr)))
                return r
            shadow = np.random.RandomState()
            shadow.set_state(np.random.get_state())
            r = real_choice(a, size, replace, p)
            tap.u_move.append(shadow.random_sample())
            return r

        monkeypatch.setattr(np.random, "dirichlet", dirichlet)
        monkeypatch.setattr(np.random, "choice", choice)


class _SelfPlayStub(HashStub):
    """HashStub with the surface one_self_play needs to rebuild a policy (self_play_worker.py:47-51)."""

    def load_state_dict(self, sd):
        pass

    def eval(self):
        pass


SP_CASES = [  # salt, sims, c_puct, eps, alpha, temp, exploratory moves, lambda
    (31, 12, 2.0, 0.3, 1.0, 1.0, 35, 0.98),
    (32, 20, 1.2, 0.0, 1.0, 1.0, 6, 0.7),
    (33, 9, 3.0, 0.5, 0.3, 1.0, 0, 1.0),
    (34, 15, 2.0, 0.25, 1.0, 0.5, 12, 0.9),
]


@pytest.mark.parametrize("salt,sims,c_puct,eps,alpha,temp,nexp,lam", SP_CASES)
def test_one_self_play_side_by_side(ref, monkeypatch, salt, sims, c_puct, eps, alpha, temp, nexp, lam):
    """The reference's one_self_play on a fresh seed, its random draws tapped and handed to the oracle's self_play:
    same trajectory length, canonical states, value targets (lambda-returns) bit for bit; policy targets bit for bit at
    temperature 1 / 0 and within 1e-6 otherwise."""
    import oracle as O
    import self_play_worker as ref_worker
    args = {"c_puct": c_puct, "num_simulations": sims, "num_threads": 1, "dirichlet_alpha": alpha, "dirichlet_epsilon": eps,
            "mcts_temperature": temp, "num_exploratory_moves": nexp, "lambda": lam}
    tap = _RngTap(monkeypatch)
    np.random.seed(500 + salt)
    traj = ref_worker.one_self_play((8, args, (_SelfPlayStub, {"salt": salt}, {}), None))
    T = len(traj)
    assert len(tap.u_move) == T and len(tap.noise) <= 1
    # np.random.choice(best actions) is called once per temp~0 ply (MCTS_model.py:249-255), in order
    u_tie, ties = np.zeros(128), list(tap.tie)
    for t in range(T):
        if (temp if t < nexp else 0.0) < 0.1:
            k, n = ties.pop(0)
            u_tie[t] = (k + 0.5) / n
    assert not ties
    u_move = np.zeros(128)
    u_move[:T] = tap.u_move
    stub = _SelfPlayStub(salt)
    out = O.self_play(args, O.Evaluator(fn=stub.inference), tap.noise[0] if tap.noise else None, u_move, u_tie)
    assert len(out["values"]) == T
    assert np.array_equal(np.stack([t[0] for t in traj]), out["states"])
    pis = np.stack([t[1] for t in traj])
    if temp == 1.0:
        assert np.array_equal(pis, out["pis"])
    else:
        assert np.abs(pis - out["pis"]).max() <= 1e-6
    assert np.array_equal(np.array([t[2] for t in traj], np.float64), out["values"])


@pytest.mark.parametrize("seed", range(3))
def test_replay_aggregation_side_by_side(ref, seed):
    """Trainer._aggregate_duplicates (train.py:142-173, called as an unbound method on a stand-in object) against the
    oracle restatement on fresh random buffers with heavy duplication and mixed model versions."""
    import collections
    import train as ref_train
    from oracle import replay as OR

    class FakeTrainer:
        _hash_state = ref_train.Trainer._hash_state

    rs = np.random.RandomState(900 + seed)
    uniq = rs.randint(-1, 2, size=(50, 8, 8)).astype(np.int8)
    n = 700
    buf = [(uniq[rs.randint(0, 50)].copy(), rs.rand(65).astype(np.float32), float(rs.uniform(-1, 1)), int(rs.randint(0, 3)))
           for _ in range(n)]
    t = FakeTrainer()
    t.replay_buffer = collections.deque(buf)
    states, policies, values = ref_train.Trainer._aggregate_duplicates(t)
    o_states, o_policies, o_values, counts, _vers = OR.aggregate_duplicates(buf)
    assert len(states) == len(o_states) and sum(counts) == n
    assert np.array_equal(np.stack(states), np.stack(o_states))
    assert np.array_equal(np.stack(policies).astype(np.float32), np.stack(o_policies))
    assert np.array_equal(np.array(values, np.float32), np.array(o_values, np.float32))


@pytest.mark.parametrize("seed,sims,c_puct,moves", [(21, 40, 1.4, 5), (22, 15, 3.0, 30), (23, 90, 1.0, 3)])
def test_rollout_mode_side_by_side(ref, monkeypatch, seed, sims, c_puct, moves):
    """policy=None (MCTS_model.py:276-303, 332-335) live: the reference's playouts draw np.random.choice(possible_actions);
    the returned indices are tapped and handed to the oracle, which must then build the identical tree -- uniform
    np.ones(65) priors masked and normalised, playout value seen from the leaf's side to move, forced passes drawing
    nothing -- over several moves with tree re-use."""
    import oracle as O
    g = ref["Game"](8)
    real_choice = np.random.choice
    picks = []

    def choice(a, *k, **kw):
        r = real_choice(a, *k, **kw)
        arr = [int(x) for x in np.asarray(a)]
        if arr != [64]:
            picks.append(arr.index(int(r)))
        return r

    monkeypatch.setattr(np.random, "choice", choice)
    np.random.seed(seed)
    m = ref["MCTS"](g, {"c_puct": c_puct, "num_simulations": sims, "num_threads": 1}, None)
    s, player = g.get_initial_state(), 1
    per_move = []
    for mv in range(moves):
        probs = m.policy_improve_step(s, player, temp=1.0)
        counts = np.zeros(65, np.int32)
        cval = np.zeros(65)
        for a, ch in m.root.children.items():
            counts[a], cval[a] = ch.visit_count, ch.value
        a = int(np.argmax(probs))
        per_move.append((s.copy(), player, counts, cval, m.root.value, m.root.visit_count, len(picks), a, probs.copy()))
        m.make_move(a)
        s = g.get_next_state(s, a, player)
        if g.get_value_and_terminated(s, a, player)[1]:
            break
        player = -player
    om = O.OracleMCTS(c_puct, sims, None, picks=np.array(picks, np.int32))
    for s, player, counts, cval, rv, rn, used, a, probs in per_move:
        assert np.array_equal(om.policy_improve_step(s, player, temp=1.0), probs)
        r = om.root_stats()
        assert np.array_equal(r["counts"], counts) and np.array_equal(r["child_value"], cval)
        assert r["root_value"] == rv and r["root_n"] == rn and om.picks_used == used
        om.make_move(a)
    assert om.picks_used == len(picks)
