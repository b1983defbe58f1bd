"""Generate tests/golden/ref_golden.npz by RUNNING THE REFERENCE ITSELF.

Run in the dev container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Nothing from the reference is copied: this script imports
/root/reference/{envs/othello.py, MCTS_model.py, self_play_worker.py}, drives
them through their public API with deterministic stub policies, and stores the
RESULTS (masks, states, visit counts, policy targets, value targets) together
with the random draws the reference consumed (captured by wrapping
np.random.dirichlet / np.random.choice), so the C oracle and the CUDA engine
can be replayed on identical inputs.

The stub policies are defined here in numpy and, by construction, produce
bit-identical float32 outputs to oracle/othello_oracle.c's orc_stub_* and the
CUDA engine's device stubs.
"""
import os
import random
import sys

import numpy as np

REF = os.environ.get("OTHELLO_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from envs.othello import OthelloGameNew, OthelloGame, get_random_symmetry  # noqa: E402
from MCTS_model import MCTS  # noqa: E402
import self_play_worker  # noqa: E402

M64 = (1 << 64) - 1
OUT = {}


# ------------------------------------------------------------ stub policies
def mix64(x):
    x &= M64
    x ^= x >> 33
    x = (x * 0xff51afd7ed558ccd) & M64
    x ^= x >> 33
    x = (x * 0xc4ceb9fe1a85ec53) & M64
    x ^= x >> 33
    return x


def canon_bits(state, player):
    own = opp = 0
    flat = (np.asarray(state).astype(np.int64) * int(player)).ravel()
    for i, v in enumerate(flat):
        if v == 1:
            own |= 1 << (63 - i)
        elif v == -1:
            opp |= 1 << (63 - i)
    return own, opp


class StubA:
    def inference(self, state, player):
        return np.full(65, 1.0 / 65.0, dtype=np.float32), 0.0


class StubB:
    W = (np.arange(64, dtype=np.int64) + 1).reshape(8, 8)

    def inference(self, state, player):
        c = (player * np.asarray(state)).astype(np.int64)
        h = int((c * self.W).sum())
        raw = np.array([((7 * a + h) % 11) + 1 for a in range(65)], dtype=np.float32)
        return raw / raw.sum(), ((h % 17) - 8) / 16.0


class StubH:
    def __init__(self, salt=0):
        self.salt = salt

    def inference(self, state, player):
        own, opp = canon_bits(state, player)
        h = mix64(own ^ mix64(opp ^ self.salt))
        raw = np.array([1 + (mix64(h + a) % 251) for a in range(65)], dtype=np.float32)
        pri = raw / raw.sum()
        assert pri.dtype == np.float32
        v = int(mix64(h ^ 0x9e3779b97f4a7c15) % 2001) - 1000
        return pri, float(np.float32(v) / np.float32(1000))

    # surface one_self_play needs (self_play_worker.py:47-51)
    def load_state_dict(self, sd):
        pass

    def eval(self):
        pass


# ------------------------------------------------------------------- env ---
def gen_env():
    g = OthelloGameNew(8)
    old = OthelloGame(8)

    # perft 1..6 through get_valid_moves / get_next_state
    def perft(state, player, depth):
        if depth == 0:
            return 1
        v, t = g.get_value_and_terminated(state, None, player)
        if t:
            return 1
        n = 0
        for a in np.nonzero(g.get_valid_moves(state, player))[0]:
            n += perft(g.get_next_state(state, int(a), player), -player, depth - 1)
        return n

    OUT["env_perft"] = np.array([perft(g.get_initial_state(), 1, d) for d in range(1, 7)], np.int64)

    # games: game 0 = lowest-index playout, games 1.. = seeded random
    def play(picker):
        s = g.get_initial_state()
        player = 1
        acts, masks, states, vals, terms, players = [], [], [], [], [], []
        while True:
            m = g.get_valid_moves(s, player)
            assert np.array_equal(m, old.get_valid_moves(s, player))
            a = picker(np.nonzero(m)[0])
            ns = g.get_next_state(s, a, player)
            assert np.array_equal(ns, old.get_next_state(s, a, player))
            v, t = g.get_value_and_terminated(ns, a, player)
            assert (v, t) == old.get_value_and_terminated(ns, a, player)
            acts.append(a); masks.append(m); states.append(ns); vals.append(v); terms.append(t); players.append(player)
            s = ns
            if t:
                break
            player = -player
        return acts, masks, states, vals, terms, players, g.get_score(s, 1)

    games = [play(lambda legal: int(legal[0]))]
    for seed in range(15):
        rng = random.Random(0xC0FFEE + seed)
        games.append(play(lambda legal: int(rng.choice(list(legal)))))
    T = max(len(x[0]) for x in games)
    n = len(games)
    A = np.full((n, T), -1, np.int32)
    Mk = np.zeros((n, T, 65), np.uint8)
    S = np.zeros((n, T, 8, 8), np.int8)
    V = np.zeros((n, T), np.int8)
    Tm = np.zeros((n, T), np.uint8)
    P = np.zeros((n, T), np.int8)
    L = np.zeros(n, np.int32)
    SC = np.zeros(n, np.int32)
    for i, (acts, masks, states, vals, terms, players, sc) in enumerate(games):
        k = len(acts)
        L[i] = k; SC[i] = sc
        A[i, :k] = acts; Mk[i, :k] = masks; S[i, :k] = states; V[i, :k] = vals; Tm[i, :k] = terms; P[i, :k] = players
    OUT.update(env_game_len=L, env_game_actions=A, env_game_masks=Mk, env_game_states=S, env_game_values=V,
               env_game_terms=Tm, env_game_players=P, env_game_scores=SC)

    # random (mostly unreachable) boards: masks + terminal test for both sides,
    # next state for every legal move of +1 and -1 (checks wrap-around guards)
    rs = np.random.RandomState(1234)
    boards = []
    for i in range(300):
        fill = rs.uniform(0.1, 1.0)
        b = rs.choice([-1, 0, 1], size=(8, 8), p=[fill / 2, 1 - fill, fill / 2]).astype(np.int8)
        boards.append(b)
    # hand-made edge cases: full boards, single colour, empty
    boards.append(np.ones((8, 8), np.int8)); boards.append(-np.ones((8, 8), np.int8))
    boards.append(np.zeros((8, 8), np.int8))
    e = np.ones((8, 8), np.int8); e[0, :] = -1; e[7, 7] = 0; boards.append(e)
    e = np.zeros((8, 8), np.int8); e[0, 0:7] = [1, -1, -1, -1, -1, -1, -1]; boards.append(e)  # 6-disc run to the edge
    e = np.zeros((8, 8), np.int8); e[0, :] = [0, -1, -1, -1, -1, -1, -1, 1]; boards.append(e)
    e = np.zeros((8, 8), np.int8); e[3, 7] = 1; e[4, 0] = -1; e[4, 1] = 0; boards.append(e)  # E/W wrap guard
    B = np.stack(boards)
    nb = len(B)
    masks = np.zeros((nb, 2, 65), np.uint8)
    vt = np.zeros((nb, 2, 2), np.int8)
    nxt = np.zeros((nb, 2, 65, 8, 8), np.int8)
    scores = np.zeros((nb, 2), np.int32)
    for i, b in enumerate(B):
        for j, pl in enumerate((1, -1)):
            m = g.get_valid_moves(b, pl)
            masks[i, j] = m
            v, t = g.get_value_and_terminated(b, None, pl)
            vt[i, j] = (v, t)
            scores[i, j] = g.get_score(b, pl)
            for a in np.nonzero(m)[0]:
                nxt[i, j, a] = g.get_next_state(b, int(a), pl)
            # every non-legal board action must raise ValueError
            for a in range(64):
                if not m[a]:
                    try:
                        g.get_next_state(b, a, pl)
                        raise SystemExit("reference accepted an illegal move")
                    except ValueError:
                        pass
    OUT.update(env_rand_boards=B, env_rand_masks=masks, env_rand_vt=vt, env_rand_next=nxt, env_rand_scores=scores)

    # symmetries
    ns = 24
    sb = B[rs.choice(300, ns, replace=False)]
    spi = rs.rand(ns, 65).astype(np.float32)
    sym_s = np.zeros((ns, 8, 8, 8), np.float32)
    sym_pi = np.zeros((ns, 8, 65), np.float32)
    rnd_k = np.zeros(ns, np.int32); rnd_f = np.zeros(ns, np.uint8)
    rnd_s = np.zeros((ns, 1, 8, 8), np.float32); rnd_pi = np.zeros((ns, 65), np.float32)
    for i in range(ns):
        for j, (b2, p2) in enumerate(old.get_symmetries(sb[i], spi[i])):
            sym_s[i, j] = b2
            sym_pi[i, j] = np.asarray(p2, np.float32)
        np.random.seed(1000 + i)
        sh = np.random.RandomState(1000 + i)
        rnd_k[i] = sh.randint(4)
        rnd_f[i] = sh.rand() < 0.5
        s2, p2 = get_random_symmetry(sb[i], spi[i])
        assert s2.dtype == np.float32 and p2.dtype == np.float32
        rnd_s[i] = s2; rnd_pi[i] = p2
    OUT.update(sym_boards=sb, sym_pi=spi, sym_all_s=sym_s, sym_all_pi=sym_pi, sym_rnd_k=rnd_k, sym_rnd_flip=rnd_f,
               sym_rnd_s=rnd_s, sym_rnd_pi=rnd_pi)


# ------------------------------------------------------------------ MCTS ---
class RngTap:
    """Wraps np.random.dirichlet / choice to log what the reference consumed."""

    def __init__(self):
        self.noise, self.u_move, self.tie = [], [], []

    def __enter__(self):
        self.o_dir, self.o_choice = np.random.dirichlet, np.random.choice
        tap = self

        def dirichlet(alpha, size=None):
            r = tap.o_dir(alpha, size)
            tap.noise.append(np.array(r, np.float64))
            return r

        def choice(a, size=None, replace=True, p=None):
            if p is not None:
                st = np.random.get_state()
                r = tap.o_choice(a, size, replace, p)
                sh = np.random.RandomState()
                sh.set_state(st)
                u = sh.random_sample()
                p64 = np.array(p, dtype=np.float64)
                cdf = p64.cumsum()
                cdf /= cdf[-1]
                assert int(cdf.searchsorted(u, side="right")) == int(r)
                tap.u_move.append(u)
                return r
            r = tap.o_choice(a)
            arr = list(np.asarray(a))
            tap.tie.append((arr.index(r), len(arr)))
            return r

        np.random.dirichlet, np.random.choice = dirichlet, choice
        return self

    def __exit__(self, *exc):
        np.random.dirichlet, np.random.choice = self.o_dir, self.o_choice


def gen_mcts():
    g = OthelloGameNew(8)
    cases = [  # name, stub id, policy, sims, c_puct, eps, alpha, temp, n_moves, move rule, seed
        ("A", 0, StubA(), 100, 2.0, 0.0, 1.0, 1.0, 3, "argmax", 1),
        ("B", 1, StubB(), 100, 2.0, 0.0, 1.0, 1.0, 3, "argmax", 2),
        ("H_noise", 2, StubH(0), 200, 2.0, 0.3, 1.0, 1.0, 10, "sample", 3),
        ("H_t0", 2, StubH(7), 64, 3.0, 0.0, 1.0, 0.0, 200, "sample", 4),
        ("H_alpha", 2, StubH(11), 150, 1.25, 0.25, 0.3, 0.5, 12, "sample", 5),
        ("B_long", 1, StubB(), 40, 2.0, 0.3, 1.0, 1.0, 200, "sample", 6),
    ]
    names = []
    for name, sid, pol, sims, c, eps, alpha, temp, nm, rule, seed in cases:
        np.random.seed(seed)
        m = MCTS(g, {"c_puct": c, "num_simulations": sims, "num_threads": 1}, pol, dirichlet_alpha=alpha,
                 dirichlet_epsilon=eps)
        s = g.get_initial_state(); pl = 1
        rec = dict(counts=[], root_value=[], root_n=[], cval=[], cpri=[], probs=[], action=[], player=[], state=[],
                   u_tie=[], noise_used=[])
        noise_all = []
        with RngTap() as tap:
            for mv in range(nm):
                n_noise0, n_tie0 = len(tap.noise), len(tap.tie)
                probs = m.policy_improve_step(s, pl, temp=temp)
                assert probs.dtype == np.float32
                counts = np.zeros(65, np.int32); cval = np.zeros(65); cpri = np.zeros(65)
                for a, ch in m.root.children.items():
                    counts[a] = ch.visit_count; cval[a] = ch.value; cpri[a] = float(ch.prior)
                rec["state"].append(s.copy()); rec["player"].append(pl)
                rec["counts"].append(counts); rec["cval"].append(cval); rec["cpri"].append(cpri)
                rec["root_value"].append(m.root.value); rec["root_n"].append(m.root.visit_count)
                rec["probs"].append(probs.copy())
                if len(tap.tie) > n_tie0:
                    k, n = tap.tie[-1]
                    rec["u_tie"].append((k + 0.5) / n)
                else:
                    rec["u_tie"].append(0.0)
                if len(tap.noise) > n_noise0:
                    rec["noise_used"].append(1); noise_all.append(tap.noise[-1])
                else:
                    rec["noise_used"].append(0); noise_all.append(np.zeros(65))
                if rule == "argmax":
                    a = int(np.argmax(probs))
                else:
                    a = int(np.random.RandomState(seed * 1000 + mv).choice(65, p=probs / probs.sum()))
                rec["action"].append(a)
                m.make_move(a)
                s = g.get_next_state(s, a, pl)
                v, t = g.get_value_and_terminated(s, a, pl)
                if t:
                    break
                pl = -pl
        pre = f"mcts_{name}_"
        OUT[pre + "cfg"] = np.array([sid, sims, c, eps, alpha, temp, getattr(pol, "salt", 0)], np.float64)
        OUT[pre + "state"] = np.stack(rec["state"]).astype(np.int8)
        OUT[pre + "player"] = np.array(rec["player"], np.int8)
        OUT[pre + "counts"] = np.stack(rec["counts"])
        OUT[pre + "cval"] = np.stack(rec["cval"]); OUT[pre + "cpri"] = np.stack(rec["cpri"])
        OUT[pre + "root_value"] = np.array(rec["root_value"], np.float64)
        OUT[pre + "root_n"] = np.array(rec["root_n"], np.int64)
        OUT[pre + "probs"] = np.stack(rec["probs"])
        OUT[pre + "action"] = np.array(rec["action"], np.int32)
        OUT[pre + "u_tie"] = np.array(rec["u_tie"], np.float64)
        OUT[pre + "noise_used"] = np.array(rec["noise_used"], np.uint8)
        OUT[pre + "noise"] = np.stack(noise_all)
        names.append(name)
        print(name, "moves", len(rec["action"]), "root_n", rec["root_n"][:3], "noise plies", sum(rec["noise_used"]))
    OUT["mcts_cases"] = np.array(names)


# ------------------------------------------------------------- self-play ---
def gen_selfplay():
    cases = [  # name, salt, sims, c_puct, eps, alpha, temp, exploratory, lambda, seed
        ("sp0", 21, 25, 2.0, 0.3, 1.0, 1.0, 35, 0.98, 10),
        ("sp1", 22, 50, 2.0, 0.3, 1.0, 1.0, 8, 0.98, 11),
        ("sp2", 23, 16, 1.5, 0.0, 1.0, 1.0, 4, 0.5, 12),
        ("sp3", 24, 30, 2.5, 0.25, 0.5, 0.8, 20, 1.0, 13),
    ]
    names = []
    for name, salt, sims, c, eps, alpha, temp, nexp, lam, seed in cases:
        args = {"c_puct": c, "num_simulations": sims, "num_threads": 1, "dirichlet_alpha": alpha,
                "dirichlet_epsilon": eps, "mcts_temperature": temp, "num_exploratory_moves": nexp, "lambda": lam}
        np.random.seed(seed)
        with RngTap() as tap:
            traj = self_play_worker.one_self_play((8, args, (StubH, {"salt": salt}, {}), None))
        T = len(traj)
        assert len(tap.u_move) == T
        # one tie entry per temp~0 ply, in order; map to per-ply u_tie
        u_tie = np.zeros(T)
        ties = list(tap.tie)
        for t in range(T):
            if (temp if t < nexp else 0.0) < 0.1:
                k, n = ties.pop(0)
                u_tie[t] = (k + 0.5) / n
        assert not ties
        pre = f"sp_{name}_"
        OUT[pre + "cfg"] = np.array([salt, sims, c, eps, alpha, temp, nexp, lam], np.float64)
        OUT[pre + "states"] = np.stack([t[0] for t in traj]).astype(np.int8)
        OUT[pre + "pis"] = np.stack([t[1] for t in traj]).astype(np.float32)
        OUT[pre + "values"] = np.array([t[2] for t in traj], np.float64)
        OUT[pre + "noise"] = tap.noise[0] if tap.noise else np.zeros(65)
        assert len(tap.noise) <= 1
        OUT[pre + "u_move"] = np.array(tap.u_move, np.float64)
        OUT[pre + "u_tie"] = u_tie
        names.append(name)
        print(name, "plies", T, "G0", traj[0][2])
    OUT["sp_cases"] = np.array(names)


# ---------------------------------------------------------- replay ingest ---
def gen_replay():
    """Trainer._aggregate_duplicates (train.py:142-173) run on a replay buffer built from the
    self-play fixtures above (shared openings => real duplicates), via the unbound method."""
    import collections
    import train as ref_train

    class FakeTrainer:
        _hash_state = ref_train.Trainer._hash_state

    buf = []
    order = [("sp0", 0), ("sp1", 1), ("sp2", 0), ("sp3", 1), ("sp0", 0), ("sp2", 1)]
    for name, ver in order:
        pre = f"sp_{name}_"
        for s, pi, v in zip(OUT[pre + "states"], OUT[pre + "pis"], OUT[pre + "values"]):
            buf.append((s.astype(np.int8), pi.astype(np.float32).copy(), float(v), ver))
    t = FakeTrainer()
    t.replay_buffer = collections.deque(buf)
    states, policies, values = ref_train.Trainer._aggregate_duplicates(t)
    OUT["rp_in_states"] = np.stack([b[0] for b in buf]).astype(np.int8)
    OUT["rp_in_pis"] = np.stack([b[1] for b in buf]).astype(np.float32)
    OUT["rp_in_values"] = np.array([b[2] for b in buf], np.float64)
    OUT["rp_in_versions"] = np.array([b[3] for b in buf], np.int32)
    OUT["rp_out_states"] = np.stack(states).astype(np.int8)
    OUT["rp_out_pis"] = np.stack(policies).astype(np.float32)
    OUT["rp_out_values"] = np.array(values, np.float32)
    print("replay", len(buf), "->", len(states), "unique")


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    if "--update" in sys.argv:  # keep existing fixtures, (re)generate only the named groups
        OUT.update(np.load(dst))
        for g in sys.argv[sys.argv.index("--update") + 1:]:
            globals()["gen_" + g]()
    else:
        gen_env()
        gen_mcts()
        gen_selfplay()
        gen_replay()
    np.savez_compressed(dst, **OUT)
    print("wrote", dst, os.path.getsize(dst), "bytes")
