"""Pins the CPU oracle (oracle/othello_oracle.c) against results of the
reference itself (tests/golden/ref_golden.npz, made by make_golden.py)."""
import numpy as np
import pytest

import oracle as O


@pytest.fixture(scope="module")
def game():
    return O.OracleGame(8)


def test_perft_matches_reference(game, golden):
    def perft(s, p, d):
        if d == 0:
            return 1
        if game.get_value_and_terminated(s, None, p)[1]:
            return 1
        return sum(perft(game.get_next_state(s, int(a), p), -p, d - 1)
                   for a in np.nonzero(game.get_valid_moves(s, p))[0])
    got = [perft(game.get_initial_state(), 1, d) for d in range(1, 6)]
    assert got == list(golden["env_perft"][:5]) == [4, 12, 56, 244, 1396]


def test_reference_games_replay_bit_exact(game, golden):
    L = golden["env_game_len"]
    for i in range(len(L)):
        s, pl = game.get_initial_state(), 1
        for t in range(L[i]):
            a = int(golden["env_game_actions"][i, t])
            assert pl == golden["env_game_players"][i, t]
            assert np.array_equal(game.get_valid_moves(s, pl), golden["env_game_masks"][i, t])
            s = game.get_next_state(s, a, pl)
            assert np.array_equal(s, golden["env_game_states"][i, t])
            v, term = game.get_value_and_terminated(s, a, pl)
            assert v == golden["env_game_values"][i, t] and term == bool(golden["env_game_terms"][i, t])
            pl = -pl
        assert term and game.get_score(s, 1) == golden["env_game_scores"][i]


def test_random_boards_masks_terminal_next_illegal(game, golden):
    B = golden["env_rand_boards"]
    for i, b in enumerate(B):
        for j, pl in enumerate((1, -1)):
            m = game.get_valid_moves(b, pl)
            assert np.array_equal(m, golden["env_rand_masks"][i, j]), (i, pl)
            v, t = game.get_value_and_terminated(b, None, pl)
            assert (v, int(t)) == tuple(golden["env_rand_vt"][i, j])
            assert game.get_score(b, pl) == golden["env_rand_scores"][i, j]
            for a in range(64):
                if m[a]:
                    assert np.array_equal(game.get_next_state(b, a, pl), golden["env_rand_next"][i, j, a])
                else:
                    with pytest.raises(ValueError):
                        game.get_next_state(b, a, pl)
            # pass is never legality-checked and returns a copy (envs/othello.py:415-416)
            assert np.array_equal(game.get_next_state(b, 64, pl), b)


def test_symmetries(golden):
    for i, (b, pi) in enumerate(zip(golden["sym_boards"], golden["sym_pi"])):
        s8, p8 = O.symmetries(b, pi)
        assert np.array_equal(s8, golden["sym_all_s"][i])
        assert np.array_equal(p8, golden["sym_all_pi"][i])
        s, p = O.symmetry(b, pi, golden["sym_rnd_k"][i], golden["sym_rnd_flip"][i])
        assert s.dtype == np.float32 and s.shape == (1, 8, 8)
        assert np.array_equal(s, golden["sym_rnd_s"][i]) and np.array_equal(p, golden["sym_rnd_pi"][i])


def _evaluator(cfg):
    sid = int(cfg[0])
    return O.Evaluator(stub=sid, salt=int(cfg[6]) if len(cfg) > 6 else 0)


@pytest.mark.parametrize("case", ["A", "B", "H_noise", "H_t0", "H_alpha", "B_long"])
def test_mcts_visit_counts_bit_exact(golden, case):
    pre = f"mcts_{case}_"
    cfg = golden[pre + "cfg"]
    sims, c, eps, temp = int(cfg[1]), cfg[2], cfg[3], cfg[5]
    m = O.OracleMCTS(c, sims, _evaluator(cfg), dirichlet_epsilon=eps)
    n = len(golden[pre + "action"])
    for t in range(n):
        noise = golden[pre + "noise"][t] if golden[pre + "noise_used"][t] else None
        probs = m.policy_improve_step(golden[pre + "state"][t], int(golden[pre + "player"][t]), temp, noise,
                                      golden[pre + "u_tie"][t])
        st = m.root_stats()
        assert np.array_equal(st["counts"], golden[pre + "counts"][t]), (case, t)
        assert st["root_n"] == golden[pre + "root_n"][t]
        assert st["root_value"] == golden[pre + "root_value"][t]
        assert np.array_equal(st["child_value"], golden[pre + "cval"][t])
        assert np.array_equal(st["child_prior"], golden[pre + "cpri"][t])
        if temp == 1.0 or abs(temp) < 0.1:
            assert np.array_equal(probs, golden[pre + "probs"][t])
        else:  # counts**(1/temp): libm powf vs numpy's, policy targets within 1e-6
            assert np.abs(probs - golden[pre + "probs"][t]).max() <= 1e-6
        m.make_move(int(golden[pre + "action"][t]))


def test_survey_appendix_a3_known_answers(golden):
    # SURVEY.md Appendix A3, independently recorded by the survey
    assert list(golden["mcts_A_root_n"]) == [101, 125, 142]
    assert list(golden["mcts_B_root_n"]) == [101, 144, 168]
    assert golden["mcts_B_root_value"][0] == 0.04888613861386139
    c = golden["mcts_B_counts"][2]
    assert {int(a): int(c[a]) for a in np.nonzero(c)[0]} == {19: 47, 26: 95, 44: 20, 53: 5}


@pytest.mark.parametrize("case", ["sp0", "sp1", "sp2", "sp3"])
def test_self_play_trajectory_bit_exact(golden, case):
    pre = f"sp_{case}_"
    salt, sims, c, eps, alpha, temp, nexp, lam = golden[pre + "cfg"]
    args = {"c_puct": c, "num_simulations": int(sims), "dirichlet_epsilon": eps, "mcts_temperature": temp,
            "num_exploratory_moves": int(nexp), "lambda": lam}
    T = len(golden[pre + "values"])
    u_move = np.zeros(128); u_move[:T] = golden[pre + "u_move"]
    u_tie = np.zeros(128); u_tie[:T] = golden[pre + "u_tie"]
    out = O.self_play(args, O.Evaluator(stub=O.STUB_H, salt=int(salt)), golden[pre + "noise"], u_move, u_tie)
    assert len(out["values"]) == T
    assert np.array_equal(out["states"], golden[pre + "states"])
    if temp == 1.0:
        assert np.array_equal(out["pis"], golden[pre + "pis"])
    else:
        assert np.abs(out["pis"] - golden[pre + "pis"]).max() <= 1e-6
    assert np.array_equal(out["values"], golden[pre + "values"])


def test_choice_model(golden):
    rs = np.random.RandomState(5)
    for _ in range(500):
        p = rs.rand(65).astype(np.float32)
        p[rs.rand(65) < 0.7] = 0
        if p.sum() == 0:
            p[3] = 1
        p = (p / p.sum()).astype(np.float32)
        st = rs.get_state()
        ref = rs.choice(65, p=p)
        sh = np.random.RandomState(); sh.set_state(st)
        assert O.choice(p, sh.random_sample()) == ref


def test_philox_known_answer():
    # Random123 known-answer vector for Philox4x32-10: counter = key = 0
    assert [hex(int(x)) for x in O.philox(0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # all-ones counter and key (Random123 kat_vectors)
    full = (1 << 64) - 1
    assert [hex(int(x)) for x in O.philox(full, full, 0xffffffff, 0xffffffff)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


# ------------------------------------------------------------ round 2: M13 rollouts and the arena, pinned to the reference
@pytest.mark.parametrize("case", ["ro0", "ro1", "ro2"])
def test_rollout_mode_matches_reference_on_its_own_picks(golden_r2, case):
    """policy=None (MCTS_model.py:276-303, 332-335): uniform priors, value = outcome of a random playout seen from the
    leaf's side to move.  The oracle consumes the indices the reference's np.random.choice calls returned (forced
    passes draw nothing) and must reproduce every visit count, child value and root value of the reference run."""
    g = golden_r2
    sims, c = g[case + "_cfg"]
    m = O.OracleMCTS(float(c), int(sims), None, picks=g[case + "_picks"])
    for t in range(len(g[case + "_action"])):
        m.search(g[case + "_state"][t], int(g[case + "_player"][t]))
        r = m.root_stats()
        assert np.array_equal(r["counts"], g[case + "_counts"][t]), (case, t)
        assert np.array_equal(r["child_value"], g[case + "_cval"][t]), (case, t)
        assert r["root_value"] == g[case + "_root_value"][t] and r["root_n"] == g[case + "_root_n"][t]
        assert m.picks_used == g[case + "_picks_after"][t]  # the same number of draws, search by search
        m.make_move(int(g[case + "_action"][t]))
    assert m.picks_used == len(g[case + "_picks"])


def oracle_play_match(first, second, u_tie, max_plies=128):
    """eval.play_match (eval.py:134-178) on two OracleMCTS trees with the reference's tie picks injected: the mover's
    tree searches with temp=0, arg-max move, BOTH trees follow it.  Returns (result, actions) -- the golden test
    below checks this loop against what the reference's own play_match / _run_one_match returned."""
    game = O.OracleGame(8)
    state, player = game.get_initial_state(), 1
    actions = []
    searched = {id(first): False, id(second): False}
    for ply in range(max_plies):
        if game.get_valid_moves(state, player).sum() == 0:
            return "Draw", actions
        tree = first if player == 1 else second
        probs = tree.policy_improve_step(state, player, temp=0.0, u_tie=float(u_tie[ply]))
        searched[id(tree)] = True
        action = int(np.argmax(probs))
        actions.append(action)
        state = game.get_next_state(state, action, player)
        reward, done = game.get_value_and_terminated(state, action, player)
        if done:
            if reward == 1:
                return ("A" if player == 1 else "B"), actions
            if reward == -1:
                return ("B" if player == 1 else "A"), actions
            return "Draw", actions
        for t in (first, second):
            if searched[id(t)]:  # make_move before the first own search is a no-op (MCTS_model.py:209-211)
                t.make_move(action)
        player = -player
    raise AssertionError("match did not end")


def _arena_trees(cfg, salt_first, salt_second):
    sims, c = int(cfg[0]), float(cfg[1])
    return (O.OracleMCTS(c, sims, O.Evaluator(stub=O.STUB_H, salt=int(salt_first))),
            O.OracleMCTS(c, sims, O.Evaluator(stub=O.STUB_H, salt=int(salt_second))))


def test_arena_matches_reference_results_and_actions(golden_r2):
    """eval._run_one_match (eval.py:86-131): even match index -> the candidate's tree plays +1, odd -> the incumbent's,
    and the result is inverted back to the candidate's point of view.  Actions of every ply and the result string of
    11 reference matches (even and odd indices, wins for both sides, one drawn game)."""
    g = golden_r2
    res = list(g["ar_results"])
    assert "Draw" in res and "A" in res and "B" in res
    for j, (idx, sa, sb, _seed, plies) in enumerate(g["ar_meta"]):
        assert idx == j
        first, second = _arena_trees(g["ar_cfg"], sa if idx % 2 == 0 else sb, sb if idx % 2 == 0 else sa)
        r, acts = oracle_play_match(first, second, g["ar_u_tie"][j])
        if idx % 2 == 1 and r != "Draw":
            r = "B" if r == "A" else "A"
        assert acts == list(g["ar_actions"][j, :plies]), j
        assert r == res[j], j
    first, second = _arena_trees(g["ar_cfg"], *g["ar_direct_salts"])  # play_match alone: A = the tree that plays +1
    r, acts = oracle_play_match(first, second, g["ar_direct_u_tie"])
    assert r == g["ar_direct_result"][0] and acts == list(g["ar_direct_actions"])
