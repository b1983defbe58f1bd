"""GPU parity of the Othello env kernels, through the C ABI, against
(a) results of the reference itself (tests/golden) and (b) the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def game():
    from alphazero_othello_b200.envs.othello import OthelloGameNew
    return OthelloGameNew(8)


def test_initial_state_and_sizes(game):
    s = game.get_initial_state()
    assert s.dtype == np.int8 and s.shape == (8, 8)
    assert s[3, 4] == s[4, 3] == 1 and s[3, 3] == s[4, 4] == -1 and np.abs(s).sum() == 4
    assert game.action_size == 65 and game.state_size == 64
    m = game.get_valid_moves(s, 1)
    assert m.dtype == np.uint8 and list(np.nonzero(m)[0]) == [19, 26, 37, 44]  # envs/test_equivalence_game.py:57-71
    with pytest.raises(AssertionError):
        type(game)(6)


def test_reference_games_bit_exact(game, golden):
    L = golden["env_game_len"]
    for i in range(len(L)):
        n = int(L[i])
        after = golden["env_game_states"][i, :n]
        before = np.concatenate([game.get_initial_state()[None], after[:-1]])
        players = golden["env_game_players"][i, :n]
        acts = golden["env_game_actions"][i, :n]
        assert np.array_equal(game.valid_moves_batch(before, players), golden["env_game_masks"][i, :n])
        assert np.array_equal(game.next_state_batch(before, acts, players), after)
        v, t = game.value_and_terminated_batch(after, players)
        assert np.array_equal(v, golden["env_game_values"][i, :n])
        assert np.array_equal(t, golden["env_game_terms"][i, :n].astype(bool))
        assert game.get_score(after[-1], 1) == golden["env_game_scores"][i]


def test_single_state_api_matches_reference_calls(game, golden):
    s, pl = game.get_initial_state(), 1
    for t in range(int(golden["env_game_len"][1])):
        a = int(golden["env_game_actions"][1, t])
        assert np.array_equal(game.get_valid_moves(s, pl), golden["env_game_masks"][1, t])
        s2 = game.get_next_state(s, a, pl)
        assert s2 is not s and s2.dtype == np.int8 and np.array_equal(s2, golden["env_game_states"][1, t])
        assert game.get_value_and_terminated(s2, a, pl) == (int(golden["env_game_values"][1, t]),
                                                            bool(golden["env_game_terms"][1, t]))
        s, pl = s2, game.get_opponent(pl)


def test_random_boards_masks_terminal_next_illegal(game, golden):
    B = golden["env_rand_boards"]
    for j, pl in enumerate((1, -1)):
        players = np.full(len(B), pl, np.int8)
        masks = game.valid_moves_batch(B, players)
        assert np.array_equal(masks, golden["env_rand_masks"][:, j])
        v, t = game.value_and_terminated_batch(B, players)
        assert np.array_equal(v, golden["env_rand_vt"][:, j, 0]) and np.array_equal(t, golden["env_rand_vt"][:, j, 1] != 0)
        bi, ai = np.nonzero(masks[:, :64])
        nxt = game.next_state_batch(B[bi], ai, players[bi])
        assert np.array_equal(nxt, golden["env_rand_next"][bi, j, ai])
        # pass copies the board unchecked (envs/othello.py:415-416)
        assert np.array_equal(game.next_state_batch(B, np.full(len(B), 64), players), B)
        # every non-legal board action raises ValueError like the reference (envs/othello.py:419-421)
        bi, ai = np.nonzero(masks[:, :64] == 0)
        sel = np.random.RandomState(0).choice(len(bi), 400, replace=False)
        for k in sel:
            with pytest.raises(ValueError, match=f"Illegal move: {ai[k]}"):
                game.get_next_state(B[bi[k]], int(ai[k]), pl)


def test_symmetries_match_reference(game, golden):
    from alphazero_othello_b200.envs.othello import get_random_symmetry
    for i, (b, pi) in enumerate(zip(golden["sym_boards"], golden["sym_pi"])):
        syms = game.get_symmetries(b, pi)
        assert len(syms) == 8
        for j, (s2, p2) in enumerate(syms):
            assert np.array_equal(s2, golden["sym_all_s"][i, j].astype(np.int8))
            assert np.array_equal(np.asarray(p2, np.float32), golden["sym_all_pi"][i, j])
        np.random.seed(1000 + i)  # same draws as the reference run that made the fixture
        s, p = get_random_symmetry(b, pi)
        assert s.dtype == np.float32 and s.shape == (1, 8, 8) and p.dtype == np.float32 and p.shape == (65,)
        assert np.array_equal(s, golden["sym_rnd_s"][i]) and np.array_equal(p, golden["sym_rnd_pi"][i])


def _bits_to_state(black, white):
    s = np.zeros(64, np.int8)
    for i in range(64):
        if (int(black) >> i) & 1:
            s[i] = 1
        elif (int(white) >> i) & 1:
            s[i] = -1
    return s.reshape(8, 8)


def test_rollout_traces_replay_through_oracle():
    import torch
    import oracle as O
    from alphazero_othello_b200.envs.othello import BatchedOthello
    env = BatchedOthello()
    n, n_trace = 8192, 1024
    r = env.rollout(n, seed=7, n_trace=n_trace)
    torch.cuda.synchronize()
    acts = r["trace_actions"].cpu().numpy()
    moves = r["trace_moves"].cpu().numpy().view(np.uint64)
    plies = r["plies"].cpu().numpy()
    score = r["score"].cpu().numpy()
    final = r["final"].cpu().numpy().view(np.uint64)
    og = O.OracleGame()
    n_pass = 0
    for g in range(n_trace):
        s, pl = og.get_initial_state(), 1
        for t in range(plies[g]):
            a = int(acts[g, t])
            mask = og.get_valid_moves(s, pl)
            got = np.array([(int(moves[g, t]) >> i) & 1 for i in range(64)] + [int(moves[g, t] == 0)], np.uint8)
            assert np.array_equal(mask, got), (g, t)
            assert mask[a] == 1
            n_pass += a == 64
            s = og.get_next_state(s, a, pl)
            v, term = og.get_value_and_terminated(s, a, pl)
            assert term == (t == plies[g] - 1)
            pl = -pl
        assert acts[g, plies[g]] == 0xFF
        assert np.array_equal(s, _bits_to_state(final[g, 0], final[g, 1]))
        assert og.get_score(s, 1) == score[g]
    assert int(r["counters"][0]) == plies.sum()
    assert 58 < plies.mean() < 61 and n_pass > 0


def test_rollout_full_size_properties():
    """Size-independent properties at 2^20 games: outcomes are terminal positions, discs are
    consistent, and the result does not depend on launch geometry (game ids key the RNG)."""
    import torch
    from alphazero_othello_b200.envs.othello import BatchedOthello
    env = BatchedOthello()
    n = 1 << 20
    r = env.rollout(n, seed=3)
    black, white = r["final"][:, 0], r["final"][:, 1]
    assert int((black & white).count_nonzero()) == 0
    pc = lambda x: sum(((x >> i) & 1) for i in range(64))
    assert torch.equal((pc(black) - pc(white)).to(torch.int32), r["score"])
    assert int(env.legal_moves(black, white).count_nonzero()) == 0
    assert int(env.legal_moves(white, black).count_nonzero()) == 0
    assert int(r["counters"][0]) == int(r["plies"].sum())
    sub = env.rollout(1000, seed=3, game_id_base=5000)
    assert torch.equal(sub["score"], r["score"][5000:6000]) and torch.equal(sub["final"], r["final"][5000:6000])


def test_device_step_and_legal_moves_vs_oracle():
    import torch
    import oracle as O
    from alphazero_othello_b200.envs.othello import BatchedOthello
    env = BatchedOthello()
    og = O.OracleGame()
    rs = np.random.RandomState(11)
    # reachable positions: random playouts by the oracle, every ply recorded
    S, P, A = [], [], []
    for g in range(12):
        s, pl = og.get_initial_state(), 1
        while True:
            m = og.get_valid_moves(s, pl)
            a = int(rs.choice(np.nonzero(m)[0]))
            S.append(s.copy()); P.append(pl); A.append(a)
            s = og.get_next_state(s, a, pl)
            if og.get_value_and_terminated(s, a, pl)[1]:
                break
            pl = -pl
    S = np.stack(S); P = np.array(P, np.int8); A = np.array(A, np.uint8)
    n = len(S) - (len(S) % 2) - 1  # odd count exercises the scalar tail
    S, P, A = S[:n], P[:n], A[:n]
    st = torch.from_numpy(S).cuda(); pt = torch.from_numpy(P).cuda(); at = torch.from_numpy(A).cuda()
    own, opp = env.pack(st, pt)
    assert torch.equal(env.unpack(own, opp, pt), st)
    lm = env.legal_moves(own, opp).cpu().numpy().view(np.uint64)
    no, np_, nm, fl = env.step(own, opp, at)
    nxt = env.unpack(np_, no, pt).cpu().numpy()  # after the move the mover's discs are `opp`
    fl = fl.cpu().numpy(); nm = nm.cpu().numpy().view(np.uint64)
    for i in range(n):
        mask = og.get_valid_moves(S[i], P[i])
        assert [(int(lm[i]) >> k) & 1 for k in range(64)] == list(mask[:64])
        ref = og.get_next_state(S[i], A[i], P[i])
        assert np.array_equal(nxt[i], ref)
        v, term = og.get_value_and_terminated(ref, A[i], P[i])
        assert bool(fl[i] & 2) == term
        assert ((fl[i] & 4) != 0) == (term and v > 0) and ((fl[i] & 8) != 0) == (term and v < 0)
        nmask = og.get_valid_moves(ref, -P[i])
        assert [(int(nm[i]) >> k) & 1 for k in range(64)] == list(nmask[:64])
        assert bool(fl[i] & 16) == (nmask[64] == 1 and not term)
    # illegal actions are flagged and leave the position unchanged
    bad = torch.zeros(n, dtype=torch.uint8, device="cuda")  # square 0 is never legal early on
    no2, np2, _, fl2 = env.step(own[:8], opp[:8], bad[:8])
    assert bool((fl2 & 1).all()) and torch.equal(no2, own[:8]) and torch.equal(np2, opp[:8])


def test_empty_batches(game):
    assert game.valid_moves_batch(np.zeros((0, 8, 8), np.int8), np.zeros(0, np.int8)).shape == (0, 65)
    assert game.next_state_batch(np.zeros((0, 8, 8), np.int8), np.zeros(0, np.int32), np.zeros(0, np.int8)).shape == (0, 8, 8)


def test_device_symmetry_batch_vs_oracle():
    import torch
    import oracle as O
    from alphazero_othello_b200.envs.othello import BatchedOthello
    env = BatchedOthello()
    rs = np.random.RandomState(2)
    n = 4097
    S = rs.randint(-1, 2, size=(n, 8, 8)).astype(np.int8)
    P = rs.rand(n, 65).astype(np.float32)
    ks = rs.randint(0, 4, n).astype(np.int32)
    fl = (rs.rand(n) < 0.5).astype(np.uint8)
    s2, p2 = env.symmetry(torch.from_numpy(S).cuda(), torch.from_numpy(P).cuda(), torch.from_numpy(ks).cuda(), torch.from_numpy(fl).cuda())
    s2, p2 = s2.cpu().numpy(), p2.cpu().numpy()
    assert s2.shape == (n, 1, 8, 8) and s2.dtype == np.float32
    for i in rs.choice(n, 300, replace=False):
        rs_, rp = O.symmetry(S[i], P[i], ks[i], fl[i])
        assert np.array_equal(s2[i], rs_) and np.array_equal(p2[i], rp)
    # identity and involution properties on the whole batch
    z = torch.zeros(n, dtype=torch.int32, device="cuda")
    s0, p0 = env.symmetry(torch.from_numpy(S).cuda(), torch.from_numpy(P).cuda(), z, z.to(torch.uint8))
    assert np.array_equal(s0.cpu().numpy()[:, 0], S.astype(np.float32)) and np.array_equal(p0.cpu().numpy(), P)
    r1, q1 = env.random_symmetry(torch.from_numpy(S).cuda(), torch.from_numpy(P).cuda())
    assert torch.equal(r1.abs().sum((1, 2, 3)), torch.from_numpy(np.abs(S).sum((1, 2)).astype(np.float32)).cuda())
    assert torch.allclose(q1.sum(1), torch.from_numpy(P.sum(1)).cuda(), atol=1e-4)


def test_rollout_rng_contract_philox_keyed_by_game_and_ply():
    """include/othello_b200.h: game g draws from Philox4x32-10 keyed (seed; game_id_base+g, ply>>2), word ply&3,
    and plays the floor(u32 * n_legal / 2^32)-th legal square in ascending order.  Restated on the host."""
    import torch
    import oracle as O
    from alphazero_othello_b200.envs.othello import BatchedOthello
    env = BatchedOthello()
    seed, base = 123456789, 777
    r = env.rollout(64, seed=seed, game_id_base=base, n_trace=64)
    torch.cuda.synchronize()
    acts = r["trace_actions"].cpu().numpy(); moves = r["trace_moves"].cpu().numpy().view(np.uint64); plies = r["plies"].cpu().numpy()
    for g in range(64):
        for t in range(plies[g]):
            m = int(moves[g, t])
            if m == 0:
                assert acts[g, t] == 64
                continue
            legal = [i for i in range(64) if (m >> i) & 1]
            word = int(O.philox(seed, base + g, t >> 2, 0)[t & 3])
            assert acts[g, t] == legal[(word * len(legal)) >> 32], (g, t)
