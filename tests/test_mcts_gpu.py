"""GPU parity of the MCTS / self-play kernels through the C ABI: bit-exact visit counts,
root values and value targets against results of the reference itself (tests/golden)
and against the CPU oracle on the engine's own recorded randomness."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LANES = [32, 16, 8]


def _pack(state, player):
    own = opp = 0
    for i, v in enumerate((np.asarray(state).astype(np.int64) * int(player)).ravel()):
        if v == 1:
            own |= 1 << i
        elif v == -1:
            opp |= 1 << i
    return own, opp


def _i64(x):
    return np.array([x], np.uint64).view(np.int64)


def _manual_engine(cfg, lanes, n_slots=1, **kw):
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    sid, sims, c, eps, alpha = int(cfg[0]), int(cfg[1]), cfg[2], cfg[3], cfg[4]
    args = {"c_puct": c, "num_simulations": sims, "dirichlet_alpha": alpha, "dirichlet_epsilon": eps}
    return MctsEngine(n_slots, args, self_play=False, eval_kind=[_lib.EVAL_STUB_A, _lib.EVAL_STUB_B, _lib.EVAL_STUB_H][sid],
                      inject_random=True, lanes=lanes, stub_salt=int(cfg[6]), max_inline_sims=1000, **kw)


def _search(e):
    from alphazero_othello_b200 import _lib
    e.begin_search()
    for _ in range(64):
        e.step()
        if (e.ctl()["phase"] != _lib.PH_RUN).all():
            break
    e.raise_on_error()
    assert (e.ctl()["phase"] == _lib.PH_IDLE).all()


@pytest.mark.parametrize("lanes", LANES)
@pytest.mark.parametrize("case", ["A", "B", "H_noise", "H_t0", "H_alpha", "B_long"])
def test_visit_counts_match_reference(golden, case, lanes):
    import torch
    pre = f"mcts_{case}_"
    cfg = golden[pre + "cfg"]
    e = _manual_engine(cfg, lanes)
    own, opp = _pack(golden[pre + "state"][0], golden[pre + "player"][0])
    e.set_roots(torch.from_numpy(_i64(own)).cuda(), torch.from_numpy(_i64(opp)).cuda(),
                torch.tensor([int(golden[pre + "player"][0])], dtype=torch.int8, device="cuda"))
    n = len(golden[pre + "action"])
    for t in range(n):
        if golden[pre + "noise_used"][t]:
            e.noise[0] = torch.from_numpy(golden[pre + "noise"][t]).cuda()
        _search(e)
        st = {k: v.cpu().numpy()[0] for k, v in e.root_stats().items()}
        assert np.array_equal(st["counts"], golden[pre + "counts"][t]), (case, t)
        assert st["root_n"] == golden[pre + "root_n"][t]
        assert st["root_value"] == golden[pre + "root_value"][t]
        assert np.array_equal(st["child_value"], golden[pre + "cval"][t])
        assert np.array_equal(st["child_prior"], golden[pre + "cpri"][t])
        assert tuple(st["root_board"].view(np.uint64)) == _pack(golden[pre + "state"][t], golden[pre + "player"][t])
        e.advance(torch.tensor([int(golden[pre + "action"][t])], dtype=torch.int32, device="cuda"))
        e.raise_on_error()


def test_make_move_on_unexpanded_action_is_keyerror(golden):
    import torch
    e = _manual_engine(golden["mcts_A_cfg"], 32)
    own, opp = _pack(golden["mcts_A_state"][0], 1)
    e.set_roots(torch.from_numpy(_i64(own)).cuda(), torch.from_numpy(_i64(opp)).cuda(),
                torch.tensor([1], dtype=torch.int8, device="cuda"))
    _search(e)
    e.advance(torch.tensor([0], dtype=torch.int32, device="cuda"))  # square 0 is not a legal first move
    with pytest.raises(KeyError):
        e.raise_on_error()


def _selfplay_engine(args, lanes, n_slots, inject, salt, seed=0, games_per_slot=1, **kw):
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    return MctsEngine(n_slots, args, self_play=True, eval_kind=_lib.EVAL_STUB_H, inject_random=inject, lanes=lanes,
                      stub_salt=salt, seed=seed, games_per_slot=games_per_slot, **kw)


def _run_to_done(e, max_launches=20000):
    from alphazero_othello_b200 import _lib
    e.reset()
    for i in range(max_launches):
        e.step()
        if i % 16 == 15:
            c = e.counters()
            if c["errors"]:
                e.raise_on_error()
            if c["active"] == 0:
                return
    raise AssertionError("self-play did not finish")


@pytest.mark.parametrize("lanes", LANES)
@pytest.mark.parametrize("case", ["sp0", "sp1", "sp2", "sp3"])
def test_self_play_matches_reference(golden, case, lanes):
    import torch
    from alphazero_othello_b200.engine import split_games
    pre = f"sp_{case}_"
    salt, sims, c, eps, alpha, temp, nexp, lam = golden[pre + "cfg"]
    args = {"c_puct": c, "num_simulations": int(sims), "dirichlet_alpha": alpha, "dirichlet_epsilon": eps,
            "mcts_temperature": temp, "num_exploratory_moves": int(nexp), "lambda": lam}
    e = _selfplay_engine(args, lanes, 1, True, int(salt))
    T = len(golden[pre + "values"])
    e.noise[0] = torch.from_numpy(golden[pre + "noise"]).cuda()
    e.u_move[0, :T] = torch.from_numpy(golden[pre + "u_move"]).cuda()
    e.u_tie[0, :T] = torch.from_numpy(golden[pre + "u_tie"]).cuda()
    _run_to_done(e)
    (traj,) = split_games(e.drain())
    assert len(traj) == T
    assert np.array_equal(np.stack([t[0] for t in traj]), golden[pre + "states"])
    pis = np.stack([t[1] for t in traj])
    if temp == 1.0:
        assert np.array_equal(pis, golden[pre + "pis"])
    else:
        assert np.abs(pis - golden[pre + "pis"]).max() <= 1e-6  # policy targets within 1e-6 (north_star)
    assert np.array_equal(np.array([t[2] for t in traj]), golden[pre + "values"])


@pytest.mark.parametrize("lanes", LANES)
def test_self_play_vs_oracle_on_engine_randomness(lanes):
    """Many concurrent games with the engine's own Philox noise / uniforms; the oracle replays
    each game from the recorded draws and must reproduce every tuple bit for bit."""
    import oracle as O
    from alphazero_othello_b200.engine import split_games
    args = {"c_puct": 2.0, "num_simulations": 48, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 20, "lambda": 0.98}
    n = 96
    e = _selfplay_engine(args, lanes, n, False, 99, seed=1234)
    _run_to_done(e)
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    out = e.drain()
    games = split_games(out)
    ids = sorted(int(g[0]) for g in out["games"].numpy())
    assert ids == list(range(n)) and len(games) == n
    c = e.counters()
    assert c["games"] == n and c["errors"] == 0
    tot_sims = 0
    for g in range(n):
        ref = O.self_play(args, O.Evaluator(stub=O.STUB_H, salt=99), noise[g], um[g], ut[g])
        traj = games[g]
        assert len(traj) == len(ref["values"]), g
        assert np.array_equal(np.stack([t[0] for t in traj]), ref["states"]), g
        assert np.array_equal(np.stack([t[1] for t in traj]), ref["pis"]), g
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"]), g
        tot_sims += ref["counters"]["sims"]
    assert c["sims"] == tot_sims
    assert not np.array_equal(noise[0], noise[1])


def test_restart_and_game_ids_are_shard_independent():
    """Slots restart with game_id += stride; a game's content depends only on its id, so two
    'ranks' with half the slots each reproduce the single-engine run."""
    from alphazero_othello_b200.engine import split_games
    args = {"c_puct": 2.0, "num_simulations": 12, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 10, "lambda": 0.9}

    def run(n_slots, base, stride, gps):
        e = _selfplay_engine(args, 32, n_slots, False, 5, seed=77, games_per_slot=gps, game_id_base=base, game_id_stride=stride)
        _run_to_done(e)
        out = e.drain()
        return {int(g[0]): t for g, t in zip(sorted(map(tuple, out["games"].numpy())), split_games(out))}

    whole = run(8, 0, 8, 2)
    assert sorted(whole) == list(range(16))
    half0, half1 = run(4, 0, 8, 2), run(4, 4, 8, 2)
    merged = {**half0, **half1}
    assert sorted(merged) == sorted(whole)
    for gid in whole:
        a, b = whole[gid], merged[gid]
        assert len(a) == len(b) and all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2]
                                        for x, y in zip(a, b))


def test_external_evaluator_record_and_replay():
    """The network path (WAIT_EVAL hand-off): a small torch module evaluates leaf batches; every
    (position -> priors, value) it produced is recorded and served to the oracle."""
    import torch
    import oracle as O
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine, split_games
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 96), torch.nn.Tanh(), torch.nn.Linear(96, 66)).cuda()
    args = {"c_puct": 2.0, "num_simulations": 24, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 12, "lambda": 0.98}
    n = 16
    e = MctsEngine(n, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1, seed=5)
    table = O.EvalTable(1 << 18)
    e.reset()
    e.step()
    w = (1 << (63 - np.arange(64, dtype=np.uint64)))  # the reference's bit numbering
    for it in range(40000):
        ph = e.ctl()["phase"]
        if not ((ph == _lib.PH_RUN) | (ph == _lib.PH_WAIT_EVAL)).any():
            break
        with torch.no_grad():
            y = net(e.nn_input.view(n, 64))
            e.priors.copy_(torch.softmax(y[:, :65], -1))
            e.values.copy_(torch.tanh(y[:, 65]))
        x = e.nn_input.view(n, 64).cpu().numpy()
        pr, va = e.priors.cpu().numpy(), e.values.cpu().numpy()
        for s in np.nonzero(ph == _lib.PH_WAIT_EVAL)[0]:
            own = int((w * (x[s] == 1)).sum()); opp = int((w * (x[s] == -1)).sum())
            assert table.put(own, opp, pr[s], va[s]) in (0, 1)
        e.step()
    e.raise_on_error()
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    games = split_games(e.drain())
    assert len(games) == n
    for g in range(n):
        ref = O.self_play(args, O.Evaluator(table=table), noise[g], um[g], ut[g])
        traj = games[g]
        assert len(traj) == len(ref["values"])
        assert np.array_equal(np.stack([t[0] for t in traj]), ref["states"])
        assert np.array_equal(np.stack([t[1] for t in traj]), ref["pis"])
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"])
    assert table.misses == 0


def test_node_overflow_fails_loudly():
    from alphazero_othello_b200 import _lib
    args = {"c_puct": 2.0, "num_simulations": 200, "dirichlet_epsilon": 0.0}
    e = _selfplay_engine(args, 32, 2, False, 1, node_cap=64)
    with pytest.raises(_lib.OthelloB200Error, match="overflow"):
        _run_to_done(e, 400)


@pytest.mark.parametrize("lanes", LANES)
def test_rollout_evaluator_policy_none_vs_oracle(lanes):
    """policy=None mode (MCTS_model.py:276-303, 332-335): uniform priors + a random playout per leaf,
    playouts drawn from Philox on both sides -> bit-exact visit counts and root values."""
    import torch
    import oracle as O
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    args = {"c_puct": 1.4, "num_simulations": 60, "dirichlet_epsilon": 0.0}
    n = 6
    e = MctsEngine(n, args, self_play=False, eval_kind=_lib.EVAL_ROLLOUT, lanes=lanes, seed=4242, max_inline_sims=1000, game_id_base=100)
    g = O.OracleGame()
    s0 = g.get_initial_state()
    own, opp = _pack(s0, 1)
    e.set_roots(torch.from_numpy(np.repeat(_i64(own), n)).cuda(), torch.from_numpy(np.repeat(_i64(opp), n)).cuda(),
                torch.ones(n, dtype=torch.int8, device="cuda"))
    refs = [O.OracleMCTS(1.4, 60, None, rollout_seed=4242, game_id=100 + i) for i in range(n)]
    states, players = [s0.copy() for _ in range(n)], [1] * n
    for mv in range(8):
        _search(e)
        st = {k: v.cpu().numpy() for k, v in e.root_stats().items()}
        acts = []
        for i in range(n):
            refs[i].search(states[i], players[i])
            r = refs[i].root_stats()
            assert np.array_equal(st["counts"][i], r["counts"]), (mv, i)
            assert st["root_value"][i] == r["root_value"] and st["root_n"][i] == r["root_n"]
            a = int(np.argmax(r["counts"]))
            acts.append(a)
            refs[i].make_move(a)
            states[i] = g.get_next_state(states[i], a, players[i])
            players[i] = -players[i]
        e.advance(torch.tensor(acts, dtype=torch.int32, device="cuda"))
        e.raise_on_error()
    assert len({tuple(st["counts"][i]) for i in range(n)}) > 1  # different game ids -> different playouts


def test_output_ring_and_path_overflow_fail_loudly():
    from alphazero_othello_b200 import _lib
    args = {"c_puct": 2.0, "num_simulations": 6, "dirichlet_epsilon": 0.0, "mcts_temperature": 1.0, "num_exploratory_moves": 60}
    e = _selfplay_engine(args, 8, 4, False, 1, out_pos_cap=40, out_game_cap=8)  # a game has ~60 positions
    with pytest.raises(_lib.OthelloB200Error, match="output ring overflow"):
        _run_to_done(e, 4000)
    args = {"c_puct": 2.0, "num_simulations": 400, "dirichlet_epsilon": 0.0}
    e = _selfplay_engine(args, 8, 2, False, 1, path_cap=3)
    with pytest.raises(_lib.OthelloB200Error, match="path overflow"):
        _run_to_done(e, 4000)


def test_many_simulations_deep_tree_matches_oracle():
    """400 simulations per move (the C4 setting) on a few games: deep paths, big arenas, many re-roots."""
    import oracle as O
    from alphazero_othello_b200.engine import split_games
    args = {"c_puct": 2.0, "num_simulations": 400, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 35, "lambda": 0.98}
    n = 6
    e = _selfplay_engine(args, 8, n, False, 7, seed=31)
    _run_to_done(e, 200000)
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    c = e.counters()
    games = split_games(e.drain())
    assert c["max_depth"] >= 8 and c["max_top"] > 3000
    for g in range(n):
        ref = O.self_play(args, O.Evaluator(stub=O.STUB_H, salt=7), noise[g], um[g], ut[g])
        traj = games[g]
        assert len(traj) == len(ref["values"])
        assert np.array_equal(np.stack([t[1] for t in traj]), ref["pis"]), g
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"]), g


def test_fused_softmax_path_record_and_replay():
    """oth_mcts_step_fused: the kernel applies softmax / tanh to raw (bf16, strided) head outputs itself; the
    priors and values it used are written back, recorded and replayed through the oracle."""
    import torch
    import oracle as O
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine, split_games
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(64, 96), torch.nn.Tanh(), torch.nn.Linear(96, 128)).cuda()
    args = {"c_puct": 2.0, "num_simulations": 20, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 12, "lambda": 0.98}
    n = 12
    e = MctsEngine(n, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1, seed=6)
    table = O.EvalTable(1 << 18)
    e.reset()
    e.step()
    w = (1 << (63 - np.arange(64, dtype=np.uint64)))
    for it in range(40000):
        ph = e.ctl()["phase"]
        if not ((ph == _lib.PH_RUN) | (ph == _lib.PH_WAIT_EVAL)).any():
            break
        x = e.nn_input.view(n, 64).cpu().numpy()
        with torch.no_grad():
            y = (3.0 * net(e.nn_input.view(n, 64))).to(torch.bfloat16)  # [n,128]: logits in cols 0..64, value pre-activation col 100
        e.step_fused(y[:, :65], y[:, 100:101], record=True)
        pr, va = e.priors.cpu().numpy(), e.values.cpu().numpy()
        ref_p = torch.softmax(y[:, :65].float(), -1).cpu().numpy()
        for s in np.nonzero(ph == _lib.PH_WAIT_EVAL)[0]:
            assert np.abs(pr[s] - ref_p[s]).max() < 1e-6 and abs(va[s] - np.tanh(float(y[s, 100]))) < 1e-6
            own = int((w * (x[s] == 1)).sum()); opp = int((w * (x[s] == -1)).sum())
            assert table.put(own, opp, pr[s], va[s]) in (0, 1)
    e.raise_on_error()
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    games = split_games(e.drain())
    assert len(games) == n
    for g in range(n):
        ref = O.self_play(args, O.Evaluator(table=table), noise[g], um[g], ut[g])
        traj = games[g]
        assert len(traj) == len(ref["values"])
        assert np.array_equal(np.stack([t[1] for t in traj]), ref["pis"])
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"])
    assert table.misses == 0


def test_move_uniforms_follow_the_documented_philox_streams():
    """u_move / u_tie of ply t in game id G = u01_53(Philox4x32-10(seed; G, t, purpose 1 / 2)) (csrc/philox.cuh)."""
    import oracle as O
    args = {"c_puct": 2.0, "num_simulations": 8, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 5, "lambda": 0.98}
    e = _selfplay_engine(args, 8, 5, False, 3, seed=2024, game_id_base=40)
    _run_to_done(e)
    um, ut = e.u_move.cpu().numpy(), e.u_tie.cpu().numpy()
    out = e.drain()
    for gid, first, n, _w in out["games"].numpy():
        slot = int(gid) - 40
        for t in range(int(n)):
            for purpose, got in ((1, um[slot, t]), (2, ut[slot, t])):
                w = O.philox(2024, int(gid), t, purpose)
                exp = ((int(w[0]) >> 5) * 67108864.0 + (int(w[1]) >> 6)) / 9007199254740992.0
                assert got == exp, (gid, t, purpose)


@pytest.mark.parametrize("alpha", [1.0, 0.3, 0.03])
def test_dirichlet_noise_is_a_dirichlet_draw(alpha):
    """Root noise is drawn in-kernel (Marsaglia-Tsang gammas from Philox, normalised): check the moments of
    Dirichlet(alpha * 1_65) over 16 384 independent slots -- np.random.dirichlet([alpha]*65), MCTS_model.py:341."""
    args = {"c_puct": 2.0, "num_simulations": 4, "dirichlet_alpha": alpha, "dirichlet_epsilon": 0.25}
    e = _selfplay_engine(args, 8, 16384, False, 0, seed=77)
    e.reset()
    nz = e.noise.cpu().numpy()
    assert nz.shape == (16384, 65) and (nz >= 0).all() and np.allclose(nz.sum(1), 1.0, atol=1e-12)
    a0 = 65 * alpha
    var = alpha * (a0 - alpha) / (a0 * a0 * (a0 + 1))
    assert abs(nz.mean() - 1 / 65) < 1e-12 + 1e-9  # rows sum to one exactly
    assert np.abs(nz.mean(0) - 1 / 65).max() < 6 * np.sqrt(var / 16384)
    assert abs(nz.var(0).mean() / var - 1) < 0.03
    # independent across slots and reproducible from (seed, game id)
    e2 = _selfplay_engine(args, 32, 64, False, 0, seed=77)
    e2.reset()
    assert np.array_equal(e2.noise.cpu().numpy(), nz[:64]) and not np.array_equal(nz[0], nz[1])


def test_full_size_c4_whole_games_invariants_and_sampled_oracle_replay():
    """BASELINE configs[3] geometry -- 16 384 concurrent games, 400 simulations per move, the training
    hyper-parameters -- played to the end with the device stub evaluator (the network is not part of
    this check).  Size-independent properties over ALL ~10^6 positions, and 24 games sampled across the
    slot range replayed through the oracle from the recorded draws: every tuple bit for bit."""
    import torch
    import oracle as O
    from alphazero_othello_b200.envs.othello import BatchedOthello
    args = {"c_puct": 2.0, "num_simulations": 400, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3,
            "mcts_temperature": 1.0, "num_exploratory_moves": 35, "lambda": 0.98}
    n = 16384
    e = _selfplay_engine(args, 8, n, False, 7, seed=2024, max_inline_sims=16, out_pos_cap=n * 72, out_game_cap=n + 16)
    _run_to_done(e, max_launches=40000)
    c = e.counters()
    assert c["games"] == n and c["errors"] == 0
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    out = e.drain(to_host=False)
    games = out["games"].cpu().numpy()
    assert sorted(games[:, 0].tolist()) == list(range(n))                 # every game id exactly once
    npos = int(out["values"].numel())
    assert npos == int(games[:, 2].sum()) == c["moves"]                   # one tuple per move
    assert c["sims"] == 400 * c["moves"]                                  # num_simulations per search, every search
    assert 4 <= games[:, 2].min() and games[:, 2].max() <= 128
    # ranges of the descriptors tile the output ring without gaps or overlaps
    order = np.argsort(games[:, 1])
    assert games[order[0], 1] == 0 and np.array_equal(games[order, 1][1:], (games[order, 1] + games[order, 2])[:-1])
    # policy targets: non-negative, sum to 1, zero outside the legal set of that position (env kernel as checker)
    pis = out["pis"]
    assert float(pis.min()) >= 0.0 and float((pis.sum(1) - 1).abs().max()) <= 1e-5
    env = BatchedOthello("cuda:0")
    own, opp = out["boards"][:, 0].contiguous(), out["boards"][:, 1].contiguous()
    lm = env.legal_moves(own, opp)
    bits = ((lm.unsqueeze(1) >> torch.arange(64, device=lm.device)) & 1).bool()
    legal65 = torch.cat([bits, (lm == 0).unsqueeze(1)], 1)
    assert not bool((pis > 0)[~legal65].any())
    assert bool(((pis > 0) & legal65).any(1).all())
    assert int((own & opp).count_nonzero()) == 0
    # value targets: |G| <= 1, and the last position of a game carries z in {-1, 0, +1}
    v = out["values"]
    assert float(v.abs().max()) <= 1.0
    last = torch.from_numpy(games[:, 1] + games[:, 2] - 1).to(v.device)
    assert bool(torch.isin(v[last], torch.tensor([-1.0, 0.0, 1.0], dtype=v.dtype, device=v.device)).all())
    # sampled games through the oracle (ids spread over the slot range)
    states, pis_h, v_h = out["states"].cpu().numpy(), pis.cpu().numpy(), v.cpu().numpy()
    desc = {int(r[0]): (int(r[1]), int(r[2])) for r in games}
    for g in list(range(0, n, n // 20)) + [1, n // 2 + 3, n - 2, n - 1]:
        ref = O.self_play(args, O.Evaluator(stub=O.STUB_H, salt=7), noise[g], um[g], ut[g])
        first, T = desc[g]
        assert T == len(ref["values"]), g
        assert np.array_equal(states[first:first + T], ref["states"]), g
        assert np.array_equal(pis_h[first:first + T], ref["pis"]), g
        assert np.array_equal(v_h[first:first + T], ref["values"]), g


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("OTH_FUZZ_SEEDS", "10"))))  # more seeds: one-off fuzz runs
def test_randomised_hyper_parameters_vs_oracle(seed):
    """Differential fuzz: random c_puct / simulations / Dirichlet eps, alpha / temperature / exploratory moves / lambda /
    lane width / inline-simulation budget; 24 concurrent games on the engine's own randomness, every game replayed
    through the oracle.  Bit-exact states, value targets and (temperature 1 or ~0) policy targets; 1e-6 otherwise."""
    import oracle as O
    from alphazero_othello_b200.engine import split_games
    rs = np.random.RandomState(4242 + seed)
    temp = float(rs.choice([0.25, 0.5, 1.0, 1.0, 2.0]))
    args = {"c_puct": float(rs.uniform(0.5, 4.0)), "num_simulations": int(rs.choice([1, 2, 7, 33, 64, 120])),
            "dirichlet_alpha": float(rs.choice([0.03, 0.3, 1.0])), "dirichlet_epsilon": float(rs.choice([0.0, 0.1, 0.3, 0.6])),
            "mcts_temperature": temp, "num_exploratory_moves": int(rs.randint(0, 41)), "lambda": float(rs.uniform(0.5, 1.0))}
    lanes = int(rs.choice(LANES))
    salt = int(rs.randint(1, 1 << 30))
    n = 24
    # kernel variant: monolithic device-evaluator kernel, or the production pair (step + move kernel) with the stub in the
    # network's place, move kernel launched by the host or from the device, path tail forced into HBM or not
    variant = dict(split_stub=bool(rs.randint(2)), move_launch=int(rs.randint(2)), hot_path=int(rs.choice([0, 0, 3, 9])))
    e = _selfplay_engine(args, lanes, n, False, salt, seed=int(rs.randint(1 << 30)), max_inline_sims=int(rs.choice([1, 4, 16, 1000])),
                         **variant)
    _run_to_done(e, max_launches=60000)
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    out = e.drain()
    games = split_games(out)
    assert sorted(int(g[0]) for g in out["games"].numpy()) == list(range(n))
    c = e.counters()
    assert c["games"] == n and c["errors"] == 0
    tot_sims = 0
    for g in range(n):
        ref = O.self_play(args, O.Evaluator(stub=O.STUB_H, salt=salt), noise[g], um[g], ut[g])
        traj = games[g]
        assert len(traj) == len(ref["values"]), (args, g)
        assert np.array_equal(np.stack([t[0] for t in traj]), ref["states"]), (args, g)
        pis = np.stack([t[1] for t in traj])
        if temp == 1.0:
            assert np.array_equal(pis, ref["pis"]), (args, g)
        else:
            assert np.abs(pis - ref["pis"]).max() <= 1e-6, (args, g)
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"]), (args, g)
        tot_sims += ref["counters"]["sims"]
    assert c["sims"] == tot_sims
