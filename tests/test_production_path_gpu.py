"""Parity of the PRODUCTION kernel pair -- the kernels bench.py times and collect_self_play_games runs:
k_mcts_step_fused (softmax / tanh fused, hot record hand-off across the network call) + the move kernel
(host-launched flag scan or device tail launch), behind the real bf16 network twin, as CUDA-graph replays.

* record / replay: every (leaf -> priors, value) the step kernel consumed is recorded and every game is replayed
  through the CPU oracle (oracle/record_replay.py): states, policy targets, value targets bit for bit, simulation
  totals equal -- at the BASELINE settings (400 simulations per move on the big net, 200 on the small one), whole
  games, all three lane widths, graph on and off, paths deeper than the 8 entries that travel with the control
  block, and (hot_path knob) paths that spill into the HBM path tail;
* the same kernel pair with a device stub in the network's place (split_stub) against the oracle on hundreds of
  games and at the full C4 geometry;
* non-finite network outputs stop a slot with OTH_ERR_NONFINITE instead of corrupting its tree.

Reference semantics: MCTS_model.py:129-169, 200-274, 325-395; self_play_worker.py:8-88.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TRAIN_ARGS = {"c_puct": 2.0, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3, "mcts_temperature": 1.0,
              "num_exploratory_moves": 35, "lambda": 0.98}  # train.py:399-423


def _runner(kind, n_slots, sims, lanes, graph, seed, hot_path=0, move_launch=None, dtype="bf16", dedup=False):
    import torch
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.Models import AlphaZeroNet, FastOthelloNet, fold_for_inference
    from alphazero_othello_b200.engine import BatchedPolicy, MctsEngine, SelfPlayRunner
    torch.manual_seed(seed)
    net = (AlphaZeroNet(8, 65, 5, 128) if kind == "big" else FastOthelloNet(8, 65)).cuda().eval()
    args = dict(TRAIN_ARGS, num_simulations=sims)
    e = MctsEngine(n_slots, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1, seed=seed, lanes=lanes,
                   hot_path=hot_path, move_launch=move_launch)
    ev = BatchedPolicy(fold_for_inference(net, torch.bfloat16 if dtype == "bf16" else torch.float32), "cuda:0", torch.float32)
    run = SelfPlayRunner(e, ev, use_graph=graph, record=True, dedup=dedup)
    assert run.fused, "the production path is the fused one"
    return e, run, args


CASES = [
    # kind, slots, sims, lanes, graph, hot_path, move_launch
    ("big", 96, 400, 8, True, 0, 1),
    ("big", 80, 400, 16, False, 0, 0),
    ("big", 80, 400, 32, True, 0, 1),
    ("small", 128, 200, 32, True, 0, 1),
    ("small", 64, 200, 8, True, 0, 0),
    ("small", 48, 120, 16, True, 3, 1),   # hot_path = 3: every path deeper than 3 goes through the HBM path tail
]


@pytest.mark.parametrize("kind,n_slots,sims,lanes,graph,hot_path,move_launch", CASES)
def test_fused_step_and_move_kernels_with_the_real_network_replay_through_the_oracle(kind, n_slots, sims, lanes, graph,
                                                                                      hot_path, move_launch):
    import oracle as O
    from oracle.record_replay import record_self_play, replay_and_compare
    e, run, args = _runner(kind, n_slots, sims, lanes, graph, seed=1000 + lanes + sims, hot_path=hot_path,
                           move_launch=move_launch)
    tape = O.EvalTape(n_slots, sims * 70 + 128)
    table = O.EvalTable(1 << 22) if n_slots * sims <= 80 * 400 and lanes != 8 else None  # position-keyed twin: batch-row invariance
    iters = record_self_play(run, tape, table)
    e.raise_on_error()
    c = e.counters()
    assert c["games"] == n_slots and c["errors"] == 0 and c["sims"] == sims * c["moves"]
    assert iters >= sims * 9 and run.graph is not None if graph else run.graph is None
    # the hot record carries 8 path entries with the control block and up to 52 in all: deeper paths took the other branches
    # (random-init small net: shallower trees; the big net at 400 simulations goes past 8, the hot_path case past 3)
    assert c["max_depth"] >= (9 if kind == "big" else 5), c["max_depth"]
    checked, oracle_sims = replay_and_compare(e, args, tape)
    assert checked == n_slots and oracle_sims == c["sims"] and tape.served == c["evals"]
    if table is not None:
        # did the network answer identical positions identically whatever their batch row?  (what an evaluation
        # de-duplication would need in order to be exact; reported)
        inserted, repeated, conflicts = (int(x) for x in table.stats)
        print(f"\n[row invariance] {kind} net, {n_slots} slots: {inserted} distinct positions, {repeated} identical repeats, "
              f"{conflicts} repeats with a different output")
        assert inserted + repeated + conflicts == c["evals"] and repeated + conflicts > 0  # the games share their openings


@pytest.mark.parametrize("graph,force_bucket", [(True, None), (False, None), (True, 256)])
def test_evaluation_dedup_replays_through_the_oracle(graph, force_bucket):
    """oth_mcts_dedup + oth_mcts_step_fused_mapped: the network runs on the distinct pending positions of a batch (bucketed
    batch sizes, one CUDA graph each), every slot reads its outputs through eval_map.  512 whole games: each game's tape of
    consumed evaluations replays bit-exactly through the oracle -- also with the bucket forced to 256 rows all game long,
    when half of the waiting slots miss every launch and are served later."""
    import oracle as O
    from oracle.record_replay import record_self_play, replay_and_compare
    n_slots, sims = 512, 48
    e, run, args = _runner("small", n_slots, sims, 8, graph, seed=77, dedup=True)
    assert run.dedup and run.buckets == [448, 384, 320, 256]
    run.force_bucket = force_bucket
    # at 512 slots of the small net a compacted batch is not faster than the whole one, and the measured-time policy knows it:
    # use the (fixed + rows) cost model with a fixed part small enough that the buckets get exercised
    run.fixed_cost_rows, run.use_measured_times = 16, False
    tape = O.EvalTape(n_slots, sims * 70 + 128)
    record_self_play(run, tape, poll_every=64)
    e.raise_on_error()
    c = e.counters()
    assert c["games"] == n_slots and c["errors"] == 0 and c["sims"] == sims * c["moves"]
    used = run.bucket_iterations
    assert used[256] > sims * 3, used            # the opening was played on compacted batches (smallest bucket) ...
    assert force_bucket or sum(used[b] for b in (448, 384, 320)) > 0, used  # ... the transition on the partial ones, with overflow
    assert not (graph and force_bucket is None) or set(run.iteration_ms) == {"plain", 0, 448, 384, 320, 256}  # graphs were timed
    assert force_bucket or used[0] > sims * 20   # ... the middle game on whole batches
    assert run.rows_evaluated < sum(used.values()) * n_slots
    checked, oracle_sims = replay_and_compare(e, args, tape)
    assert checked == n_slots and oracle_sims == c["sims"] and tape.served == c["evals"]


def test_dedup_kernel_maps_identical_positions_to_one_row():
    """oth_mcts_dedup on a fresh engine: every slot waits for the initial position -> one distinct row, eval_map all 0,
    the row's planes are the canonical start position; bucket 0 only counts."""
    import torch
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    import ctypes as C
    n = 1000
    e = MctsEngine(n, dict(TRAIN_ARGS, num_simulations=8), self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1, lanes=8)
    e.reset()
    e.priors.fill_(1.0 / 65); e.values.zero_()
    e.step()
    nb = C.c_int64(0)
    _lib.check(e.L.oth_mcts_dedup_workspace_bytes(n, C.byref(nb)))
    ws = torch.zeros(nb.value, dtype=torch.uint8, device="cuda")
    x = torch.full((256, 1, 8, 8), 7.0, device="cuda")
    emap = torch.full((n,), -5, dtype=torch.int32, device="cuda")
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    e.dedup(0, None, None, stats, ws)
    assert stats.tolist() == [1, n]
    e.dedup(256, x, emap, stats, ws)
    assert stats.tolist() == [1, n] and (emap == 0).all()
    assert torch.equal(x[0], e.nn_input[0]) and (x[1:] == 0).all()
    # a few iterations later the slots have diverged (different root noise): several distinct rows, every waiting slot mapped
    for _ in range(6):
        e.step()
    e.dedup(256, x, emap, stats, ws)
    u, w = stats.tolist()
    assert 1 < u <= 256 and w == n and int(emap.max()) == u - 1 and int(emap.min()) == 0
    planes = e.nn_input.view(n, 64)
    assert torch.equal(x.view(256, 64)[emap.long()], planes)  # each slot's row holds exactly its own position


def _stub_engine(n, sims, lanes, split, seed, salt, move_launch=None, hot_path=0, **kw):
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    args = dict(TRAIN_ARGS, num_simulations=sims)
    e = MctsEngine(n, args, self_play=True, eval_kind=_lib.EVAL_STUB_H, games_per_slot=1, seed=seed, stub_salt=salt, lanes=lanes,
                   split_stub=split, move_launch=move_launch, hot_path=hot_path, **kw)
    return e, args


def _run_to_done(e, max_launches=400000, every=64):
    e.reset()
    for i in range(max_launches):
        e.step()
        if i % every == every - 1:
            c = e.counters()
            if c["errors"]:
                e.raise_on_error()
            if c["active"] == 0:
                return i + 1
    raise AssertionError("self-play did not finish")


def _compare_all(e, args, salt, games=None):
    import oracle as O
    from alphazero_othello_b200.engine import split_games
    noise = e.noise.cpu().numpy(); um = e.u_move.cpu().numpy(); ut = e.u_tie.cpu().numpy()
    out = e.drain()
    trajs = split_games(out)
    assert sorted(int(g[0]) for g in out["games"].numpy()) == list(range(e.n_slots))
    sims = 0
    for g in (range(e.n_slots) if games is None else games):
        ref = O.self_play(args, O.Evaluator(stub=O.STUB_H, salt=salt), noise[g], um[g], ut[g])
        t = trajs[g]
        assert len(t) == len(ref["values"]), g
        assert np.array_equal(np.stack([x[0] for x in t]), ref["states"]), g
        assert np.array_equal(np.stack([x[1] for x in t]), ref["pis"]), g
        assert np.array_equal(np.array([x[2] for x in t]), ref["values"]), g
        sims += ref["counters"]["sims"]
    return sims


@pytest.mark.parametrize("lanes,move_launch,hot_path", [(8, 1, 0), (16, 0, 0), (32, 1, 0), (8, 1, 2), (32, 0, 5)])
def test_split_kernel_pair_with_device_stub_400_sims_vs_oracle(lanes, move_launch, hot_path):
    """The production split (64/72-register step kernel + move kernel) with the hash stub where the network would be:
    400 simulations per move, 32 whole games, every tuple against the oracle; with hot_path = 2 / 5 every deeper path
    entry goes through OTH_BUF_PATH."""
    n, sims, salt = 32, 400, 17
    e, args = _stub_engine(n, sims, lanes, True, seed=5 + lanes, salt=salt, move_launch=move_launch, hot_path=hot_path)
    launches = _run_to_done(e)
    c = e.counters()
    assert c["games"] == n and c["errors"] == 0 and c["max_depth"] >= 9
    assert launches >= c["evals"] // n  # one evaluation per slot per launch: the split, not the monolithic kernel
    assert _compare_all(e, args, salt) == c["sims"]


def test_split_and_monolithic_kernels_agree_on_every_game():
    """Same seed, same stub: the monolithic device-evaluator kernel and the production pair emit identical replay tuples."""
    from alphazero_othello_b200.engine import split_games
    outs = []
    for split in (False, True):
        e, args = _stub_engine(256, 64, 8, split, seed=99, salt=4, max_inline_sims=16)
        _run_to_done(e)
        out = e.drain()
        outs.append({int(g[0]): t for g, t in zip(sorted(map(tuple, out["games"].numpy())), split_games(out))})
    a, b = outs
    assert sorted(a) == sorted(b) == list(range(256))
    for gid in a:
        assert len(a[gid]) == len(b[gid])
        for x, y in zip(a[gid], b[gid]):
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2]


def test_full_size_c4_geometry_on_the_production_kernel_pair():
    """BASELINE configs[3] geometry -- 16 384 concurrent games, 400 simulations per move -- played to the end by the
    production kernel pair (device stub in the network's place, device-launched move kernel): whole-run invariants
    and 24 games sampled across the slot range against the oracle."""
    n, sims, salt = 16384, 400, 7
    e, args = _stub_engine(n, sims, None, True, seed=2025, salt=salt, move_launch=1, out_pos_cap=n * 72, out_game_cap=n + 16)
    assert e.cfg.lanes == 8
    _run_to_done(e, every=512)
    c = e.counters()
    assert c["games"] == n and c["errors"] == 0 and c["sims"] == sims * c["moves"] and c["max_depth"] >= 10
    sample = list(range(0, n, n // 20)) + [1, n // 2 + 3, n - 2, n - 1]
    _compare_all(e, args, salt, games=sample)


@pytest.mark.parametrize("what", ["prior_nan", "value_nan", "logit_inf_fused", "value_nan_fused"])
def test_non_finite_network_outputs_stop_the_slot_instead_of_corrupting_the_tree(what):
    """A diverged network (NaN / inf logits, priors or values) must not reach PUCT: with every child score NaN the
    arg-max has no winner and the next node index would be garbage.  The slot is stopped with
    OTH_ERR_NONFINITE at expansion, the other slots keep playing, the host gets a loud error."""
    import torch
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    args = dict(TRAIN_ARGS, num_simulations=16)
    n = 8
    e = MctsEngine(n, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1, seed=3, lanes=8)
    e.reset()
    e.priors.fill_(1.0 / 65)
    e.values.zero_()
    e.step()
    logits = torch.zeros((n, 65), dtype=torch.bfloat16, device="cuda")
    vpre = torch.zeros((n, 1), dtype=torch.bfloat16, device="cuda")
    for it in range(40):
        fused = what.endswith("fused")
        if it == 20:  # slot 5 gets a poisoned evaluation once
            if what == "prior_nan":
                e.priors[5, 19] = float("nan")
            elif what == "value_nan":
                e.values[5] = float("nan")
            elif what == "logit_inf_fused":
                logits[5, 3] = float("inf")
            else:
                vpre[5, 0] = float("nan")
        if fused:
            e.step_fused(logits, vpre)
        else:
            e.step()
        if it == 20:
            e.priors.fill_(1.0 / 65); e.values.zero_(); logits.zero_(); vpre.zero_()
    c = e.ctl()
    assert c["error"][5] == 32 and c["phase"][5] == _lib.PH_ERROR
    assert (np.delete(c["error"], 5) == 0).all() and (np.delete(c["sims_done"] + 16 * c["ply"], 5) > 16).all()
    with pytest.raises(_lib.OthelloB200Error, match="non-finite"):
        e.raise_on_error()


def test_launch_profile_is_per_engine():
    """othello_b200_experimental.h: a profile handle attached to one engine's buffers times that engine's launches
    only; another engine stepping in between is not recorded."""
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.engine import MctsEngine
    args = dict(TRAIN_ARGS, num_simulations=8)
    a = MctsEngine(64, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1)
    b = MctsEngine(64, args, self_play=True, eval_kind=_lib.EVAL_STUB_H, games_per_slot=1, split_stub=True)
    for e in (a, b):
        e.priors.fill_(1.0 / 65)
        e.reset()
    a.profile_begin(5)
    for _ in range(7):
        a.step()
        b.step()
    step_ms, move_ms = a.profile_end()
    assert len(step_ms) == len(move_ms) == 5  # recording stops at the cap
    assert all(0.0 < t < 50.0 for t in step_ms) and all(0.0 <= t < 50.0 for t in move_ms)
    b.profile_begin(3)
    b.step()
    s2, m2 = b.profile_end()
    assert len(s2) == 1
    a.step()  # launches keep working with profiling off
    a.raise_on_error(); b.raise_on_error()


def test_debug_build_asserts_hold_on_a_short_run():
    """The -DOTH_DEBUG build (arena index assertions on every tree access; OTH_B200_DEBUG=1 loads it) is exercised by
    running this whole file under that variable once per round; here: whichever build is loaded reports no
    OTH_ERR_DEBUG over a few hundred launches of the production pair."""
    e, args = _stub_engine(128, 32, 8, True, seed=1, salt=2, move_launch=1)
    _run_to_done(e)
    assert (e.ctl()["error"] == 0).all()
    assert os.environ.get("OTH_B200_DEBUG", "") in ("", "0") or "debug" in os.path.basename(_lib_path())


def _lib_path():
    from alphazero_othello_b200 import _lib
    return _lib.lib()._name
