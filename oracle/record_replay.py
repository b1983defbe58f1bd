"""Record / replay checker for the engine's network path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The parity contract is "bit-exact given identical network outputs" (BASELINE.json north_star).  This
module drives the PRODUCTION pipeline -- ``SelfPlayRunner`` with a real network twin, ``oth_mcts_step_fused`` +
the move kernel, optionally replayed as a CUDA graph -- one iteration at a time, records for every pending leaf the
priors / value the step kernel consumed (``record=True``: what its own softmax / tanh produced), and then replays
every game through the CPU oracle from the recorded draws: states, policy targets and value targets must be
bit-identical, simulation totals equal.  The record is a TAPE per game slot -- the k-th evaluation the oracle asks for
is answered with the k-th output the engine consumed for that game, and the positions must agree -- so the check also
pins the ORDER in which leaves are evaluated, and it does not assume that the network answers one position identically
in every batch row (a position-keyed table, kept as an option, measures exactly that).

Used by tests/test_production_path_gpu.py and __graft_entry__.smoke(); never imported by the product.
"""
import numpy as np

from . import Evaluator, self_play


def record_self_play(runner, tape, table=None, poll_every=32, max_iterations=200000):
    """Play every slot's game to the end, one iteration (network forward + step kernel + move kernel) per host step,
    appending every consumed (leaf -> priors, value) to the slot's ``tape`` (and, optionally, to a position-keyed
    ``table`` whose stats[2] then counts positions the network answered differently in different batch rows).
    Returns the number of iterations."""
    from alphazero_othello_b200 import _lib
    e = runner.e
    assert runner.record and runner.external
    runner.warm_start()
    n = e.n_slots
    ctl_view = e._t[_lib.BUF_CTL][: n * 8]
    snap = {}
    done = [0]

    def before():  # which slots wait for an evaluation, and for which position (.cpu() synchronises the current stream)
        snap["waiting"] = ctl_view.cpu().numpy().view(np.int32).reshape(n, 16)[:, 0] == _lib.PH_WAIT_EVAL
        snap["planes"] = e.nn_input.view(n, 64).cpu().numpy()

    def after():   # what the step kernel consumed for them
        if runner.last_map is not None:  # de-duplicated batch: slots whose position missed the bucket were not served
            snap["waiting"] = snap["waiting"] & (runner.last_map.cpu().numpy() >= 0)
        pr, va = e.priors.cpu().numpy(), e.values.cpu().numpy()
        tape.put_planes(snap["planes"], snap["waiting"], pr, va)
        if table is not None:
            table.put_planes(snap["planes"], snap["waiting"], pr, va)
        done[0] += 1

    runner.before_iteration, runner.after_iteration = before, after  # also sees the warm-up iterations of the graph capture
    next_poll = poll_every
    try:
        while done[0] < max_iterations:
            runner.run_iterations(1)
            if done[0] >= next_poll:
                next_poll = done[0] + poll_every
                c = e.counters()
                if c["errors"]:
                    e.raise_on_error()
                if c["active"] == 0:
                    return done[0]
    finally:
        runner.before_iteration = runner.after_iteration = None
    raise AssertionError("self-play did not finish")


def replay_and_compare(engine, args, tape, out=None, games=None):
    """Replay finished games through the oracle, slot s served by its own tape.  ``games``: iterable of slots (one game
    per slot), default all.  Returns (games checked, total oracle simulations)."""
    from alphazero_othello_b200.engine import split_games
    noise = engine.noise.cpu().numpy()
    um, ut = engine.u_move.cpu().numpy(), engine.u_tie.cpu().numpy()
    out = engine.drain() if out is None else out
    trajs = split_games(out)
    ids = sorted(int(g[0]) for g in out["games"].numpy())
    base = int(engine.cfg.game_id_base)
    assert ids == list(range(base, base + engine.n_slots)), "one finished game per slot expected"
    checked = sims = 0
    for g in (range(engine.n_slots) if games is None else games):
        ref = self_play(args, Evaluator(tape=tape, slot=g), noise[g], um[g], ut[g])
        traj = trajs[g]
        assert tape.mismatches == 0, f"slot {g}: the oracle asked for a position out of the engine's evaluation order"
        assert tape.consumed(g) == tape.recorded(g), (g, tape.consumed(g), tape.recorded(g))
        assert len(traj) == len(ref["values"]), (g, len(traj), len(ref["values"]))
        assert np.array_equal(np.stack([t[0] for t in traj]), ref["states"]), g
        assert np.array_equal(np.stack([t[1] for t in traj]), ref["pis"]), g
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"]), g
        sims += ref["counters"]["sims"]
        checked += 1
    return checked, sims
