"""Record / replay checker for the engine's network path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The parity contract is "bit-exact given identical network outputs" (BASELINE.json north_star).  This
module drives the PRODUCTION pipeline -- ``SelfPlayRunner`` with a real network twin, ``oth_mcts_step_fused`` +
the move kernel, optionally replayed as a CUDA graph -- one iteration at a time, records for every pending leaf the
priors / value the step kernel consumed (``record=True``: what its own softmax / tanh produced), and then replays
every game through the CPU oracle from the recorded draws: states, policy targets and value targets must be
bit-identical, simulation totals equal, no table miss.  It also checks that the network answered identical positions
identically whatever their batch row (``table.stats[2] == 0``).

Used by tests/test_production_path_gpu.py and __graft_entry__.smoke(); never imported by the product.
"""
import numpy as np

from . import EvalTable, Evaluator, self_play


def record_self_play(runner, table, poll_every=32, max_iterations=200000):
    """Play every slot's games to the end, one iteration (network forward + step kernel + move kernel) per host
    step, feeding every consumed (leaf -> priors, value) into ``table``.  Returns the number of iterations."""
    import torch
    from alphazero_othello_b200 import _lib
    e = runner.e
    assert runner.record and runner.external
    runner.warm_start()
    n = e.n_slots
    it = 0
    ctl_view = e._t[_lib.BUF_CTL][: n * 8]
    while it < max_iterations:
        phase = ctl_view.cpu().numpy().view(np.int32).reshape(n, 16)[:, 0]
        planes = e.nn_input.view(n, 64).cpu().numpy()
        runner.run_iterations(1)
        torch.cuda.synchronize(e.device)
        table.put_planes(planes, phase == _lib.PH_WAIT_EVAL, e.priors.cpu().numpy(), e.values.cpu().numpy())
        it += 1
        if it % poll_every == 0:
            c = e.counters()
            if c["errors"]:
                e.raise_on_error()
            if c["active"] == 0:
                return it
    raise AssertionError("self-play did not finish")


def replay_and_compare(engine, args, table, out=None, games=None):
    """Replay finished games through the oracle with the table as evaluator.  ``games``: iterable of game ids
    (= slot index for one game per slot), default all.  Returns (games checked, total oracle simulations)."""
    from alphazero_othello_b200.engine import split_games
    noise = engine.noise.cpu().numpy()
    um, ut = engine.u_move.cpu().numpy(), engine.u_tie.cpu().numpy()
    out = engine.drain() if out is None else out
    trajs = split_games(out)
    ids = sorted(int(g[0]) for g in out["games"].numpy())
    base = int(engine.cfg.game_id_base)
    assert ids == list(range(base, base + engine.n_slots)), "one finished game per slot expected"
    ev = Evaluator(table=table)
    checked = sims = 0
    for g in (range(engine.n_slots) if games is None else games):
        ref = self_play(args, ev, noise[g], um[g], ut[g])
        traj = trajs[g]
        assert len(traj) == len(ref["values"]), (g, len(traj), len(ref["values"]))
        assert np.array_equal(np.stack([t[0] for t in traj]), ref["states"]), g
        assert np.array_equal(np.stack([t[1] for t in traj]), ref["pis"]), g
        assert np.array_equal(np.array([t[2] for t in traj]), ref["values"]), g
        sims += ref["counters"]["sims"]
        checked += 1
    assert table.misses == 0, "the oracle asked for a position the engine never evaluated"
    return checked, sims
