"""Batched arena (eval.py:12-178) on the GPU against the oracle: every network output the two
engines consumed is recorded and the matches are replayed tree by tree on the CPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _net(seed):
    import torch

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(seed)
            self.body = torch.nn.Sequential(torch.nn.Linear(64, 80), torch.nn.Tanh(), torch.nn.Linear(80, 66))

        def forward(self, x):
            y = self.body(x.reshape(x.size(0), 64))
            return y[:, :65], torch.tanh(y[:, 65:66])
    return Net()


def test_batched_arena_replays_through_oracle():
    import torch
    import oracle as O
    from alphazero_othello_b200 import _lib
    from alphazero_othello_b200.eval import play_matches_batched
    args = {"c_puct": 3.0, "num_simulations": 20}
    tables = {}
    w = (1 << (63 - np.arange(64, dtype=np.uint64)))

    def record(side):
        t = tables.setdefault(id(side), O.EvalTable(1 << 18))
        e = side.eng
        ph = e.ctl()["phase"]
        x = e.nn_input.view(-1, 64).cpu().numpy()
        pr, va = e.priors.cpu().numpy(), e.values.cpu().numpy()
        for s in np.nonzero(ph == _lib.PH_WAIT_EVAL)[0]:
            assert t.put(int((w * (x[s] == 1)).sum()), int((w * (x[s] == -1)).sum()), pr[s], va[s]) in (0, 1)
        side._table = t

    np.random.seed(5)
    n = 10
    na, nb = _net(1), _net(2)
    results, log = play_matches_batched(na, nb, args, n, dtype=None, record=record)
    assert all(r in ("A", "B", "Draw") for r in results) and len(results) == n
    ta, tb = list(tables.values())  # insertion order: side A searches first (ply 0, even matches)
    g = O.OracleGame()
    for i in range(n):
        first_t, second_t = (ta, tb) if i % 2 == 0 else (tb, ta)
        trees = {1: O.OracleMCTS(3.0, 20, O.Evaluator(table=first_t)), -1: O.OracleMCTS(3.0, 20, O.Evaluator(table=second_t))}
        searched = {1: False, -1: False}
        s, pl = g.get_initial_state(), 1
        res = None
        for ply in range(200):
            probs = trees[pl].policy_improve_step(s, pl, 0.0, None, log["u_tie"][i, ply])
            searched[pl] = True
            a = int(np.argmax(probs))
            assert a == log["actions"][i, ply], (i, ply)
            s = g.get_next_state(s, a, pl)
            rew, done = g.get_value_and_terminated(s, a, pl)
            if done:
                res = "Draw" if rew == 0 else ("A" if (rew == 1) == (pl == 1) else "B")
                break
            for p2 in (1, -1):
                if searched[p2]:
                    trees[p2].make_move(a)
            pl = -pl
        if i % 2 == 1 and res != "Draw":
            res = "B" if res == "A" else "A"
        assert res == results[i] and log["plies"][i] == ply + 1
    assert ta.misses == 0 and tb.misses == 0


def test_evaluate_models_parallel_signature():
    from alphazero_othello_b200.Models import FastOthelloNet
    from alphazero_othello_b200.eval import evaluate_models_parallel
    import torch
    torch.manual_seed(0)
    a, b = FastOthelloNet(8, 65), FastOthelloNet(8, 65)
    wa, wb = evaluate_models_parallel(8, {"c_puct": 2.0, "num_simulations": 6},
                                      (FastOthelloNet, a.get_config(), a.state_dict()),
                                      (FastOthelloNet, b.get_config(), b.state_dict()), n_matches=8)
    assert 0.0 <= wa <= 1.0 and 0.0 <= wb <= 1.0 and wa + wb <= 1.0


def test_batched_arena_reproduces_reference_matches(golden_r2):
    """The match loop pinned to the reference itself: 11 matches that eval._run_one_match (eval.py:86-131, 134-178)
    played on hash-stub policies -- even and odd match indices, wins for both sides, one drawn game -- with the tie
    picks its np.random.choice(best_actions) calls made.  The batched arena, given the same device stubs and tie picks,
    must return the same action at every ply and the same result string.  (Each reference match used its own pair of
    stub salts; an engine has one salt, so match j is replayed as row j of a batch of j+1 concurrent matches.)"""
    from alphazero_othello_b200.eval import DeviceStubPolicy, play_matches_batched
    g = golden_r2
    args = {"c_puct": float(g["ar_cfg"][1]), "num_simulations": int(g["ar_cfg"][0])}
    seen = set()
    for j, (idx, sa, sb, _seed, plies) in enumerate(g["ar_meta"]):
        n = j + 1
        u = np.zeros((n, 128))
        u[j] = g["ar_u_tie"][j]
        results, log = play_matches_batched(DeviceStubPolicy("H", sa), DeviceStubPolicy("H", sb), args, n, u_tie=u,
                                            lanes=[8, 16, 32][j % 3])
        assert log["plies"][j] == plies and list(log["actions"][j, :plies]) == list(g["ar_actions"][j, :plies]), j
        assert results[j] == str(g["ar_results"][j]), j
        seen.add(results[j])
    assert seen == {"A", "B", "Draw"}
    # play_match alone (first tree = +1): match 0 of a one-match batch
    sa, sb = g["ar_direct_salts"]
    u = g["ar_direct_u_tie"][None, :]
    u = np.concatenate([u, np.zeros((1, 128 - u.shape[1]))], 1)
    results, log = play_matches_batched(DeviceStubPolicy("H", sa), DeviceStubPolicy("H", sb), args, 1, u_tie=u)
    k = len(g["ar_direct_actions"])
    assert list(log["actions"][0, :k]) == list(g["ar_direct_actions"]) and results[0] == str(g["ar_direct_result"][0])
