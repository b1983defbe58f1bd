"""How many of the leaves a network batch evaluates are duplicates of one another?  (VERDICT r1 task 5.)

Plays whole games of a BASELINE config through the production pipeline and, every `--every` iterations, counts the
distinct canonical positions among the slots that wait for an evaluation (exact: packed (own, opp) bitboards,
torch.unique on the device).  Prints one JSON object: per-probe rows and the mean duplicate fraction, overall and by
game phase.  Evaluation de-duplication can save at most that fraction of the network's work.

    python tools/dedup_probe.py --workload c4 --every 25 > gpurun_out/dedup_c4.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table, network construction)
from alphazero_othello_b200 import _lib  # noqa: E402
from alphazero_othello_b200.Models import fold_for_inference  # noqa: E402
from alphazero_othello_b200.engine import BatchedPolicy, MctsEngine, SelfPlayRunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--every", type=int, default=25)
    ap.add_argument("--games", type=int, default=0)
    a = ap.parse_args()
    desc, kind, G, sims = bench.WORKLOADS[a.workload]
    G = a.games or G
    dev = torch.device("cuda:0")
    args = dict(bench.TRAIN_ARGS, num_simulations=sims)
    eng = MctsEngine(G, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=1, device=dev, out_pos_cap=G * 80,
                     out_game_cap=G + 64)
    run = SelfPlayRunner(eng, BatchedPolicy(fold_for_inference(bench.make_net(kind).to(dev), torch.bfloat16), dev, torch.float32))
    run.warm_start()
    sh = torch.arange(64, device=dev, dtype=torch.int64)
    ctl = eng._t[_lib.BUF_CTL][: G * 8].view(torch.int32).view(G, 16)
    rows, it = [], 0
    while True:
        run.run_iterations(a.every)
        it += a.every
        x = eng.nn_input.view(G, 64)
        own = (((x > 0.5).to(torch.int64)) << sh).sum(1)
        opp = (((x < -0.5).to(torch.int64)) << sh).sum(1)
        waiting = ctl[:, 0] == _lib.PH_WAIT_EVAL
        keys = torch.stack([own, opp], 1)[waiting]
        nw = int(keys.size(0))
        nu = int(torch.unique(keys, dim=0).size(0)) if nw else 0
        ply = float(ctl[:, 4][waiting].float().mean()) if nw else -1.0
        rows.append((it, nw, nu, round(ply, 2)))
        if it % (a.every * 40) == 0:
            c = eng.counters()
            if c["errors"]:
                eng.raise_on_error()
            if c["active"] == 0:
                break
    tot_w = sum(r[1] for r in rows)
    tot_u = sum(r[2] for r in rows)
    by_phase = {}
    for lo, hi in ((0, 4), (4, 8), (8, 16), (16, 32), (32, 48), (48, 128)):
        sel = [r for r in rows if lo <= r[3] < hi]
        w, u = sum(r[1] for r in sel), sum(r[2] for r in sel)
        by_phase[f"ply {lo}-{hi}"] = {"probes": len(sel), "duplicate_fraction": (1 - u / w) if w else None}
    print(json.dumps({"workload": f"{a.workload}: {desc}", "games": G, "sims_per_move": sims, "probe_every_iterations": a.every,
                      "probes": len(rows), "leaves_probed": tot_w, "distinct": tot_u, "duplicate_fraction": 1 - tot_u / max(tot_w, 1),
                      "by_game_phase": by_phase, "rows_it_waiting_unique_meanply": rows[:: max(1, len(rows) // 200)]}))


if __name__ == "__main__":
    main()
