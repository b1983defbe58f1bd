"""Diagnosis: does collect_self_play_games-style self-play (restarting games) finish for every kernel variant?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from alphazero_othello_b200 import _lib
from alphazero_othello_b200.Models import FastOthelloNet, fold_for_inference
from alphazero_othello_b200.engine import BatchedPolicy, MctsEngine, SelfPlayRunner

args = {"c_puct": 2.0, "num_simulations": 8, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3, "mcts_temperature": 1.0,
        "num_exploratory_moves": 35, "lambda": 0.98}
torch.manual_seed(0)
net = fold_for_inference(FastOthelloNet(8, 65).cuda().eval(), torch.bfloat16)
for lanes in (32, 8):
    for ml in (1, 0):
        for gps in (1, 2):
            for graph in (True, False):
                e = MctsEngine(32, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=gps, lanes=lanes, move_launch=ml)
                run = SelfPlayRunner(e, BatchedPolicy(net, "cuda:0", torch.float32), use_graph=graph)
                run.warm_start()
                ok = False
                for k in range(60):
                    run.run_iterations(64)
                    c = e.counters()
                    if c["active"] == 0:
                        ok = True
                        break
                ctl = e.ctl()
                ml_buf = e._t[_lib.BUF_MOVE_LIST][:8].cpu().numpy().view(np.int32)
                flags = e._t[_lib.BUF_MOVE_FLAGS][:4].cpu().numpy().view(np.uint8)
                print(f"lanes {lanes} move_launch {ml} gps {gps} graph {graph}: finished={ok} after {(k + 1) * 64} iterations, games {c['games']}, "
                      f"phases {np.bincount(ctl['phase'], minlength=6).tolist()} errors {c['errors']} list_head {ml_buf[:6].tolist()} "
                      f"flags_set {int(flags.sum())}", flush=True)
                if not ok:
                    stuck = np.nonzero(ctl["phase"] == _lib.PH_MOVE)[0]
                    print("   stuck slots", stuck.tolist(), "sims_done", ctl["sims_done"][stuck].tolist(), "ply", ctl["ply"][stuck].tolist(),
                          "games_left", ctl["games_left"][stuck].tolist())
