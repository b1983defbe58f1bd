"""Complete self-play batches (every game played to the end) for C3 / C4: whole-game throughput,
arena high-water mark, depth -- the numbers DESIGN.md quotes beside the short default bench."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from alphazero_othello_b200 import _lib
from alphazero_othello_b200.Models import fold_for_inference
from alphazero_othello_b200.engine import BatchedPolicy, MctsEngine, SelfPlayRunner
from bench import TRAIN_ARGS, WORKLOADS, make_net

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
gps = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # games per slot: > 1 shows the desynchronised steady state
dedup = (sys.argv[3] != "0") if len(sys.argv) > 3 else "auto"  # evaluation de-duplication: auto (default) / 1 / 0
desc, kind, G, sims = WORKLOADS[wl]
args = dict(TRAIN_ARGS, num_simulations=sims)
dev = torch.device("cuda:0")
net = fold_for_inference(make_net(kind).to(dev), torch.bfloat16)
eng = MctsEngine(G, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=gps, device=dev, seed=1,
                 out_pos_cap=G * gps * 72, out_game_cap=G * gps + 16)
run = SelfPlayRunner(eng, BatchedPolicy(net, dev, torch.float32), dedup=dedup)
torch.cuda.synchronize()
t0 = time.perf_counter()
run.warm_start()
it, max_top = 0, 0
trace = []  # (seconds, sims, games finished) every 1024 iterations
while True:
    run.run_iterations(256)
    it += 256
    c = eng.counters()
    max_top = max(max_top, c["max_top"])
    if it % 1024 == 0:
        torch.cuda.synchronize()
        trace.append((round(time.perf_counter() - t0, 3), c["sims"], c["games"]))
    if c["errors"]:
        eng.raise_on_error()
    if c["active"] == 0:
        break
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out = eng.drain()
c = eng.counters()
print(json.dumps({"workload": f"{wl}: {desc}", "evaluation_dedup": bool(run.dedup),
                  "iterations_by_bucket": getattr(run, "bucket_iterations", None), "network_rows_evaluated": run.rows_evaluated,
                  "iteration_ms_by_variant": {str(k): round(v, 4) for k, v in getattr(run, "iteration_ms", {}).items()} or None,
                  "complete_games": int(out["games"].shape[0]), "positions": int(out["values"].numel()),
                  "seconds": dt, "iterations": it, "sims": c["sims"], "sims_per_s": c["sims"] / dt, "positions_per_s": out["values"].numel() / dt,
                  "terminal_sim_fraction": c["terminal_sims"] / c["sims"], "mean_plies": out["values"].numel() / out["games"].shape[0],
                  "arena_high_water": max_top, "node_cap": eng.cfg.node_cap, "max_depth": c["max_depth"],
                  "nodes_created": c["nodes"], "nodes_copied_by_reroot": c["copied"],
                  "interval_sims_per_s": [round((b[1] - a[1]) / (b[0] - a[0])) for a, b in zip(trace, trace[1:])],
                  "games_finished_at": [t[2] for t in trace],
                  "mean_levels_per_sim": c["levels"] / c["sims"], "mean_children_per_level": c["children"] / max(c["levels"], 1)}))
