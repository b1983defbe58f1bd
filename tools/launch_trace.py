"""Per-launch trace of the MCTS step inside the C4 pipeline: launch time + counter deltas."""
import json, sys
import torch
sys.path.insert(0, ".")
from alphazero_othello_b200 import _lib
from alphazero_othello_b200.Models import fold_for_inference
from alphazero_othello_b200.engine import BatchedPolicy, MctsEngine, SelfPlayRunner
from bench import TRAIN_ARGS, make_net
dev = torch.device("cuda:0")
G, sims = 16384, 400
net = fold_for_inference(make_net("big").to(dev), torch.bfloat16)
ev = BatchedPolicy(net, dev, torch.float32)
eng = MctsEngine(G, dict(TRAIN_ARGS, num_simulations=sims), self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=-1, device=dev,
                 out_pos_cap=G * 80, out_game_cap=G + 64)
run = SelfPlayRunner(eng, ev)
run.warm_start()
run.run_iterations(400 * 6 + 380)
rows = []
prev = eng.counters()
for i in range(140):
    ev(eng.nn_input, eng.priors, eng.values)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(); e1.record()
    torch.cuda.synchronize()
    c = eng.counters()
    rows.append((round(e0.elapsed_time(e1) * 1000), c["sims"] - prev["sims"], c["levels"] - prev["levels"], c["children"] - prev["children"],
                 c["nodes"] - prev["nodes"], c["moves"] - prev["moves"], c["copied"] - prev["copied"], c["terminal_sims"] - prev["terminal_sims"]))
    prev = c
print("us, sims, levels, children, nodes, moves, copied, terminal")
for r in rows[::3]:
    print(r)
