"""Experiment: does splitting the concurrent games of one GPU into several engines (chunks) pay?

  python tools/chunk_probe.py [c4|c3] [iters]

For each chunk count the same G games are played by `chunks` MctsEngine instances of G/chunks slots
(game ids stay global, so the games are the same ones).  Per iteration every chunk runs
{network forward over its leaf batch; oth_mcts_step_fused}.  Two layouts:
  seq     all chunks back to back on one stream (activations of a chunk stay L2-resident)
  2stream chunks alternate between two streams inside one CUDA graph (the HBM-bound epilogue /
          MCTS kernels of one chunk can overlap the tensor-core convolutions of the other)
Prints simulations/s per layout; the single-engine line is what bench.py runs today.
"""
import sys
import json

sys.path.insert(0, ".")
import torch

from bench import TRAIN_ARGS, WORKLOADS, make_net
from alphazero_othello_b200 import _lib
from alphazero_othello_b200.Models import fold_for_inference
from alphazero_othello_b200.engine import BatchedPolicy, MctsEngine


def run(kind, G, sims, chunks, n_streams, iters, dev):
    args = dict(TRAIN_ARGS, num_simulations=sims)
    net = make_net(kind).to(dev)
    Gc = G // chunks
    engs = [MctsEngine(Gc, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=-1, device=dev,
                       game_id_base=j * Gc, game_id_stride=G, out_pos_cap=Gc * 80, out_game_cap=Gc + 64) for j in range(chunks)]
    evs = [BatchedPolicy(fold_for_inference(net, torch.bfloat16), dev, torch.float32) for _ in range(n_streams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    for e in engs:
        e.reset()
        e.step()

    def iteration(main):
        for s in streams:
            s.wait_stream(main)
        for j, e in enumerate(engs):
            k = j % n_streams
            with torch.cuda.stream(streams[k]):
                lg, v = evs[k].raw(e.nn_input)
                e.step_fused(lg, v)
        for s in streams:
            main.wait_stream(s)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            iteration(side)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        iteration(torch.cuda.current_stream(dev))
    for _ in range(40):
        g.replay()
    torch.cuda.synchronize(dev)
    c0 = sum(e.counters()["sims"] for e in engs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    c1 = sum(e.counters()["sims"] for e in engs)
    for e in engs:
        e.raise_on_error()
    return (c1 - c0) / (ms * 1e-3), ms / iters


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    _, kind, G, sims = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    res = []
    for chunks, ns in ((1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (2, 2), (4, 2), (8, 2), (16, 2), (8, 4)):
        if G // chunks < 128:
            continue
        try:
            r, ms = run(kind, G, sims, chunks, ns, iters, dev)
        except Exception as ex:  # keep going: this is an experiment
            print("chunks", chunks, "streams", ns, "FAILED", repr(ex)[:200], flush=True)
            continue
        torch.cuda.empty_cache()
        res.append({"workload": wl, "chunks": chunks, "streams": ns, "sims_per_s": r, "ms_per_iteration": ms})
        print(json.dumps(res[-1]), flush=True)
    json.dump(res, open(f"gpurun_out/chunk_probe_{wl}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
