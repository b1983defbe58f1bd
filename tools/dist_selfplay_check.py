"""2+ GPU check of the sharded self-play plumbing (run with torchrun under `gpurun --gpus N`):
weights broadcast over NCCL, games sharded by id with no collective on the search path, replay
tuples gathered to rank 0 -- and the union must equal what ONE GPU produces for the same game ids."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from alphazero_othello_b200 import _lib, parallel
from alphazero_othello_b200.Models import FastOthelloNet
from alphazero_othello_b200.engine import MctsEngine


def play(n_slots, base, stride, gps, dev):
    args = {"c_puct": 2.0, "num_simulations": 24, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3, "mcts_temperature": 1.0,
            "num_exploratory_moves": 12, "lambda": 0.98}
    e = MctsEngine(n_slots, args, self_play=True, eval_kind=_lib.EVAL_STUB_H, games_per_slot=gps, device=dev, seed=9, stub_salt=2,
                   game_id_base=base, game_id_stride=stride)
    e.reset()
    while True:
        for _ in range(64):
            e.step()
        if e.counters()["active"] == 0:
            break
    e.raise_on_error()
    return e.drain(to_host=False)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # 1. weight broadcast: every rank starts from different weights, ends with rank 0's
    torch.manual_seed(100 + rank)
    net = FastOthelloNet(8, 65).to(dev)
    ver = parallel.broadcast_weights(net, src=0, version=3)
    flat, _ = parallel.flatten_state(net.state_dict())
    sums = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(sums, flat.double().sum().reshape(1))
    assert ver == 3 and all(float(s) == float(sums[0]) for s in sums), "weights differ after broadcast"
    # 2. sharded self-play + gather
    n_slots, gps = 48, 2
    base, stride = parallel.shard_game_ids(rank, world, n_slots)
    merged = parallel.gather_replay(play(n_slots, base, stride, gps, dev), dst=0, device=dev)
    if rank == 0:
        whole = play(n_slots * world, 0, n_slots * world, gps, dev)  # the same game ids on one GPU
        def by_game(o):
            d = {}
            for gid, first, n, w in o["games"].cpu().numpy():
                d[int(gid)] = (o["boards"][first:first + n].cpu().numpy(), o["pis"][first:first + n].cpu().numpy(),
                               o["values"][first:first + n].cpu().numpy(), int(w))
            return d
        a, b = by_game(merged), by_game(whole)
        assert sorted(a) == sorted(b) == list(range(n_slots * world * gps)), (len(a), len(b))
        for gid in a:
            for x, y in zip(a[gid][:3], b[gid][:3]):
                assert np.array_equal(x, y), gid
            assert a[gid][3] == b[gid][3]
        print(f"dist check ok: world={world}, {len(a)} games, {int(merged['values'].numel())} positions gathered to rank 0, "
              f"identical to the single-GPU run", flush=True)
    # 3. the packaged entry point with a real network
    from alphazero_othello_b200.self_play_worker import collect_self_play_games_distributed
    args = {"c_puct": 2.0, "num_simulations": 8, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3, "mcts_temperature": 1.0,
            "num_exploratory_moves": 35, "lambda": 0.98}
    games = collect_self_play_games_distributed(net, args, 32, n_slots=32)
    if rank == 0:
        assert len(games) == 32 * world and all(9 <= len(g) <= 128 and abs(g[0][1].sum() - 1) < 1e-5 for g in games)
        print(f"collect_self_play_games_distributed ok: {len(games)} games on rank 0", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
