"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): game-id sharding, the weight
broadcast and the replay gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_drain(rank, n_games):
    """What MctsEngine.drain returns, synthesised: games of different lengths tagged by id."""
    rs = np.random.RandomState(100 + rank)
    games, pis, values, meta, boards = [], [], [], [], []
    first = 0
    for j in range(n_games):
        gid = rank * 10 + j
        n = int(rs.randint(3, 9))
        games.append([gid, first, n, int(rs.randint(-1, 2))])
        for t in range(n):
            pis.append(rs.rand(65).astype(np.float32))
            values.append(rs.randn())
            meta.append((gid << 16) | (t << 8) | 1)
            boards.append([int(rs.randint(0, 1 << 62)), int(rs.randint(0, 1 << 62))])
        first += n
    return dict(pis=torch.tensor(np.array(pis)), values=torch.tensor(values, dtype=torch.float64),
                meta=torch.tensor(meta, dtype=torch.int64), boards=torch.tensor(boards, dtype=torch.int64),
                games=torch.tensor(games, dtype=torch.int64))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alphazero_othello_b200 import parallel
    from alphazero_othello_b200.Models import FastOthelloNet
    torch.manual_seed(rank)  # different initial weights per rank
    net = FastOthelloNet(8, 65)
    ver = parallel.broadcast_weights(net, src=0, version=7 if rank == 0 else -1)
    flat, layout = parallel.flatten_state(net.state_dict())
    merged = parallel.gather_replay(_fake_drain(rank, 2 + rank), dst=0)
    q.put((rank, ver, float(flat.double().sum()), flat.numel(), None if merged is None else {k: v.numpy() for k, v in merged.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_replay_gather_gloo():
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, v0, s0, n0, m0), (r1, v1, s1, n1, m1) = res
    assert v0 == v1 == 7 and n0 == n1 == 641026 and s0 == s1  # 640 514 params + BN running stats
    assert m1 is None and m0 is not None
    exp = [_fake_drain(0, 2), _fake_drain(1, 3)]
    assert len(m0["games"]) == 5 and len(m0["values"]) == sum(len(e["values"]) for e in exp)
    # every game's slice of the merged arrays equals what its rank produced
    for gid, first, n, w in m0["games"]:
        e = exp[gid // 10]
        row = [g for g in e["games"].numpy() if g[0] == gid][0]
        sl = slice(row[1], row[1] + row[2])
        assert n == row[2] and w == row[3]
        assert np.array_equal(m0["pis"][first:first + n], e["pis"].numpy()[sl])
        assert np.array_equal(m0["values"][first:first + n], e["values"].numpy()[sl])
        assert np.array_equal(m0["boards"][first:first + n], e["boards"].numpy()[sl])
        assert np.array_equal(m0["meta"][first:first + n] >> 16, np.full(n, gid))


def test_game_id_sharding_covers_every_game_once():
    from alphazero_othello_b200.parallel import shard_game_ids
    for world in (1, 2, 4, 8):
        n_slots, rounds = 16, 3
        ids = []
        for r in range(world):
            base, stride = shard_game_ids(r, world, n_slots)
            ids += [base + s + k * stride for s in range(n_slots) for k in range(rounds)]
        assert sorted(ids) == list(range(world * n_slots * rounds))


def test_models_match_reference_when_available():
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "Models.py")):
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_models", os.path.join(ref, "Models.py"))
    R = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(R)
    import alphazero_othello_b200.Models as M
    for rc, mc in ((R.FastOthelloNet, M.FastOthelloNet), (R.AlphaZeroNet, M.AlphaZeroNet)):
        torch.manual_seed(0)
        r = rc(8, 65).eval()
        m = mc(8, 65).eval()
        m.load_state_dict(r.state_dict())  # same parameter names: checkpoints are interchangeable
        x = torch.randint(-1, 2, (4, 1, 8, 8)).float()
        (a1, b1), (a2, b2) = r(x), m(x)
        assert torch.equal(a1, a2) and torch.equal(b1, b2) and r.get_config() == m.get_config()
        s = x[0, 0].numpy().astype(np.int8)
        p1, v1 = r.inference(s, -1)
        p2, v2 = m.inference(s, -1)
        assert np.array_equal(p1, p2) and v1 == v2


def test_dedup_bucket_policy_host_logic():
    """SelfPlayRunner._choose_bucket (host side of evaluation de-duplication): cost per served evaluation, from the
    (fixed + rows) model before the graphs are timed and from measured iteration times afterwards."""
    from types import SimpleNamespace
    from alphazero_othello_b200.engine import SelfPlayRunner, default_lanes
    n = 16384
    buckets = sorted({n * k // 8 for k in range(1, 8)} | {n // 16, n // 32}, reverse=True)
    me = SimpleNamespace(force_bucket=None, e=SimpleNamespace(n_slots=n), _recent=[n, n], iteration_ms={}, use_measured_times=True,
                         DEDUP_BLOCK=16, buckets=buckets, fixed_cost_rows=n // 10)
    choose = lambda u: (setattr(me, "_recent", [u, u]), SelfPlayRunner._choose_bucket(me))[1]
    assert choose(n) == 0 and choose(16000) == 0          # no duplicates: the whole batch
    assert choose(1) == 512 and choose(400) == 512        # opening: the smallest bucket
    assert choose(5400) in (4096, 6144) and choose(9000) == 8192
    me._recent = [100, 9000]                              # the larger of the last two counts decides
    assert SelfPlayRunner._choose_bucket(me) == 8192
    # measured times: a compacted batch that is not faster than the whole one is never chosen
    me.iteration_ms = {"plain": 0.130, 0: 0.160, **{b: 0.135 for b in buckets}}
    assert choose(300) == 0
    me.iteration_ms = {"plain": 2.87, 0: 2.90, **{b: 0.25 + 2.62 * b / n for b in buckets}}
    assert choose(300) == 512 and choose(n) == 0 and choose(14800) in (14336, 0)
    me.force_bucket = 2048
    assert choose(n) == 2048
    # lanes: a warp per slot while that is one wave, 8 lanes at the C4 geometry
    assert default_lanes(1) == 32 and default_lanes(4096) == 32 and default_lanes(8192) == 16 and default_lanes(16384) == 8
