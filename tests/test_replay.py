"""Replay ingest: duplicate aggregation, train.py:142-173.
CPU: the numpy oracle against the reference's own output.  GPU: the CUDA pipeline against the
reference's output (bit-exact) and against the oracle on a large synthetic buffer."""
import numpy as np
import pytest

from oracle import replay as OR


def _buffer(golden):
    return [(s, p.copy(), float(v), int(ver)) for s, p, v, ver in zip(golden["rp_in_states"], golden["rp_in_pis"],
                                                                     golden["rp_in_values"], golden["rp_in_versions"])]


def test_oracle_matches_reference(golden):
    states, policies, values, counts, vers = OR.aggregate_duplicates(_buffer(golden))
    assert np.array_equal(np.stack(states), golden["rp_out_states"])
    assert np.array_equal(np.stack(policies), golden["rp_out_pis"])
    assert np.array_equal(np.array(values, np.float32), golden["rp_out_values"])
    assert sum(counts) == len(golden["rp_in_values"]) and max(counts) >= 3


@pytest.mark.gpu
def test_gpu_matches_reference_bit_exact(golden):
    import torch
    from alphazero_othello_b200 import replay
    agg = replay.aggregate_duplicates(replay.pack_states(golden["rp_in_states"]), torch.from_numpy(golden["rp_in_pis"]),
                                      torch.from_numpy(golden["rp_in_values"]), torch.from_numpy(golden["rp_in_versions"]))
    states, policies, values = replay.to_training_arrays(agg)
    assert states.dtype == torch.float32 and values.shape == (len(golden["rp_out_values"]), 1)
    assert np.array_equal(states.cpu().numpy().astype(np.int8), golden["rp_out_states"])
    assert np.array_equal(policies.cpu().numpy(), golden["rp_out_pis"])
    assert np.array_equal(values.cpu().numpy().ravel(), golden["rp_out_values"])
    assert int(agg["counts"].sum()) == len(golden["rp_in_values"])


@pytest.mark.gpu
def test_gpu_matches_oracle_on_large_buffer_with_heavy_duplication():
    import torch
    from alphazero_othello_b200 import replay
    rs = np.random.RandomState(3)
    uniq = rs.randint(-1, 2, size=(700, 8, 8)).astype(np.int8)
    n = 40000
    which = np.minimum((rs.pareto(0.8, n) * 3).astype(np.int64), 699)  # a few very hot states, long tail
    states = uniq[which]
    pis = rs.rand(n, 65).astype(np.float32)
    pis /= pis.sum(1, keepdims=True)
    values = rs.uniform(-1, 1, n)
    vers = rs.randint(0, 3, n).astype(np.int32)
    ref = OR.aggregate_duplicates([(s, p.copy(), float(v), int(k)) for s, p, v, k in zip(states, pis, values, vers)])
    agg = replay.aggregate_duplicates(replay.pack_states(states), torch.from_numpy(pis), torch.from_numpy(values), torch.from_numpy(vers))
    st, po, va = replay.to_training_arrays(agg)
    assert len(ref[0]) == st.shape[0]
    assert np.array_equal(st.cpu().numpy().astype(np.int8), np.stack(ref[0]))
    assert np.array_equal(po.cpu().numpy(), np.stack(ref[1]))
    assert np.array_equal(va.cpu().numpy().ravel(), np.array(ref[2], np.float32))
    assert np.array_equal(agg["counts"].cpu().numpy(), np.array(ref[3])) and np.array_equal(agg["versions"].cpu().numpy(), np.array(ref[4]))
    assert max(ref[3]) > 1000  # the hot bucket is walked sequentially, in buffer order


@pytest.mark.gpu
def test_gpu_empty_and_singleton():
    import torch
    from alphazero_othello_b200 import replay
    e = replay.aggregate_duplicates(torch.zeros((0, 2), dtype=torch.int64), torch.zeros((0, 65)), torch.zeros(0, dtype=torch.float64),
                                    torch.zeros(0, dtype=torch.int32))
    assert e["values"].numel() == 0
    pi = torch.rand(1, 65)
    one = replay.aggregate_duplicates(torch.tensor([[5, 9]]), pi, torch.tensor([0.25], dtype=torch.float64), torch.tensor([4], dtype=torch.int32))
    assert one["counts"].tolist() == [1] and one["versions"].tolist() == [4] and one["boards"].tolist() == [[5, 9]]
    p = pi[0].numpy().copy()
    p = (p / np.float32(1)) ; p /= p.sum() + 1e-12
    assert np.array_equal(one["pis"].cpu().numpy()[0], p.astype(np.float32))


@pytest.mark.gpu
def test_replay_buffer_matches_deque_semantics():
    """extend + maxlen eviction + aggregate == the reference's deque + _aggregate_duplicates (oracle restatement)."""
    import collections
    import torch
    from alphazero_othello_b200 import replay
    rs = np.random.RandomState(9)
    uniq = rs.randint(-1, 2, size=(40, 8, 8)).astype(np.int8)
    cap = 500
    buf = replay.ReplayBuffer(cap)
    ref = collections.deque(maxlen=cap)
    for it in range(7):
        n = int(rs.randint(50, 200))
        states = uniq[rs.randint(0, 40, n)]
        pis = rs.rand(n, 65).astype(np.float32)
        values = rs.uniform(-1, 1, n)
        buf.extend(dict(boards=replay.pack_states(states), pis=torch.from_numpy(pis), values=torch.from_numpy(values)), version=it // 2)
        ref.extend((s, p.copy(), float(v), it // 2) for s, p, v in zip(states, pis, values))
    assert len(buf) == len(ref) == cap
    exp = OR.aggregate_duplicates(list(ref))
    st, po, va = replay.to_training_arrays(buf.aggregate())
    assert np.array_equal(st.cpu().numpy().astype(np.int8), np.stack(exp[0]))
    assert np.array_equal(po.cpu().numpy(), np.stack(exp[1])) and np.array_equal(va.cpu().numpy().ravel(), np.array(exp[2], np.float32))


@pytest.mark.gpu
def test_replay_buffer_pickle_round_trip_in_the_reference_format(tmp_path):
    """save() writes what Trainer._save_replay_buffer writes (train.py:104-111): a pickled
    deque(maxlen) of (int8[8,8], float32[65], float, int); load() reads such a file back
    (train.py:113-134), returns the latest model version and honours this buffer's capacity."""
    import collections
    import pickle
    import torch
    from alphazero_othello_b200 import replay
    rs = np.random.RandomState(3)
    n = 300
    states = rs.randint(-1, 2, size=(n, 8, 8)).astype(np.int8)
    pis = rs.rand(n, 65).astype(np.float32)
    values = rs.uniform(-1, 1, n)
    buf = replay.ReplayBuffer(256)
    buf.extend(dict(boards=replay.pack_states(states[:200]), pis=torch.from_numpy(pis[:200]), values=torch.from_numpy(values[:200])), 3)
    buf.extend(dict(boards=replay.pack_states(states[200:]), pis=torch.from_numpy(pis[200:]), values=torch.from_numpy(values[200:])), 4)
    path = str(tmp_path / "replay.pkl")
    buf.save(path)
    with open(path, "rb") as f:
        dq = pickle.load(f)
    assert isinstance(dq, collections.deque) and dq.maxlen == 256 and len(dq) == 256
    for k, (s, p, v, ver) in enumerate(dq):  # the newest 256 of the 300, oldest first
        i = n - 256 + k
        assert s.dtype == np.int8 and s.shape == (8, 8) and np.array_equal(s, states[i])
        assert p.dtype == np.float32 and np.array_equal(p, pis[i])
        assert isinstance(v, float) and v == values[i] and isinstance(ver, int) and ver == (3 if i < 200 else 4)
    # a file as the reference writes it (plain deque of tuples, one legacy 3-tuple) into a smaller buffer
    ref = collections.deque(maxlen=500000)
    ref.extend((states[i], pis[i], float(values[i]), 7 + i // 100) for i in range(n))
    ref.append((states[0], pis[0], 0.5))
    with open(path, "wb") as f:
        pickle.dump(ref, f, protocol=pickle.HIGHEST_PROTOCOL)
    small = replay.ReplayBuffer(100)
    assert small.load(path) == 9 and len(small) == 100
    back = list(small.to_reference_tuples())
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] for a, b in zip(back, list(ref)[-100:]))
    assert [t[3] for t in back] == [9] * 99 + [0]
    small.extend(dict(boards=replay.pack_states(states[:5]), pis=torch.from_numpy(pis[:5]), values=torch.from_numpy(values[:5])), 10)
    assert len(small) == 100 and [t[3] for t in small.to_reference_tuples()][-6:] == [0, 10, 10, 10, 10, 10]
