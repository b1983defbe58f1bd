"""CPU oracle for the replay-ingest row -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates Trainer._aggregate_duplicates (train.py:142-173) with the same numpy operations in the
same order (float32 ``sum_pi += pi``, Python-float ``sum_v += v``, ``avg_pi /= avg_pi.sum() +
1e-12``), keyed by the board bytes themselves instead of their sha1 (train.py:100-102)."""
import numpy as np


def aggregate_duplicates(replay):
    """replay: iterable of (state int8[8,8], pi float32[65], v float, version)."""
    buckets = {}
    for (s, pi, v, ver) in replay:
        key = (np.asarray(s, dtype=np.int8).tobytes(), ver)
        b = buckets.get(key)
        if b is None:
            buckets[key] = {"state": s, "sum_pi": pi.copy(), "sum_v": v, "count": 1, "ver": ver}
        else:
            b["sum_pi"] += pi
            b["sum_v"] += v
            b["count"] += 1
    states, policies, values, counts, vers = [], [], [], [], []
    for b in buckets.values():
        cnt = b["count"]
        avg_pi = b["sum_pi"] / cnt
        avg_pi /= avg_pi.sum() + 1e-12
        states.append(b["state"])
        policies.append(avg_pi.astype(np.float32))
        values.append(np.float32(b["sum_v"] / cnt))
        counts.append(cnt)
        vers.append(b["ver"])
    return states, policies, values, counts, vers
