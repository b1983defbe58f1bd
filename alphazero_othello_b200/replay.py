"""Replay-buffer ingest on the GPU: the caller of the self-play path.

``aggregate_duplicates`` is ``Trainer._aggregate_duplicates`` (train.py:142-173) as one CUDA
pipeline over packed replay tuples -- the form ``MctsEngine.drain`` already produces -- instead
of a Python dict keyed by sha1 hashes.  ``to_training_arrays`` gives the float32 arrays
``Trainer.setup_dataloader`` (train.py:175-197) builds from it.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def aggregate_duplicates(boards, pis, values, versions, device="cuda:0"):
    """boards int64[n,2] (own, opp canonical), pis f32[n,65], values f64[n], versions int32[n]
    (torch tensors, host or device).  Returns dict(boards int64[m,2], pis f32[m,65],
    values f32[m], versions int32[m], counts int32[m]) on the device, in first-occurrence order."""
    _lib.require_device()
    dev = torch.device(device)
    boards = boards.to(dev, torch.int64).contiguous()
    pis = pis.to(dev, torch.float32).contiguous()
    values = values.to(dev, torch.float64).contiguous()
    versions = versions.to(dev, torch.int32).contiguous()
    n = int(values.numel())
    assert boards.shape == (n, 2) and pis.shape == (n, 65) and versions.numel() == n
    L = _lib.lib()
    nb = C.c_int64(0)
    _lib.check(L.oth_replay_aggregate_workspace_bytes(n, C.byref(nb)))
    ws = torch.empty(max(nb.value, 8), dtype=torch.uint8, device=dev)
    ob = torch.empty((max(n, 1), 2), dtype=torch.int64, device=dev)
    op = torch.empty((max(n, 1), 65), dtype=torch.float32, device=dev)
    ov = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    over = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    oc = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    om = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.oth_replay_aggregate(boards.data_ptr(), pis.data_ptr(), values.data_ptr(), versions.data_ptr(), n,
                                          ws.data_ptr(), ws.numel(), ob.data_ptr(), op.data_ptr(), ov.data_ptr(), over.data_ptr(),
                                          oc.data_ptr(), om.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "oth_replay_aggregate")
    m = int(om.item())
    return dict(boards=ob[:m], pis=op[:m], values=ov[:m], versions=over[:m], counts=oc[:m])


def to_training_arrays(agg):
    """(states f32[m,8,8], policies f32[m,65], values f32[m,1]) as train.py:184-186 builds them."""
    m = agg["values"].numel()
    dev = agg["values"].device
    states = torch.empty((m, 8, 8), dtype=torch.int8, device=dev)
    if m:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().oth_unpack_canonical(agg["boards"].contiguous().data_ptr(), states.data_ptr(), m,
                                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return states.float(), agg["pis"], agg["values"].reshape(-1, 1)


def augment_batch(states, policies, device="cuda:0", generator=None):
    """Training-time augmentation of a COLLATED batch in the training process -- the replacement of the
    per-sample ``get_random_symmetry`` call in ``RandomSymmetryDataset.__getitem__`` (train.py:26-42), which
    the reference runs inside forked DataLoader workers where CUDA cannot be used.  The dataset yields the raw
    ``(state [8,8], policy [65], value)`` rows (any worker count); the training loop calls this on each batch:

        for state, pi, v in dataloader:                       # raw rows, workers never touch CUDA
            state, pi = augment_batch(state, pi, device)      # one random dihedral image per sample, on the GPU

    states: [B,8,8] or [B,1,8,8] (any real dtype, values in {-1,0,1}); policies: float [B,65].
    Returns (float32 [B,1,8,8], float32 [B,65]) on ``device`` -- the shapes/dtypes the per-sample path collates to."""
    from .envs.othello import BatchedOthello
    dev = torch.device(device)
    st = states.to(dev).reshape(-1, 8, 8).to(torch.int8).contiguous()
    pi = policies.to(dev, torch.float32).contiguous()
    return BatchedOthello(dev).random_symmetry(st, pi, generator=generator)


def pack_states(states):
    """int8 canonical boards [n,8,8] (numpy) -> int64[n,2] (own, opp), bit i = row*8+col."""
    s = np.asarray(states).reshape(len(states), 64)
    w = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    own = ((s == 1) * w).sum(1, dtype=np.uint64)
    opp = ((s == -1) * w).sum(1, dtype=np.uint64)
    return torch.from_numpy(np.stack([own, opp], 1).view(np.int64))


class ReplayBuffer:
    """GPU-resident replay buffer with the semantics of the reference's ``deque(maxlen=...)`` of
    ``(state, pi, value, model_version)`` tuples (train.py:77-82, 136-140): ``extend`` appends the
    drained self-play output tagged with the current model version, the oldest tuples fall out when
    the capacity is exceeded, ``aggregate`` is ``_aggregate_duplicates`` over the buffer in
    chronological order (oldest first, which fixes the bucket order)."""

    def __init__(self, capacity, device="cuda:0"):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.boards = torch.zeros((self.capacity, 2), dtype=torch.int64, device=self.device)
        self.pis = torch.zeros((self.capacity, 65), dtype=torch.float32, device=self.device)
        self.values = torch.zeros(self.capacity, dtype=torch.float64, device=self.device)
        self.versions = torch.zeros(self.capacity, dtype=torch.int32, device=self.device)
        self.head = 0  # next write position
        self.size = 0

    def __len__(self):
        return self.size

    def extend(self, drained, version):
        """``drained``: output of ``MctsEngine.drain`` (or any dict with boards / pis / values)."""
        b = drained["boards"].to(self.device, torch.int64)
        p = drained["pis"].to(self.device, torch.float32)
        v = drained["values"].to(self.device, torch.float64)
        n = int(v.numel())
        if n > self.capacity:  # only the newest `capacity` tuples can survive
            b, p, v, n = b[-self.capacity:], p[-self.capacity:], v[-self.capacity:], self.capacity
        idx = (self.head + torch.arange(n, device=self.device)) % self.capacity
        self.boards[idx], self.pis[idx], self.values[idx] = b, p, v
        self.versions[idx] = int(version)
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def chronological(self):
        """(boards, pis, values, versions) oldest first."""
        start = (self.head - self.size) % self.capacity
        idx = (start + torch.arange(self.size, device=self.device)) % self.capacity
        return self.boards[idx], self.pis[idx], self.values[idx], self.versions[idx]

    def aggregate(self):
        return aggregate_duplicates(*self.chronological(), device=self.device)

    # ------------------------------------------------------- on-disk format
    def to_reference_tuples(self):
        """The buffer as the reference keeps it in memory and on disk: ``deque(maxlen=capacity)`` of
        ``(state int8[8,8] canonical, policy_target float32[65], value_target float, version int)``
        (train.py:77-82, README.md:85-88), oldest first."""
        from collections import deque
        b, p, v, ver = self.chronological()
        n = int(v.numel())
        states = torch.empty((n, 8, 8), dtype=torch.int8, device=self.device)
        if n:
            _lib.require_device()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().oth_unpack_canonical(b.contiguous().data_ptr(), states.data_ptr(), n,
                                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        states, p, v, ver = states.cpu().numpy(), p.cpu().numpy(), v.cpu().numpy(), ver.cpu().numpy()
        return deque(((states[i], p[i], float(v[i]), int(ver[i])) for i in range(n)), maxlen=self.capacity)

    def save(self, path):
        """``Trainer._save_replay_buffer`` (train.py:104-111): pickle of the whole deque, atomic replace."""
        import os
        import pickle
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            pickle.dump(self.to_reference_tuples(), f, protocol=pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, path)

    def load(self, path):
        """``Trainer._load_replay_buffer`` (train.py:113-134): replaces the contents with the pickled
        tuples (keeping this buffer's capacity: the newest survive) and returns the latest model
        version found (0 for an empty file or 3-tuples).  Accepts replay files written by the reference."""
        import pickle
        with open(path, "rb") as f:
            loaded = list(pickle.load(f))
        self.head = self.size = 0
        if not loaded:
            return 0
        loaded = loaded[-self.capacity:]
        has_ver = [isinstance(t, (tuple, list)) and len(t) >= 4 for t in loaded]
        vers = np.array([int(t[3]) if h else 0 for t, h in zip(loaded, has_ver)], np.int32)
        n = len(loaded)
        self.boards[:n] = pack_states(np.stack([np.asarray(t[0]).reshape(8, 8) for t in loaded])).to(self.device)
        self.pis[:n] = torch.from_numpy(np.stack([np.asarray(t[1], np.float32) for t in loaded])).to(self.device)
        self.values[:n] = torch.tensor([float(t[2]) for t in loaded], dtype=torch.float64, device=self.device)
        self.versions[:n] = torch.from_numpy(vers).to(self.device)
        self.head, self.size = n % self.capacity, n
        return int(max((int(t[3]) for t, h in zip(loaded, has_ver) if h), default=0))
