"""Multi-GPU plumbing: one process per GPU (torchrun), games sharded by id, NO collective on the
search path.  NCCL (gloo in the CPU tests) is used only where the reference ships data between
processes: the network weights going out (pickled ``state_dict`` per task, train.py:207-217) and
the replay tuples coming back (``Trainer._extend_buffer``, train.py:136-140)."""
import torch
import torch.distributed as dist


def shard_game_ids(rank, world, n_slots):
    """(game_id_base, game_id_stride) for a rank: slot s of rank r plays global games
    r*n_slots + s + k*(n_slots*world), k = 0, 1, ...  A game's RNG streams are keyed by its global
    id, so the set of games is independent of the world size."""
    return rank * n_slots, n_slots * world


def flatten_state(state_dict):
    """All floating-point entries of a state_dict as one flat float32 tensor (+ the layout)."""
    items = [(k, v) for k, v in state_dict.items() if v.dtype.is_floating_point]
    flat = torch.cat([v.detach().reshape(-1).float() for _, v in items])
    return flat, [(k, tuple(v.shape)) for k, v in items]


def unflatten_into(flat, module):
    """Copy a flat weight buffer (flatten_state order) into ``module``'s parameters and buffers."""
    off = 0
    with torch.no_grad():
        for _k, v in module.state_dict().items():
            if v.dtype.is_floating_point:
                n = v.numel()
                v.copy_(flat[off:off + n].view_as(v))
                off += n
    assert off == flat.numel()


def broadcast_weights(module, src=0, version=0):
    """One broadcast of the flattened weights (2.56 MB small / 6.03 MB big net) + model version
    from ``src``; every rank then holds identical parameters.  Returns the version."""
    flat, _ = flatten_state(module.state_dict())
    dev = flat.device
    meta = torch.tensor([float(version)], dtype=torch.float32, device=dev)
    buf = torch.cat([flat, meta])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(buf, src)
    unflatten_into(buf[:-1], module)
    return int(buf[-1].item())


def gather_replay(out, dst=0, device=None):
    """Gather drained replay tuples (engine.MctsEngine.drain output) from every rank to ``dst``:
    sizes first, then one padded payload per rank.  Returns the merged dict on ``dst`` (None
    elsewhere); positions keep their (game_id, ply) tags, games their descriptors."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    device = device or out["values"].device
    n = torch.tensor([out["values"].numel(), out["games"].shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    max_n = max(int(s[0]) for s in sizes)
    max_g = max(int(s[1]) for s in sizes)
    k, g = int(n[0]), int(n[1])
    # float64 payload: pi (65) | value | meta | own | opp ; ints carried bit-exactly as int64 views
    pay_f = torch.zeros((max(max_n, 1), 66), dtype=torch.float64, device=device)
    pay_i = torch.zeros((max(max_n, 1), 3), dtype=torch.int64, device=device)
    pay_g = torch.zeros((max(max_g, 1), 4), dtype=torch.int64, device=device)
    if k:
        pay_f[:k, :65] = out["pis"].to(device).double()
        pay_f[:k, 65] = out["values"].to(device)
        pay_i[:k, 0] = out["meta"].to(device)
        pay_i[:k, 1:3] = out["boards"].to(device)
    if g:
        pay_g[:g] = out["games"].to(device)
    lf = [torch.zeros_like(pay_f) for _ in range(world)] if rank == dst else None
    li = [torch.zeros_like(pay_i) for _ in range(world)] if rank == dst else None
    lg = [torch.zeros_like(pay_g) for _ in range(world)] if rank == dst else None
    dist.gather(pay_f, lf, dst)
    dist.gather(pay_i, li, dst)
    dist.gather(pay_g, lg, dst)
    if rank != dst:
        return None
    pis, values, meta, boards, games = [], [], [], [], []
    base = 0
    for r in range(world):
        kr, gr = int(sizes[r][0]), int(sizes[r][1])
        pis.append(lf[r][:kr, :65].float())
        values.append(lf[r][:kr, 65])
        meta.append(li[r][:kr, 0])
        boards.append(li[r][:kr, 1:3])
        gd = lg[r][:gr].clone()
        gd[:, 1] += base  # first-position offsets now index the merged arrays
        games.append(gd)
        base += kr
    return dict(pis=torch.cat(pis), values=torch.cat(values), meta=torch.cat(meta), boards=torch.cat(boards),
                games=torch.cat(games))
