"""Self-play on the B200 engine behind the reference's worker surface.

* ``one_self_play(args_tuple)`` -- same signature, RNG call order and return value as
  ``self_play_worker.py:38-88`` (one game, host loop over the GPU-resident tree).
* ``collect_self_play_games(policy, args, n_games, ...)`` -- the batched replacement of
  ``Trainer.collect_self_play_games`` (train.py:199-225): thousands of concurrent games,
  one network evaluation per simulation per game, moves sampled / trees re-rooted /
  lambda-returns formed in-kernel; returns the same per-game lists of
  ``(state int8[8,8], pi float32[65], value float)`` that ``Trainer._extend_buffer``
  (train.py:136-140) consumes.
"""
import numpy as np
import torch

from . import _lib
from .MCTS_model import MCTS
from .engine import BatchedPolicy, MctsEngine, SelfPlayRunner, split_games
from .envs.othello import OthelloGameNew as OthelloGame


def get_training_data(trajectory, winning_player, lambd=1.0):
    """Backward lambda-return over one game (self_play_worker.py:8-35); the batched engine
    computes the same recurrence in-kernel (emit_game in csrc/mcts_kernels.cu)."""
    out = [None] * len(trajectory)
    g_next, next_player = None, None
    for t in range(len(trajectory) - 1, -1, -1):
        state, pi, player, v_root = trajectory[t]
        z = 0.0 if winning_player == 0 else (1.0 if player == winning_player else -1.0)
        if g_next is None:
            g = z
        else:
            sign = 1.0 if player == next_player else -1.0
            g = (1.0 - lambd) * v_root + lambd * sign * g_next
        out[t] = (state, pi, g)
        g_next, next_player = g, player
    return out


@torch.no_grad()
def one_self_play(args_tuple):
    board_size, args, policy_state, inference_cache = args_tuple
    env = OthelloGame(board_size)
    policy_class, policy_config, policy_state_dict = policy_state
    policy = policy_class(**policy_config)
    policy.load_state_dict(policy_state_dict)
    policy.eval()
    mcts = MCTS(env, args, policy, dirichlet_alpha=args["dirichlet_alpha"], dirichlet_epsilon=args["dirichlet_epsilon"],
                inference_cache=inference_cache)
    trajectory = []
    state = env.get_initial_state()
    player, is_terminal = 1, False
    while not is_terminal:
        temperature = args["mcts_temperature"] if len(trajectory) < args["num_exploratory_moves"] else 0.0
        action_probs = mcts.policy_improve_step(state, player, temp=temperature)
        trajectory.append((state.copy() * player, action_probs.copy(), player, mcts.root.value))
        action = np.random.choice(env.action_size, p=action_probs)
        mcts.make_move(action)
        state = env.get_next_state(state, action, player)
        reward, is_terminal = env.get_value_and_terminated(state, action, player)
        if is_terminal:
            winner = player if reward > 0 else (env.get_opponent(player) if reward < 0 else 0)
            return get_training_data(trajectory, winner, args["lambda"])
        player = env.get_opponent(player)


def collect_self_play_games(policy, args, n_games, *, n_slots=None, device="cuda:0", dtype=torch.bfloat16, seed=0,
                            game_id_base=0, fold=True, use_graph=True, return_raw=False):
    """Play ``n_games`` self-play games with ``policy`` (a torch module) on one GPU."""
    from .Models import fold_for_inference
    n_slots = int(n_slots or min(n_games, 4096))
    gps = -(-n_games // n_slots)
    eng = MctsEngine(n_slots, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=gps, device=device,
                     seed=seed, game_id_base=game_id_base)
    net = policy.to(device).eval()
    if fold:
        net = fold_for_inference(net, dtype)
        ev = BatchedPolicy(net, device, torch.float32)
    else:
        ev = BatchedPolicy(net, device, dtype)
    out = SelfPlayRunner(eng, ev, use_graph=use_graph).play()
    eng.raise_on_error()
    return out if return_raw else split_games(out)[:n_games]


def collect_self_play_games_distributed(policy, args, n_games_per_rank, *, n_slots=None, dtype=torch.bfloat16, seed=0, version=0):
    """Sharded ``collect_self_play_games`` (launch with torchrun, one process per GPU, NCCL initialised):
    rank 0's weights are broadcast, every rank plays its own game ids with no collective on the
    search path, and the replay tuples are gathered to rank 0, which gets the per-game lists
    (other ranks get None).  BASELINE config C5 is this with 16 384 slots per rank on 8 GPUs."""
    import torch.distributed as dist
    from . import parallel
    from .Models import fold_for_inference
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    net = policy.to(dev).eval()
    parallel.broadcast_weights(net, src=0, version=version)
    n_slots = int(n_slots or min(n_games_per_rank, 4096))
    gps = -(-n_games_per_rank // n_slots)
    base, stride = parallel.shard_game_ids(rank, world, n_slots)
    eng = MctsEngine(n_slots, args, self_play=True, eval_kind=_lib.EVAL_EXTERNAL, games_per_slot=gps, device=dev, seed=seed,
                     game_id_base=base, game_id_stride=stride)
    out = SelfPlayRunner(eng, BatchedPolicy(fold_for_inference(net, dtype), dev, torch.float32)).play()
    eng.raise_on_error()
    merged = parallel.gather_replay({k: v.to(dev) for k, v in out.items() if k != "states"}, dst=0, device=dev)
    if merged is None:
        return None
    n = merged["values"].numel()
    states = torch.empty((n, 8, 8), dtype=torch.int8, device=dev)
    if n:
        import ctypes as C
        _lib.check(_lib.lib().oth_unpack_canonical(merged["boards"].contiguous().data_ptr(), states.data_ptr(), n,
                                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    merged["states"] = states
    return split_games({k: v.cpu() for k, v in merged.items()})
