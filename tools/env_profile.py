"""Stepwise env API under ncu: k_legal_moves and k_step on 2^24 mid-game positions (packed boards resident in HBM)."""
import sys
sys.path.insert(0, ".")
import torch
from alphazero_othello_b200.envs.othello import BatchedOthello
env = BatchedOthello()
n = 1 << 24
own, opp = env.initial(n)
gen = torch.Generator(device="cuda")
gen.manual_seed(0)
act = None
for ply in range(12):  # mid-game positions: 12 plies, each the r-th lowest legal square (r random per game)
    lm = env.legal_moves(own, opp)
    r = torch.randint(0, 4, (n,), device="cuda", generator=gen)
    pick = lm
    for _ in range(3):
        nxt = pick & (pick - 1)
        pick = torch.where((r > 0) & (nxt != 0), nxt, pick)
        r = r - 1
    bit = pick & (-pick)
    act = torch.where(lm == 0, torch.full_like(bit, 64), (torch.log2(bit.double().abs()) + 0.5).long() % 64)
    act = torch.where((bit < 0) & (lm != 0), torch.full_like(act, 63), act).to(torch.uint8)
    own, opp, _, fl = env.step(own, opp, act)
for _ in range(3):
    env.legal_moves(own, opp)
    env.step(own, opp, act)
torch.cuda.synchronize()
print("positions", n)
