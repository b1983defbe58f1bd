"""Config C1 under ncu: one k_rollout launch of 2^22 games (plus the INT32 probe for reference)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import torch
from alphazero_othello_b200 import _lib
from alphazero_othello_b200.envs.othello import BatchedOthello
env = BatchedOthello()
for rep in range(3):
    r = env.rollout(1 << 22, seed=rep)
torch.cuda.synchronize()
print("plies", int(r["counters"][0]))
ips, ms = C.c_double(0), C.c_float(0)
_lib.check(_lib.lib().oth_host_int32_peak(C.byref(ips), C.byref(ms)))
print("int32 probe", ips.value)
