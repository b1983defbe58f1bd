"""Which torch ops the network twin spends its time in (B200, batch = leaf batch)."""
import sys
import torch
sys.path.insert(0, ".")
from alphazero_othello_b200.Models import AlphaZeroNet, FastOthelloNet, fold_for_inference
kind, B = sys.argv[1], int(sys.argv[2])
torch.manual_seed(0)
net = (AlphaZeroNet(8, 65) if kind == "big" else FastOthelloNet(8, 65)).cuda().eval()
f = fold_for_inference(net, torch.bfloat16)
x = torch.randint(-1, 2, (B, 1, 8, 8), device="cuda").float()
for _ in range(3):
    f(x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(5):
        f(x)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
