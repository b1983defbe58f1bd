"""Search-only throughput (production kernel pair, device stub in the network's place) for a config: python tools/search_only.py c4|c3"""
import sys
sys.path.insert(0, ".")
import torch
from bench import WORKLOADS, aux_search_only
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
_, _, G, sims = WORKLOADS[wl]
print(wl, aux_search_only(torch.device("cuda:0"), G, sims))
