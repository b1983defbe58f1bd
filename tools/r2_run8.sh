#!/bin/bash
set -u
mkdir -p gpurun_out
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from alphazero_othello_b200.Models import AlphaZeroNet, FastOthelloNet, fold_for_inference
for kind in ("small", "big"):
    torch.manual_seed(1)
    net = (AlphaZeroNet(8, 65, 5, 128) if kind == "big" else FastOthelloNet(8, 65)).cuda().eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    x = torch.randint(-1, 2, (4096, 1, 8, 8), device="cuda").float()
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        lr, vr = net(x)
        l32, v32 = fold_for_inference(net, torch.float32)(x)
        torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
        lt, vt = fold_for_inference(net, torch.float32)(x)
        l16, v16 = fold_for_inference(net, torch.bfloat16)(x)
    pr = torch.softmax(lr, -1)
    for name, l, v in (("fp32 twin", l32, v32), ("tf32 twin", lt, vt), ("bf16 twin", l16, v16)):
        p = torch.softmax(l.float(), -1)
        print(kind, name, "priors max %.2e mean %.2e | values max %.2e mean %.2e | argmax agree %.4f | KL mean %.2e" % (
            (p - pr).abs().max(), (p - pr).abs().mean(), (v.float() - vr).abs().max(), (v.float() - vr).abs().mean(),
            (p.argmax(-1) == pr.argmax(-1)).float().mean(), (pr * (pr.clamp_min(1e-12).log() - p.clamp_min(1e-12).log())).sum(-1).mean()))
PY
