"""Does cuDNN have a fast fused conv+bias+residual+ReLU for [16384,128,8,8] bf16 channels-last? (benchmark mode on/off)"""
import torch, time
x = torch.randn(16384, 128, 8, 8, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
z = torch.randn_like(x)
w = torch.randn(128, 128, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
b = torch.randn(128, device="cuda", dtype=torch.bfloat16)

def t(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000

for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    print("benchmark", bench,
          "conv_relu %.0f us" % t(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)),
          "conv_add_relu %.0f us" % t(lambda: torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, (1, 1), (1, 1), (1, 1), 1)),
          "conv2d %.0f us" % t(lambda: torch.nn.functional.conv2d(x, w, None, 1, 1)))
