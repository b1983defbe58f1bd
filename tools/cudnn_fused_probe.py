"""Experiment: cuDNN graph API (frontend) conv3x3 + bias + residual + ReLU as ONE engine for the big net's
second convolution of a residual block, [B,128,8,8] bf16 channels-last, against what the twin runs today
(plain cuDNN convolution + k_bias_add_relu_bf16).   python tools/cudnn_fused_probe.py [B] [C]
"""
import ctypes as C
import sys

sys.path.insert(0, ".")
import torch
import cudnn

from alphazero_othello_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
print("cudnn frontend", cudnn.__version__, "backend", cudnn.backend_version(), "torch cudnn", torch.backends.cudnn.version())
torch.manual_seed(0)
x = torch.randn(B, CH, 8, 8, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
z = torch.randn_like(x).relu_()
w = (torch.randn(CH, CH, 3, 3, device=dev, dtype=torch.bfloat16) * 0.03).contiguous(memory_format=torch.channels_last)
b = torch.randn(CH, device=dev, dtype=torch.bfloat16)
b4 = b.view(1, CH, 1, 1).contiguous(memory_format=torch.channels_last)
y = torch.empty_like(x)


def t(fn, n=30):
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


def today():
    o = torch.nn.functional.conv2d(x, w, None, 1, 1)
    _lib.check(_lib.lib().oth_nn_bias_add_relu_bf16(o.data_ptr(), z.data_ptr(), b.data_ptr(), o.numel(), CH,
                                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return o


ref = today().float()
print("today: conv2d + k_bias_add_relu %.1f us | conv2d alone %.1f us | cudnn_convolution_relu %.1f us | torch add_relu %.1f us" % (
    t(today), t(lambda: torch.nn.functional.conv2d(x, w, None, 1, 1)),
    t(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)),
    t(lambda: torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, (1, 1), (1, 1), (1, 1), 1))), flush=True)

handle = cudnn.create_handle()
cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream(dev).cuda_stream)


def build(order):
    g = cudnn.pygraph(io_data_type=cudnn.data_type.BFLOAT16, intermediate_data_type=cudnn.data_type.FLOAT,
                      compute_data_type=cudnn.data_type.FLOAT, handle=handle)
    X, W, Bt = g.tensor_like(x), g.tensor_like(w), g.tensor_like(b4)
    o = g.conv_fprop(image=X, weight=W, padding=[1, 1], stride=[1, 1], dilation=[1, 1])
    Z = None
    for op in order:
        if op == "bias":
            o = g.bias(input=o, bias=Bt)
        elif op == "addb":
            o = g.add(a=o, b=Bt)
        elif op == "res":
            Z = Z or g.tensor_like(z)
            o = g.add(a=o, b=Z)
    Y = g.relu(input=o)
    Y.set_output(True).set_data_type(cudnn.data_type.BFLOAT16)
    g.validate()
    g.build_operation_graph()
    g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.B, cudnn.heur_mode.FALLBACK])
    g.check_support()
    return g, X, W, Bt, Z, Y


for order in (("bias", "res"), ("res", "bias"), ("res", "addb"), ("addb", "res"), ("bias",)):
    with_res = "res" in order
    try:
        g, X, W, Bt, Z, Y = build(order)
    except Exception as ex:
        print(order, "build failed", repr(ex)[:300], flush=True)
        continue
    n = g.get_execution_plan_count()
    pack = {X: x, W: w, Bt: b4, Y: y}
    if with_res:
        pack[Z] = z
    want = ref if with_res else torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1).float()
    rows = []
    seen = set()
    for i in range(n):
        try:
            name = g.get_plan_name_at_index(i)
            if name in seen:
                continue
            seen.add(name)
            g.build_plan_at_index(i)
            ws = torch.empty(max(int(g.get_workspace_size_plan_at_index(i)), 16), dtype=torch.uint8, device=dev)
            run = lambda: g.execute_plan_at_index(pack, ws, i, handle=handle)
            y.zero_()
            run()
            torch.cuda.synchronize()
            err = float((y.float() - want).abs().max())
            us = t(run, 3)
            if us < 400:
                us = t(run, 30)
            rows.append((us, i, err, ws.numel(), name))
        except Exception as ex:
            rows.append((1e9, i, -1, 0, "failed: " + repr(ex)[:100]))
    rows.sort()
    print(order, ": plans", n, "distinct", len(seen), flush=True)
    for r in rows[:4]:
        print("   %.1f us  plan %d  max|diff| %.4f  ws %d  %s" % r, flush=True)
    engines = sorted({r[4].split("_")[0] for r in rows})
    print("   engines:", engines, flush=True)
