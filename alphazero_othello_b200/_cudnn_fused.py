"""conv3x3 + residual + bias + ReLU of the network twin as ONE cuDNN graph (cuDNN graph API through the
``cudnn`` frontend package), used by ``Models._FusedConv`` for the second convolution of a residual block.

Network-boundary plumbing, like the rest of the twin (the network stays PyTorch / cuDNN, north_star):
PyTorch's own ``cudnn_convolution_add_relu`` lands on a slow legacy engine for bf16 channels-last on
sm_100 (376 us at [16384,128,8,8]) and the twin's previous path -- plain convolution + one
``k_bias_add_relu_bf16`` pass -- costs 220 + 111 us; the graph ``relu(bias(conv(x, w) + z))`` (in THIS
operation order, which is the one the sm_100 engines accept) runs as a single kernel in ~255 us
(measurements: tools/cudnn_fused_probe.py, tools/cudnn_knob_search.py).

If the frontend package is missing, or cuDNN offers no plan for a shape, ``conv_res_bias_relu`` returns
None and the caller keeps the two-kernel path.
"""
import torch

_MIN_BATCH = 256        # below this the two-kernel path is as good and no plan has to be built
_AUTOTUNE_BATCH = 2048  # from here on the candidate plans of heuristic mode A are timed once per shape
_state = {"cudnn": None, "handles": {}, "plans": {}}


def _frontend():
    if _state["cudnn"] is None:
        try:
            import cudnn
            _state["cudnn"] = cudnn
        except Exception:  # not installed / backend library not loadable
            _state["cudnn"] = False
    return _state["cudnn"]


class _Plan:
    def __init__(self, cudnn, handle, x, w, b4, z):
        dt = cudnn.data_type
        g = cudnn.pygraph(io_data_type=dt.BFLOAT16, intermediate_data_type=dt.FLOAT, compute_data_type=dt.FLOAT, handle=handle)
        X, W, Bt, Z = g.tensor_like(x), g.tensor_like(w), g.tensor_like(b4), g.tensor_like(z)
        pad = w.size(-1) // 2
        o = g.conv_fprop(image=X, weight=W, padding=[pad, pad], stride=[1, 1], dilation=[1, 1])
        o = g.bias(input=g.add(a=o, b=Z), bias=Bt)
        Y = g.relu(input=o)
        Y.set_output(True).set_data_type(dt.BFLOAT16)
        g.validate()
        g.build_operation_graph()
        g.create_execution_plans([cudnn.heur_mode.A])
        g.check_support()
        self.g, self.t, self.handle, self.cudnn = g, (X, W, Bt, Z, Y), handle, cudnn
        self.index, self.ws = None, None
        self.timings = []

    def _pack(self, x, w, b4, z, y):
        X, W, Bt, Z, Y = self.t
        return {X: x.data_ptr(), W: w.data_ptr(), Bt: b4.data_ptr(), Z: z.data_ptr(), Y: y.data_ptr()}

    def choose(self, x, w, b4, z, autotune):
        """Build the first supported plan of heuristic mode A, or (autotune) time its first few distinct
        candidates on the real operands and keep the fastest."""
        g, dev = self.g, x.device
        y = torch.empty_like(x)
        pack = self._pack(x, w, b4, z, y)
        seen, best = set(), None
        for i in range(g.get_execution_plan_count()):
            name = g.get_plan_name_at_index(i)
            if name in seen:
                continue
            seen.add(name)
            try:
                g.build_plan_at_index(i)
                nb = int(g.get_workspace_size_plan_at_index(i))
                if nb > (64 << 20):
                    continue
                ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
                run = lambda: g.execute_plan_at_index(pack, ws, i, handle=self.handle)
                run()
                if not autotune:
                    best = (0.0, i, ws)
                    break
                for _ in range(2):
                    run()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream(dev))
                for _ in range(5):
                    run()
                e1.record(torch.cuda.current_stream(dev))
                e1.synchronize()
                ms = e0.elapsed_time(e1) / 5
                self.timings.append((name, ms))
                if best is None or ms < best[0]:
                    best = (ms, i, ws)
            except Exception:
                continue
            if len(seen) >= 6:
                break
        if best is None:
            raise RuntimeError("no cuDNN plan")
        _, self.index, self.ws = best

    def run(self, x, w, b4, z):
        y = torch.empty_like(x)
        self.cudnn.set_stream(handle=self.handle, stream=torch.cuda.current_stream(x.device).cuda_stream)
        self.g.execute_plan_at_index(self._pack(x, w, b4, z, y), self.ws, self.index, handle=self.handle)
        return y


def conv_res_bias_relu(x, w, b, z):
    """relu(conv2d(x, w, padding=k//2) + z + b) for bf16 channels-last ``x``/``z`` [B,C,H,W], ``w`` [K,C,k,k]
    (channels-last) and ``b`` [K]; a new channels-last tensor, or None when this path is unavailable."""
    if x.dtype != torch.bfloat16 or not x.is_cuda or x.size(0) < _MIN_BATCH:
        return None
    cudnn = _frontend()
    if not cudnn:
        return None
    dev = x.device
    key = (dev.index, tuple(x.shape), tuple(w.shape))
    plan = _state["plans"].get(key)
    if plan is False:
        return None
    b4 = b.view(1, -1, 1, 1)
    if plan is None:
        if torch.cuda.is_current_stream_capturing():  # plans are built eagerly (warm-up iterations), never while capturing
            return None
        try:
            with torch.cuda.device(dev):
                h = _state["handles"].get(dev.index)
                if h is None:
                    h = _state["handles"][dev.index] = cudnn.create_handle()
                cudnn.set_stream(handle=h, stream=torch.cuda.current_stream(dev).cuda_stream)
                plan = _Plan(cudnn, h, x, w, b4, z)
                plan.choose(x, w, b4, z, autotune=x.size(0) >= _AUTOTUNE_BATCH)
        except Exception:
            _state["plans"][key] = False
            return None
        _state["plans"][key] = plan
    with torch.cuda.device(dev):
        return plan.run(x, w, b4, z)


def chosen_plans():
    """{(device, x shape, w shape): (plan name, [(candidate, ms), ...])} -- for logs and the bench line."""
    out = {}
    for k, p in _state["plans"].items():
        if p:
            out[str(k)] = (p.g.get_plan_name_at_index(p.index), p.timings)
    return out
