"""Policy/value networks at the engine boundary -- these STAY PyTorch (north_star).

Architectures and parameter names follow the reference's ``Models.py`` so its
checkpoints / ``state_dict()``s load unchanged (``FastOthelloNet`` :93-161, the
"small" net; ``AlphaZeroNet`` :164-221, the "big" net) and so
``one_self_play``'s ``policy_class(**policy_config)`` reconstruction works.
The engine only needs ``policy(x[B,1,8,8]) -> (logits[B,65], value[B,1])``.

``fold_for_inference`` builds the fast eval-mode twin the batched engine runs:
BatchNorm folded into the convolutions, channels_last bf16, cuDNN's fused
conv+bias+ReLU / conv+add+ReLU epilogues.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Inference:
    """Single-position evaluation, the signature MCTS programs against (Models.py:11-31)."""

    @torch.no_grad()
    def inference(self, state, current_player):
        x = torch.from_numpy((current_player * np.asarray(state)).astype(np.float32))[None]
        x = x.to(next(self.parameters()).device)
        self.eval()
        logits, value = self(x)
        return self.softmax(logits)[0].cpu().numpy(), value[0, 0].cpu().numpy().item()


def _conv_bn_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU())


class ResidualBlock(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


class FastOthelloNet(nn.Module, Inference):
    """Small net: conv-bn-relu, one residual block, conv-bn-relu, linear heads (640 514 parameters)."""

    def __init__(self, board_size, action_size):
        super().__init__()
        self.board_size, self.action_size = board_size, action_size
        self.initial_conv = _conv_bn_relu(1, 64)
        self.res_block = ResidualBlock(64)
        self.conv_add = _conv_bn_relu(64, 64)
        flat = 64 * board_size * board_size
        self.fc_policy = nn.Linear(flat, action_size)
        self.fc_value1 = nn.Linear(flat, 64)
        self.fc_value2 = nn.Linear(64, 1)
        self.softmax = nn.Softmax(dim=-1)

    def get_config(self):
        return {"board_size": self.board_size, "action_size": self.action_size}

    def forward(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        h = self.conv_add(self.res_block(self.initial_conv(x)))
        h = h.reshape(h.size(0), -1)
        return self.fc_policy(h), torch.tanh(self.fc_value2(F.relu(self.fc_value1(h))))


class AlphaZeroNet(nn.Module, Inference):
    """Big net: 3x3 stem, n residual blocks, 1x1-conv policy and value heads (1 505 480 parameters at 5x128)."""

    def __init__(self, board_size, action_size, n_res_blocks=5, channels=128):
        super().__init__()
        self.board_size, self.action_size = board_size, action_size
        self.n_res_blocks, self.channels = n_res_blocks, channels
        self.conv0 = nn.Conv2d(1, channels, 3, padding=1, bias=False)
        self.bn0 = nn.BatchNorm2d(channels)
        self.res = nn.Sequential(*[ResidualBlock(channels) for _ in range(n_res_blocks)])
        self.pol_conv = nn.Conv2d(channels, 2, 1, bias=False)
        self.pol_bn = nn.BatchNorm2d(2)
        self.pol_fc = nn.Linear(2 * board_size * board_size, action_size)
        self.val_conv = nn.Conv2d(channels, 1, 1, bias=False)
        self.val_bn = nn.BatchNorm2d(1)
        self.val_fc1 = nn.Linear(board_size * board_size, 256)
        self.val_fc2 = nn.Linear(256, 1)
        self.softmax = nn.Softmax(dim=-1)

    def get_config(self):
        return {"board_size": self.board_size, "action_size": self.action_size, "n_res_blocks": self.n_res_blocks,
                "channels": self.channels}

    def forward(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        h = self.res(F.relu(self.bn0(self.conv0(x))))
        p = F.relu(self.pol_bn(self.pol_conv(h)))
        v = F.relu(self.val_bn(self.val_conv(h)))
        p = self.pol_fc(p.reshape(p.size(0), -1))
        v = torch.tanh(self.val_fc2(F.relu(self.val_fc1(v.reshape(v.size(0), -1)))))
        return p, v


# ----------------------------------------------------------- fast eval twin --
def _fold(conv, bn):
    """Eval-mode BatchNorm folded into the preceding convolution: (weight, bias) float32."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.size(0), device=w.device)
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * s.view(-1, 1, 1, 1), (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()


class _FusedConv(nn.Module):
    """3x3 / 1x1 convolution + bias + ReLU in one cuDNN call; with a residual, one cuDNN graph
    ``relu(bias(conv + residual))`` (``_cudnn_fused``, batches >= 256), else cuDNN's plain convolution
    followed by one in-place ``relu(x + bias + residual)`` pass (``oth_nn_bias_add_relu_bf16``) -- both
    faster than ``torch.cudnn_convolution_add_relu``, which lands on a legacy engine here."""

    graph_fusion = True  # class-wide switch (tests / A-B measurements)

    def __init__(self, w, b, dtype):
        super().__init__()
        self.pad = w.size(-1) // 2
        self.register_buffer("w", w.to(dtype).contiguous(memory_format=torch.channels_last))
        self.register_buffer("b", b.to(dtype))

    def forward(self, x, residual=None):
        p = (self.pad, self.pad)
        if residual is None:
            return torch.cudnn_convolution_relu(x, self.w, self.b, (1, 1), p, (1, 1), 1)
        if x.dtype != torch.bfloat16:
            return torch.cudnn_convolution_add_relu(x, self.w, residual, 1.0, self.b, (1, 1), p, (1, 1), 1)
        if self.graph_fusion and x.is_contiguous(memory_format=torch.channels_last) \
                and residual.is_contiguous(memory_format=torch.channels_last):
            from . import _cudnn_fused
            y = _cudnn_fused.conv_res_bias_relu(x, self.w, self.b, residual)  # one cuDNN graph (large batches)
            if y is not None:
                return y
        y = F.conv2d(x, self.w, None, 1, self.pad)
        assert y.is_contiguous(memory_format=torch.channels_last) and residual.is_contiguous(memory_format=torch.channels_last)
        import ctypes as C
        from . import _lib
        _lib.check(_lib.lib().oth_nn_bias_add_relu_bf16(y.data_ptr(), residual.data_ptr(), self.b.data_ptr(), y.numel(),
                                                        y.size(1), C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)))
        return y


class _StemGemm(nn.Module):
    """The 1-channel 3x3 stem as im2col + GEMM with a fused bias+ReLU epilogue: the output is
    written straight in channels-last layout (cuDNN's 1-input-channel path is slow and NCHW)."""

    def __init__(self, w, b, dtype):
        super().__init__()
        cout = w.size(0)
        wt = torch.zeros(16, cout, dtype=torch.float32, device=w.device)  # K padded 9 -> 16 for alignment
        wt[:9] = w.reshape(cout, 9).t()
        self.register_buffer("wt", wt.to(dtype).contiguous())
        self.register_buffer("b", b.to(dtype))
        self.cout = cout

    def forward(self, x):  # x [B,1,8,8]: float32 (engine path) or already in the compute dtype
        B = x.size(0)
        # one im2col buffer PER batch size, kept for the life of the module: CUDA graphs captured at different batch
        # sizes (evaluation de-duplication buckets) each hold the address of theirs
        cache = self.__dict__.setdefault("_cols", {})
        key = (B, self.wt.dtype, x.device)
        cols = cache.get(key)
        if cols is None:
            cols = cache[key] = torch.zeros(B, 8, 8, 16, dtype=self.wt.dtype, device=x.device)  # K columns 9..15 stay zero
        if x.dtype == torch.float32 and cols.dtype == torch.bfloat16 and x.is_contiguous():
            import ctypes as C
            from . import _lib
            _lib.check(_lib.lib().oth_nn_stem_im2col_bf16(x.data_ptr(), cols.data_ptr(), B,
                                                          C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        else:
            taps = F.pad(x[:, 0].to(cols.dtype), (1, 1, 1, 1)).unfold(1, 3, 1).unfold(2, 3, 1)  # [B,8,8,3,3] view
            cols[..., :9].unflatten(-1, (3, 3)).copy_(taps)  # one strided copy = im2col
        y = torch._addmm_activation(self.b, cols.view(B * 64, 16), self.wt, use_gelu=False)  # relu epilogue
        return y.view(B, 8, 8, self.cout).permute(0, 3, 1, 2)  # NHWC memory == channels_last [B,C,8,8]


def _pad_rows(w, b, mult=64):
    """Zero-pad a linear layer's output rows to a multiple of `mult`: odd sizes (65, 129, 1) push
    cuBLAS onto slow legacy kernels."""
    n = w.size(0)
    m = ((n + mult - 1) // mult) * mult
    wp = torch.zeros(m, w.size(1), dtype=w.dtype, device=w.device)
    bp = torch.zeros(m, dtype=b.dtype, device=b.device)
    wp[:n], bp[:n] = w, b
    return wp, bp


class FoldedNet(nn.Module):
    """Inference-only twin of FastOthelloNet / AlphaZeroNet (same function up to dtype rounding):
    BatchNorm folded, channels-last activations, and both heads as ONE GEMM over the NHWC-flattened
    trunk output (policy logits | value hidden) followed by the tiny value output layer."""

    def __init__(self, net, dtype=torch.bfloat16):
        super().__init__()
        self.dtype = dtype
        self.kind = "big" if isinstance(net, AlphaZeroNet) or hasattr(net, "pol_conv") else "small"
        mk = lambda c, bn: _FusedConv(*_fold(c, bn), dtype)
        stem = (lambda c, bn: _StemGemm(*_fold(c, bn), dtype)) if dtype != torch.float32 else mk
        f32 = lambda t: t.detach().float()
        if self.kind == "small":
            self.stem = stem(net.initial_conv[0], net.initial_conv[1])
            self.blocks = nn.ModuleList([nn.ModuleList([mk(net.res_block.conv1, net.res_block.bn1),
                                                        mk(net.res_block.conv2, net.res_block.bn2)])])
            self.tail = mk(net.conv_add[0], net.conv_add[1])
            wp, wv = f32(net.fc_policy.weight), f32(net.fc_value1.weight)
            w = torch.cat([wp, wv], 0).view(-1, 64, 8, 8).permute(0, 2, 3, 1).reshape(wp.size(0) + wv.size(0), -1)
            hb = torch.cat([f32(net.fc_policy.bias), f32(net.fc_value1.bias)])
            self.n_hidden = wv.size(0)
            v2w, v2b = f32(net.fc_value2.weight), f32(net.fc_value2.bias)
        else:
            self.stem = stem(net.conv0, net.bn0)
            self.blocks = nn.ModuleList([nn.ModuleList([mk(b.conv1, b.bn1), mk(b.conv2, b.bn2)]) for b in net.res])
            # both 1x1 head convolutions (2 policy planes, 1 value plane) as one GEMM over the NHWC trunk
            # output: [B*64, C] x [C, 16] (3 real columns, zero-padded) with the bias+ReLU epilogue fused ...
            pw, pb = _fold(net.pol_conv, net.pol_bn)
            vw, vb = _fold(net.val_conv, net.val_bn)
            C_in = pw.size(1)
            hw1 = torch.zeros(C_in, 16, device=pw.device)
            hw1[:, 0:2] = pw.view(2, C_in).t()
            hw1[:, 2] = vw.view(C_in)
            hb1 = torch.zeros(16, device=pw.device)
            hb1[0:2], hb1[2] = pb, vb[0]
            self.register_buffer("heads_w", hw1.to(dtype).contiguous())
            self.register_buffer("heads_b", hb1.to(dtype))
            # ... and both first linear layers as one matrix over its flattening [hw*16 + c]
            wpol, wval = f32(net.pol_fc.weight), f32(net.val_fc1.weight)  # [65, 2*64] (c*64+hw), [256, 64] (hw)
            no, nh = wpol.size(0), wval.size(0)
            w = torch.zeros(no + nh, 64, 16, device=wpol.device)
            w[:no, :, 0:2] = wpol.view(no, 2, 64).permute(0, 2, 1)
            w[no:, :, 2] = wval
            w = w.view(no + nh, 64 * 16)
            hb = torch.cat([f32(net.pol_fc.bias), f32(net.val_fc1.bias)])
            self.n_hidden = nh
            v2w, v2b = f32(net.val_fc2.weight), f32(net.val_fc2.bias)
        w, hb = _pad_rows(w, hb)
        v2w, v2b = _pad_rows(v2w, v2b, 8)
        self.register_buffer("head_w", w.to(dtype).contiguous())
        self.register_buffer("head_b", hb.to(dtype))
        self.register_buffer("v2_w", v2w.to(dtype).contiguous())
        self.register_buffer("v2_b", v2b.to(dtype))
        self.n_actions = net.action_size
        self.supports_raw = True

    @torch.no_grad()
    def forward(self, x, raw=False):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        h = self.stem(x if isinstance(self.stem, _StemGemm) else x.to(self.dtype))
        if not h.is_contiguous(memory_format=torch.channels_last):
            h = h.contiguous(memory_format=torch.channels_last)
        for c1, c2 in self.blocks:
            h = c2(c1(h), residual=h)
        B = h.size(0)
        if self.kind == "small":
            t = self.tail(h)
            feat = t.permute(0, 2, 3, 1).reshape(B, -1)  # NHWC flattening: a view
        else:
            hf = h.permute(0, 2, 3, 1).reshape(B * 64, -1)  # [B*64, C] view of the channels-last trunk output
            feat = torch._addmm_activation(self.heads_b, hf, self.heads_w, use_gelu=False).view(B, -1)
        y = F.linear(feat, self.head_w, self.head_b)
        na, nh = self.n_actions, self.n_hidden
        v = F.linear(F.relu(y[:, na:na + nh]), self.v2_w, self.v2_b)[:, :1]
        if raw:  # engine path: (logits, value pre-activation) views in the compute dtype; softmax / tanh are
            return y[:, :na], v  # applied downstream (fused into oth_mcts_step_fused, or by BatchedPolicy)
        return y[:, :na].float(), torch.tanh(v.float())


@torch.no_grad()
def refold_(folded, net):
    """Refresh a FoldedNet in place from (updated) weights of ``net`` -- same buffers, so a
    captured CUDA graph that runs ``folded`` stays valid."""
    fresh = FoldedNet(net.eval(), folded.dtype).to(next(net.parameters()).device)
    for old, new in zip(folded.buffers(), fresh.buffers()):
        old.copy_(new)
    return folded


def fold_for_inference(net, dtype=torch.bfloat16):
    return FoldedNet(net.eval(), dtype).to(next(net.parameters()).device).eval()
