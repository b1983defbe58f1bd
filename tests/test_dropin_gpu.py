"""The reference-facing Python surfaces (MCTS class, one_self_play, collect_self_play_games,
network twins) on the GPU, checked like the reference's own call sites use them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _StubH:
    """Python twin of the oracle's hash stub with the Models.Inference.inference signature."""

    def __init__(self, salt=0):
        self.salt = salt

    def load_state_dict(self, sd):
        pass

    def eval(self):
        pass

    def inference(self, state, player):
        import oracle as O
        pri = np.zeros(65, np.float32)
        val = np.zeros(1, np.float64)
        import ctypes as C
        salt = C.c_uint64(self.salt)
        s = np.ascontiguousarray(state, dtype=np.int8)
        fn = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)(O.lib().orc_get_stub(O.STUB_H))
        fn(C.addressof(salt), s.ctypes.data, int(player), pri.ctypes.data, val.ctypes.data)
        return pri, float(val[0])


@pytest.mark.parametrize("case", ["H_noise", "H_t0", "H_alpha"])
def test_mcts_class_matches_reference_including_rng_stream(golden, case):
    """Same np.random seed as the reference run that made the fixture -> same Dirichlet draw and
    tie picks, bit-identical visit counts / root values / policy targets (even for temp != 1)."""
    from alphazero_othello_b200.MCTS_model import MCTS
    from alphazero_othello_b200.envs.othello import OthelloGameNew
    pre = f"mcts_{case}_"
    sid, sims, c, eps, alpha, temp, salt = (float(x) for x in golden[pre + "cfg"])  # Python floats, as the reference's callers pass
    seed = {"H_noise": 3, "H_t0": 4, "H_alpha": 5}[case]
    env = OthelloGameNew(8)
    np.random.seed(seed)
    m = MCTS(env, {"c_puct": c, "num_simulations": int(sims), "num_threads": 1}, _StubH(int(salt)), dirichlet_alpha=alpha,
             dirichlet_epsilon=eps)
    m.make_move(3)  # before any search: silent no-op (MCTS_model.py:209-211)
    n = min(len(golden[pre + "action"]), 12)
    for t in range(n):
        s, pl = golden[pre + "state"][t], int(golden[pre + "player"][t])
        probs = m.policy_improve_step(s, pl, temp=temp)
        assert probs.dtype == np.float32 and np.array_equal(probs, golden[pre + "probs"][t]), (case, t)
        assert m.root.value == golden[pre + "root_value"][t] and m.root.visit_count == golden[pre + "root_n"][t]
        for a, ch in m.root.children.items():
            assert ch.visit_count == golden[pre + "counts"][t][a] and ch.value == golden[pre + "cval"][t][a]
            assert float(ch.prior) == golden[pre + "cpri"][t][a]
        assert sorted(m.root.children) == list(np.nonzero(env.get_valid_moves(s, pl))[0])
        m.make_move(int(golden[pre + "action"][t]))
    with pytest.raises(AssertionError):  # tree/game state mismatch (MCTS_model.py:231-232)
        m.policy_improve_step(golden[pre + "state"][0], int(golden[pre + "player"][0]))
    with pytest.raises(KeyError):
        bad = next(a for a in range(64) if a not in m.root.children)
        m.make_move(bad)


def test_one_self_play_matches_reference(golden):
    from alphazero_othello_b200.self_play_worker import one_self_play
    pre = "sp_sp0_"
    salt, sims, c, eps, alpha, temp, nexp, lam = (float(x) for x in golden[pre + "cfg"])
    args = {"c_puct": c, "num_simulations": int(sims), "num_threads": 1, "dirichlet_alpha": alpha, "dirichlet_epsilon": eps,
            "mcts_temperature": temp, "num_exploratory_moves": int(nexp), "lambda": lam}
    np.random.seed(10)  # the seed of the reference run (tests/golden/make_golden.py)
    traj = one_self_play((8, args, (_StubH, {"salt": int(salt)}, {}), None))
    assert len(traj) == len(golden[pre + "values"])
    for t, (s, pi, v) in enumerate(traj):
        assert s.dtype == np.int8 and np.array_equal(s, golden[pre + "states"][t])
        assert pi.dtype == np.float32 and np.array_equal(pi, golden[pre + "pis"][t])
        assert v == golden[pre + "values"][t]


@pytest.mark.parametrize("kind", ["small", "big"])
def test_folded_network_twin(kind):
    import torch
    from alphazero_othello_b200.Models import AlphaZeroNet, FastOthelloNet, fold_for_inference, refold_
    torch.manual_seed(1)
    net = (AlphaZeroNet(8, 65) if kind == "big" else FastOthelloNet(8, 65)).cuda().eval()
    for m in net.modules():  # non-trivial BN statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    x = torch.randint(-1, 2, (257, 1, 8, 8), device="cuda").float()
    with torch.no_grad():
        lr, vr = net(x)
        l32, v32 = fold_for_inference(net, torch.float32)(x)
        f16 = fold_for_inference(net, torch.bfloat16)
        l16, v16 = f16(x)
    assert (l32 - lr).abs().max() < 2e-3 and (v32 - vr).abs().max() < 2e-3
    pr, p16 = torch.softmax(lr, -1), torch.softmax(l16, -1)
    # measured on a B200 (tools/r2_run8.sh, 4096 positions): bf16 twin |dp| max 1.0e-4 / 6.3e-5, |dv| max 1.4e-3 / 9.6e-4
    # (small / big net), arg-max agreement 0.99 / 0.98; TF32 twin |dp| 6e-6, |dv| 1e-4.  Bounds = 20x / 7x that.
    assert (p16 - pr).abs().max() < 2e-3 and (v16 - vr).abs().max() < 1e-2
    assert (p16.argmax(-1) == pr.argmax(-1)).float().mean() > 0.9
    with torch.no_grad():
        net.val_fc2.bias.add_(0.5) if kind == "big" else net.fc_value2.bias.add_(0.5)
        v_new = refold_(f16, net)(x)[1]
    assert (v_new - v16).abs().max() > 1e-3


def test_collect_self_play_games_small_net():
    import torch
    from alphazero_othello_b200.Models import FastOthelloNet
    from alphazero_othello_b200.self_play_worker import collect_self_play_games
    torch.manual_seed(0)
    args = {"c_puct": 2.0, "num_simulations": 8, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3, "mcts_temperature": 1.0,
            "num_exploratory_moves": 35, "lambda": 0.98}
    games = collect_self_play_games(FastOthelloNet(8, 65), args, 48, n_slots=32)
    assert len(games) == 48
    for g in games:
        assert 9 <= len(g) <= 128
        s0, pi0, v0 = g[0]
        assert s0.dtype == np.int8 and s0.shape == (8, 8) and np.abs(s0).sum() == 4
        assert pi0.dtype == np.float32 and abs(pi0.sum() - 1) < 1e-5 and isinstance(v0, float)
        assert set(np.nonzero(pi0)[0]) <= {19, 26, 37, 44}
        assert g[-1][2] in (-1.0, 0.0, 1.0)
        for (s, pi, v) in g:
            assert abs(pi.sum() - 1) < 1e-5 and -1.0 <= v <= 1.0


def test_mcts_class_rollout_mode_finds_forced_win():
    """The reference's own MCTS tests use policy=None (test_MCTS.py:7-36, TicTacToe): a forced win must
    get the visits.  Othello analogue: a position where one move wipes the opponent out."""
    from alphazero_othello_b200.MCTS_model import MCTS
    from alphazero_othello_b200.envs.othello import OthelloGameNew
    env = OthelloGameNew(8)
    s = np.zeros((8, 8), np.int8)
    s[3, 2] = 1; s[3, 3] = -1; s[3, 4] = -1          # playing (3,5) captures every -1 disc: immediate win
    s[0, 0] = 1; s[1, 0] = -1; s[5, 5] = 1; s[5, 6] = -1; s[2, 0] = 0
    m = MCTS(env, {"c_puct": 1.4, "num_simulations": 300, "num_threads": 1}, None, seed=1)
    probs = m.policy_improve_step(s, 1, temp=1.0)
    assert m.use_rollout and abs(probs.sum() - 1) < 1e-5
    legal = np.nonzero(env.get_valid_moves(s, 1))[0]
    assert set(np.nonzero(probs)[0]) <= set(legal)
    assert m.root.visit_count == 301


def test_residual_conv_as_one_cudnn_graph_matches_the_two_kernel_path():
    """Network boundary: relu(conv(x) + residual + bias) as one cuDNN graph (autotuned plan, also when
    replayed from a CUDA graph) against cuDNN conv + k_bias_add_relu_bf16 and against float32."""
    import torch
    from alphazero_othello_b200 import _cudnn_fused
    from alphazero_othello_b200.Models import _FusedConv
    torch.manual_seed(3)
    B, C = 2048, 128
    conv = _FusedConv(torch.randn(C, C, 3, 3) * 0.03, torch.randn(C) * 0.1, torch.bfloat16).cuda()
    x = torch.randn(B, C, 8, 8, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    z = torch.randn_like(x).relu_()
    want = torch.relu(torch.nn.functional.conv2d(x.float(), conv.w.float(), conv.b.float(), 1, 1) + z.float())
    try:
        _FusedConv.graph_fusion = False
        two = conv(x, residual=z)
    finally:
        _FusedConv.graph_fusion = True
    one = conv(x, residual=z)
    assert any(str((B, C, 8, 8)) in k for k in _cudnn_fused.chosen_plans()), "cuDNN graph path not taken"
    assert one.is_contiguous(memory_format=torch.channels_last) and one.dtype == torch.bfloat16
    tol = 0.02 * float(want.abs().max())  # bf16 output rounding
    assert float((one.float() - want).abs().max()) <= tol and float((two.float() - want).abs().max()) <= 2 * tol
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cap = conv(x, residual=z)
    cap.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(cap, one)


class _RawRows:
    """What RandomSymmetryDataset (train.py:26-42) becomes: raw rows, no CUDA in __getitem__."""

    def __init__(self, states, pis, vs):
        self.s, self.p, self.v = states, pis, vs

    def __len__(self):
        return len(self.s)

    def __getitem__(self, i):
        import torch
        return torch.from_numpy(self.s[i]), torch.from_numpy(self.p[i]), torch.tensor(self.v[i], dtype=torch.float32)


class _SwappedRows(_RawRows):
    """The swap INTEGRATION.md warns against: the GPU get_random_symmetry inside a worker."""

    def __getitem__(self, i):
        from alphazero_othello_b200.envs.othello import get_random_symmetry
        return get_random_symmetry(self.s[i], self.p[i])


def test_augment_batch_with_dataloader_workers():
    """Training-time symmetry augmentation with DataLoader(num_workers > 0): workers yield raw rows, the
    training process augments the collated batch on the GPU; every output row is one of the 8 dihedral images of its
    input row (checked against the oracle's get_symmetries).  Calling the GPU get_random_symmetry from a worker fails
    with a message that names the replacement instead of a CUDA initialisation error."""
    import torch
    import oracle as O
    from torch.utils.data import DataLoader
    from alphazero_othello_b200.replay import augment_batch
    rs = np.random.RandomState(3)
    n = 64
    states = rs.randint(-1, 2, size=(n, 8, 8)).astype(np.int8)
    pis = rs.dirichlet([0.5] * 65, size=n).astype(np.float32)
    vs = rs.uniform(-1, 1, n).astype(np.float32)
    seen = 0
    for s, p, v in DataLoader(_RawRows(states, pis, vs), batch_size=16, shuffle=False, num_workers=2):
        s2, p2 = augment_batch(s, p, "cuda:0")
        assert s2.shape == (16, 1, 8, 8) and s2.dtype == torch.float32 and p2.shape == (16, 65)
        s2, p2 = s2.cpu().numpy(), p2.cpu().numpy()
        for j in range(16):
            imgs_s, imgs_p = O.symmetries(states[seen + j], pis[seen + j])
            assert any(np.array_equal(s2[j, 0], imgs_s[k]) and np.array_equal(p2[j], imgs_p[k]) for k in range(8))
        seen += 16
    assert seen == n
    with pytest.raises(RuntimeError, match="augment_batch"):
        next(iter(DataLoader(_SwappedRows(states, pis, vs), batch_size=4, num_workers=1)))


def test_callers_module_is_left_alone():
    """collect_self_play_games / MCTS work on a private copy: the trainer's module stays on its device, in its mode."""
    import torch
    from alphazero_othello_b200.MCTS_model import MCTS
    from alphazero_othello_b200.Models import FastOthelloNet
    from alphazero_othello_b200.envs.othello import OthelloGameNew
    from alphazero_othello_b200.self_play_worker import collect_self_play_games
    torch.manual_seed(1)
    net = FastOthelloNet(8, 65).train()
    args = {"c_puct": 2.0, "num_simulations": 4, "dirichlet_alpha": 1.0, "dirichlet_epsilon": 0.3, "mcts_temperature": 1.0,
            "num_exploratory_moves": 5, "lambda": 0.98}
    games = collect_self_play_games(net, args, 8, n_slots=8)
    assert len(games) == 8
    env = OthelloGameNew(8)
    m = MCTS(env, dict(args, num_threads=4), net)  # num_threads is accepted; the search is sequential (docstring)
    probs = m.policy_improve_step(env.get_initial_state(), 1, temp=1.0)
    assert abs(probs.sum() - 1) < 1e-5 and m.root.visit_count == 5
    assert net.training and next(net.parameters()).device.type == "cpu"
