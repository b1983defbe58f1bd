"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/othello_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from alphazero_othello_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    names = set()
    for h in ("othello_b200.h", "othello_b200_experimental.h"):  # the boundary + the measurement hooks
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(oth_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    assert _lib.lib().oth_abi_version() == _lib.ABI_VERSION == 2
    hdr = open(os.path.join(ROOT, "include", "othello_b200.h")).read()
    assert re.search(r"#define OTH_ABI_VERSION 2\b", hdr)


def test_every_header_symbol_is_exported_and_bound():
    L = ctypes.CDLL(build.build())
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in othello_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.MctsCtl) == 64
    assert ctypes.sizeof(_lib.MctsConfig) == 16 * 4 + 2 * 8 + 5 * 8 + 4 * 8
    assert ctypes.sizeof(_lib.MctsBuffers) == 8 * (_lib.BUF_COUNT + 1)
    # field order of the header's struct == the ctypes mirror
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "othello_b200.h")).read(), flags=re.S)
    body = re.search(r"typedef struct oth_mcts_config \{(.*?)\} oth_mcts_config;", hdr, re.S).group(1)
    fields = re.findall(r"\b(?:u?int(?:32|64)_t|double)\s+(\w+);", body)
    assert fields == [f[0].rstrip("_") for f in _lib.MctsConfig._fields_]
    buf_enum = re.search(r"enum \{\s*OTH_BUF_NODES = 0,(.*?)OTH_BUF_COUNT\b", hdr, re.S).group(1)
    assert len(re.findall(r"OTH_BUF_\w+", buf_enum)) + 1 == _lib.BUF_COUNT


def test_argument_validation_without_device():
    L = _lib.lib()
    assert L.oth_legal_moves(None, None, None, -1, None) == _lib.OTH_E_ARG
    assert L.oth_legal_moves(None, None, None, 0, None) == _lib.OTH_OK
    cfg = _lib.MctsConfig()
    sizes = (ctypes.c_int64 * _lib.BUF_COUNT)()
    assert L.oth_mcts_buffer_bytes(ctypes.byref(cfg), sizes) == _lib.OTH_E_ARG
    cfg.n_slots, cfg.node_cap, cfg.path_cap, cfg.num_simulations, cfg.max_inline_sims, cfg.lanes = 4, 256, 64, 10, 4, 32
    cfg.self_play, cfg.out_pos_cap, cfg.out_game_cap = 1, 1000, 10
    assert L.oth_mcts_buffer_bytes(ctypes.byref(cfg), sizes) == 0
    assert sizes[_lib.BUF_NODES] == 4 * 2 * 256 * 32 and sizes[_lib.BUF_BOARDS] == 4 * 2 * 256 * 16
    assert L.oth_error_string(_lib.OTH_E_ILLEGAL).decode() == "Illegal move"
    cfg.hot_path = 53
    assert L.oth_mcts_buffer_bytes(ctypes.byref(cfg), sizes) == _lib.OTH_E_ARG   # at most 52 entries fit the hot record
    cfg.hot_path, cfg.split_stub = 4, 1
    assert L.oth_mcts_buffer_bytes(ctypes.byref(cfg), sizes) == _lib.OTH_E_ARG   # split_stub needs a device stub evaluator
    cfg.eval_kind = _lib.EVAL_STUB_H
    assert L.oth_mcts_buffer_bytes(ctypes.byref(cfg), sizes) == 0
    assert sizes[_lib.BUF_MOVE_LIST] == (4 + 4) * 4
    h = ctypes.c_void_p()
    assert L.oth_mcts_profile_create(0, ctypes.byref(h)) == _lib.OTH_E_ARG
    assert L.oth_mcts_profile_read(None, None, None, None) == _lib.OTH_E_ARG


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alphazero_othello_b200.envs.othello import OthelloGameNew
    with pytest.raises(_lib.OthelloB200Error):
        OthelloGameNew(8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "alphazero_othello_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_cudnn_graph_path_declines_cpu_and_small_batches():
    """The fused residual convolution is an optimisation of the network twin only: without a CUDA bf16
    batch of >= 256 it returns None and the caller keeps its two-kernel path."""
    import torch
    from alphazero_othello_b200 import _cudnn_fused
    x = torch.zeros(512, 8, 8, 8, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.zeros(8, 8, 3, 3, dtype=torch.bfloat16)
    assert _cudnn_fused.conv_res_bias_relu(x, w, torch.zeros(8, dtype=torch.bfloat16), x) is None
    assert _cudnn_fused.conv_res_bias_relu(x.float(), w.float(), torch.zeros(8), x.float()) is None
    assert _cudnn_fused.chosen_plans() == {}
