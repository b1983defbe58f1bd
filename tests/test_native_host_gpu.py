"""A host with no Python and no PyTorch in the loop: examples/selfplay_native.cpp drives the C ABI with
cudaMalloc'ed buffers (what a C++ maintainer of a trainer would do)."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_native_cpp_host_plays_complete_games():
    from alphazero_othello_b200 import build
    build.build()
    exe = os.path.join(ROOT, "examples", "selfplay_native")
    src = exe + ".cpp"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    hdr = os.path.join(ROOT, "include", "othello_b200.h")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call([nvcc, "-O2", "-o", exe, src, "-L" + os.path.join(ROOT, "alphazero_othello_b200"), "-lothello_b200",
                               "-Xlinker", "-rpath", "-Xlinker", os.path.join(ROOT, "alphazero_othello_b200")])
    out = subprocess.run([exe, "512", "32", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["games"] == 1024 and d["positions"] >= 9 * 1024
    assert d["white_wins"] + d["draws"] + d["black_wins"] == 1024
    assert d["sims"] >= 32 * d["positions"] * 0.9 and -1.0 <= d["mean_value_target"] <= 1.0
