"""The header-only C++ facade (include/loik_b200/first_order_loik_optimized.hpp) compiles with plain g++, links
against libloik_b200.so and behaves: without a GPU the constructor throws (no CPU fallback), with one it solves."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from loik_b200 import build
    build.build()
    exe = os.path.join(str(tmp_path), "facade_smoke")
    libdir = os.path.join(ROOT, "loik_b200")
    subprocess.check_call(["g++", "-std=c++17", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", "facade_smoke.cpp"),
                           "-o", exe, f"-L{libdir}", "-lloik_b200", f"-Wl,-rpath,{libdir}"])
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_facade_solves_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "facade ok" in r.stdout, r.stdout + r.stderr
