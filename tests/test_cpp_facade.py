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


def _build_adapter(tmp_path):
    """The reference-shaped facade against the stand-in Pinocchio / Eigen headers, with the oracle linked in."""
    from loik_b200 import build
    build.build()
    exe = os.path.join(str(tmp_path), "pinocchio_adapter_test")
    libdir = os.path.join(ROOT, "loik_b200")
    obj = os.path.join(str(tmp_path), "loik_oracle.o")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", "loik_oracle.c"), "-o", obj])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", f"-I{ROOT}/tests/cpp/stub", f"-I{ROOT}/include",
                           os.path.join(ROOT, "tests", "cpp", "pinocchio_adapter_test.cpp"), obj, "-o", exe, f"-L{libdir}", "-lloik_b200",
                           f"-Wl,-rpath,{libdir}", "-lm", "-lpthread"])
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_pinocchio_shaped_facade_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build_adapter(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "pinocchio adapter ok" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_pinocchio_shaped_facade_writes_the_oracles_results_into_ikiddata(tmp_path):
    """ctor (const Model&, IkIdData&), Eigen / aligned-vector arguments, results in the
    caller-owned IkIdData -- compared with the oracle field by field at 1e-10 (tests/cpp/pinocchio_adapter_test.cpp)."""
    exe = _build_adapter(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "pinocchio adapter ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_facade_solves_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "facade ok" in r.stdout, r.stdout + r.stderr
