"""The C-ABI library loads and exports every symbol include/loik_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "loik_b200.h")).read()
    return sorted(set(re.findall(r"LOIK_API\s+[\w\s\*]+?\b(loik_\w+)\s*\(", src)))


def test_header_symbols_exported():
    from loik_b200 import build, solver
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/loik_b200.h but not exported"
    assert sorted(solver.EXPORTS) == names
    lib.loik_abi_version.restype = ctypes.c_int32
    assert lib.loik_abi_version() == 2  # (2: A_per_instance in loik_solve_init / _full / _task)


def test_enums_match_header():
    from loik_b200 import solver
    src = open(os.path.join(ROOT, "include", "loik_b200.h")).read()
    body = src[src.index("typedef enum loik_norm_index"):src.index("} loik_norm_index")]
    names = re.findall(r"LOIK_N_\w+", body)
    assert len(names) == len(solver.NORM_NAMES)
    body = src[src.index("typedef enum loik_field"):src.index("} loik_field")]
    fields = re.findall(r"^\s*(LOIK_F_\w+)", body, flags=re.M)
    assert len(fields) == 23 and fields[solver.F_RESIDUALS] == "LOIK_F_RESIDUALS" and fields[solver.F_Z] == "LOIK_F_Z"


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path fails loudly (it must never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from loik_b200 import problems, robots, solver
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        solver.make_solver(robots.panda(), problems.bench_params(1), 8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "loik_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle-vs", "").lower() or f == "robots.py" or "import oracle" not in txt, f
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_model_validation_needs_no_gpu():
    """loik_create validates the model before it touches CUDA: the reference-style error texts come back on any box."""
    import numpy as np
    from loik_b200 import problems, robots, solver
    P = problems.bench_params(1)
    m = robots.panda()
    m.jtype = m.jtype.copy(); m.jtype[3] = 99
    with pytest.raises(RuntimeError, match="unsupported joint type"):
        solver.make_solver(m, P, 4)
    m = robots.panda()
    m.parent = m.parent.copy(); m.parent[2] = 5
    with pytest.raises(RuntimeError, match="parents"):
        solver.make_solver(m, P, 4)
    many = robots._build("many_md", [(f"s{i}", i, "S", None, (0.1, 0, 0), (0, 0, 0), None, None, 1.0) for i in range(17)])  # kMaxMd = 16
    with pytest.raises(RuntimeError, match="multi-DoF"):
        solver.make_solver(many, P, 4)
    with pytest.raises(RuntimeError, match="equality constraint dimension is not 6"):
        solver.make_solver(robots.panda(), dict(P, eq_c_dim=3), 4)


def test_header_is_plain_c_and_cxx(tmp_path):
    """The boundary is a C ABI: the header must compile on its own as C99 and as C++ (no CUDA / torch types)."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    cases = [("gcc", "t.c", ["-std=c99", "-pedantic", "-Werror"]), ("g++", "t.cpp", ["-std=c++17", "-Werror"])]
    for cc, fname, flags in cases:
        if not shutil.which(cc):
            pytest.skip(f"{cc} not installed")
        src = tmp_path / fname
        src.write_text('#include "loik_b200.h"\nint main(void) { loik_params p; loik_model_desc m; (void)p; (void)m; return (int)sizeof(loik_solver*) == 0; }\n')
        subprocess.check_call([cc, *flags, "-Wall", "-fsyntax-only", "-I", inc, str(src)])
