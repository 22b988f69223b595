#!/usr/bin/env python
"""BASELINE.json configs[0] / SURVEY.md section 8(d)(1): ONE Panda problem instance, the reference's own timing protocol
(tests/loik-loid.cpp:987-1032: SolveInit once, then N x Solve()) on one host thread with the CPU restatement
(oracle/loik_oracle.c, rebuilt -march=native), at max_iter = 2 (one iteration, the reference's setting) and at
max_iter = 200 (to convergence) -- and, when a GPU is present, the latency of the same single instance through the
batched CUDA path (batch = 1), for which the GPU is the wrong tool: it is a throughput device."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots  # noqa: E402
from oracle import recursion  # noqa: E402
from tests.helpers import ctor_kwargs, prob_args  # noqa: E402


def main():
    model = robots.get_robot(sys.argv[1] if len(sys.argv) > 1 else "panda")
    pr = problems.fixture_problem(model, 4.0)
    pb = problems.random_batch(model, 8, seed=0)
    try:
        out = os.path.join("/tmp", f"libloik_oracle_native_{os.getpid()}.so")
        recursion.build(out=out, march="native")
        lib = recursion.load(out)
    except Exception:
        lib = recursion.load()
    lib.lo_time_solve.restype = C.c_double
    lib.lo_time_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    res = {"robot": model.name, "protocol": "SolveInit once, N x Solve(), one thread (tests/loik-loid.cpp:987-1032)", "cpu": []}
    cases = [("fixture (neutral q, b = (0,0,.5,0,0,0), bounds +-4)", prob_args(pr))]
    cases += [(f"random instance {i}", (pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][i], pb["lb"], pb["ub"])) for i in range(3)]
    for what, args in cases:
        for max_iter in (2, 200):
            S = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(dict(problems.FIXTURE_PARAMS, max_iter=max_iter)), lib=lib)
            S.SolveInit(*args)
            it = C.c_int(0)
            lib.lo_time_solve(S._h, 1000, C.byref(it))
            n = 100000 if max_iter == 2 else 20000
            sec = lib.lo_time_solve(S._h, n, C.byref(it))
            res["cpu"].append({"instance": what, "max_iter": max_iter, "iterations": it.value, "us_per_solve": 1e6 * sec / n,
                               "us_per_iteration": 1e6 * sec / n / max(it.value, 1)})
    try:
        import torch
        if torch.cuda.is_available():
            from loik_b200 import solver as lk
            res["gpu_batch_1"] = []
            for what, args in cases[:2]:
                for max_iter in (2, 200):
                    G = lk.make_solver(model, dict(problems.FIXTURE_PARAMS, max_iter=max_iter), 1)
                    G.SolveInit(np.asarray(args[0])[None], args[1], args[2], args[3], args[4], np.asarray(args[5])[None], args[6], args[7])
                    st = torch.cuda.Stream()
                    with torch.cuda.stream(st):
                        for _ in range(20):
                            G.Solve()
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        for _ in range(200):
                            G.Solve()
                        torch.cuda.synchronize()
                        dt = (time.perf_counter() - t0) / 200
                    res["gpu_batch_1"].append({"instance": what, "max_iter": max_iter, "iterations": int(G.get_iter()[0]), "us_per_solve": 1e6 * dt})
                    G.close()
    except ImportError:
        pass
    print(json.dumps(res))


if __name__ == "__main__":
    main()
