"""Analysis aid (not a test of product code): warp-iteration model of the batched solve's launch schedule (DESIGN.md sections 2 and 9: "lanes idling next to
unfinished neighbours inside a launch cost ~10 %").

`plan` restates run_schedule (loik_b200/csrc/loik_solver.cu): `dense` sweeps on the home arena, then migrating launches
of 1,1,2,2,4,4,...,64 iterations (`reps` launches per chunk size, chunk x `growth`, capped at 64) until `budget` =
max_iter sweeps are covered; after every launch the survivors are packed into full tiles.  A warp runs a launch for as
long as its slowest lane needs (at most the launch's iteration count), so the schedule spends
sum over launches and warps of max over lanes of min(chunk, remaining) warp-iterations against the ideal
sum(iterations) / 32.  The per-instance iteration counts come from the CPU oracle on the bench's Panda batch.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots  # noqa: E402


def plan(budget, dense=4, reps=2, growth=2.0, cap=64):
    out, done = [], 0
    if dense > 0:
        out.append(min(dense, budget))
        done = out[0]
    chunk, r = 1, 0
    while done < budget:
        c = min(chunk, budget - done)
        out.append(c)
        done += c
        r += 1
        if r == reps:
            r = 0
            if chunk < cap:
                chunk = max(chunk + 1, int(chunk * growth))
    return out


def warp_iterations(iters, sched):
    rem = np.asarray(iters, np.int64).copy()  # iterations every instance still needs, in slot order
    total = 0
    for c in sched:
        if rem.size == 0:
            break
        r = np.concatenate([rem, np.zeros((-rem.size) % 32, np.int64)]).reshape(-1, 32)
        total += int(np.minimum(r, c).max(axis=1).sum())
        rem = rem - c
        rem = rem[rem > 0]  # the survivors claim the dense prefix of the next launch, order preserved
    return total, int(rem.size)


def _panda_iteration_counts(B=16384):
    from oracle import recursion
    model = robots.panda()
    pb = problems.random_batch(model, B, seed=0)
    params = problems.bench_params(1)
    ref = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"],
                                pb["ub"], nthreads=4, want_outputs=True)
    return ref["iters"], params["max_iter"]


if __name__ == "__main__":
    iters, max_iter = _panda_iteration_counts()
    ideal = iters.sum() / 32
    spent, left = warp_iterations(iters, plan(max_iter))
    never, _ = warp_iterations(iters, [max_iter])
    print(f"default schedule {plan(max_iter)}: {spent / ideal:.3f} x the ideal warp-iterations "
          f"(mean {iters.mean():.2f} iterations per instance, {left} instances left); one launch of {max_iter} iterations: {never / ideal:.2f} x")
    for d in (3, 4, 5):
        for r in (2, 3):
            print(f"  dense {d}, reps {r}: {warp_iterations(iters, plan(max_iter, dense=d, reps=r))[0] / ideal:.3f}")
