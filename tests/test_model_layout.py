"""Host-side bookkeeping of loik_create, checked without a GPU through loik_model_layout: how the kinematic tree is cut
into register-carried chains (segments), which edges hand their contribution over through pending blocks, the level /
warp schedule of the segment-parallel kernel, the spans of the one-warp-per-tile kernel and the tile-record size.

The invariants are what the kernels rely on (loik_b200/csrc/loik_solver.cu, k_iterate / k_iterate_seg): one writer per
pending block, children swept before parents on the way to the root (backward levels) and parents before children on
the way out (forward levels), every joint in exactly one segment and one span.
"""
import numpy as np
import pytest

from loik_b200 import problems, robots, solver

JR_ROWS, TR_ROWS, PR_ROWS, FR_ROWS, GR_ROWS = 89, 81, 33, 127, 49  # loik_device.cuh


def _layout(model, nc=1):
    return solver.model_layout(model, problems.bench_params(nc))


def _check_invariants(model, L, nc):
    nb, parent = model.nb, model.parent
    J, segs, spans = L["joints"], L["segs"], L["spans"]
    nvj = [0] + [model.nv_joint(i) for i in range(1, model.nj)]
    nchild = np.bincount(parent[1:], minlength=model.nj)
    # ---- carry: the contribution travels in registers iff the parent is the previous joint, has no other child and
    # neither end is a multi-DoF joint; every other edge to a non-root parent owns one pending block
    pouts = []
    for i in range(1, model.nj):
        j = J[i - 1]
        expect = parent[i] > 0 and parent[i] == i - 1 and nchild[parent[i]] == 1 and nvj[i] == 1 and nvj[parent[i]] == 1
        assert j["carry"] == int(expect), i
        if parent[i] > 0 and not expect:
            assert j["pout"] >= 0
            pouts.append(j["pout"])
        else:
            assert j["pout"] == -1
        assert j["npin"] == sum(1 for c in range(1, model.nj) if parent[c] == i and not J[c - 1]["carry"])
        assert (j["mblk"] >= 0) == (nvj[i] > 1)
    assert sorted(pouts) == list(range(L["npend"]))  # single writer per block, no gaps
    assert sorted(j["mblk"] for j in J if j["mblk"] >= 0) == list(range(L["nmd"]))
    # ---- spans: consecutive, cover 1..nb, a multi-DoF joint is a span of its own
    assert spans[0]["lo"] == 1 and spans[-1]["hi"] == nb
    for a, b in zip(spans, spans[1:]):
        assert b["lo"] == a["hi"] + 1
    for s in spans:
        if s["md"]:
            assert s["lo"] == s["hi"] and nvj[s["lo"]] == s["md"]
        else:
            assert all(nvj[i] == 1 for i in range(s["lo"], s["hi"] + 1))
    # ---- tile record size
    rows = GR_ROWS + JR_ROWS * nb + TR_ROWS * max(nc, 1) + PR_ROWS * max(L["npend"], 1) + FR_ROWS * L["nmd"]  # (the debug residual vectors live in an arena of their own)
    assert L["rows"] == rows
    # ---- segments
    if L["nwarp"] == 1:  # one warp sweeps the whole tree in joint order
        assert L["nseg"] == 1 and (segs[0]["lo"], segs[0]["hi"]) == (1, nb)
        return
    assert 2 <= L["nwarp"] <= 4
    seg_of = {}
    for g, s in enumerate(segs):
        assert J[s["lo"] - 1]["carry"] == 0
        for i in range(s["lo"], s["hi"] + 1):
            assert i not in seg_of
            seg_of[i] = g
            if i > s["lo"]:
                assert J[i - 1]["carry"] == 1 and parent[i] == i - 1
        assert 0 <= s["bwarp"] < L["nwarp"] and 0 <= s["fwarp"] < L["nwarp"]
        assert 0 <= s["blevel"] < L["nblevel"] and 0 <= s["flevel"] < L["nflevel"]
    assert sorted(seg_of) == list(range(1, nb + 1))
    has_child = set()
    for g, s in enumerate(segs):
        p = parent[s["lo"]]
        if p > 0:
            pg = seg_of[p]
            has_child.add(pg)
            assert segs[pg]["blevel"] > s["blevel"]  # children are swept before their parent on the way to the root
            assert s["flevel"] == segs[pg]["flevel"] + 1  # and after it on the way out
        else:
            assert s["flevel"] == 0
    for g, s in enumerate(segs):
        if g not in has_child:
            assert s["blevel"] == 0
    # a level's segments are spread over the warps: no warp idles while another holds two segments of that level
    for key, wkey, nlev in (("blevel", "bwarp", L["nblevel"]), ("flevel", "fwarp", L["nflevel"])):
        for lv in range(nlev):
            ws = [s[wkey] for s in segs if s[key] == lv]
            assert ws, (key, lv)
            assert len(set(ws)) == min(len(ws), L["nwarp"])


@pytest.mark.parametrize("name", sorted(robots.ROBOTS))
def test_layout_of_the_robots(name):
    model = robots.get_robot(name)
    nc = len(robots.TASK_JOINTS[name])
    L = _layout(model, nc)
    _check_invariants(model, L, nc)
    if name in ("panda", "ur10", "ur10c"):  # serial chains: everything in registers
        assert (L["npend"], L["nseg"], L["nwarp"]) == (0, 1, 1)
        assert all(j["carry"] for j in L["joints"][1:])


def test_talos_segments():
    """Fixed-base Talos: legs 6 + 6, torso 2, arms 8 + 8, head 2; the torso waits for the arms and
    the head, which hang off its last joint through three pending blocks."""
    L = _layout(robots.talos(), 2)
    assert [(s["lo"], s["hi"]) for s in L["segs"]] == [(1, 6), (7, 12), (13, 14), (15, 22), (23, 30), (31, 32)]
    assert (L["npend"], L["nwarp"], L["nblevel"], L["nflevel"]) == (3, 4, 2, 2)
    assert [s["blevel"] for s in L["segs"]] == [0, 0, 1, 0, 0, 0]
    assert [s["flevel"] for s in L["segs"]] == [0, 0, 0, 1, 1, 1]
    assert L["joints"][14 - 1]["npin"] == 3


@pytest.mark.parametrize("seed", range(12))
def test_layout_of_random_trees(seed):
    model = robots.random_tree(8 + 4 * (seed % 6), seed, branching=0.15 + 0.05 * (seed % 4), continuous=0.2 * (seed % 2),
                               multidof=0.25 * (seed % 3))
    L = _layout(model, 2)
    _check_invariants(model, L, 2)


def test_too_many_branching_children():
    star = robots._build("star", [("root", 0, "R", "z", (0, 0, 0), (0, 0, 0), -1, 1, 1.0)] +
                         [(f"c{i}", 1, "R", "x", (0.1 * i, 0, 0), (0, 0, 0), -1, 1, 1.0) for i in range(7)])
    with pytest.raises(RuntimeError, match="kMaxPin"):
        _layout(star)
    ok = robots._build("star6", [("root", 0, "R", "z", (0, 0, 0), (0, 0, 0), -1, 1, 1.0)] +
                       [(f"c{i}", 1, "R", "x", (0.1 * i, 0, 0), (0, 0, 0), -1, 1, 1.0) for i in range(6)])
    L = _layout(ok)
    assert L["npend"] == 6 and L["joints"][0]["npin"] == 6
    _check_invariants(ok, L, 1)


# ---- the step table of the lane-parallel kernel's wide geometry (k_iterate_lane<4>, loik_lane.cuh) ---------------------
WF_VALID, WF_FIRST, WF_GIVE, WF_ROOT, WF_PINS = 1, 2, 4, 8, 16
LJ_ROWS, LJ_SLOT_WIDE = 96, 108


def _check_wide_table(model, nc=1):
    params = problems.bench_params(nc)
    L = solver.model_layout(model, params)
    W = solver.wide_table(model, params)
    J, parent = L["joints"], model.parent
    seg_of, c = {}, -1  # chains = maximal runs of register-carried joints
    for i in range(1, model.nj):
        c += 0 if J[i - 1]["carry"] else 1
        seg_of[i] = c
    for direction in ("backward", "forward"):
        at = {}  # joint -> (step, group)
        prev = [None] * 4
        for s, step in enumerate(W[direction]):
            assert len(step) == 4
            for g, e in enumerate(step):
                if not e["flags"] & WF_VALID:
                    assert e["joint"] == 1 and e["flags"] == 0  # a group without work: joint 1, nothing stored
                    prev[g] = None
                    continue
                i = e["joint"]
                assert i not in at, "every joint is swept exactly once"
                at[i] = (s, g)
                assert e["parent"] == parent[i]
                assert bool(e["flags"] & WF_ROOT) == (parent[i] == 0)
                assert bool(e["flags"] & WF_PINS) == (J[i - 1]["npin"] > 0)
                assert bool(e["flags"] & WF_GIVE) == (parent[i] > 0 and not J[i - 1]["carry"])
                if e["flags"] & WF_GIVE:
                    assert e["pout"] == J[i - 1]["pout"]
                # a chain continues on the same group in the next step; its first step is flagged
                cont = prev[g] is not None and seg_of[prev[g]] == seg_of[i] and prev[g] == (i + 1 if direction == "backward" else i - 1)
                assert bool(e["flags"] & WF_FIRST) == (not cont)
                if not e["flags"] & WF_FIRST:  # inside a chain the hand-over is the register carry
                    assert J[i if direction == "backward" else i - 1]["carry"] == 1
                prev[g] = i
        assert sorted(at) == list(range(1, model.nj))
        for i in range(1, model.nj):  # dependencies: strictly earlier steps, or the previous step of the same group (registers)
            p = parent[i]
            if p == 0:
                continue
            (si, gi), (sp, gp) = at[i], at[p]
            if direction == "backward":
                assert si < sp
                if J[i - 1]["carry"]:
                    assert gi == gp and sp == si + 1
            else:
                assert sp < si
    # ---- blocks of the joints in the shared-memory record: inside their slot, 16 B aligned, 4 x group (mod 16) doubles
    offs = {}
    for step in W["backward"]:
        for g, e in enumerate(step):
            if e["flags"] & WF_VALID:
                i = e["joint"]
                offs[i] = e["loff"]
                assert LJ_SLOT_WIDE * (i - 1) <= e["loff"] and e["loff"] + LJ_ROWS <= LJ_SLOT_WIDE * i
                assert e["loff"] % 2 == 0 and e["loff"] % 16 == 4 * g
    for step in W["forward"]:
        for e in step:
            if e["flags"] & WF_VALID:
                assert e["loff"] == offs[e["joint"]]
                if e["parent"] > 0:
                    assert e["parent_loff"] == offs[e["parent"]]
    return W


@pytest.mark.parametrize("name", ["panda", "ur10", "talos", "talos_ff"])
def test_wide_table_robots(name):
    model = robots.get_robot(name)
    if any(model.nv_joint(i) > 1 for i in range(1, model.nj)):
        pytest.skip("the lane-parallel kernel covers trees of 1-DoF joints")
    W = _check_wide_table(model)
    if name == "talos":  # the critical path: an arm (8 joints) and the torso (2), with the legs and the head alongside
        assert len(W["backward"]) == 10 and len(W["forward"]) == 10
    if name == "panda":
        assert len(W["backward"]) == 7


@pytest.mark.parametrize("seed", range(12))
def test_wide_table_random_trees(seed):
    model = robots.random_tree(4 + 5 * seed, seed=100 + seed, branching=0.35)
    _check_wide_table(model, nc=2)
