"""ctypes front-end of ``oracle/loik_oracle.c`` (oracle B, the recursion restatement).

TEST INFRASTRUCTURE ONLY -- see the header of ``loik_oracle.c``.  The class mirrors the public
surface of ``FirstOrderLoikOptimizedTpl`` (``loik-loid-optimized.hpp:129-755``) so the parity tests
read like ``/root/reference/tests/loik-loid.cpp``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libloik_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False, march: str | None = None, out: str | None = None, fp_contract: str = "off") -> str:
    """Compile the oracle with gcc (``make -C oracle``).  Returns the .so path.  ``fp_contract="fast"`` (with
    ``march="native"``: FMA) builds the same source with a different rounding, for the sensitivity test."""
    if out is not None:
        src = os.path.join(_HERE, "loik_oracle.c")
        flags = ["-O3", f"-march={march or 'native'}", "-std=c99", "-fPIC", "-fvisibility=hidden", f"-ffp-contract={fp_contract}"]
        subprocess.check_call(["gcc", *flags, "-shared", "-o", out, src, "-lm", "-lpthread"])
        return out
    src = os.path.join(_HERE, "loik_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        args = ["make", "-C", _HERE] + (["-B"] if force else [])
        if march:
            args.append(f"MARCH={march}")
        subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _as_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _as_i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(_dp if a.dtype == np.float64 else _ip)


def load(path: str | None = None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    lib = C.CDLL(path or build())
    lib.lo_create.restype = C.c_void_p
    lib.lo_create.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                              C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
    lib.lo_destroy.argtypes = [C.c_void_p]
    for fn in ("lo_reset_inf_norms", "lo_update_prev", "lo_reset_solver_public", "lo_fwd_pass1", "lo_bwd_pass",
               "lo_fwd_pass2", "lo_box_proj", "lo_dual_update", "lo_compute_residuals", "lo_check_convergence",
               "lo_check_feasibility"):
        getattr(lib, fn).argtypes = [C.c_void_p]
        getattr(lib, fn).restype = None
    lib.lo_update_mu.argtypes = [C.c_void_p]
    lib.lo_update_mu.restype = C.c_int
    lib.lo_fwd_pass_init.argtypes = [C.c_void_p, _dp]
    lib.lo_update_references.argtypes = [C.c_void_p, _dp, _dp]
    lib.lo_solve_init.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, _ip, _dp, _dp, _dp, _dp]
    lib.lo_solve_init.restype = C.c_int
    lib.lo_solve.argtypes = [C.c_void_p]
    lib.lo_solve.restype = C.c_int
    lib.lo_solve_full.argtypes = lib.lo_solve_init.argtypes
    lib.lo_solve_full.restype = C.c_int
    lib.lo_solve_task.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp]
    lib.lo_solve_task.restype = C.c_int
    lib.lo_last_error.argtypes = [C.c_void_p]
    lib.lo_last_error.restype = C.c_char_p
    lib.lo_set_max_iter.argtypes = [C.c_void_p, C.c_int]
    lib.lo_set_warm_start.argtypes = [C.c_void_p, C.c_int]
    lib.lo_array.argtypes = [C.c_void_p, C.c_char_p, _ip]
    lib.lo_array.restype = _dp
    lib.lo_scalar.argtypes = [C.c_void_p, C.c_char_p]
    lib.lo_scalar.restype = C.c_double
    lib.lo_batch_solve.restype = C.c_long
    lib.lo_batch_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int,
                                   _dp, _dp, _dp, _ip, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   _dp, _dp, _dp, _dp, _ip, _ip, _dp]
    if path is None:
        _lib = lib
    return lib


# The reference fixture's hyper-parameters (tests/loik-loid.cpp:91-105).
FIXTURE_PARAMS = dict(max_iter=2, tol_abs=1e-3, tol_rel=1e-3, tol_primal_inf=1e-2, tol_dual_inf=1e-2, rho=1e-5, mu=1e-2,
                      mu_equality_scale_factor=1e4, mu_update_strat=0, num_eq_c=1, eq_c_dim=6, warm_start=False,
                      tol_tail_solve=1e-1)

_SHAPES6 = {"vis", "vis_prev", "pis", "pis_aba", "fis", "delta_fis", "fis_diff_plus_Aty", "delta_fis_diff_plus_Aty",
            "Href_v", "Hv", "yis", "delta_yis", "Aty", "Av_minus_b", "Atb"}
_SHAPES36 = {"His", "His_aba", "H_refs", "AtA", "U_full", "UDinv_full", "S_full", "Dinv_full"}
# pinocchio JointData of 1-DoF joints: U / UDinv / S are the first column, Dinv the (0,0) entry of the per-joint blocks
_FIRST_COL = {"U": "U_full", "UDinv": "UDinv_full", "S": "S_full"}


class FirstOrderLoikOptimized:
    """Oracle B.  Same ctor argument order as ``loik-loid-optimized.hpp:129-134`` (model replaces
    ``pinocchio::Model``; the IkIdData lives inside the object; verbose/logging are dropped)."""

    def __init__(self, model, max_iter, tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu,
                 mu_equality_scale_factor, mu_update_strat=0, num_eq_c=1, eq_c_dim=6, warm_start=False,
                 tol_tail_solve=1e-1, lib=None):
        self._lib = lib or load()
        self.model = model
        if eq_c_dim != 6:
            raise RuntimeError("[IkProblemFormulation::IkProblemFormulation]: equality constraint dimension is not 6, "
                               "problem formulation not supported !!!")
        self._keep = [_as_i(model.parent), _as_i(model.jtype), _as_d(model.axis), _as_d(model.placement_R),
                      _as_d(model.placement_p)]
        k = self._keep
        self._h = self._lib.lo_create(model.nj, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), int(max_iter), tol_abs,
                                      tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_equality_scale_factor,
                                      int(mu_update_strat), int(num_eq_c), int(eq_c_dim), int(bool(warm_start)),
                                      tol_tail_solve)
        if not self._h:
            raise RuntimeError("lo_create failed")
        self.num_eq_c = num_eq_c

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lo_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._lib.lo_last_error(self._h).decode())

    # ---- reference interface -------------------------------------------------------------
    def _problem_args(self, q, H_ref, v_ref, ids, Ais, bis, lb, ub):
        ids = _as_i(ids)
        a = [_as_d(q), _as_d(H_ref), _as_d(v_ref), ids, _as_d(Ais), _as_d(bis), _as_d(lb), _as_d(ub)]
        if a[4].reshape(-1, 36).shape[0] != ids.shape[0] or a[5].reshape(-1, 6).shape[0] != ids.shape[0]:
            raise RuntimeError("[IkProblemFormulation::UpdateEqConstraints]: task_constraint_ids, Ais, and bis have "
                               "different size !!!")
        if a[6].shape != a[7].shape:
            raise RuntimeError("[IkProblemFormulation::UpdateIneqConstraints]: lower bound and upper bound have "
                               "different dimensions!!!")
        if a[6].shape[0] != self.model.nv:
            raise RuntimeError("IkProblemFormulation::UpdateIneqConstraints]: inequality constraint dimension has "
                               "changed, this is not supported currently!!!")
        return a

    def SolveInit(self, q, H_ref, v_ref, active_task_constraint_ids, Ais, bis, lb, ub):
        a = self._problem_args(q, H_ref, v_ref, active_task_constraint_ids, Ais, bis, lb, ub)
        self._check(self._lib.lo_solve_init(self._h, _p(a[0]), _p(a[1]), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]),
                                            _p(a[5]), _p(a[6]), _p(a[7])))

    def Solve(self, *args):
        if len(args) == 0:
            self._check(self._lib.lo_solve(self._h))
        elif len(args) == 8:
            a = self._problem_args(*args)
            self._check(self._lib.lo_solve_full(self._h, _p(a[0]), _p(a[1]), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]),
                                                _p(a[5]), _p(a[6]), _p(a[7])))
        elif len(args) == 4:
            q, c_id, Ai, bi = args
            q, Ai, bi = _as_d(q), _as_d(Ai), _as_d(bi)
            self._check(self._lib.lo_solve_task(self._h, _p(q), int(c_id), _p(Ai), _p(bi)))
        else:
            raise TypeError("Solve() takes 0, 4 or 8 arguments")

    def ResetSolver(self):
        self._lib.lo_reset_solver_public(self._h)

    def UpdatePrev(self):
        self._lib.lo_update_prev(self._h)

    def ResetInfNorms(self):
        self._lib.lo_reset_inf_norms(self._h)

    def FwdPassInit(self, q):
        q = _as_d(q)
        self._lib.lo_fwd_pass_init(self._h, _p(q))

    def FwdPass1(self):
        self._lib.lo_fwd_pass1(self._h)

    def BwdPassOptimizedVisitor(self):
        self._lib.lo_bwd_pass(self._h)

    def FwdPass2OptimizedVisitor(self):
        self._lib.lo_fwd_pass2(self._h)

    def BoxProj(self):
        self._lib.lo_box_proj(self._h)

    def DualUpdate(self):
        self._lib.lo_dual_update(self._h)

    def ComputeResiduals(self):
        self._lib.lo_compute_residuals(self._h)

    def CheckConvergence(self):
        self._lib.lo_check_convergence(self._h)

    def CheckFeasibility(self):
        self._lib.lo_check_feasibility(self._h)

    def UpdateMu(self):
        self._check(self._lib.lo_update_mu(self._h))

    def UpdateReferences(self, H_refs, v_refs):
        H_refs, v_refs = _as_d(H_refs), _as_d(v_refs)
        if H_refs.size != 36 * self.model.nj or v_refs.size != 6 * self.model.nj:
            raise RuntimeError("[IkProblemFormulation::UpdateReferences]: input arguments 'H_refs', 'v_refs' have wrong size!!")
        self._lib.lo_update_references(self._h, _p(H_refs), _p(v_refs))

    def set_max_iter(self, m):
        self._lib.lo_set_max_iter(self._h, int(m))

    def set_warm_start(self, ws):
        self._lib.lo_set_warm_start(self._h, int(bool(ws)))

    # ---- state access (the reference exposes these as public IkIdData members / getters) ----
    def array(self, name) -> np.ndarray:
        if name in _FIRST_COL:
            return self.array(_FIRST_COL[name])[:, :, 0].copy()
        if name == "Dinv":
            return self.array("Dinv_full")[:, 0, 0].copy()
        n = C.c_int(0)
        ptr = self._lib.lo_array(self._h, name.encode(), C.byref(n))
        if not ptr:
            raise KeyError(name)
        a = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0)
        if name in _SHAPES6:
            return a.reshape(-1, 6)
        if name in _SHAPES36:
            return a.reshape(-1, 6, 6)
        if name.endswith("_R"):
            return a.reshape(-1, 3, 3)
        if name.endswith("_p"):
            return a.reshape(-1, 3)
        return a

    def scalar(self, name) -> float:
        v = self._lib.lo_scalar(self._h, name.encode())
        if np.isnan(v):
            raise KeyError(name)
        return v

    def __getattr__(self, name):
        # data members: solver.z, solver.nu, solver.vis ...; getters: solver.get_iter() ...
        if name.startswith("get_"):
            key = name[4:]
            alias = {"convergence_status": "converged", "primal_infeasibility_status": "primal_infeasible",
                     "dual_infeasibility_status": "dual_infeasible", "delta_z_qp_inf_norm": "delta_z_inf_norm"}
            key = alias.get(key, key)
            if key in ("primal_residual_vec", "dual_residual_vec"):
                return lambda: self.array(key)
            if key in ("iter", "tail_solve_iter"):
                return lambda: int(self.scalar(key))
            if key in ("converged", "primal_infeasible", "dual_infeasible", "primal_infeasibility_cond_1",
                       "primal_infeasibility_cond_2"):
                return lambda: bool(self.scalar(key))
            return lambda: self.scalar(key)
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self.array(name)
        except KeyError:
            try:
                return self.scalar(name)
            except KeyError:
                raise AttributeError(name) from None


def batch_solve(model, params: dict, q, H_ref, v_ref, ids, Ais, bis, lb, ub, *, mode=0, fixed_iters=0, nthreads=1,
                want_outputs=True, lib=None):
    """CPU baseline: solve a batch with one oracle solver per thread.

    q [B,nq]; bis [B,nc,6] or [nc,6]; lb/ub [nv] or [B,nv].  Returns dict(z,nu,w,y,iters,status,mu,total_iters).
    """
    lib = lib or load()
    q = _as_d(q)
    B = q.shape[0]
    nv, nc = model.nv, len(ids)
    bis = _as_d(bis)
    lb, ub = _as_d(lb), _as_d(ub)
    b_per = int(bis.ndim == 3)
    bounds_per = int(lb.ndim == 2)
    ids = _as_i(ids)
    H_ref, v_ref, Ais = _as_d(H_ref), _as_d(v_ref), _as_d(Ais)
    k = [_as_i(model.parent), _as_i(model.jtype), _as_d(model.axis), _as_d(model.placement_R), _as_d(model.placement_p)]
    out = {}
    if want_outputs:
        out = dict(z=np.zeros((B, nv)), nu=np.zeros((B, nv)), w=np.zeros((B, nv)), y=np.zeros((B, nc, 6)),
                   iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32), mu=np.zeros(B))
    null_d, null_i = C.cast(None, _dp), C.cast(None, _ip)
    g = lambda name: _p(out[name]) if want_outputs else (null_i if name in ("iters", "status") else null_d)
    p = params
    tot = lib.lo_batch_solve(model.nj, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), int(p["max_iter"]), p["tol_abs"],
                             p["tol_rel"], p["tol_primal_inf"], p["tol_dual_inf"], p["rho"], p["mu"],
                             p["mu_equality_scale_factor"], int(p.get("mu_update_strat", 0)), nc, p["tol_tail_solve"], B,
                             _p(q), _p(H_ref), _p(v_ref), _p(ids), _p(Ais), _p(bis), b_per, _p(lb), _p(ub), bounds_per,
                             int(mode), int(fixed_iters), int(nthreads), g("z"), g("nu"), g("w"), g("y"), g("iters"),
                             g("status"), g("mu"))
    out["total_iters"] = int(tot)
    return out


def batch_track(model, params: dict, q, H_ref, v_ref, ids, Ais, bis0, bis1, lb, ub, *, c_id, dt, steps, warm=True, nthreads=1,
                lib=None):
    """CPU baseline of the trajectory-tracking loop: per instance Solve(...), then `steps` times
    { q <- integrate(q, dt z); Solve(q, c_id, A, b_t) }, b_t blending bis0 -> bis1 (lo_batch_track).
    Returns dict(z [B,nv], q [B,nq], step_iters [B,steps], total_iters)."""
    lib = lib or load()
    q = _as_d(q)
    B = q.shape[0]
    nc = len(ids)
    bis0, bis1, lb, ub = _as_d(bis0), _as_d(bis1), _as_d(lb), _as_d(ub)
    ids = _as_i(ids)
    H_ref, v_ref, Ais = _as_d(H_ref), _as_d(v_ref), _as_d(Ais)
    k = [_as_i(model.parent), _as_i(model.jtype), _as_d(model.axis), _as_d(model.placement_R), _as_d(model.placement_p)]
    out = dict(z=np.zeros((B, model.nv)), q=np.zeros((B, model.nq)), step_iters=np.zeros((B, steps), np.int32))
    p = params
    lib.lo_batch_track.restype = C.c_long
    lib.lo_batch_track.argtypes = ([C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int] + [C.c_double] * 7 + [C.c_int, C.c_int, C.c_double,
                                   C.c_int, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_int,
                                   C.c_int, _dp, _dp, _ip])
    tot = lib.lo_batch_track(model.nj, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), int(p["max_iter"]), p["tol_abs"],
                             p["tol_rel"], p["tol_primal_inf"], p["tol_dual_inf"], p["rho"], p["mu"],
                             p["mu_equality_scale_factor"], int(p.get("mu_update_strat", 0)), nc, p["tol_tail_solve"], B,
                             _p(q), _p(H_ref), _p(v_ref), _p(ids), _p(Ais), _p(bis0), _p(bis1), _p(lb), _p(ub), int(c_id),
                             float(dt), int(steps), int(bool(warm)), int(nthreads), _p(out["z"]), _p(out["q"]),
                             _p(out["step_iters"]))
    out["total_iters"] = int(tot)
    return out
