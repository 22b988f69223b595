"""Host-side mirror of the reference solver interface over the C ABI of ``libloik_b200.so``.

``FirstOrderLoikOptimized`` here has the constructor argument order and the method names of
``loik::FirstOrderLoikOptimizedTpl`` (``/root/reference/include/loik/loik-loid-optimized.hpp:129-134,
335-338,368,475-478,596-597``) but solves a *batch* of independent instances on one B200: ``q`` is
``[B, nq]``, ``bis`` is ``[B, nc, 6]`` and the results (``z``, ``nu``, ``w``, ``yis`` ...) come back as
``[B, ...]``.  All arithmetic happens in the hand-written CUDA kernels behind ``include/loik_b200.h``;
this file only marshals pointers (numpy arrays = host buffers, torch CUDA tensors = device buffers).

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but creating a
solver without CUDA raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libloik_b200.so")
_lib = None

LOIK_HOST, LOIK_DEVICE, LOIK_HOST_PINNED = 0, 1, 2

# loik_field (include/loik_b200.h)
(F_Z, F_NU, F_W, F_Y, F_V, F_F, F_ATY, F_FDPA, F_STF_PLUS_W, F_H, F_P, F_UDINV, F_DINV, F_R, F_LIMI, F_MU, F_ITER,
 F_STATUS, F_RESIDUALS, F_NORMS, F_PRIMAL_RES_VEC, F_DUAL_RES_VEC, F_Q) = range(23)
(STEP_BACKWARD, STEP_FORWARD, STEP_RESIDUAL, STEP_UPDATE_PREV, STEP_RESET_INF_NORMS, STEP_FWD_PASS1, STEP_BWD_PASS,
 STEP_FWD_PASS2, STEP_BOX_PROJ, STEP_DUAL_UPDATE, STEP_COMPUTE_RESIDUALS, STEP_CHECK_CONVERGENCE, STEP_CHECK_FEASIBILITY,
 STEP_UPDATE_MU) = range(14)

NORM_NAMES = ["bT_delta_y_plus", "bT_delta_y_minus", "Av_inf_norm", "nu_inf_norm", "Href_v_inf_norm",
              "fis_diff_plus_Aty_inf_norm", "Stf_plus_w_inf_norm", "delta_fis_diff_plus_Aty_inf_norm",
              "delta_Stf_plus_w_inf_norm", "delta_vis_inf_norm", "delta_nu_inf_norm", "delta_z_inf_norm",
              "delta_fis_inf_norm", "delta_yis_inf_norm", "delta_w_inf_norm", "primal_residual_task",
              "primal_residual_slack", "dual_residual_v", "dual_residual_nu", "delta_y_qp_inf_norm",
              "A_qp_T_delta_y_qp_inf_norm", "ub_qp_T_delta_y_qp_plus", "lb_qp_T_delta_y_qp_minus",
              "primal_infeasibility_cond_1", "primal_infeasibility_cond_2", "delta_x_qp_inf_norm", "converged",
              "primal_infeasible"]

EXPORTS = ["loik_abi_version", "loik_last_error", "loik_create", "loik_destroy", "loik_model_layout", "loik_wide_table", "loik_solve_init",
           "loik_update_references", "loik_update_references_batch", "loik_solve", "loik_solve_full", "loik_solve_task", "loik_integrate", "loik_iterate_fixed",
           "loik_fwd_pass_init", "loik_reset_recursion", "loik_step", "loik_set_debug", "loik_set_logging", "loik_history_capacity", "loik_get_history", "loik_set_keep_workspace", "loik_get", "loik_get_stats", "loik_reduce_stats", "loik_launch_count",
           "loik_set_max_iter", "loik_set_rho", "loik_set_mu", "loik_set_mu_equality_scale_factor", "loik_set_tol_abs",
           "loik_set_tol_rel", "loik_set_tol_primal_inf", "loik_set_tol_dual_inf", "loik_set_tol_tail_solve",
           "loik_set_warm_start", "loik_get_params", "loik_get_schedule", "loik_set_schedule",
           "loik_active_count_device_ptr", "loik_solve_begin", "loik_solve_chunk", "loik_reset_solver"]


class _ModelDesc(C.Structure):
    _fields_ = [("njoints", C.c_int32), ("parents", C.POINTER(C.c_int32)), ("joint_types", C.POINTER(C.c_int32)),
                ("joint_axes", C.POINTER(C.c_double)), ("placement_R", C.POINTER(C.c_double)),
                ("placement_p", C.POINTER(C.c_double))]


class _Params(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("tol_abs", C.c_double), ("tol_rel", C.c_double),
                ("tol_primal_inf", C.c_double), ("tol_dual_inf", C.c_double), ("rho", C.c_double), ("mu", C.c_double),
                ("mu_equality_scale_factor", C.c_double), ("mu_update_strat", C.c_int32), ("num_eq_c", C.c_int32),
                ("eq_c_dim", C.c_int32), ("warm_start", C.c_int32), ("tol_tail_solve", C.c_double),
                ("verbose", C.c_int32), ("logging", C.c_int32)]


class _Schedule(C.Structure):
    _fields_ = [("dense_sweeps", C.c_int32), ("repack_reps", C.c_int32), ("repack_growth", C.c_double),
                ("hi_priority_after", C.c_int32), ("seg_after", C.c_int32), ("seg_warps", C.c_int32),
                ("lane_after", C.c_int32), ("use_graph", C.c_int32), ("small_after", C.c_int32), ("small_grid", C.c_int32),
                ("lane_warps_per_cta", C.c_int32), ("lane_groups_per_instance", C.c_int32), ("drop_workspace", C.c_int32), ("lane_hard_first_ratio", C.c_double),
                ("lane_available", C.c_int32), ("lane_warps_chosen", C.c_int32), ("lane_groups_chosen", C.c_int32), ("lane_ctas", C.c_int32),
                ("lane_smem_bytes", C.c_int32)]


def load_library(path: str | None = None):
    """dlopen libloik_b200.so (built in-tree by ``loik_b200.build``).  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("LOIK_B200_LIB") or LIB_PATH  # LOIK_B200_LIB: A/B builds of the same ABI (development)
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -m loik_b200.build` "
                           f"(libloik_b200 has no CPU fallback)")
    lib = C.CDLL(p)
    vp, i32, dp, ip = C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p
    lib.loik_abi_version.restype = i32
    lib.loik_last_error.restype = C.c_char_p
    lib.loik_create.argtypes = [C.POINTER(_ModelDesc), C.POINTER(_Params), i32, i32, C.POINTER(vp)]
    lib.loik_destroy.argtypes = [vp]
    lib.loik_destroy.restype = None
    prob = [dp, dp, dp, i32, ip, dp, i32, dp, i32, dp, dp, i32, i32, vp]
    lib.loik_solve_init.argtypes = [vp] + prob
    lib.loik_solve_full.argtypes = [vp] + prob
    lib.loik_update_references.argtypes = [vp, dp, dp, vp]
    lib.loik_update_references_batch.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.loik_solve.argtypes = [vp, vp]
    lib.loik_solve_task.argtypes = [vp, dp, i32, dp, i32, dp, i32, i32, vp]
    lib.loik_iterate_fixed.argtypes = [vp, i32, i32, vp]
    lib.loik_integrate.argtypes = [vp, C.c_double, vp]
    lib.loik_reset_recursion.argtypes = [vp, vp]
    lib.loik_fwd_pass_init.argtypes = [vp, dp, i32, vp]
    lib.loik_step.argtypes = [vp, i32, vp]
    lib.loik_set_debug.argtypes = [vp, i32]
    lib.loik_set_logging.argtypes = [vp, i32]
    lib.loik_history_capacity.argtypes = [vp]
    lib.loik_history_capacity.restype = i32
    lib.loik_get_history.argtypes = [vp, vp, i32, vp]
    lib.loik_model_layout.argtypes = [vp, vp, C.POINTER(C.c_int32), i32]
    lib.loik_model_layout.restype = i32
    lib.loik_wide_table.argtypes = [vp, vp, C.POINTER(C.c_int32), i32]
    lib.loik_wide_table.restype = i32
    lib.loik_set_keep_workspace.argtypes = [vp, i32]
    lib.loik_get.argtypes = [vp, i32, vp, i32, vp]
    lib.loik_get_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.loik_reduce_stats.argtypes = [vp, vp, C.POINTER(vp)]
    lib.loik_launch_count.argtypes = [vp]
    lib.loik_launch_count.restype = C.c_int64
    lib.loik_set_max_iter.argtypes = [vp, i32]
    lib.loik_set_rho.argtypes = [vp, C.c_double]
    lib.loik_set_mu.argtypes = [vp, C.c_double]
    lib.loik_set_tol_tail_solve.argtypes = [vp, C.c_double]
    for name in ("mu_equality_scale_factor", "tol_abs", "tol_rel", "tol_primal_inf", "tol_dual_inf"):
        getattr(lib, "loik_set_" + name).argtypes = [vp, C.c_double]
    lib.loik_get_params.argtypes = [vp, C.POINTER(_Params)]
    lib.loik_get_schedule.argtypes = [vp, C.POINTER(_Schedule)]
    lib.loik_set_schedule.argtypes = [vp, C.POINTER(_Schedule)]
    lib.loik_set_warm_start.argtypes = [vp, i32]
    lib.loik_active_count_device_ptr.argtypes = [vp, C.POINTER(vp)]
    lib.loik_solve_begin.argtypes = [vp, vp]
    lib.loik_solve_chunk.argtypes = [vp, i32, vp]
    lib.loik_reset_solver.argtypes = [vp, vp]
    if path is None:
        _lib = lib
    return lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _current_stream() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_stream().cuda_stream)
    except Exception:
        pass
    return 0


class _DevArray:
    """Zero-copy view of a device buffer owned by the library."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class _Buf:
    """A pointer + where it lives; keeps the backing object alive."""

    def __init__(self, x, dtype=np.float64):
        if _is_torch(x):
            import torch
            want = torch.float64 if dtype == np.float64 else torch.int32
            if x.dtype != want or not x.is_contiguous():
                x = x.to(want).contiguous()
            self.obj = x
            self.ptr = x.data_ptr()
            self.loc = LOIK_DEVICE if x.is_cuda else (LOIK_HOST_PINNED if x.is_pinned() else LOIK_HOST)
            self.shape = tuple(x.shape)
        else:
            a = np.ascontiguousarray(x, dtype=dtype)
            self.obj = a
            self.ptr = a.ctypes.data
            self.loc = LOIK_HOST
            self.shape = a.shape


def _dry_args(model, params):
    keep = [np.ascontiguousarray(model.parent, np.int32), np.ascontiguousarray(model.jtype, np.int32),
            np.ascontiguousarray(model.axis, np.float64), np.ascontiguousarray(model.placement_R, np.float64),
            np.ascontiguousarray(model.placement_p, np.float64)]
    md = _ModelDesc(model.nj, keep[0].ctypes.data_as(C.POINTER(C.c_int32)), keep[1].ctypes.data_as(C.POINTER(C.c_int32)),
                    keep[2].ctypes.data_as(C.POINTER(C.c_double)), keep[3].ctypes.data_as(C.POINTER(C.c_double)),
                    keep[4].ctypes.data_as(C.POINTER(C.c_double)))
    P = params
    pr = _Params(int(P["max_iter"]), P["tol_abs"], P["tol_rel"], P["tol_primal_inf"], P["tol_dual_inf"], P["rho"], P["mu"],
                 P["mu_equality_scale_factor"], int(P.get("mu_update_strat", 0)), int(P["num_eq_c"]), int(P.get("eq_c_dim", 6)),
                 int(bool(P.get("warm_start", False))), P.get("tol_tail_solve", 1e-1), 0, 0)
    return keep, md, pr


def wide_table(model, params, lib=None):
    """Step table of the lane-parallel kernel's wide geometry (``loik_wide_table``; no CUDA needed): per sweep direction a
    list of steps, each a list of four dicts (one per 8-lane group)."""
    lib = lib or load_library()
    keep, md, pr = _dry_args(model, params)
    cap = 2 + 2 * 4 * 7 * model.nj
    out = (C.c_int32 * cap)()
    n = lib.loik_wide_table(C.byref(md), C.byref(pr), out, cap)
    if n < 0:
        raise RuntimeError(lib.loik_last_error().decode())
    v = list(out[:n])
    nsb, nsf = v[:2]
    names = ("joint", "parent", "flags", "loff", "pout", "sidx", "parent_loff")
    ent = [dict(zip(names, v[2 + 7 * e:2 + 7 * e + 7])) for e in range(4 * (nsb + nsf))]
    steps = [ent[4 * s:4 * s + 4] for s in range(nsb + nsf)]
    return dict(backward=steps[:nsb], forward=steps[nsb:])


def model_layout(model, params, lib=None):
    """Host-side bookkeeping ``loik_create`` derives from a model (no CUDA needed): how the tree is cut into
    register-carried chains, which edges go through pending blocks, the level / warp schedule of the segment-parallel
    kernel, the spans of the one-warp kernel and the tile-record size.  Raises like the constructor on a bad model."""
    lib = lib or load_library()
    keep = [np.ascontiguousarray(model.parent, np.int32), np.ascontiguousarray(model.jtype, np.int32),
            np.ascontiguousarray(model.axis, np.float64), np.ascontiguousarray(model.placement_R, np.float64),
            np.ascontiguousarray(model.placement_p, np.float64)]
    md = _ModelDesc(model.nj, keep[0].ctypes.data_as(C.POINTER(C.c_int32)), keep[1].ctypes.data_as(C.POINTER(C.c_int32)),
                    keep[2].ctypes.data_as(C.POINTER(C.c_double)), keep[3].ctypes.data_as(C.POINTER(C.c_double)),
                    keep[4].ctypes.data_as(C.POINTER(C.c_double)))
    P = params
    pr = _Params(int(P["max_iter"]), P["tol_abs"], P["tol_rel"], P["tol_primal_inf"], P["tol_dual_inf"], P["rho"], P["mu"],
                 P["mu_equality_scale_factor"], int(P.get("mu_update_strat", 0)), int(P["num_eq_c"]), int(P.get("eq_c_dim", 6)),
                 int(bool(P.get("warm_start", False))), P.get("tol_tail_solve", 1e-1), 0, 0)
    cap = 8 + 4 * model.nj + 6 * 64 + 3 * 64
    out = (C.c_int32 * cap)()
    n = lib.loik_model_layout(C.byref(md), C.byref(pr), out, cap)
    if n < 0:
        raise RuntimeError(lib.loik_last_error().decode())
    v = list(out[:n])
    nb = model.nj - 1
    head = dict(zip(("rows", "npend", "nseg", "nwarp", "nblevel", "nflevel", "nspan", "nmd"), v[:8]))
    o = 8
    joints = [dict(zip(("carry", "pout", "npin", "mblk"), v[o + 4 * i:o + 4 * i + 4])) for i in range(nb)]
    o += 4 * nb
    segs = [dict(zip(("lo", "hi", "bwarp", "blevel", "fwarp", "flevel"), v[o + 6 * g:o + 6 * g + 6])) for g in range(head["nseg"])]
    o += 6 * head["nseg"]
    spans = [dict(zip(("lo", "hi", "md"), v[o + 3 * g:o + 3 * g + 3])) for g in range(head["nspan"])]
    return dict(head, joints=joints, segs=segs, spans=spans)


class FirstOrderLoikOptimized:
    """Batched drop-in for ``loik::FirstOrderLoikOptimizedTpl<double>``.

    Ctor arguments follow ``loik-loid-optimized.hpp:129-134``; ``model`` is a
    :class:`loik_b200.robots.RobotModel` (the flat view of ``pinocchio::Model``), the caller-owned
    ``IkIdData`` of the reference lives inside the handle (HBM), and ``batch`` / ``device`` are new.
    """

    def __init__(self, max_iter, tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_equality_scale_factor,
                 mu_update_strat, num_eq_c, eq_c_dim, model, batch=1, warm_start=False, tol_tail_solve=1e-1,
                 verbose=False, logging=False, device=0, lib=None):
        self._lib = lib or load_library()
        self.model = model
        self.batch = int(batch)
        self.nc = int(num_eq_c)
        self.max_iter = int(max_iter)
        self._keep = [np.ascontiguousarray(model.parent, np.int32), np.ascontiguousarray(model.jtype, np.int32),
                      np.ascontiguousarray(model.axis, np.float64), np.ascontiguousarray(model.placement_R, np.float64),
                      np.ascontiguousarray(model.placement_p, np.float64)]
        k = self._keep
        md = _ModelDesc(model.nj, k[0].ctypes.data_as(C.POINTER(C.c_int32)), k[1].ctypes.data_as(C.POINTER(C.c_int32)),
                        k[2].ctypes.data_as(C.POINTER(C.c_double)), k[3].ctypes.data_as(C.POINTER(C.c_double)),
                        k[4].ctypes.data_as(C.POINTER(C.c_double)))
        pr = _Params(int(max_iter), tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_equality_scale_factor,
                     int(mu_update_strat), int(num_eq_c), int(eq_c_dim), int(bool(warm_start)), tol_tail_solve,
                     int(bool(verbose)), int(bool(logging)))
        h = C.c_void_p()
        self._h = None
        self._check(self._lib.loik_create(C.byref(md), C.byref(pr), self.batch, int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.loik_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._lib.loik_last_error().decode())

    # ---- problem set-up + solves (reference names) ----------------------------------------------
    def _prob(self, q, H_ref, v_ref, ids, Ais, bis, lb, ub):
        B, nc, nv = self.batch, self.nc, self.model.nv
        ids = np.ascontiguousarray(ids, np.int32)
        H = _Buf(np.asarray(H_ref, np.float64).reshape(36))
        vr = _Buf(np.asarray(v_ref, np.float64).reshape(6))
        # Ais: [nc, 6, 6] shared by the batch (host), or [B, nc, 6, 6]: every instance its own task matrices
        a_per = int(_is_torch(Ais) and Ais.dim() == 4 or (not _is_torch(Ais)) and np.ndim(Ais) == 4)
        if a_per:
            A = _Buf(Ais)
            if tuple(A.shape) != (B, ids.shape[0], 6, 6):
                raise RuntimeError("[IkProblemFormulation::UpdateEqConstraints]: task_constraint_ids, Ais, and bis have "
                                   "different size !!!")
        else:
            A = _Buf(np.asarray(Ais.cpu() if _is_torch(Ais) else Ais, np.float64).reshape(-1, 36))
            if A.shape[0] != ids.shape[0]:
                raise RuntimeError("[IkProblemFormulation::UpdateEqConstraints]: task_constraint_ids, Ais, and bis have "
                                   "different size !!!")
        if ids.shape[0] != nc:
            raise RuntimeError("[IkProblemFormulation::UpdateEqConstraints]: number of equality constraints doesn't "
                               "match initialization!!!")
        qb, bb, lbb, ubb = _Buf(q), _Buf(bis), _Buf(lb), _Buf(ub)
        if qb.shape not in ((B, self.model.nq),) and not (B == 1 and qb.shape == (self.model.nq,)):
            raise RuntimeError(f"q must be [batch={B}, nq={self.model.nq}]")
        nb_elems = int(np.prod(bb.shape))
        if nb_elems == B * nc * 6 and (B > 1 or len(bb.shape) == 3):
            b_per = 1
        elif nb_elems == nc * 6:
            b_per = 0
        else:
            raise RuntimeError("[IkProblemFormulation::UpdateEqConstraints]: task_constraint_ids, Ais, and bis have "
                               "different size !!!")
        if lbb.shape != ubb.shape:
            raise RuntimeError("[IkProblemFormulation::UpdateIneqConstraints]: lower bound and upper bound have "
                               "different dimensions!!!")
        if lbb.shape == (nv,):
            bd_per = 0
        elif lbb.shape == (B, nv):
            bd_per = 1
        else:
            raise RuntimeError("IkProblemFormulation::UpdateIneqConstraints]: inequality constraint dimension has "
                               "changed, this is not supported currently!!!")
        locs = {qb.loc, bb.loc} | ({lbb.loc, ubb.loc} if bd_per else set()) | ({A.loc} if a_per else set())
        if len(locs) != 1:
            raise RuntimeError("q, bis (and per-instance Ais / bounds) must all be host arrays or all be CUDA tensors")
        loc = locs.pop()
        if not bd_per and lbb.loc == LOIK_DEVICE:  # batch-shared bounds are always read on the host
            lbb, ubb = _Buf(lbb.obj.cpu().numpy()), _Buf(ubb.obj.cpu().numpy())
        keep = (qb, H, vr, ids, A, bb, lbb, ubb)
        args = (qb.ptr, H.ptr, vr.ptr, int(ids.shape[0]), ids.ctypes.data, A.ptr, a_per, bb.ptr, b_per, lbb.ptr, ubb.ptr,
                bd_per, loc, _current_stream())
        return keep, args

    def SolveInit(self, q, H_ref, v_ref, active_task_constraint_ids, Ais, bis, lb, ub):
        keep, args = self._prob(q, H_ref, v_ref, active_task_constraint_ids, Ais, bis, lb, ub)
        self._check(self._lib.loik_solve_init(self._h, *args))

    def Solve(self, *args):
        if len(args) == 0:
            self._check(self._lib.loik_solve(self._h, _current_stream()))
        elif len(args) == 8:
            keep, a = self._prob(*args)
            self._check(self._lib.loik_solve_full(self._h, *a))
        elif len(args) == 4:
            q, c_id, Ai, bi = args
            bb = _Buf(bi)
            if Ai is None:  # UpdateEqConstraint(c_id, bi) (ik-id-description-optimized.hpp:224): the task keeps its matrix
                class _Null:
                    ptr = None
                a_per, Ab = 0, _Null()
            else:
                a_per = int((Ai.dim() if _is_torch(Ai) else np.ndim(Ai)) == 3)  # [B, 6, 6]: every instance its own Ai
                Ab = _Buf(Ai) if a_per else _Buf(np.asarray(Ai.cpu() if _is_torch(Ai) else Ai, np.float64).reshape(36))
                if a_per and (tuple(Ab.shape) != (self.batch, 6, 6) or Ab.loc != bb.loc):
                    raise RuntimeError("a per-instance Ai must be [batch, 6, 6] and live where bi lives")
            b_per = int(int(np.prod(bb.shape)) == self.batch * 6 and (self.batch > 1 or len(bb.shape) == 2))
            if q is None:  # keep the device-resident configuration (after Integrate)
                self._check(self._lib.loik_solve_task(self._h, None, int(c_id), Ab.ptr, a_per, bb.ptr, b_per, bb.loc,
                                                      _current_stream()))
                return
            qb = _Buf(q)
            if qb.loc != bb.loc:
                raise RuntimeError("q and bi must both be host arrays or both be CUDA tensors")
            self._check(self._lib.loik_solve_task(self._h, qb.ptr, int(c_id), Ab.ptr, a_per, bb.ptr, b_per, qb.loc,
                                                  _current_stream()))
        else:
            raise TypeError("Solve() takes 0, 4 or 8 arguments")

    def UpdateReferences(self, H_refs, v_refs):
        """problem_.UpdateReferences(H_refs, v_refs) (ik-id-description-optimized.hpp:103-121).  ``v_refs`` of shape
        ``[batch][njoints][6]`` (numpy or a CUDA tensor) gives every instance its own reference velocities, ``H_refs`` of shape
        ``[batch][njoints][6][6]`` its own weights as well."""
        if getattr(v_refs, "ndim", 0) == 3:
            h_per = int(getattr(H_refs, "ndim", 0) == 4)  # [batch, njoints, 6, 6]: every instance its own weights
            Hb = _Buf(H_refs) if h_per else _Buf(np.asarray(H_refs, np.float64).reshape(-1))
            vb = _Buf(v_refs)
            ok_H = tuple(Hb.shape) == (self.batch, self.model.nj, 6, 6) and Hb.loc == vb.loc if h_per else Hb.shape[0] == 36 * self.model.nj
            if not ok_H or tuple(vb.shape) != (self.batch, self.model.nj, 6):
                raise RuntimeError("[IkProblemFormulation::UpdateReferences]: input arguments 'H_refs', 'v_refs' have wrong size!!")
            self._check(self._lib.loik_update_references_batch(self._h, Hb.ptr, h_per, vb.ptr, vb.loc, _current_stream()))
            return
        Hb, vb = _Buf(np.asarray(H_refs, np.float64).reshape(-1)), _Buf(np.asarray(v_refs, np.float64).reshape(-1))
        if Hb.shape[0] != 36 * self.model.nj or vb.shape[0] != 6 * self.model.nj:
            raise RuntimeError("[IkProblemFormulation::UpdateReferences]: input arguments 'H_refs', 'v_refs' have wrong size!!")
        self._check(self._lib.loik_update_references(self._h, Hb.ptr, vb.ptr, _current_stream()))

    def ResetSolver(self):
        """ResetSolver() (loik-loid-optimized.hpp:168-186): iteration counter, flags, mu and the feasibility scalars only --
        the primal / dual state (nu, z, w, vis, fis, yis ...) is kept."""
        self._check(self._lib.loik_reset_solver(self._h, _current_stream()))

    # ---- fused steps (parity tests) -------------------------------------------------------------
    def ResetRecursion(self):
        self._check(self._lib.loik_reset_recursion(self._h, _current_stream()))

    def StepBackward(self):
        """UpdatePrev + ResetInfNorms + FwdPass1 + BwdPassOptimizedVisitor."""
        self._check(self._lib.loik_step(self._h, STEP_BACKWARD, _current_stream()))

    def StepForward(self):
        """FwdPass2OptimizedVisitor + BoxProj + DualUpdate + ComputePrimalResiduals."""
        self._check(self._lib.loik_step(self._h, STEP_FORWARD, _current_stream()))

    def StepResidual(self):
        """ComputeDualResiduals + CheckConvergence + CheckFeasibility + UpdateMu + loop control."""
        self._check(self._lib.loik_step(self._h, STEP_RESIDUAL, _current_stream()))

    def Integrate(self, dt):
        """q <- q + dt * z and FwdPassInit(q) on the device (outer IK loop)."""
        self._check(self._lib.loik_integrate(self._h, float(dt), _current_stream()))

    # ---- the reference's public per-step methods, one by one (loik-loid-optimized.hpp:192-264); need set_debug(True)
    def _fine(self, step):
        self._check(self._lib.loik_step(self._h, step, _current_stream()))

    def FwdPassInit(self, q):
        qb = _Buf(q)
        self._check(self._lib.loik_fwd_pass_init(self._h, qb.ptr, qb.loc, _current_stream()))

    def UpdatePrev(self):
        self._fine(STEP_UPDATE_PREV)

    def ResetInfNorms(self):
        self._fine(STEP_RESET_INF_NORMS)

    def FwdPass1(self):
        self._fine(STEP_FWD_PASS1)

    def BwdPassOptimizedVisitor(self):
        self._fine(STEP_BWD_PASS)

    def FwdPass2OptimizedVisitor(self):
        self._fine(STEP_FWD_PASS2)

    def BoxProj(self):
        self._fine(STEP_BOX_PROJ)

    def DualUpdate(self):
        self._fine(STEP_DUAL_UPDATE)

    def ComputeResiduals(self):
        self._fine(STEP_COMPUTE_RESIDUALS)

    def CheckConvergence(self):
        self._fine(STEP_CHECK_CONVERGENCE)

    def CheckFeasibility(self):
        self._fine(STEP_CHECK_FEASIBILITY)

    def UpdateMu(self):
        self._fine(STEP_UPDATE_MU)

    def IterateFixed(self, iters, reset=True):
        self._check(self._lib.loik_iterate_fixed(self._h, int(iters), int(bool(reset)), _current_stream()))

    def set_debug(self, on=True):
        self._check(self._lib.loik_set_debug(self._h, int(bool(on))))

    def set_logging(self, on=True):
        """logging_ / LoikSolverInfo (loik-loid-optimized.hpp:47-127): keep the per-iteration log of the following solves."""
        self._check(self._lib.loik_set_logging(self._h, int(bool(on))))

    HISTORY_COLS = ("primal_residual_task", "primal_residual_slack", "dual_residual_v", "dual_residual_nu", "mu",
                    "delta_x_qp_inf_norm", "delta_z_inf_norm", "tail_solve")

    def history(self):
        """[batch][capacity][8] solver log (columns: HISTORY_COLS); rows [0, get_iter()[i]) of instance i are valid."""
        cap = int(self._lib.loik_history_capacity(self._h))
        out = np.empty((self.batch, cap, len(self.HISTORY_COLS)))
        self._check(self._lib.loik_get_history(self._h, out.ctypes.data, LOIK_HOST, _current_stream()))
        return out

    def set_keep_workspace(self, on=True):
        """His / pis / UDinv / Dinv / r of the last backward pass stay readable after Solve(), as the reference leaves
        them in ik_id_data (tests/loik-loid.cpp:597-615); off, those getters raise after a solve instead of returning
        stale rows."""
        self._check(self._lib.loik_set_keep_workspace(self._h, int(bool(on))))

    # chunked solve for the multi-GPU driver
    def SolveBegin(self):
        self._check(self._lib.loik_solve_begin(self._h, _current_stream()))

    def SolveChunk(self, iters):
        self._check(self._lib.loik_solve_chunk(self._h, int(iters), _current_stream()))

    def active_count_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.loik_active_count_device_ptr(self._h, C.byref(p)))
        return int(p.value)

    # ---- setters / getters ----------------------------------------------------------------------
    def set_max_iter(self, m):
        self._check(self._lib.loik_set_max_iter(self._h, int(m)))
        self.max_iter = int(m)

    def set_rho(self, rho):
        self._check(self._lib.loik_set_rho(self._h, float(rho)))

    def set_mu(self, mu):
        self._check(self._lib.loik_set_mu(self._h, float(mu)))

    def set_tol_tail_solve(self, tol):
        self._check(self._lib.loik_set_tol_tail_solve(self._h, float(tol)))

    def set_warm_start(self, ws):
        self._check(self._lib.loik_set_warm_start(self._h, int(bool(ws))))

    def set_mu_equality_scale_factor(self, f):
        self._check(self._lib.loik_set_mu_equality_scale_factor(self._h, float(f)))

    def set_tol_abs(self, tol):
        self._check(self._lib.loik_set_tol_abs(self._h, float(tol)))

    def set_tol_rel(self, tol):
        self._check(self._lib.loik_set_tol_rel(self._h, float(tol)))

    def set_tol_primal_inf(self, tol):
        self._check(self._lib.loik_set_tol_primal_inf(self._h, float(tol)))

    def set_tol_dual_inf(self, tol):
        self._check(self._lib.loik_set_tol_dual_inf(self._h, float(tol)))

    def get_params(self) -> dict:
        """The hyper-parameters as the solver holds them (get_max_iter ... get_tol_dual_inf, task-solver-base.hpp:87-141)."""
        pr = _Params()
        self._check(self._lib.loik_get_params(self._h, C.byref(pr)))
        return {name: getattr(pr, name) for name, _ in _Params._fields_}

    def get_rho(self):
        return self.get_params()["rho"]

    def get_max_iter(self):
        return self.get_params()["max_iter"]

    def get_tol_primal_inf(self):
        return self.get_params()["tol_primal_inf"]

    def get_tol_dual_inf(self):
        return self.get_params()["tol_dual_inf"]

    def get_schedule(self) -> dict:
        sc = _Schedule()
        self._check(self._lib.loik_get_schedule(self._h, C.byref(sc)))
        return {name: getattr(sc, name) for name, _ in _Schedule._fields_}

    def set_schedule(self, **kw):
        """Change how a batched Solve() is laid out on the GPU (include/loik_b200.h: loik_schedule); unknown keys raise."""
        sc = _Schedule()
        self._check(self._lib.loik_get_schedule(self._h, C.byref(sc)))
        for k, v in kw.items():
            if k not in dict(_Schedule._fields_) or k in ("lane_available", "lane_warps_chosen", "lane_groups_chosen", "lane_ctas", "lane_smem_bytes"):
                raise KeyError(k)
            setattr(sc, k, v)
        self._check(self._lib.loik_set_schedule(self._h, C.byref(sc)))

    def _field_shape(self, field):
        nb, nc, nv, nq = self.model.nb, self.nc, self.model.nv, self.model.nq
        return {F_Z: (nv,), F_NU: (nv,), F_W: (nv,), F_Y: (nc, 6), F_V: (nb, 6), F_F: (nb, 6), F_ATY: (nc, 6),
                F_FDPA: (nb, 6), F_STF_PLUS_W: (nv,), F_H: (nb, 6, 6), F_P: (nb, 6), F_UDINV: (nb, 6), F_DINV: (nb,),
                F_R: (nv,), F_LIMI: (nb, 12), F_MU: (), F_ITER: (), F_STATUS: (), F_RESIDUALS: (4,),
                F_NORMS: (len(NORM_NAMES),), F_PRIMAL_RES_VEC: (6 * nb + nv,), F_DUAL_RES_VEC: (6 * nb + nv,),
                F_Q: (nq,)}[field]

    def get(self, field, out=None):
        """Copy a per-instance field of the whole batch out: numpy array [B, ...] (or into a CUDA tensor)."""
        shape = (self.batch,) + self._field_shape(field)
        dtype = np.int32 if field in (F_ITER, F_STATUS) else np.float64
        if out is None:
            out = np.empty(shape, dtype)
            self._check(self._lib.loik_get(self._h, field, out.ctypes.data, LOIK_HOST, _current_stream()))
            return out
        # an output buffer is written in place: it must already have the right dtype, shape and layout (a converted
        # copy would silently receive the data instead of `out`)
        if _is_torch(out):
            import torch
            want = torch.int32 if dtype == np.int32 else torch.float64
            if out.dtype != want or not out.is_contiguous() or tuple(out.shape) != shape:
                raise RuntimeError(f"get(out=): need a contiguous {want} tensor of shape {shape}")
        else:
            if not isinstance(out, np.ndarray) or out.dtype != dtype or not out.flags.c_contiguous or out.shape != shape:
                raise RuntimeError(f"get(out=): need a C-contiguous {np.dtype(dtype).name} array of shape {shape}")
        b = _Buf(out, dtype)
        self._check(self._lib.loik_get(self._h, field, b.ptr, b.loc, _current_stream()))
        return out

    z = property(lambda self: self.get(F_Z))
    nu = property(lambda self: self.get(F_NU))
    w = property(lambda self: self.get(F_W))
    yis = property(lambda self: self.get(F_Y))
    vis = property(lambda self: self.get(F_V))
    fis = property(lambda self: self.get(F_F))
    Aty = property(lambda self: self.get(F_ATY))
    fis_diff_plus_Aty = property(lambda self: self.get(F_FDPA))
    Stf_plus_w = property(lambda self: self.get(F_STF_PLUS_W))
    His = property(lambda self: self.get(F_H))
    pis = property(lambda self: self.get(F_P))
    UDinv = property(lambda self: self.get(F_UDINV))
    Dinv = property(lambda self: self.get(F_DINV))
    r = property(lambda self: self.get(F_R))
    liMi = property(lambda self: self.get(F_LIMI))
    q = property(lambda self: self.get(F_Q))

    def get_iter(self):
        return self.get(F_ITER)

    def get_mu(self):
        return self.get(F_MU)

    def get_status(self):
        return self.get(F_STATUS)

    def get_convergence_status(self):
        return (self.get(F_STATUS) & 1).astype(bool)

    def get_primal_infeasibility_status(self):
        return ((self.get(F_STATUS) >> 1) & 1).astype(bool)

    def get_dual_infeasibility_status(self):
        return np.zeros(self.batch, bool)  # never evaluated by the optimized path

    def get_primal_residual(self):
        return self.get(F_RESIDUALS)[:, 0]

    def get_dual_residual(self):
        return self.get(F_RESIDUALS)[:, 1]

    def get_tol_primal(self):
        return self.get(F_RESIDUALS)[:, 2]

    def get_tol_dual(self):
        return self.get(F_RESIDUALS)[:, 3]

    def get_primal_residual_vec(self):
        return self.get(F_PRIMAL_RES_VEC)

    def get_dual_residual_vec(self):
        return self.get(F_DUAL_RES_VEC)

    def norms(self) -> dict:
        a = self.get(F_NORMS)
        return {n: a[:, i] for i, n in enumerate(NORM_NAMES)}

    def stats(self) -> dict:
        out = (C.c_int64 * 5)()
        self._check(self._lib.loik_get_stats(self._h, out))
        return dict(converged=out[0], primal_infeasible=out[1], max_iter=out[2], total_iters=out[3], sweeps=out[4])

    def reduce_stats_ptr(self) -> int:
        """Enqueue the status reduction on the current stream; device pointer to 4 int64 (see loik_reduce_stats)."""
        p = C.c_void_p()
        self._check(self._lib.loik_reduce_stats(self._h, _current_stream(), C.byref(p)))
        return int(p.value)

    def stats_tensor(self):
        """Device tensor (4 int64: #converged, #primal infeasible, #max_iter, sum of iterations) of the last solve: the
        reduction is enqueued on the current stream; a zero-copy view of a library buffer, valid until the next call."""
        import torch
        return torch.as_tensor(_DevArray(self.reduce_stats_ptr(), 4, "<i8"), device=f"cuda:{self.device}")

    def active_tensor(self):
        """Device tensor (1 int32): instances still active after the last SolveChunk (zero-copy view, see stats_tensor)."""
        import torch
        return torch.as_tensor(_DevArray(self.active_count_ptr(), 1, "<i4"), device=f"cuda:{self.device}")

    def launch_count(self) -> int:
        return int(self._lib.loik_launch_count(self._h))


def make_solver(model, params: dict, batch: int, device: int = 0) -> FirstOrderLoikOptimized:
    p = params
    return FirstOrderLoikOptimized(p["max_iter"], p["tol_abs"], p["tol_rel"], p["tol_primal_inf"], p["tol_dual_inf"],
                                   p["rho"], p["mu"], p["mu_equality_scale_factor"], p.get("mu_update_strat", 0),
                                   p["num_eq_c"], p.get("eq_c_dim", 6), model, batch=batch,
                                   warm_start=p.get("warm_start", False), tol_tail_solve=p["tol_tail_solve"],
                                   logging=p.get("logging", False), device=device)
