"""ctypes binding of the C ABI (include/srk/ba_c_api.h).  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsrk_ba.so")

SOLVER_AUTO, SOLVER_DENSE_CHOLESKY, SOLVER_BLOCK_PCG = 0, 1, 2
STOP_REASONS = {0: "", 1: "abs err threshold", 2: "small relative err change", 3: "hessian overflow",
                4: "err converged to limit value", 5: "", 6: "max iterations"}
TIMING_FAMILIES = ("jacobian", "frame_blocks", "schur", "solve", "backsub", "update", "residual", "allreduce", "solve_factor", "solve_trsv")


class SrkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("srk error %d: %s" % (code, msg))
        self.code = code


class _Problem(C.Structure):
    _fields_ = [("n_cams", C.c_int64), ("n_points", C.c_int64), ("n_obs", C.c_int64),
                ("obs_cam", C.c_void_p), ("obs_point", C.c_void_p), ("obs_xy", C.c_void_p),
                ("points", C.c_void_p), ("cams", C.c_void_p), ("K", C.c_void_p),
                ("shared_K", C.c_int32), ("f0", C.c_double)]


class _Options(C.Structure):
    _fields_ = [("has_err_change", C.c_int32), ("err_change", C.c_double), ("has_max_hessian_factor", C.c_int32),
                ("max_hessian_factor", C.c_double), ("unity_comp_ind", C.c_int32), ("unity_comp_value", C.c_double),
                ("max_outer_iters", C.c_int32), ("solver", C.c_int32), ("pcg_max_iters", C.c_int32), ("pcg_rel_tol", C.c_double),
                ("refine_steps", C.c_int32)]


class _Report(C.Structure):
    _fields_ = [("converged", C.c_int32), ("stop_reason", C.c_int32), ("outer_iters", C.c_int32), ("attempts", C.c_int32),
                ("err_initial", C.c_double), ("err_final", C.c_double), ("hessian_factor_final", C.c_double),
                ("seen_points", C.c_int64), ("err_trace", C.c_void_p), ("err_trace_cap", C.c_int32), ("err_trace_len", C.c_int32),
                ("attempt_trace", C.c_void_p), ("attempt_trace_cap", C.c_int32), ("attempt_trace_len", C.c_int32),
                ("gpu_launches", C.c_int64), ("solver_used", C.c_int32), ("pcg_iters_last", C.c_int32),
                ("pcg_rel_res_last", C.c_double), ("factor_failures", C.c_int32), ("reserved_", C.c_int32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

_lib = None


def load_library():
    """Loads libsrk_ba.so (built by surikatoko_b200/build.py).  No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrkError(-2, "CUDA library %s is missing: run `python -m surikatoko_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.srk_last_error.restype = C.c_char_p
    L.srk_stop_reason_string.restype = C.c_char_p
    L.srk_stop_reason_string.argtypes = [C.c_int32]
    L.srk_ba_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
    L.srk_ba_destroy.argtypes = [C.c_void_p]
    L.srk_ba_destroy.restype = None
    L.srk_ba_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.srk_ba_solve.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Options), C.POINTER(_Report)]
    L.srk_ba_reproj_error.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.srk_ba_bind.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Options)]
    L.srk_ba_run.argtypes = [C.c_void_p, C.POINTER(_Options), C.POINTER(_Report)]
    L.srk_ba_reset.argtypes = [C.c_void_p]
    L.srk_ba_fetch.argtypes = [C.c_void_p, C.POINTER(_Problem)]
    L.srk_ba_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p, C.c_int, C.c_int]
    L.srk_ba_debug_derivs_and_solve.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 8
    L.srk_ba_debug_derivs_and_solve_ex.argtypes = [C.c_void_p, C.c_double, C.c_int32] + [C.c_void_p] * 8 + [C.POINTER(C.c_int32)]
    L.srk_ba_debug_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.srk_ba_debug_apply.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.srk_ba_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.srk_ba_get_timing.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.srk_ba_solve_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.srk_ba_solve_order.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.srk_ba_solve_levels.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.srk_nccl_unique_id.argtypes = [C.c_void_p]
    L.srk_ba_nccl_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.srk_ba_set_nccl_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.srk_ba_pcg_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.srk_ba_default_options.argtypes = [C.POINTER(_Options)]
    L.srk_ba_default_options.restype = None
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise SrkError(rc, load_library().srk_last_error().decode())
    return rc


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def nccl_unique_id():
    """128-byte ncclUniqueId (srk_nccl_unique_id); create on rank 0 and hand to every rank."""
    buf = (C.c_ubyte * 128)()
    _check(load_library().srk_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


class BAProblem:
    """Flat problem (srk_ba_problem).  Arrays are kept alive and contiguous; points/cams are refined in place."""

    def __init__(self, obs_cam, obs_point, obs_xy, points, cams, K, shared_K, f0):
        self.obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
        self.obs_point = np.ascontiguousarray(obs_point, dtype=np.int32)
        self.obs_xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(-1, 2)
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.cams = np.ascontiguousarray(cams, dtype=np.float64).reshape(-1, 12)
        self.K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1, 9)
        self.shared_K = bool(shared_K)
        self.f0 = float(f0)
        if self.obs_cam.shape[0] != self.obs_point.shape[0] or self.obs_cam.shape[0] != self.obs_xy.shape[0]:
            raise ValueError("observation arrays disagree in length")
        if self.K.shape[0] != (1 if self.shared_K else self.cams.shape[0]):
            raise ValueError("Provide either shared K or separate K for each camera frame")

    @property
    def n_cams(self): return self.cams.shape[0]
    @property
    def n_points(self): return self.points.shape[0]
    @property
    def n_obs(self): return self.obs_cam.shape[0]

    def c_struct(self):
        return _Problem(self.n_cams, self.n_points, self.n_obs, _ptr(self.obs_cam), _ptr(self.obs_point), _ptr(self.obs_xy),
                        _ptr(self.points), _ptr(self.cams), _ptr(self.K), 1 if self.shared_K else 0, self.f0)

    def copy(self):
        return BAProblem(self.obs_cam.copy(), self.obs_point.copy(), self.obs_xy.copy(), self.points.copy(), self.cams.copy(),
                         self.K.copy(), self.shared_K, self.f0)


class BAOptions:
    def __init__(self, err_change=None, max_hessian_factor=None, unity_comp_ind=1, unity_comp_value=1.0, max_outer_iters=0,
                 solver=SOLVER_AUTO, pcg_max_iters=0, pcg_rel_tol=0.0, refine_steps=1):
        self.err_change, self.max_hessian_factor = err_change, max_hessian_factor
        self.unity_comp_ind, self.unity_comp_value = unity_comp_ind, unity_comp_value
        self.max_outer_iters, self.solver = max_outer_iters, solver
        self.pcg_max_iters, self.pcg_rel_tol, self.refine_steps = pcg_max_iters, pcg_rel_tol, refine_steps

    def c_struct(self):
        return _Options(self.err_change is not None, self.err_change or 0.0, self.max_hessian_factor is not None,
                        self.max_hessian_factor or 0.0, self.unity_comp_ind, self.unity_comp_value, self.max_outer_iters, self.solver,
                        self.pcg_max_iters, self.pcg_rel_tol, self.refine_steps)


class BAReport:
    def __init__(self, rep, err_trace, attempts):
        self.converged = bool(rep.converged)
        self.stop_reason_code = rep.stop_reason
        self.stop_reason = STOP_REASONS.get(rep.stop_reason, "")
        self.outer_iters, self.attempts_count = rep.outer_iters, rep.attempts
        self.err_initial, self.err_final = rep.err_initial, rep.err_final
        self.hessian_factor_final = rep.hessian_factor_final
        self.seen_points = rep.seen_points
        self.err_trace = err_trace[:rep.err_trace_len].copy()
        self.attempts = attempts[:rep.attempt_trace_len].copy()   # [n,4]: hessian_factor, err_new, accepted, skipped_points
        self.gpu_launches = rep.gpu_launches
        self.solver_used = rep.solver_used
        self.pcg_iters_last = rep.pcg_iters_last
        self.pcg_rel_res_last = rep.pcg_rel_res_last
        self.factor_failures = rep.factor_failures

    def __repr__(self):
        return ("BAReport(converged=%s, stop=%r, outer_iters=%d, attempts=%d, err %.9g -> %.9g, launches=%d)" %
                (self.converged, self.stop_reason, self.outer_iters, self.attempts_count, self.err_initial, self.err_final, self.gpu_launches))


class Engine:
    """One handle = one CUDA device + stream (srk_ba_create)."""

    def __init__(self, device=0, trace_cap=4096):
        self._lib = load_library()
        self._h = C.c_void_p()
        dev = C.c_int(device)
        _check(self._lib.srk_ba_create(C.byref(self._h), C.byref(dev), 1))
        self._trace_cap = trace_cap
        self._cb = None
        self._bound = None

    def close(self):
        if self._h:
            self._lib.srk_ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(self._lib.srk_ba_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_allreduce(self, fn, rank, world):
        """fn(dev_ptr: int, count: int, stream: int) -> None, in-place sum over ranks of `count` doubles."""
        def _tramp(user, dev, count, stream):
            try:
                fn(int(dev), int(count), int(stream or 0))
                return 0
            except Exception:  # pragma: no cover
                import traceback
                traceback.print_exc()
                return 1
        self._cb = ALLREDUCE_FN(_tramp)
        _check(self._lib.srk_ba_set_allreduce(self._h, self._cb, None, rank, world))

    def nccl_init(self, unique_id, rank, world):
        """The library's own NCCL exchange (srk_ba_nccl_init): no Python callback per all-reduce.  unique_id: 128 bytes from nccl_unique_id()
        on rank 0, handed to every rank; collective."""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        _check(self._lib.srk_ba_nccl_init(self._h, C.cast(buf, C.c_void_p), rank, world))

    def _new_report(self):
        rep = _Report()
        tr = np.zeros(self._trace_cap)
        att = np.zeros((self._trace_cap, 4))
        rep.err_trace = tr.ctypes.data; rep.err_trace_cap = self._trace_cap
        rep.attempt_trace = att.ctypes.data; rep.attempt_trace_cap = self._trace_cap
        return rep, tr, att

    def solve(self, problem, options=None):
        options = options or BAOptions()
        rep, tr, att = self._new_report()
        ps, os_ = problem.c_struct(), options.c_struct()
        _check(self._lib.srk_ba_solve(self._h, C.byref(ps), C.byref(os_), C.byref(rep)))
        return BAReport(rep, tr, att)

    def reproj_error(self, problem):
        err, seen = C.c_double(), C.c_int64()
        ps = problem.c_struct()
        _check(self._lib.srk_ba_reproj_error(self._h, C.byref(ps), C.byref(err), C.byref(seen)))
        return err.value, seen.value

    def bind(self, problem, options=None):
        options = options or BAOptions()
        ps, os_ = problem.c_struct(), options.c_struct()
        rc = _check(self._lib.srk_ba_bind(self._h, C.byref(ps), C.byref(os_)))
        self._bound = problem
        return rc == 0

    def run(self, options=None):
        options = options or BAOptions()
        rep, tr, att = self._new_report()
        os_ = options.c_struct()
        _check(self._lib.srk_ba_run(self._h, C.byref(os_), C.byref(rep)))
        return BAReport(rep, tr, att)

    def reset(self):
        _check(self._lib.srk_ba_reset(self._h))

    def fetch(self, problem=None):
        problem = problem or self._bound
        ps = problem.c_struct()
        _check(self._lib.srk_ba_fetch(self._h, C.byref(ps)))
        return problem

    def debug_derivs_and_solve(self, c=None, solver=SOLVER_DENSE_CHOLESKY):
        pr = self._bound
        N, M, O = pr.n_points, pr.n_cams, pr.n_obs
        nf = 10 * M - 7
        out = dict(gradE=np.zeros(3 * N + 10 * M), E=np.zeros((N, 3, 3)), G=np.zeros((M, 10, 10)), F=np.zeros((O, 3, 10)))
        if c is not None:
            out.update(S=np.zeros((nf, nf)), rhs=np.zeros(nf), skipped=np.zeros(N, dtype=np.uint8), corrections=np.zeros(3 * N + 10 * M))
        iters = C.c_int32(0)
        _check(self._lib.srk_ba_debug_derivs_and_solve_ex(self._h, -1.0 if c is None else float(c), solver, _ptr(out["gradE"]), _ptr(out["E"]),
                                                          _ptr(out["G"]), _ptr(out["F"]), _ptr(out.get("S")), _ptr(out.get("rhs")),
                                                          _ptr(out.get("skipped")), _ptr(out.get("corrections")), C.byref(iters)))
        out["pcg_iters"] = iters.value
        if "S" in out:
            out["S"] = out["S"].T.copy()  # column-major -> [row, col]
        return out

    def debug_get_state(self):
        pr = self._bound
        pts = np.zeros((pr.n_points, 3)); cams = np.zeros((pr.n_cams, 12))
        _check(self._lib.srk_ba_debug_get_state(self._h, _ptr(pts), _ptr(cams)))
        return pts, cams

    def debug_apply(self, corrections):
        corr = np.ascontiguousarray(corrections, dtype=np.float64)
        err = C.c_double()
        _check(self._lib.srk_ba_debug_apply(self._h, _ptr(corr), C.byref(err)))
        return err.value

    def solve_stats(self):
        """Structure of the last dense Cholesky factor: n_f, 64-wide block rows, non-zero tiles of L, flops executed."""
        nf, nb, nz, fl = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        _check(self._lib.srk_ba_solve_stats(self._h, C.byref(nf), C.byref(nb), C.byref(nz), C.byref(fl)))
        on, parts, mp, sb = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(self._lib.srk_ba_solve_order(self._h, C.byref(on), C.byref(parts), C.byref(mp), C.byref(sb)))
        mids, mm = C.c_int64(), C.c_int64()
        _check(self._lib.srk_ba_solve_levels(self._h, C.byref(mids), C.byref(mm)))
        return dict(n_f=nf.value, block_rows=nb.value, nonzero_tiles=nz.value, factor_flops=fl.value, ordered_n=on.value, parts=parts.value,
                    max_part_blocks=mp.value, separator_blocks=sb.value, mid_separators=mids.value, max_mid_blocks=mm.value)

    def pcg_stats(self):
        """Stored 10x10 blocks of the block-sparse reduced system and the PCG iterations executed since set_timing(True)."""
        nb, it = C.c_int64(), C.c_int64()
        _check(self._lib.srk_ba_pcg_stats(self._h, C.byref(nb), C.byref(it)))
        return dict(nnz_blocks=nb.value, iters=it.value)

    def set_timing(self, enabled):
        _check(self._lib.srk_ba_set_timing(self._h, 1 if enabled else 0))

    def get_timing(self):
        out = {}
        for name in TIMING_FAMILIES:
            last, tot, n = C.c_double(), C.c_double(), C.c_int64()
            _check(self._lib.srk_ba_get_timing(self._h, name.encode(), C.byref(last), C.byref(tot), C.byref(n)))
            out[name] = dict(ms_last=last.value, ms_total=tot.value, count=n.value)
        return out
