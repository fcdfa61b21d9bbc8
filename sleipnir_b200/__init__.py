"""B200-native interior-point Newton step with a Sleipnir-compatible front end.

This Python package is only a thin ctypes loader used by tests/, bench.py and
__graft_entry__.py. The product is native:

* ``lib/libslpb.so``      CUDA kernels for sm_100a behind the C ABI declared in
                          ``include/slpb.h``;
* ``lib/libslpb_host.so`` the host side in C++ (``slp::Problem`` DSL, the IPM
                          driver, the benchmark problem builders).

There is no CPU fallback: without the built libraries import fails, and without
a CUDA device every solve raises ``DeviceError``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DEVICE = os.path.join(_HERE, "lib", "libslpb.so")
LIB_HOST = os.path.join(_HERE, "lib", "libslpb_host.so")

EXIT_STATUS = {
    0: "SUCCESS", 1: "CALLBACK_REQUESTED_STOP", -1: "TOO_FEW_DOFS",
    -2: "LOCALLY_INFEASIBLE", -3: "GLOBALLY_INFEASIBLE",
    -4: "FACTORIZATION_FAILED", -5: "LINE_SEARCH_FAILED",
    -6: "FEASIBILITY_RESTORATION_FAILED", -7: "NONFINITE_INITIAL_GUESS",
    -8: "DIVERGING_ITERATES", -9: "MAX_ITERATIONS_EXCEEDED", -10: "TIMEOUT",
}

ORDER_NESTED_DISSECTION, ORDER_AMD, ORDER_NATURAL, ORDER_CUSTOM = 0, 1, 2, 3
# slpb_factor_arithmetic (include/slpb.h)
ARITH_REFERENCE, ARITH_TENSOR = 0, 1

# enum slpb_output / slpb_array (include/slpb.h)
OUT_F, OUT_G, OUT_H_F, OUT_H_C, OUT_C_E, OUT_A_E, OUT_C_I, OUT_A_I = range(8)
(ARR_X, ARR_S, ARR_Y, ARR_Z, ARR_G, ARR_C_E, ARR_C_I, ARR_A_E_VAL,
 ARR_A_I_VAL, ARR_H_VAL, ARR_KKT_VAL, ARR_D, ARR_RHS, ARR_P_X, ARR_P_S,
 ARR_P_Y, ARR_P_Z, ARR_TRIAL_X, ARR_TRIAL_S, ARR_TRIAL_Y, ARR_TRIAL_Z,
 ARR_TRIAL_C_E, ARR_TRIAL_C_I) = range(23)


class DeviceError(RuntimeError):
    """The CUDA library failed (no device, CUDA error, unsupported graph)."""


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class PointInfo(C.Structure):
    _fields_ = [("f", C.c_double), ("ce_l1", C.c_double),
                ("cis_l1", C.c_double), ("log_s_sum", C.c_double),
                ("finite", C.c_int32), ("ci_all_positive", C.c_int32)]


class KktStats(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "r_inf", "r_l1", "y_l1", "z_l1", "sz_min", "sz_max", "sz_mu_l1",
        "ce_inf", "ce_l1", "cis_inf", "cis_l1", "u_r_inf", "u_y_l1", "u_z_l1",
        "u_sz_min", "u_sz_max", "u_ce_inf", "u_cis_inf", "aetce_l2", "ce_l2",
        "aitcip_l2", "cip_l2", "x_inf", "s_inf")] + [
        ("xs_finite", C.c_int32), ("pad", C.c_int32)]


class FactorInfo(C.Structure):
    _fields_ = [("n_pos", C.c_int32), ("n_neg", C.c_int32),
                ("n_zero", C.c_int32), ("zero_pivot", C.c_int32),
                ("min_abs_d", C.c_double)]


class StepInfo(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "alpha_max", "alpha_z", "g_dot_px", "sinv_dot_ps", "px_inf", "ps_inf",
        "py_inf", "pz_inf")] + [("finite", C.c_int32), ("pad", C.c_int32)]


class SymbolicStats(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nnz_kkt", C.c_int64),
                ("nnz_l", C.c_int64), ("nnz_l_stored", C.c_int64),
                ("n_supernodes", C.c_int32), ("n_levels", C.c_int32),
                ("max_front", C.c_int32), ("etree_height", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "kernel_launches", "factorizations", "solves", "evals_full",
        "evals_values", "tape_nodes", "program_bytes", "n_clusters",
        "n_program_classes", "h2d_bytes", "d2h_bytes",
        "factorizations_completed")]


# Every symbol include/slpb.h declares (tests check that the library exports
# all of them).
ABI_SYMBOLS = [
    "slpb_create", "slpb_destroy", "slpb_last_error", "slpb_comm_unique_id",
    "slpb_comm_init", "slpb_comm_agree", "slpb_upload_tape",
    "slpb_upload_rows", "slpb_finalize", "slpb_set_scaling",
    "slpb_set_ignore_constraint_hessian", "slpb_set_factor_arithmetic",
    "slpb_analyze",
    "slpb_get_permutation", "slpb_set_iterate", "slpb_get_iterate",
    "slpb_eval_current", "slpb_kkt_stats_current", "slpb_kkt_stats_trial",
    "slpb_factor", "slpb_factor_pair", "slpb_select_factor",
    "slpb_prepare_rhs", "slpb_solve",
    "slpb_soc_begin", "slpb_soc_iterate",
    "slpb_trial", "slpb_solve_trial", "slpb_accept_relinearize",
    "slpb_probe_point", "slpb_multiplier_estimate", "slpb_accept", "slpb_array_size", "slpb_download",
    "slpb_pattern", "slpb_get_counters", "slpb_get_timers", "slpb_get_comm_stats",
    "slpb_last_device_ms", "slpb_flush_l2", "slpb_stream",
    "slpb_batch_create", "slpb_batch_destroy", "slpb_batch_size",
    "slpb_batch_set_system", "slpb_batch_capture", "slpb_batch_factor",
    "slpb_batch_solve", "slpb_batch_get", "slpb_batch_last_ms",
    "slpb_batch_bytes",
    "slpb_group_create", "slpb_group_destroy", "slpb_group_join",
    "slpb_group_abandon", "slpb_group_pause", "slpb_group_resume",
    "slpb_group_leave", "slpb_group_stats",
]

_dev = None
_host = None


def device_lib() -> C.CDLL:
    global _dev
    if _dev is None:
        if not os.path.exists(LIB_DEVICE):
            raise ImportError(
                f"{LIB_DEVICE} is missing: build it with `make -C sleipnir_b200`"
                " (or `python -c 'import __graft_entry__ as g; g.build()'`)")
        L = C.CDLL(LIB_DEVICE, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        L.slpb_last_error.restype = C.c_char_p
        L.slpb_last_error.argtypes = [vp]
        L.slpb_set_scaling.argtypes = [vp, C.c_double, _dp, _dp]
        L.slpb_analyze.argtypes = [vp, C.c_int, _ip, C.POINTER(SymbolicStats)]
        L.slpb_get_permutation.argtypes = [vp, _ip]
        L.slpb_set_iterate.argtypes = [vp, _dp, _dp, _dp, _dp]
        L.slpb_get_iterate.argtypes = [vp, _dp, _dp, _dp, _dp]
        L.slpb_eval_current.argtypes = [vp, C.c_int, C.POINTER(PointInfo)]
        L.slpb_kkt_stats_current.argtypes = [vp, C.c_double, C.POINTER(KktStats)]
        L.slpb_kkt_stats_trial.argtypes = [vp, C.c_double, C.POINTER(KktStats)]
        L.slpb_factor.argtypes = [vp, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(FactorInfo)]
        L.slpb_factor_pair.argtypes = [vp, _dp, _dp, C.c_int,
                                       C.POINTER(FactorInfo)]
        L.slpb_select_factor.argtypes = [vp, C.c_int]
        L.slpb_prepare_rhs.argtypes = [vp, C.c_double]
        L.slpb_solve.argtypes = [vp, C.c_double, C.c_double, C.POINTER(StepInfo)]
        L.slpb_soc_begin.argtypes = [vp]
        L.slpb_soc_iterate.argtypes = [vp, C.c_double, C.c_double, C.c_double,
                                       C.POINTER(StepInfo)]
        L.slpb_trial.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.POINTER(PointInfo)]
        L.slpb_accept.argtypes = [vp, C.c_double]
        L.slpb_solve_trial.argtypes = [vp, C.c_double, C.c_double, C.c_int,
                                       C.c_int, C.POINTER(StepInfo),
                                       C.POINTER(PointInfo)]
        L.slpb_accept_relinearize.argtypes = [vp, C.c_double,
                                              C.POINTER(C.c_int32),
                                              C.POINTER(KktStats)]
        L.slpb_probe_point.argtypes = [vp, _dp, _dp, C.POINTER(PointInfo)]
        L.slpb_multiplier_estimate.argtypes = [vp, C.c_double,
                                               C.POINTER(FactorInfo)]
        L.slpb_array_size.argtypes = [vp, C.c_int, _lp]
        L.slpb_download.argtypes = [vp, C.c_int, _dp]
        L.slpb_pattern.argtypes = [vp, C.c_int, _ip, _ip, _lp, _ip, _ip]
        L.slpb_get_counters.argtypes = [vp, C.POINTER(Counters)]
        L.slpb_last_device_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
        L.slpb_stream.restype = vp
        L.slpb_stream.argtypes = [vp]
        L.slpb_batch_create.argtypes = [vp, C.c_int32, C.POINTER(vp)]
        L.slpb_batch_destroy.argtypes = [vp]
        L.slpb_batch_destroy.restype = None
        L.slpb_batch_size.argtypes = [vp, _ip, _ip]
        L.slpb_batch_set_system.argtypes = [vp, C.c_int32, _dp, _dp]
        L.slpb_batch_capture.argtypes = [vp, C.c_int32, vp]
        L.slpb_batch_factor.argtypes = [vp, _dp, _dp, C.POINTER(FactorInfo)]
        L.slpb_batch_solve.argtypes = [vp]
        L.slpb_batch_get.argtypes = [vp, C.c_int32, C.c_int, _dp]
        L.slpb_batch_last_ms.argtypes = [vp, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float)]
        L.slpb_batch_bytes.argtypes = [vp, _lp, _lp]
        _dev = L
    return _dev


def comm_unique_id() -> bytes:
    """A fresh ncclUniqueId (rank 0 creates it and sends it to the others)."""
    buf = C.create_string_buffer(128)
    rc = device_lib().slpb_comm_unique_id(buf)
    if rc != 0:
        raise DeviceError(f"slpb_comm_unique_id failed (status {rc})")
    return buf.raw


def multistart(name, N, p0, p1=None, perturb=0.0, max_concurrency=0,
               max_iterations=5000, device=0, n_vars=None):
    """slp::multistart (reference multistart.hpp:44-73) over len(p0) starts of
    a named problem, every start on its own host thread and device stream.
    Returns dict(status, cost, index, x, wall_s, starts=[(status, cost,
    iterations, seconds), …])."""
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    p1 = np.zeros_like(p0) if p1 is None else \
        np.ascontiguousarray(p1, dtype=np.float64)
    count = len(p0)
    if n_vars is None:
        P = Problem(name, N, float(p0[0]), float(p1[0]))
        n_vars = P.n
        P.close()
    best, best_x = np.zeros(3), np.zeros(n_vars)
    per = np.zeros(4 * count)
    wall = host_lib().slpbh_multistart(
        name.encode(), N, count, _d(p0), _d(p1), float(perturb),
        int(max_concurrency), int(max_iterations), int(device), _d(best),
        _d(best_x), n_vars, _d(per))
    if wall < 0:
        raise DeviceError("slpbh_multistart: a device call failed")
    per = per.reshape(count, 4)
    return dict(status=int(best[0]), cost=float(best[1]), index=int(best[2]),
                x=best_x, wall_s=wall,
                starts=[(int(r[0]), float(r[1]), int(r[2]), float(r[3]))
                        for r in per])


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        device_lib()
        if not os.path.exists(LIB_HOST):
            raise ImportError(
                f"{LIB_HOST} is missing: build it with `make -C sleipnir_b200`")
        L = C.CDLL(LIB_HOST)
        vp = C.c_void_p
        L.slpbh_problem_create.restype = vp
        L.slpbh_problem_create.argtypes = [C.c_char_p, C.c_int, C.c_double,
                                           C.c_double]
        L.slpbh_problem_destroy.argtypes = [vp]
        L.slpbh_error.restype = C.c_char_p
        L.slpbh_error.argtypes = [vp]
        L.slpbh_dims.argtypes = [vp, _ip, _ip, _ip]
        L.slpbh_types.argtypes = [vp, _ip, _ip, _ip]
        L.slpbh_initial_guess.argtypes = [vp, _dp]
        L.slpbh_set_guess.argtypes = [vp, _dp]
        L.slpbh_solve.restype = C.c_int
        L.slpbh_solve.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_int, _ip, C.c_int]
        L.slpbh_trace_rows.restype = C.c_int
        L.slpbh_trace_rows.argtypes = [vp]
        L.slpbh_trace_get.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.slpbh_solution.argtypes = [vp, _dp, _dp, _dp, _dp]
        L.slpbh_loop_seconds.restype = C.c_double
        L.slpbh_loop_seconds.argtypes = [vp]
        L.slpbh_multistart.restype = C.c_double
        L.slpbh_multistart.argtypes = [C.c_char_p, C.c_int, C.c_int, _dp, _dp,
                                       C.c_double, C.c_int, C.c_int, C.c_int,
                                       _dp, _dp, C.c_int, _dp]
        L.slpbh_set_diagnostics.argtypes = [vp, C.c_int, C.c_int]
        L.slpbh_solver_kind.restype = C.c_int
        L.slpbh_solver_kind.argtypes = [vp]
        L.slpbh_set_flush_l2.argtypes = [vp, C.c_int]
        L.slpbh_set_factor_arithmetic.argtypes = [vp, C.c_int]
        L.slpbh_flush_seconds.restype = C.c_double
        L.slpbh_flush_seconds.argtypes = [vp]
        L.slpbh_phase_seconds.argtypes = [vp, _dp]
        L.slpbh_set_timeout.argtypes = [vp, C.c_double]
        L.slpbh_set_comm.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        L.slpbh_add_recording_callback.argtypes = [vp, C.c_int, C.c_int]
        L.slpbh_clear_callbacks.argtypes = [vp]
        L.slpbh_callback_log.restype = C.c_int
        L.slpbh_callback_log.argtypes = [vp, _dp, C.c_int, _dp]
        L.slpbh_symbolic_stats.argtypes = [vp, _lp]
        L.slpbh_counters.argtypes = [vp, _lp]
        L.slpbh_timers.argtypes = [vp, _dp]
        L.slpbh_comm_stats.argtypes = [vp, _dp]
        L.slpbh_device_open.restype = vp
        L.slpbh_device_open.argtypes = [vp, C.c_int]
        L.slpbh_device_close.argtypes = [vp]
        _host = L
    return _host


@dataclass
class IterationRecord:
    iteration: int
    type: int
    error: float
    cost: float
    infeasibility: float
    complementarity: float
    mu: float
    delta: float
    gamma: float
    alpha: float
    alpha_max: float
    alpha_z: float
    factorizations: int
    solves: int
    trials: int
    t_end: float = 0.0
    x: np.ndarray | None = None
    s: np.ndarray | None = None
    y: np.ndarray | None = None
    z: np.ndarray | None = None


class DeviceSession:
    """Direct access to the C ABI for a problem that was uploaded already."""

    def __init__(self, raw, n, me, mi):
        self.L = device_lib()
        self.raw = C.c_void_p(raw)
        self.n, self.me, self.mi = n, me, mi

    def _check(self, rc, what):
        if rc != 0:
            raise DeviceError(
                f"{what}: {self.L.slpb_last_error(self.raw).decode()} ({rc})")

    def set_scaling(self, d_f, d_ce, d_ci):
        d_ce = np.ascontiguousarray(d_ce, dtype=np.float64)
        d_ci = np.ascontiguousarray(d_ci, dtype=np.float64)
        self._check(self.L.slpb_set_scaling(self.raw, d_f, _d(d_ce), _d(d_ci)),
                    "slpb_set_scaling")

    def set_factor_arithmetic(self, mode):
        """ARITH_REFERENCE (default) or ARITH_TENSOR (fused Schur updates, FP64
        tensor cores on frontal matrices of order ≥ 16)."""
        self._check(self.L.slpb_set_factor_arithmetic(self.raw, int(mode)),
                    "slpb_set_factor_arithmetic")

    def analyze(self, ordering=ORDER_NESTED_DISSECTION, perm=None):
        st = SymbolicStats()
        p = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        self._check(self.L.slpb_analyze(self.raw, ordering, _i(p), C.byref(st)),
                    "slpb_analyze")
        return st

    def permutation(self):
        p = np.zeros(self.n + self.me, dtype=np.int32)
        self._check(self.L.slpb_get_permutation(self.raw, _i(p)),
                    "slpb_get_permutation")
        return p

    def set_iterate(self, x, s, y, z):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, s, y, z)]
        self._check(self.L.slpb_set_iterate(self.raw, *[_d(v) for v in a]),
                    "slpb_set_iterate")

    def get_iterate(self):
        x, s = np.zeros(self.n), np.zeros(self.mi)
        y, z = np.zeros(self.me), np.zeros(self.mi)
        self._check(self.L.slpb_get_iterate(self.raw, _d(x), _d(s), _d(y), _d(z)),
                    "slpb_get_iterate")
        return x, s, y, z

    def eval_current(self, derivatives=1):
        info = PointInfo()
        self._check(self.L.slpb_eval_current(self.raw, derivatives,
                                             C.byref(info)), "slpb_eval_current")
        return info

    def kkt_stats(self, mu, trial=False):
        st = KktStats()
        fn = self.L.slpb_kkt_stats_trial if trial else self.L.slpb_kkt_stats_current
        self._check(fn(self.raw, mu, C.byref(st)), "slpb_kkt_stats")
        return st

    def factor(self, delta, gamma, reassemble=True):
        info = FactorInfo()
        self._check(self.L.slpb_factor(self.raw, delta, gamma, int(reassemble),
                                       C.byref(info)), "slpb_factor")
        return info

    def factor_pair(self, delta, gamma, reassemble=True):
        """Two regularisations factored side by side; returns (info0, info1)."""
        info = (FactorInfo * 2)()
        d = np.ascontiguousarray(delta, dtype=np.float64)
        g = np.ascontiguousarray(gamma, dtype=np.float64)
        self._check(self.L.slpb_factor_pair(self.raw, _d(d), _d(g),
                                            int(reassemble), info),
                    "slpb_factor_pair")
        return info[0], info[1]

    def prepare_rhs(self, mu):
        """Lets the next factorisation carry the forward substitution."""
        self._check(self.L.slpb_prepare_rhs(self.raw, mu), "slpb_prepare_rhs")

    def select_factor(self, which):
        self._check(self.L.slpb_select_factor(self.raw, which),
                    "slpb_select_factor")

    def solve(self, mu, tau):
        info = StepInfo()
        self._check(self.L.slpb_solve(self.raw, mu, tau, C.byref(info)),
                    "slpb_solve")
        return info

    def soc_begin(self):
        self._check(self.L.slpb_soc_begin(self.raw), "slpb_soc_begin")

    def soc_iterate(self, mu, tau, alpha_soc):
        info = StepInfo()
        self._check(self.L.slpb_soc_iterate(self.raw, mu, tau, alpha_soc,
                                            C.byref(info)), "slpb_soc_iterate")
        return info

    def trial(self, alpha, alpha_z, which_step=0, slack_from_ci=0):
        info = PointInfo()
        self._check(self.L.slpb_trial(self.raw, alpha, alpha_z, which_step,
                                      slack_from_ci, C.byref(info)), "slpb_trial")
        return info

    def probe_point(self, x, s):
        info = PointInfo()
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.ascontiguousarray(s if self.mi else np.zeros(1), dtype=np.float64)
        self._check(self.L.slpb_probe_point(self.raw, _d(x), _d(s),
                                            C.byref(info)), "slpb_probe_point")
        return info

    def multiplier_estimate(self, mu):
        info = FactorInfo()
        self._check(self.L.slpb_multiplier_estimate(self.raw, mu, C.byref(info)),
                    "slpb_multiplier_estimate")
        return info

    def accept(self, mu):
        self._check(self.L.slpb_accept(self.raw, mu), "slpb_accept")

    def solve_trial(self, mu, tau, dual_uses_primal_alpha=0, slack_from_ci=0):
        """slpb_solve + slpb_trial at the full step in one host round trip."""
        step, trial = StepInfo(), PointInfo()
        self._check(self.L.slpb_solve_trial(
            self.raw, mu, tau, dual_uses_primal_alpha, slack_from_ci,
            C.byref(step), C.byref(trial)), "slpb_solve_trial")
        return step, trial

    def accept_relinearize(self, mu):
        """slpb_accept + slpb_eval_current(2) + slpb_kkt_stats_current in one
        host round trip; returns (finite bits of the derivatives, KktStats)."""
        finite, st = C.c_int32(), KktStats()
        self._check(self.L.slpb_accept_relinearize(
            self.raw, mu, C.byref(finite), C.byref(st)),
            "slpb_accept_relinearize")
        return finite.value, st

    def download(self, which):
        cnt = C.c_int64()
        self._check(self.L.slpb_array_size(self.raw, which, C.byref(cnt)),
                    "slpb_array_size")
        out = np.zeros(max(cnt.value, 1))
        self._check(self.L.slpb_download(self.raw, which, _d(out)),
                    "slpb_download")
        return out[:cnt.value]

    def pattern(self, which):
        r, c, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self.L.slpb_pattern(self.raw, which, C.byref(r), C.byref(c),
                                        C.byref(nnz), None, None), "slpb_pattern")
        colptr = np.zeros(c.value + 1, dtype=np.int32)
        rowidx = np.zeros(max(nnz.value, 1), dtype=np.int32)
        self._check(self.L.slpb_pattern(self.raw, which, C.byref(r), C.byref(c),
                                        C.byref(nnz), _i(colptr), _i(rowidx)),
                    "slpb_pattern")
        return r.value, c.value, colptr, rowidx[:nnz.value]

    def counters(self):
        c = Counters()
        self.L.slpb_get_counters(self.raw, C.byref(c))
        return c

    def last_device_ms(self, which):
        ms = C.c_float()
        self._check(self.L.slpb_last_device_ms(self.raw, which, C.byref(ms)),
                    "slpb_last_device_ms")
        return ms.value


class Batch:
    """Many KKT systems over the symbolic structure of one DeviceSession,
    factored and solved side by side (slpb_batch_*, lane = instance)."""

    SOLUTION, D = 0, 1

    def __init__(self, session: DeviceSession, batch: int):
        self.L = device_lib()
        self.session = session
        self.batch = batch
        self.dim = session.n + session.me
        h = C.c_void_p()
        session._check(self.L.slpb_batch_create(session.raw, batch, C.byref(h)),
                       "slpb_batch_create")
        self.h = h

    def close(self):
        if self.h:
            self.L.slpb_batch_destroy(self.h)
            self.h = None

    def set_system(self, instance, kkt_val=None, rhs=None):
        k = None if kkt_val is None else np.ascontiguousarray(kkt_val, dtype=np.float64)
        r = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
        self.session._check(self.L.slpb_batch_set_system(self.h, instance, _d(k), _d(r)),
                            "slpb_batch_set_system")

    def capture(self, instance, source: DeviceSession = None):
        src = (source or self.session).raw
        self.session._check(self.L.slpb_batch_capture(self.h, instance, src),
                            "slpb_batch_capture")

    def factor(self, delta, gamma):
        d = np.ascontiguousarray(np.broadcast_to(delta, (self.batch,)), dtype=np.float64)
        g = np.ascontiguousarray(np.broadcast_to(gamma, (self.batch,)), dtype=np.float64)
        info = (FactorInfo * self.batch)()
        self.session._check(self.L.slpb_batch_factor(self.h, _d(d), _d(g), info),
                            "slpb_batch_factor")
        return list(info)

    def solve(self):
        self.session._check(self.L.slpb_batch_solve(self.h), "slpb_batch_solve")

    def get(self, instance, what=0):
        out = np.zeros(self.dim)
        self.session._check(self.L.slpb_batch_get(self.h, instance, what, _d(out)),
                            "slpb_batch_get")
        return out

    def last_ms(self):
        f, s = C.c_float(), C.c_float()
        self.L.slpb_batch_last_ms(self.h, C.byref(f), C.byref(s))
        return f.value, s.value

    def stored_entries(self):
        p, u = C.c_int64(), C.c_int64()
        self.L.slpb_batch_bytes(self.h, C.byref(p), C.byref(u))
        return p.value, u.value


class Problem:
    """A named benchmark / test problem built by the C++ host side."""

    def __init__(self, name: str, N: int = 0, p0: float = 0.0, p1: float = 0.0):
        self.H = host_lib()
        self.h = self.H.slpbh_problem_create(name.encode(), N, p0, p1)
        if not self.h:
            raise ValueError(f"unknown problem {name!r}")
        n, me, mi = C.c_int32(), C.c_int32(), C.c_int32()
        self.H.slpbh_dims(self.h, C.byref(n), C.byref(me), C.byref(mi))
        self.n, self.me, self.mi = n.value, me.value, mi.value
        self._keep = False

    def close(self):
        if self.h:
            self.H.slpbh_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def types(self):
        f, ce, ci = C.c_int32(), C.c_int32(), C.c_int32()
        self.H.slpbh_types(self.h, C.byref(f), C.byref(ce), C.byref(ci))
        return f.value, ce.value, ci.value

    def initial_guess(self):
        x = np.zeros(self.n)
        self.H.slpbh_initial_guess(self.h, _d(x))
        return x

    def set_guess(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.H.slpbh_set_guess(self.h, _d(x))

    def solve(self, tolerance=1e-8, max_iterations=5000, feasible_ipm=False,
              device=0, ordering=ORDER_NESTED_DISSECTION, perm=None,
              keep_iterates=False) -> int:
        p = None
        if perm is not None:
            p = np.ascontiguousarray(perm, dtype=np.int32)
            ordering = ORDER_CUSTOM
        self._keep = keep_iterates
        st = self.H.slpbh_solve(self.h, tolerance, max_iterations,
                                int(feasible_ipm), device, ordering, _i(p),
                                int(keep_iterates))
        if st == -100:
            raise DeviceError(self.H.slpbh_error(self.h).decode())
        return st

    def trace(self):
        rows = []
        for r in range(self.H.slpbh_trace_rows(self.h)):
            sc = np.zeros(16)
            x = s = y = z = None
            if self._keep:
                # restoration rows (type 1) carry the enlarged problem's vectors
                extra = 2 * self.me + 2 * self.mi
                x, s = np.zeros(self.n + extra), np.zeros(self.mi + extra + 1)
                y, z = np.zeros(max(self.me, 1)), np.zeros(self.mi + extra + 1)
            self.H.slpbh_trace_get(self.h, r, _d(sc), _d(x), _d(s), _d(y), _d(z))
            if self._keep:
                if int(sc[1]) == 0:
                    x, s, z = x[:self.n], s[:self.mi], z[:self.mi]
                else:
                    s, z = s[:self.mi + extra], z[:self.mi + extra]
                y = y[:self.me]
            rows.append(IterationRecord(int(sc[0]), int(sc[1]), *sc[2:12],
                                        int(sc[12]), int(sc[13]), int(sc[14]),
                                        float(sc[15]), x, s, y, z))
        return rows

    def solution(self):
        x, s = np.zeros(self.n), np.zeros(max(self.mi, 1))
        y, z = np.zeros(max(self.me, 1)), np.zeros(max(self.mi, 1))
        self.H.slpbh_solution(self.h, _d(x), _d(s), _d(y), _d(z))
        return x, s[:self.mi], y[:self.me], z[:self.mi]

    def loop_seconds(self):
        return self.H.slpbh_loop_seconds(self.h)

    def set_diagnostics(self, diagnostics=True, spy=False):
        """Options::diagnostics and the spy flag of solve() for the next solves."""
        self.H.slpbh_set_diagnostics(self.h, int(diagnostics), int(spy))

    def solver_kind(self):
        """Branch of the last solve(): "IPM", "SQP" or "NEWTON"
        (reference problem.hpp:512, :403, :335)."""
        return ("IPM", "SQP", "NEWTON")[self.H.slpbh_solver_kind(self.h)]

    def set_comm(self, rank, world, unique_id: bytes):
        """Multi-GPU sharded solve: this process is `rank` of `world`; every
        rank solves the same problem in lockstep. unique_id: the 128 bytes of
        comm_unique_id() from rank 0."""
        assert world == 1 or len(unique_id) == 128
        self.H.slpbh_set_comm(self.h, int(rank), int(world), unique_id)

    def set_timeout(self, seconds):
        """Options::timeout of the following solves (negative: none)."""
        self.H.slpbh_set_timeout(self.h, float(seconds))

    def add_callback(self, stop_at=-1, persistent=False):
        """Registers an iteration callback (Problem::add_callback /
        add_persistent_callback) that records the IterationInfo it receives and
        returns true once iteration `stop_at` is reached."""
        self.H.slpbh_add_recording_callback(self.h, int(stop_at), int(persistent))

    def clear_callbacks(self):
        self.H.slpbh_clear_callbacks(self.h)

    def callback_log(self):
        """(records, last_x): one row per callback invocation — iteration, n,
        ‖x‖∞, |s|, |y|, |z|, nnz(H)+nnz(A_e)+nnz(A_i), |g|."""
        out = np.zeros(8 * 4096)
        last_x = np.zeros(max(self.n, 1))
        k = self.H.slpbh_callback_log(self.h, _d(out), 4096, _d(last_x))
        return out[:8 * min(k, 4096)].reshape(-1, 8), last_x[:self.n]

    def set_factor_arithmetic(self, mode):
        """DeviceOptions::factor_arithmetic of the following solve() calls
        (ARITH_REFERENCE / ARITH_TENSOR; −1: the library default)."""
        self.H.slpbh_set_factor_arithmetic(self.h, int(mode))

    def set_flush_l2(self, on=True):
        """Benchmark hygiene: evict the device L2 before every iteration of the
        following solves; the flushes are excluded from the iteration times."""
        self.H.slpbh_set_flush_l2(self.h, int(on))

    def flush_seconds(self):
        return self.H.slpbh_flush_seconds(self.h)

    def phase_seconds(self):
        """Host wall time of the phases of the last solve() call."""
        out = np.zeros(9)
        self.H.slpbh_phase_seconds(self.h, _d(out))
        keys = ("build_graphs", "flatten", "device_create", "upload_compile",
                "scaling", "analyze", "newton_loop", "write_back", "teardown")
        return dict(zip(keys, (float(v) for v in out)))

    def symbolic_stats(self):
        out = np.zeros(8, dtype=np.int64)
        self.H.slpbh_symbolic_stats(self.h, out.ctypes.data_as(_lp))
        keys = ("dim", "nnz_kkt", "nnz_l", "nnz_l_stored", "n_supernodes",
                "n_levels", "max_front", "etree_height")
        return dict(zip(keys, (int(v) for v in out)))

    def counters(self):
        out = np.zeros(12, dtype=np.int64)
        self.H.slpbh_counters(self.h, out.ctypes.data_as(_lp))
        keys = [k for k, _ in Counters._fields_]
        return dict(zip(keys, (int(v) for v in out)))

    def timers(self):
        """Device time of the kernel groups of the last solve: total_ms over
        the `count` SAMPLED runs (one in SLPB_TIMER_EVERY, default 8),
        `launches` = all runs; mean_ms = total_ms / count."""
        out = np.zeros(15)
        self.H.slpbh_timers(self.h, _d(out))
        names = ("eval_full", "eval_values", "assemble", "factor", "solve")
        return {k: {"total_ms": float(out[i]), "count": int(out[5 + i]),
                    "launches": int(out[10 + i]),
                    "mean_ms": float(out[i]) / max(int(out[5 + i]), 1)}
                for i, k in enumerate(names)}

    def comm_stats(self):
        """Collectives of the last sharded solve: per kind (derivative rows,
        subtree roots, solution pieces) calls, bytes contributed, device ms."""
        out = np.zeros(9)
        self.H.slpbh_comm_stats(self.h, _d(out))
        names = ("derivative_rows", "subtree_roots", "solution")
        return {k: {"count": int(out[i]), "bytes": int(out[3 + i]),
                    "total_ms": float(out[6 + i])} for i, k in enumerate(names)}

    def open_device(self, device=0) -> DeviceSession:
        raw = self.H.slpbh_device_open(self.h, device)
        if not raw:
            raise DeviceError(self.H.slpbh_error(self.h).decode())
        return DeviceSession(raw, self.n, self.me, self.mi)

    def close_device(self):
        self.H.slpbh_device_close(self.h)
