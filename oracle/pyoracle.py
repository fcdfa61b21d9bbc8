"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for the CPU oracle (oracle/liboracle.so, and the tier-A build
oracle/_ref/liboracle_ref.so when present). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package (sleipnir_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_RESTATED = os.path.join(_HERE, "liboracle.so")
LIB_REFERENCE = os.path.join(_HERE, "_ref", "liboracle_ref.so")

EXIT_STATUS = {
    0: "SUCCESS", 1: "CALLBACK_REQUESTED_STOP", -1: "TOO_FEW_DOFS",
    -2: "LOCALLY_INFEASIBLE", -3: "GLOBALLY_INFEASIBLE",
    -4: "FACTORIZATION_FAILED", -5: "LINE_SEARCH_FAILED",
    -6: "FEASIBILITY_RESTORATION_FAILED", -7: "NONFINITE_INITIAL_GUESS",
    -8: "DIVERGING_ITERATES", -9: "MAX_ITERATIONS_EXCEEDED", -10: "TIMEOUT",
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _bind(lib):
    lib.orc_backend_name.restype = C.c_char_p
    lib.orc_problem_create.restype = C.c_void_p
    lib.orc_problem_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double]
    lib.orc_problem_destroy.argtypes = [C.c_void_p]
    lib.orc_problem_dims.argtypes = [C.c_void_p, _ip, _ip, _ip]
    lib.orc_problem_types.argtypes = [C.c_void_p, _ip, _ip, _ip]
    lib.orc_problem_initial_guess.argtypes = [C.c_void_p, _dp]
    lib.orc_problem_set_guess.argtypes = [C.c_void_p, _dp]
    lib.orc_problem_solve.restype = C.c_int
    lib.orc_problem_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int,
                                      C.c_int, _ip, C.c_int, C.c_int]
    lib.orc_solve_seconds.restype = C.c_double
    lib.orc_solve_seconds.argtypes = [C.c_void_p]
    lib.orc_trace_rows.restype = C.c_int
    lib.orc_trace_rows.argtypes = [C.c_void_p]
    lib.orc_trace_get.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp]
    lib.orc_solution.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.orc_eval_setup.restype = C.c_int
    lib.orc_eval_setup.argtypes = [C.c_void_p]
    lib.orc_eval_scaling.argtypes = [C.c_void_p, _dp, _dp, _dp]
    lib.orc_eval_f.restype = C.c_double
    lib.orc_eval_f.argtypes = [C.c_void_p, _dp]
    lib.orc_eval_vector.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.orc_eval_matrix.restype = C.c_int
    lib.orc_eval_matrix.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    lib.orc_last_matrix.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _dp]
    lib.orc_time_graph_walk.restype = C.c_double
    lib.orc_time_graph_walk.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
    lib.orc_amd.argtypes = [C.c_int, _ip, _ip, _ip]
    lib.orc_ldlt.restype = C.c_int
    lib.orc_ldlt.argtypes = [C.c_int, _ip, _ip, _dp, _ip, _dp, _dp, _dp, _ip]
    return lib


_libs: dict[str, C.CDLL] = {}


def load(which: str = "restated") -> C.CDLL:
    """which: 'restated' (always available once built) or 'reference'."""
    if which not in _libs:
        path = LIB_RESTATED if which == "restated" else LIB_REFERENCE
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing; run `make -C oracle`"
                + (" ref" if which == "reference" else ""))
        _libs[which] = _bind(C.CDLL(path))
    return _libs[which]


def have_reference() -> bool:
    return os.path.exists(LIB_REFERENCE)


@dataclass
class CscMatrix:
    rows: int
    cols: int
    colptr: np.ndarray
    rowidx: np.ndarray
    val: np.ndarray

    def todense(self) -> np.ndarray:
        m = np.zeros((self.rows, self.cols))
        for c in range(self.cols):
            for k in range(self.colptr[c], self.colptr[c + 1]):
                m[self.rowidx[k], c] += self.val[k]
        return m


@dataclass
class TraceRow:
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


class OracleProblem:
    """One problem instance living inside the oracle library."""

    def __init__(self, name: str, N: int = 0, p0: float = 0.0, p1: float = 0.0,
                 backend: str = "restated"):
        self.lib = load(backend)
        self.h = self.lib.orc_problem_create(name.encode(), N, p0, p1)
        if not self.h:
            raise ValueError(f"oracle does not know problem {name!r}")
        n, me, mi = C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_problem_dims(self.h, C.byref(n), C.byref(me), C.byref(mi))
        self.n, self.me, self.mi = n.value, me.value, mi.value
        self._setup = False

    def close(self):
        if self.h:
            self.lib.orc_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def types(self):
        f, ce, ci = C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_problem_types(self.h, C.byref(f), C.byref(ce), C.byref(ci))
        return f.value, ce.value, ci.value

    def initial_guess(self) -> np.ndarray:
        x = np.zeros(self.n)
        self.lib.orc_problem_initial_guess(self.h, _d(x))
        return x

    def set_guess(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.lib.orc_problem_set_guess(self.h, _d(x))

    def solve(self, tolerance=1e-8, max_iterations=5000, feasible_ipm=False,
              ordering="amd", perm=None, force_sparse=-1, keep_iterates=True):
        code = {"amd": 0, "natural": 1, "custom": 2}[ordering]
        p = None
        if perm is not None:
            p = np.ascontiguousarray(perm, dtype=np.int32)
            assert p.size == self.n + self.me
            code = 2
        st = self.lib.orc_problem_solve(self.h, tolerance, max_iterations,
                                        int(feasible_ipm), code, _i(p),
                                        force_sparse, int(keep_iterates))
        self._keep = keep_iterates
        return st

    def solve_seconds(self) -> float:
        return self.lib.orc_solve_seconds(self.h)

    def trace(self) -> list[TraceRow]:
        rows = []
        for r in range(self.lib.orc_trace_rows(self.h)):
            sc = np.zeros(16)
            x = s = y = z = None
            if self._keep:
                # restoration rows carry the enlarged problem's vectors
                big = self.n + 2 * self.me + 2 * self.mi
                x = np.zeros(big)
                s = np.zeros(self.mi + 2 * self.me + 2 * self.mi)
                y = np.zeros(self.me)
                z = np.zeros(self.mi + 2 * self.me + 2 * self.mi)
            self.lib.orc_trace_get(self.h, r, _d(sc), _d(x), _d(s), _d(y), _d(z))
            if self._keep and int(sc[1]) == 0:
                x, s, z = x[:self.n], s[:self.mi], z[:self.mi]
            rows.append(TraceRow(int(sc[0]), int(sc[1]), *sc[2:12],
                                 int(sc[12]), int(sc[13]), int(sc[14]),
                                 float(sc[15]), x, s, y, z))
        return rows

    def solution(self):
        x, s = np.zeros(self.n), np.zeros(self.mi)
        y, z = np.zeros(self.me), np.zeros(self.mi)
        self.lib.orc_solution(self.h, _d(x), _d(s), _d(y), _d(z))
        return x, s, y, z

    # ---- callback-level evaluation -------------------------------------
    def eval_setup(self) -> int:
        r = self.lib.orc_eval_setup(self.h)
        self._setup = True
        return r

    def scaling(self):
        assert self._setup
        df = C.c_double()
        dce, dci = np.zeros(max(self.me, 1)), np.zeros(max(self.mi, 1))
        self.lib.orc_eval_scaling(self.h, C.byref(df), _d(dce), _d(dci))
        return df.value, dce[:self.me], dci[:self.mi]

    def f(self, x) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.orc_eval_f(self.h, _d(x))

    def _vec(self, which, x, m):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(max(m, 1))
        self.lib.orc_eval_vector(self.h, which, _d(x), _d(out))
        return out[:m]

    def g(self, x):
        return self._vec(0, x, self.n)

    def c_e(self, x):
        return self._vec(1, x, self.me)

    def c_i(self, x):
        return self._vec(2, x, self.mi)

    def _mat(self, which, x, y=None, z=None) -> CscMatrix:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(max(self.me, 1)) if y is None else np.ascontiguousarray(y, dtype=np.float64)
        z = np.zeros(max(self.mi, 1)) if z is None else np.ascontiguousarray(z, dtype=np.float64)
        nnz = self.lib.orc_eval_matrix(self.h, which, _d(x), _d(y), _d(z))
        cols = self.n
        colptr = np.zeros(cols + 1, dtype=np.int32)
        rowidx = np.zeros(max(nnz, 1), dtype=np.int32)
        val = np.zeros(max(nnz, 1))
        r, c = C.c_int(), C.c_int()
        self.lib.orc_last_matrix(self.h, C.byref(r), C.byref(c), _i(colptr),
                                 _i(rowidx), _d(val))
        return CscMatrix(r.value, c.value, colptr, rowidx[:nnz], val[:nnz])

    def A_e(self, x):
        return self._mat(0, x)

    def A_i(self, x):
        return self._mat(1, x)

    def H(self, x, y, z):
        return self._mat(2, x, y, z)

    def H_c(self, x, y, z):
        return self._mat(3, x, y, z)

    def time_graph_walk(self, x, y, z, reps=3) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        return self.lib.orc_time_graph_walk(self.h, _d(x), _d(y), _d(z), reps)


def amd(n, colptr, rowidx, backend="restated") -> np.ndarray:
    colptr = np.ascontiguousarray(colptr, dtype=np.int32)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    perm = np.zeros(n, dtype=np.int32)
    load(backend).orc_amd(n, _i(colptr), _i(rowidx), _i(perm))
    return perm


def ldlt(n, colptr, rowidx, val, rhs=None, perm=None, backend="restated"):
    """Returns (nnz_L or -1, D, x, etree_height)."""
    colptr = np.ascontiguousarray(colptr, dtype=np.int32)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    D, x = np.zeros(n), np.zeros(n)
    h = C.c_int()
    p = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    r = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
    nnz = load(backend).orc_ldlt(n, _i(colptr), _i(rowidx), _d(val), _i(p),
                                 _d(r), _d(D), _d(x), C.byref(h))
    return nnz, D, x, h.value
