"""Pins the CPU oracle (oracle/) — the checker everything else is compared
with — against (1) the reference's own known-answer tests, (2) golden vectors
produced by the reference's OWN autodiff core (tests/golden/make_golden.py) and
(3), where oracle/_ref exists, that core directly."""
import glob
import os

import numpy as np
import pytest

from oracle.pyoracle import (EXIT_STATUS, OracleProblem, amd, have_reference,
                             ldlt)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# (problem, expected status, expected x, abs tol) — citations are the
# reference's tests under test/src/optimization.
KNOWN = [
    ("lp_maximize", "SUCCESS", (375, 250), 1e-6),            # linear_problem_test.cpp:14-40
    ("quartic", "SUCCESS", (1,), 1e-6),                      # nonlinear_problem_test.cpp:19-37
    ("wachter_biegler", "SUCCESS", (1, 0, 0.5), 1e-6),       # :167-201
    ("qp_inequality_2d", "SUCCESS", (3 + 1 / 3, 1 + 2 / 3), 1e-6),  # quadratic_problem_test.cpp:164-186
    ("conflicting_bounds", "GLOBALLY_INFEASIBLE", None, 0),  # nonlinear_problem_test.cpp:145-165
    ("locally_infeasible_ineq", "LOCALLY_INFEASIBLE", None, 0),    # exit_status_test.cpp:97-117
    ("nonfinite_ineq", "NONFINITE_INITIAL_GUESS", None, 0),        # :160-166
    ("nonfinite_ineq_jacobian", "NONFINITE_INITIAL_GUESS", None, 0),  # :169-175
    # the Newton (problem.hpp:335) and SQP (:403) branches
    ("unconstrained_1d", "SUCCESS", (3,), 1e-6),             # quadratic_problem_test.cpp:15-34
    ("unconstrained_2d", "SUCCESS", (0, 0), 1e-6),           # :36-58
    ("eq_maximize_xy", "SUCCESS", (18, 6), 1e-5),            # :80-141
    ("eq_pin_2d", "SUCCESS", (3, 3), 1e-5),                  # :143-161
    ("min_distance_line", "SUCCESS", (2.5, 2.5), 1e-2),      # nonlinear_problem_test.cpp:120-143
    ("min_x_squared", "SUCCESS", (0,), 1e-6),                # exit_status_test.cpp:17-27
    ("too_few_dofs", "TOO_FEW_DOFS", None, 0),               # :52-72
    ("locally_infeasible_eq", "LOCALLY_INFEASIBLE", None, 0),  # :78-95
    ("nonfinite_cost", "NONFINITE_INITIAL_GUESS", None, 0),  # :124-130
    ("nonfinite_gradient", "NONFINITE_INITIAL_GUESS", None, 0),  # :133-139
    ("nonfinite_eq", "NONFINITE_INITIAL_GUESS", None, 0),    # :142-148
    ("nonfinite_eq_jacobian", "NONFINITE_INITIAL_GUESS", None, 0),  # :151-157
    ("diverging", "DIVERGING_ITERATES", None, 0),            # :178-194
    ("spy_test", "SUCCESS", (1, 2), 1e-8),                   # problem_spy_test.cpp:64-86
    ("empty", "SUCCESS", (), 0),                             # trivial_problem_test.cpp:14-24
    ("no_cost_unconstrained", "SUCCESS", (0,) * 6, 0),       # :26-46
]


def test_arm_on_elevator_succeeds():
    # arm_on_elevator_problem_test.cpp:27-122 REQUIREs SUCCESS at N = 800
    P = OracleProblem("arm_on_elevator", 800)
    assert EXIT_STATUS[P.solve()] == "SUCCESS"
    x, *_ = P.solution()
    N = 800
    elevator = x[:2 * (N + 1)].reshape(2, N + 1)
    arm = x[2 * (N + 1) + N:2 * (N + 1) + N + 2 * (N + 1)].reshape(2, N + 1)
    assert abs(elevator[0, 0] - 1.0) < 1e-8 and abs(elevator[0, N] - 1.25) < 1e-8
    assert abs(arm[0, 0]) < 1e-8 and abs(arm[0, N] - np.pi) < 1e-8
    assert np.all(elevator[0] + np.sin(arm[0]) <= 1.8 + 1e-6)
    P.close()


def test_differential_drive_replay():
    # differential_drive_problem_test.cpp:28-138: SUCCESS, and the states are
    # the RK4 replay of the inputs to 1e-8
    N = 100
    dt = 5.0 / N
    P = OracleProblem("differential_drive", N)
    assert EXIT_STATUS[P.solve()] == "SUCCESS"
    x, *_ = P.solution()
    X = x[:5 * (N + 1)].reshape(5, N + 1)
    U = x[5 * (N + 1):].reshape(2, N)
    assert np.all(np.abs(U) <= 12.0 + 1e-9)
    Kv_l, Ka_l, Kv_a, Ka_a, trackwidth = 3.02, 0.642, 1.382, 0.08495, 0.699
    A1 = -(Kv_l / Ka_l + Kv_a / Ka_a) / 2.0
    A2 = -(Kv_l / Ka_l - Kv_a / Ka_a) / 2.0
    B1 = 0.5 / Ka_l + 0.5 / Ka_a
    B2 = 0.5 / Ka_l - 0.5 / Ka_a
    A, B = np.array([[A1, A2], [A2, A1]]), np.array([[B1, B2], [B2, B1]])

    def f(s, u):
        v = (s[3] + s[4]) / 2.0
        return np.concatenate([[v * np.cos(s[2]), v * np.sin(s[2]),
                                (s[4] - s[3]) / trackwidth], A @ s[3:] + B @ u])
    state = np.zeros(5)
    for k in range(N):
        np.testing.assert_allclose(X[:, k], state, atol=1e-8)
        u = U[:, k]
        k1 = f(state, u)
        k2 = f(state + dt * 0.5 * k1, u)
        k3 = f(state + dt * 0.5 * k2, u)
        k4 = f(state + dt * k3, u)
        state = state + dt / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    np.testing.assert_allclose(X[:, N], (1.0, 1.0, 0.0, 0.0, 0.0), atol=1e-8)
    P.close()


def test_double_integrator_profile():
    # double_integrator_problem_test.cpp:27-127: accelerate, coast, brake
    N = 700
    dt = 3.5 / N
    P = OracleProblem("double_integrator", N)
    assert EXIT_STATUS[P.solve()] == "SUCCESS"
    x, *_ = P.solution()
    X, U = x[:2 * (N + 1)].reshape(2, N + 1), x[2 * (N + 1):]
    assert abs(X[0, 0]) < 1e-8 and abs(X[1, 0]) < 1e-8
    state = np.zeros(2)
    for k in range(N):
        assert abs(X[0, k] - state[0]) < 1e-2 and abs(X[1, k] - state[1]) < 1e-2
        t = k * dt
        u = 1.0 if t < 1 else 0.0 if t < 2.05 else -1.0 if t < 3.275 else 1.0
        if not (0 < k < N - 1 and abs(U[k - 1] - U[k + 1]) >= 1 - 1e-2):
            assert abs(U[k] - u) < 1e-4
        state = np.array([state[0] + dt * state[1] + 0.5 * dt * dt * u,
                          state[1] + dt * u])
    assert abs(X[0, N] - 2.0) < 1e-8 and abs(X[1, N]) < 1e-8
    P.close()



@pytest.mark.parametrize("name,status,expect,tol", KNOWN)
def test_known_answers(name, status, expect, tol):
    P = OracleProblem(name)
    st = P.solve()
    assert EXIT_STATUS[st] == status
    if expect is not None:
        x, *_ = P.solution()
        np.testing.assert_allclose(x, expect, atol=tol)
    P.close()


def test_multistart_mishra_bird():
    # multistart_test.cpp:17-55: two starts, the one with the lower cost wins
    runs = []
    for guess in ((-3.0, -8.0), (-3.0, -1.5)):
        P = OracleProblem("mishra_bird", 0, *guess)
        st = P.solve()
        x, *_ = P.solution()
        J = (np.sin(x[1]) * np.exp((1 - np.cos(x[0])) ** 2) +
             np.cos(x[0]) * np.exp((1 - np.sin(x[1])) ** 2) + (x[0] - x[1]) ** 2)
        runs.append((EXIT_STATUS[st] != "SUCCESS", J, x))
        P.close()
    _, _, best = min(runs, key=lambda r: (r[0], r[1]))
    np.testing.assert_allclose(best, (-3.13024680, -1.58214218), atol=1e-8)


def test_rosenbrock_disk_grid():
    # nonlinear_problem_test.cpp:84-118 sweeps a 30x30 grid of starts; a coarse
    # sub-grid keeps the CPU tier fast.
    for x0 in np.arange(-1.5, 1.5, 0.5):
        for y0 in np.arange(-1.5, 1.5, 0.5):
            P = OracleProblem("rosenbrock_disk", 0, x0, y0)
            assert EXIT_STATUS[P.solve()] == "SUCCESS"
            x, *_ = P.solution()
            np.testing.assert_allclose(x, (1, 1), atol=1e-3)
            P.close()


def test_rosenbrock_cubic_line_grid():
    # nonlinear_problem_test.cpp:39-82: local minimum (0,0), global (1,1)
    for x0 in np.arange(-1.5, 1.5, 0.6):
        for y0 in np.arange(-0.5, 2.5, 0.6):
            P = OracleProblem("rosenbrock_cubic_line", 0, x0, y0)
            assert EXIT_STATUS[P.solve()] == "SUCCESS"
            x, *_ = P.solution()
            assert (abs(x[0]) < 1e-2 or abs(x[0] - 1) < 1e-2)
            assert (abs(x[1]) < 1e-2 or abs(x[1] - 1) < 1e-2)
            P.close()


def test_cart_pole_problem_test():
    """cart_pole_problem_test.cpp:87-124, N reduced to 60: classification,
    SUCCESS, boundary conditions, bounds and RK4 dynamics residual ≤ 1e-8."""
    N, T = 60, 5.0
    P = OracleProblem("cart_pole", N)
    assert P.types() == (3, 4, 2)  # QUADRATIC, NONLINEAR, LINEAR
    assert EXIT_STATUS[P.solve(keep_iterates=False)] == "SUCCESS"
    x, *_ = P.solution()
    X = x[:4 * (N + 1)].reshape(4, N + 1)
    U = x[4 * (N + 1):]
    np.testing.assert_allclose(X[:, 0], 0, atol=1e-8)
    np.testing.assert_allclose(X[:, N], (1, np.pi, 0, 0), atol=1e-8)
    assert (X[0] >= -1e-9).all() and (X[0] <= 2 + 1e-9).all()
    assert (np.abs(U) <= 20 + 1e-9).all()

    def f(s, u):
        m_c, m_p, l, g = 5.0, 0.5, 0.5, 9.806
        th, xd, thd = s[1], s[2], s[3]
        M = np.array([[m_c + m_p, m_p * l * np.cos(th)],
                      [m_p * l * np.cos(th), m_p * l * l]])
        rhs = np.array([m_p * l * thd * thd * np.sin(th) + u,
                        -m_p * g * l * np.sin(th)])
        return np.concatenate([[xd, thd], np.linalg.solve(M, rhs)])

    h = T / N
    for k in range(N):
        s, u = X[:, k], U[k]
        k1 = f(s, u); k2 = f(s + h / 2 * k1, u)
        k3 = f(s + h / 2 * k2, u); k4 = f(s + h * k3, u)
        np.testing.assert_allclose(X[:, k + 1], s + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4),
                                   atol=1e-8)
    P.close()


def test_flywheel_problem_test():
    """flywheel_problem_test.cpp:70-122 shape (N=50): converges; bang-bang
    input at the start; final state approaches the reference r = 10."""
    P = OracleProblem("flywheel", 50)
    assert EXIT_STATUS[P.solve()] == "SUCCESS"
    x, *_ = P.solution()
    X, U = x[:51], x[51:]
    assert abs(X[0]) < 1e-8
    assert abs(U[0] - 12) < 1e-4
    assert abs(X[-1] - 10) < 2e-2
    P.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "eval_*.npz"))))
def test_golden_eval(path):
    """The restated expression core reproduces the reference core's values,
    Jacobians, gradient and Hessian bit for bit (same libm on both sides)."""
    g = np.load(path)
    name, N = os.path.basename(path)[5:-4].rsplit("_", 1)
    p0, p1 = {"rosenbrock_cubic_line": (0.3, 0.7), "rosenbrock_disk": (-0.5, 1.2)}.get(name, (0, 0))
    P = OracleProblem(name, int(N), p0, p1)
    P.eval_setup()
    d_f, d_ce, d_ci = P.scaling()
    assert d_f == g["d_f"]
    np.testing.assert_array_equal(d_ce, g["d_ce"])
    np.testing.assert_array_equal(d_ci, g["d_ci"])
    x, y, z = g["x"], g["y"], g["z"]
    assert P.f(x) == g["f"]
    np.testing.assert_array_equal(P.c_e(x), g["c_e"])
    np.testing.assert_array_equal(P.c_i(x), g["c_i"])
    np.testing.assert_array_equal(P.g(x), g["g"])
    for nm, M in (("A_e", P.A_e(x)), ("A_i", P.A_i(x)), ("H", P.H(x, y, z))):
        np.testing.assert_array_equal(M.colptr, g[nm + "_colptr"])
        np.testing.assert_array_equal(M.rowidx, g[nm + "_rowidx"])
        np.testing.assert_array_equal(M.val, g[nm + "_val"])
    P.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "solve_*.npz"))))
def test_golden_solve(path):
    g = np.load(path)
    name, N = os.path.basename(path)[6:-4].rsplit("_", 1)
    p0, p1 = {"rosenbrock_cubic_line": (0.3, 0.7), "rosenbrock_disk": (-0.5, 1.2)}.get(name, (0, 0))
    P = OracleProblem(name, int(N), p0, p1)
    st = P.solve(keep_iterates=False)
    tr = P.trace()
    assert st == int(g["status"])
    assert len(tr) == int(g["iterations"])
    np.testing.assert_array_equal([r.error for r in tr], g["error"])
    np.testing.assert_array_equal([r.alpha for r in tr], g["alpha"])
    np.testing.assert_array_equal([r.delta for r in tr], g["delta"])
    np.testing.assert_array_equal([r.factorizations for r in tr], g["factorizations"])
    np.testing.assert_array_equal(P.solution()[0], g["x"])
    P.close()


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built here")
def test_restated_core_matches_reference_core_trajectory():
    """Whole solves agree iterate by iterate, bit for bit, between expr.hpp and
    the reference's expression.hpp."""
    for name, N in (("cart_pole", 30), ("flywheel", 40)):
        a = OracleProblem(name, N, backend="restated")
        b = OracleProblem(name, N, backend="reference")
        assert a.solve() == b.solve()
        ta, tb = a.trace(), b.trace()
        assert len(ta) == len(tb)
        for ra, rb in zip(ta, tb):
            np.testing.assert_array_equal(ra.x, rb.x)
            np.testing.assert_array_equal(ra.y, rb.y)
            np.testing.assert_array_equal(ra.z, rb.z)
        a.close(); b.close()


def _random_kkt(n, me, seed):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    H = sp.random(n, n, 0.05, random_state=seed, format="csc")
    H = H + H.T + sp.identity(n) * (n * 0.2)
    A = sp.random(me, n, 0.1, random_state=seed + 1, format="csc") + \
        sp.eye(me, n, format="csc")
    K = sp.bmat([[H, None], [A, -1e-6 * sp.identity(me)]], format="csc")
    K = sp.tril(K).tocsc()
    K.sort_indices()
    return K, rng.standard_normal(n + me)


def test_ldlt_solves_and_counts_inertia():
    import scipy.sparse as sp
    n, me = 60, 25
    K, rhs = _random_kkt(n, me, 7)
    full = (K + sp.tril(K, -1).T).tocsc()
    for perm in (None, np.arange(n + me, dtype=np.int32)[::-1].copy()):
        nnzL, D, x, h = ldlt(n + me, K.indptr, K.indices, K.data, rhs, perm)
        assert nnzL > 0
        assert (D > 0).sum() == n and (D < 0).sum() == me
        assert np.abs(full @ x - rhs).max() < 1e-8


def test_amd_is_a_fill_reducing_permutation():
    n, me = 80, 30
    K, rhs = _random_kkt(n, me, 11)
    p = amd(n + me, K.indptr, K.indices)
    assert sorted(p) == list(range(n + me))
    nat = ldlt(n + me, K.indptr, K.indices, K.data, rhs,
               np.arange(n + me, dtype=np.int32))[0]
    ord_ = ldlt(n + me, K.indptr, K.indices, K.data, rhs, p)[0]
    assert ord_ <= nat


def test_zero_pivot_is_reported():
    import scipy.sparse as sp
    K = sp.csc_matrix(np.array([[0.0, 0.0], [1.0, 1.0]]))
    K = sp.tril(K).tocsc()
    K.sort_indices()
    # force an explicit zero diagonal entry
    indptr = np.array([0, 2, 3], dtype=np.int32)
    indices = np.array([0, 1, 1], dtype=np.int32)
    data = np.array([0.0, 1.0, 1.0])
    nnzL, D, x, h = ldlt(2, indptr, indices, data, np.ones(2),
                         np.arange(2, dtype=np.int32))
    assert nnzL == -1


def test_cart_pole_sensitivity_floor():
    """How far two CORRECT runs of the reference algorithm drift apart: the
    oracle against itself with two elimination orders (its own AMD vs the
    nested-dissection order the device uses). Same decisions, but the iterates
    differ by ~1e-8 after ONE cart-pole step and by ~1e-5 after ten — the
    regularised KKT system (γ = 1e-10, unpivoted factor) amplifies rounding by
    ≳ 1e8. This is the floor any implementation's iterate parity sits on, and
    why tests/test_gpu_parity.py compares cart-pole trajectories at 1e-5 and
    single steps from a common state at 1e-6."""
    import emu
    name, N = "cart_pole", 100
    E = emu.Emu(name, N)
    E.L.emu_kkt_build(E.h)
    E.analyze(0)
    nd = E.perm()
    E.close()
    traces = []
    for perm in (nd, None):
        O = OracleProblem(name, N)
        O.solve(max_iterations=10, perm=perm, force_sparse=1)
        traces.append(O.trace())
        O.close()

    def rel(a, b):
        return np.abs(a - b).max() / np.abs(b).max()

    drift = [rel(a.x, b.x) for a, b in zip(*traces)]
    for a, b in zip(*traces):
        assert a.delta == b.delta and a.factorizations == b.factorizations
        assert a.trials == b.trials
    assert 1e-10 < drift[0] < 1e-6      # one step: already ~1e-8
    assert max(drift) > 1e-7            # and it grows along the trajectory
    assert max(drift) < 1e-3
