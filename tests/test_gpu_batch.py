"""Batched many-instance LDLᵀ (slpb_batch_*, csrc/batch.cuh) against the
single-instance path: every instance of a batch must reproduce slpb_factor /
slpb_solve on the same system BIT FOR BIT (same products in the same order:
the lane-per-instance kernels run the scalar elimination of ldlt_core.hpp), and
the oracle's LDLᵀ to the single path's tolerance. Reference surface:
optimization/multistart.hpp:44-73 (many starts, one pattern)."""
import numpy as np
import pytest

import sleipnir_b200 as sb
from oracle.pyoracle import OracleProblem, ldlt

pytestmark = pytest.mark.gpu


def _systems(P, O, D, count, seed, delta, gamma):
    """`count` perturbed iterates: assembles each on the single path, keeps its
    KKT values, rhs, D and solution."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        x = O.initial_guess() + 0.02 * rng.standard_normal(P.n)
        y = 0.1 * rng.standard_normal(P.me)
        z = 0.5 + np.abs(rng.standard_normal(P.mi))
        s = 0.5 + np.abs(rng.standard_normal(P.mi))
        D.set_iterate(x, s, y, z)
        D.eval_current(1)
        fi = D.factor(delta[i], gamma[i], True)
        D.solve(0.1, 0.99)
        out.append(dict(
            kkt=D.download(sb.ARR_KKT_VAL), rhs=D.download(sb.ARR_RHS),
            D=D.download(sb.ARR_D), info=(fi.n_pos, fi.n_neg, fi.n_zero, fi.zero_pivot),
            min_abs_d=fi.min_abs_d,
            sol=np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)])))
    return out


@pytest.mark.parametrize("name,N,batch", [("cart_pole", 100, 40), ("gfold", 40, 33),
                                          ("cart_pole", 1000, 64)])
def test_batch_is_bit_identical_to_the_single_path(name, N, batch):
    P, O = sb.Problem(name, N), OracleProblem(name, N)
    O.eval_setup()
    D = P.open_device()
    D.set_scaling(*O.scaling())
    D.set_iterate(O.initial_guess(), np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
    D.eval_current(1)
    D.analyze()
    rng = np.random.default_rng(5)
    # instance-specific regularisations, some of them the reference's own pair
    delta = np.where(rng.random(batch) < 0.5, 1.0, 1e-4) * (1 + rng.random(batch))
    gamma = np.where(rng.random(batch) < 0.5, 1e-6, 1e-10)
    n_sys = min(batch, 6)
    systems = _systems(P, O, D, n_sys, 11, delta, gamma)
    B = sb.Batch(D, batch)
    for i in range(batch):
        src = systems[i % n_sys]
        if i % 2 == 0:
            B.set_system(i, src["kkt"], src["rhs"])
        else:
            B.set_system(i, src["kkt"], None)
            B.set_system(i, None, src["rhs"])
    d_b = np.array([delta[i % n_sys] for i in range(batch)])
    g_b = np.array([gamma[i % n_sys] for i in range(batch)])
    info = B.factor(d_b, g_b)
    B.solve()
    for i in range(batch):
        src = systems[i % n_sys]
        assert (info[i].n_pos, info[i].n_neg, info[i].n_zero, info[i].zero_pivot) == src["info"]
        assert info[i].min_abs_d == src["min_abs_d"]
        np.testing.assert_array_equal(B.get(i, sb.Batch.D), src["D"])
        np.testing.assert_array_equal(B.get(i, sb.Batch.SOLUTION), src["sol"])
    # … and one instance against the oracle's LDLᵀ of the same system
    _, _, cp, ri = D.pattern(-1)
    dim = P.n + P.me
    src = systems[0]
    kvr = src["kkt"].copy()
    for c in range(dim):
        k = cp[c] + np.searchsorted(ri[cp[c]:cp[c + 1]], c)
        kvr[k] += delta[0] if c < P.n else -gamma[0]
    _, Do, xo, _ = ldlt(dim, cp, ri, kvr, src["rhs"], D.permutation())
    if gamma[0] > 1e-8:
        np.testing.assert_allclose(B.get(0, sb.Batch.D), Do, rtol=1e-9)
    B.close()
    P.close_device(); P.close(); O.close()


def test_batch_capture_from_a_live_solver():
    """slpb_batch_capture: device-to-device hand-over of a solver's assembled
    lhs and rhs (what a multistart worker does between slpb_prepare_rhs and the
    factorisation)."""
    name, N = "cart_pole", 60
    P, O = sb.Problem(name, N), OracleProblem(name, N)
    O.eval_setup()
    D = P.open_device()
    D.set_scaling(*O.scaling())
    rng = np.random.default_rng(3)
    x = O.initial_guess() + 0.01 * rng.standard_normal(P.n)
    D.set_iterate(x, np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
    D.eval_current(1)
    D.analyze()
    B = sb.Batch(D, 3)
    want = []
    for i in range(3):
        D.set_iterate(x + 0.01 * i, np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
        D.eval_current(1)
        D.factor(1e-2, 1e-8, True)
        D.solve(0.1, 0.99)
        B.capture(i)
        want.append(np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)]))
    B.factor(1e-2, 1e-8)
    B.solve()
    for i in range(3):
        np.testing.assert_array_equal(B.get(i), want[i])
    B.close()
    P.close_device(); P.close(); O.close()
