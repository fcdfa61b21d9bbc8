"""The two arithmetic modes of the LDLᵀ factorisation (include/slpb.h,
slpb_set_factor_arithmetic; csrc/ldlt_core.hpp) on the device:

 * SLPB_ARITH_REFERENCE (default): the blocked warp-per-front elimination
   rounds product and difference separately, like the reference's x86-64 build;
 * SLPB_ARITH_TENSOR: fused Schur updates, the rank-4 updates of frontal
   matrices of order ≥ 16 on the FP64 tensor cores (BASELINE config 3; the
   dense kernels behind solver/util/dense_regularized_ldlt.hpp:59-136).

In BOTH modes the device must reproduce, BIT FOR BIT, the generic per-front body
of ldlt_core.hpp run on the host (tests/emu) from the same lhs values — for the
tensor mode that pins the tensor core's accumulation order (a chain of FMAs in
ascending pivot order) inside the production kernel — and the batched kernels
must reproduce the single-instance path. Against the CPU oracle the tensor mode
meets the same one-step bars as the reference mode."""
import numpy as np
import pytest

import sleipnir_b200 as sb
from emu import Emu
from oracle.pyoracle import OracleProblem, ldlt

pytestmark = pytest.mark.gpu
MODES = [sb.ARITH_REFERENCE, sb.ARITH_TENSOR]


def _point(P, O, seed):
    rng = np.random.default_rng(seed)
    x = O.initial_guess() + 0.02 * rng.standard_normal(P.n)
    y = 0.1 * rng.standard_normal(P.me)
    z = 0.5 + np.abs(rng.standard_normal(P.mi))
    s = 0.5 + np.abs(rng.standard_normal(P.mi))
    return x, s, y, z


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,N,order", [
    ("cart_pole", 100, sb.ORDER_NESTED_DISSECTION),
    ("cart_pole", 400, sb.ORDER_AMD),
    ("gfold", 60, sb.ORDER_NESTED_DISSECTION),
    ("arm_on_elevator", 100, sb.ORDER_NESTED_DISSECTION),
    # a time-step variable shared by all stages: fronts above order 32 — the
    # hybrid path (tree kernel below them, one block per front above)
    ("differential_drive_ocp", 50, sb.ORDER_NESTED_DISSECTION),
    # … and a front beyond one block's shared memory (order > 158): its
    # workspace lives in global memory
    ("differential_drive_ocp", 100, sb.ORDER_NESTED_DISSECTION),
    ("flywheel", 50, sb.ORDER_NATURAL)])
def test_device_factor_equals_the_host_emulation_bit_for_bit(name, N, order, mode):
    P, O = sb.Problem(name, N), OracleProblem(name, N)
    O.eval_setup()
    D = P.open_device()
    D.set_scaling(*O.scaling())
    x, s, y, z = _point(P, O, 7)
    D.set_iterate(x, s, y, z)
    D.eval_current(1)
    st = D.analyze(order)
    perm = D.permutation()
    E = Emu(name, N)
    E.eval(x, y, z, 1.0, np.ones(E.me), np.ones(E.mi))
    E.kkt(z / s if P.mi else np.ones(0))
    se = E.analyze(sb.ORDER_CUSTOM, perm)
    assert se["n_super"] == st.n_supernodes and se["max_front"] == st.max_front
    F, npv, _, _ = E.fronts(se["n_super"])
    D.set_factor_arithmetic(mode)
    E.set_fused(mode == sb.ARITH_TENSOR)
    for delta, gamma in ((1.0, 1e-6), (1e-4, 1e-10), (0.0, 0.0)):
        fi = D.factor(delta, gamma, True)
        E.set_kkt_values(D.download(sb.ARR_KKT_VAL))
        info, mn, De = E.factor(delta, gamma)
        assert bool(fi.zero_pivot) == bool(info[3])
        if fi.zero_pivot:
            continue   # an abandoned variant leaves its later fronts untouched
        assert (fi.n_pos, fi.n_neg, fi.n_zero) == tuple(info[:3])
        assert fi.min_abs_d == mn
        Dd = D.download(sb.ARR_D)
        assert Dd.tobytes() == De.tobytes(), np.abs(Dd - De).max()
    # the tensor-core path was actually exercised
    if mode == sb.ARITH_TENSOR and name != "flywheel":
        assert (F >= 16).sum() > 0
    if name == "differential_drive_ocp":
        assert F.max() > (158 if N == 100 else 32) and (F <= 32).sum() > 0.8 * len(F)   # hybrid
        # … and the solve through both kinds of kernels: L D Lᵀ x = rhs
        D.factor(1.0, 1e-6, True)
        E.set_kkt_values(D.download(sb.ARR_KKT_VAL))
        E.factor(1.0, 1e-6)
        D.solve(0.1, 0.99)
        x_dev = np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)])
        x_emu = E.solve(D.download(sb.ARR_RHS))
        assert x_dev.tobytes() == x_emu.tobytes(), np.abs(x_dev - x_emu).max()
    E.close(); P.close_device(); P.close(); O.close()


def test_modes_differ_only_by_rounding_and_both_match_the_oracle():
    """D and the Newton step of the two modes agree to rounding level on a
    well-conditioned regularisation, and both meet the oracle's LDLᵀ (same
    permutation) at the single path's tolerance."""
    name, N = "cart_pole", 300
    P, O = sb.Problem(name, N), OracleProblem(name, N)
    O.eval_setup()
    D = P.open_device()
    D.set_scaling(*O.scaling())
    x, s, y, z = _point(P, O, 3)
    D.set_iterate(x, s, y, z)
    D.eval_current(1)
    D.analyze()
    got = {}
    for mode in MODES:
        D.set_factor_arithmetic(mode)
        fi = D.factor(1.0, 1e-6, True)
        assert (fi.n_pos, fi.n_neg, fi.n_zero) == (P.n, P.me, 0)
        D.solve(0.1, 0.99)
        got[mode] = (D.download(sb.ARR_D),
                     np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)]))
    d0, p0 = got[sb.ARITH_REFERENCE]
    d1, p1 = got[sb.ARITH_TENSOR]
    assert d0.tobytes() != d1.tobytes()          # the modes really are different arithmetic
    np.testing.assert_allclose(d1, d0, rtol=1e-8)
    np.testing.assert_allclose(p1, p0, rtol=0, atol=1e-9 * np.abs(p0).max())
    _, _, cp, ri = D.pattern(-1)
    dim = P.n + P.me
    kvr = D.download(sb.ARR_KKT_VAL).copy()
    for c in range(dim):
        k = cp[c] + np.searchsorted(ri[cp[c]:cp[c + 1]], c)
        kvr[k] += 1.0 if c < P.n else -1e-6
    _, Do, xo, _ = ldlt(dim, cp, ri, kvr, D.download(sb.ARR_RHS), D.permutation())
    for mode in MODES:
        np.testing.assert_allclose(got[mode][0], Do, rtol=1e-9)
        np.testing.assert_allclose(got[mode][1], xo, rtol=0, atol=1e-9 * np.abs(xo).max())
    P.close_device(); P.close(); O.close()


def test_tensor_mode_pair_and_fused_forward_substitution():
    """The speculated pair and the forward substitution carried by the factor
    launch produce the same bits as single factorisations + separate solves in
    the tensor mode too."""
    name, N = "cart_pole", 200
    P, O = sb.Problem(name, N), OracleProblem(name, N)
    O.eval_setup()
    D = P.open_device()
    D.set_scaling(*O.scaling())
    D.set_factor_arithmetic(sb.ARITH_TENSOR)
    x, s, y, z = _point(P, O, 5)
    D.set_iterate(x, s, y, z)
    D.eval_current(1)
    D.analyze()
    singles = []
    for d, g in ((0.0, 0.0), (1e-2, 1e-10)):
        fi = D.factor(d, g, True)
        singles.append(((fi.n_pos, fi.n_neg, fi.n_zero, fi.zero_pivot), D.download(sb.ARR_D)))
    D.solve(0.1, 0.99)
    want = np.concatenate([D.download(sb.ARR_P_X), D.download(sb.ARR_P_Y)])
    D.prepare_rhs(0.1)
    i0, i1 = D.factor_pair([0.0, 1e-2], [0.0, 1e-10], True)
    assert (i1.n_pos, i1.n_neg, i1.n_zero, i1.zero_pivot) == singles[1][0]
    D.select_factor(1)
    assert D.download(sb.ARR_D).tobytes() == singles[1][1].tobytes()
    D.solve(0.1, 0.99)
    got = np.concatenate([D.download(sb.ARR_P_X), D.download(sb.ARR_P_Y)])
    assert got.tobytes() == want.tobytes()
    P.close_device(); P.close(); O.close()


@pytest.mark.parametrize("name,N,batch", [("cart_pole", 100, 40), ("gfold", 40, 33)])
def test_tensor_mode_batch_is_bit_identical_to_the_single_path(name, N, batch):
    P, O = sb.Problem(name, N), OracleProblem(name, N)
    O.eval_setup()
    D = P.open_device()
    D.set_scaling(*O.scaling())
    D.set_factor_arithmetic(sb.ARITH_TENSOR)
    D.set_iterate(O.initial_guess(), np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
    D.eval_current(1)
    D.analyze()
    rng = np.random.default_rng(9)
    systems = []
    for i in range(4):
        D.set_iterate(*_point(P, O, 20 + i))
        D.eval_current(1)
        fi = D.factor(1.0 + i, 1e-6, True)
        D.solve(0.1, 0.99)
        systems.append(dict(kkt=D.download(sb.ARR_KKT_VAL), rhs=D.download(sb.ARR_RHS),
                            D=D.download(sb.ARR_D), min_abs_d=fi.min_abs_d,
                            sol=np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)])))
    B = sb.Batch(D, batch)     # inherits the solver's arithmetic mode
    for i in range(batch):
        B.set_system(i, systems[i % 4]["kkt"], systems[i % 4]["rhs"])
    info = B.factor(np.array([1.0 + (i % 4) for i in range(batch)]), 1e-6)
    B.solve()
    for i in range(batch):
        src = systems[i % 4]
        assert info[i].min_abs_d == src["min_abs_d"]
        assert B.get(i, sb.Batch.D).tobytes() == src["D"].tobytes()
        assert B.get(i, sb.Batch.SOLUTION).tobytes() == src["sol"].tobytes()
    # …and the reference mode gives different bits for the same systems
    D.set_factor_arithmetic(sb.ARITH_REFERENCE)
    B2 = sb.Batch(D, 32)
    B2.set_system(0, systems[0]["kkt"], systems[0]["rhs"])
    B2.factor(1.0, 1e-6)
    assert B2.get(0, sb.Batch.D).tobytes() != systems[0]["D"].tobytes()
    B2.close(); B.close()
    P.close_device(); P.close(); O.close()


@pytest.mark.parametrize("name,N", [("flywheel", 50), ("cart_pole", 60), ("double_integrator", 200)])
def test_tensor_mode_solves_to_the_same_optimum(name, N):
    """Whole solves in the tensor mode reach SUCCESS at the optimum of the
    reference mode (the iteration path may differ, see ldlt_core.hpp)."""
    sols = {}
    for mode in MODES:
        P = sb.Problem(name, N)
        P.set_factor_arithmetic(mode)
        assert sb.EXIT_STATUS[P.solve()] == "SUCCESS"
        sols[mode] = (P.solution()[0], P.trace()[-1].cost)
        P.close()
    x0, c0 = sols[sb.ARITH_REFERENCE]
    x1, c1 = sols[sb.ARITH_TENSOR]
    assert c1 == pytest.approx(c0, rel=1e-6, abs=1e-9)
    np.testing.assert_allclose(x1, x0, atol=2e-4 * max(np.abs(x0).max(), 1.0))


def test_device_timers_are_sampled_and_counted():
    """slpb_get_timers: the kernel groups are timed on one run in
    SLPB_TIMER_EVERY (timing events between back-to-back kernels are not free);
    `launches` counts every run, `count` the sampled ones."""
    P = sb.Problem("cart_pole", 100)
    P.solve(max_iterations=40)
    tim = P.timers()
    tr = P.trace()
    assert tim["eval_full"]["launches"] >= len(tr)
    for name in ("eval_full", "factor", "solve"):
        t = tim[name]
        assert 1 <= t["count"] <= t["launches"]
        assert t["count"] <= t["launches"] // 8 + 1
        assert 0.0 < t["mean_ms"] < 5.0
    P.close()
