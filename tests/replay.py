"""One-step replay of the interior-point iteration along the ORACLE's trajectory
(test infrastructure for tests/test_gpu_replay.py and scripts/replay_report.py).

The oracle (reference algorithm, reference's own autodiff core when built, AMD
elimination order like Eigen::SimplicialLDLT) runs the first iterations of a
config and keeps every iterate. For iteration k the device is put at the
oracle's iterate k−1 with the oracle's μ and previous δ, and takes ONE Newton
iteration through the C ABI: evaluation, KKT assembly, the δ/γ inertia loop,
solve, step recovery, fraction-to-the-boundary rule, the step the oracle
accepted, commit. Compared per iteration:

  decisions   number of factorisations, final (δ, γ), inertia
  α_max, α_z  against the oracle's row
  next iterate x⁺ s⁺ y⁺ z⁺ against the oracle's iterate k (relative, ∞-norm)
  Newton step p against (a) the oracle's LDLᵀ of the SAME assembled system and
              (b) the exact solution p* of that system (SuperLU + iterative
              refinement with 80-bit residuals): e_gpu = ‖p_gpu − p*‖/‖p*‖,
              e_cpu = ‖p_cpu − p*‖/‖p*‖, cond₁ estimate of the system.

interior_point.hpp:382-863 is the loop being replayed; the δ/γ sequence is
solver/util/sparse_regularized_ldlt.hpp:64-152.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import sleipnir_b200 as sb
from oracle.pyoracle import OracleProblem, have_reference, ldlt

EPS = np.finfo(float).eps


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def oracle_rows(name, N, iterations):
    backend = "reference" if have_reference() else "restated"
    O = OracleProblem(name, N, backend=backend)
    O.solve(max_iterations=iterations, keep_iterates=True)   # AMD, default Options
    rows = O.trace()
    O.eval_setup()
    scaling = O.scaling()
    return O, rows, scaling


def regularization_loop(D, n, me, prev_delta, gamma_min=1e-10):
    """DeviceRegularizedLDLT::compute without the speculated pair; returns
    (factorisations, δ, γ, last FactorInfo)."""
    def ideal(fi):
        return (fi.n_pos, fi.n_neg, fi.n_zero) == (n, me, 0)

    count = 1
    fi = D.factor(0.0, 0.0, True)
    if not fi.zero_pivot and ideal(fi) and fi.min_abs_d >= 1e-4:
        return count, 0.0, 0.0, fi
    delta = 1e-4 if prev_delta == 0.0 else max(prev_delta / 2.0, EPS)
    gamma = gamma_min
    while True:
        count += 1
        fi = D.factor(delta, gamma, False)
        if not fi.zero_pivot:
            if ideal(fi):
                return count, delta, gamma, fi
            if fi.n_zero > 0:
                if gamma == 0.0:
                    gamma = 1e-10
                else:
                    delta *= 10.0
                    gamma *= 10.0
            elif fi.n_neg > me:
                delta *= 10.0
            elif fi.n_pos > n:
                gamma = 1e-10 if gamma == 0.0 else gamma * 10.0
        else:
            delta *= 10.0
            gamma = 1e-10 if gamma == 0.0 else gamma * 10.0
        if delta > 1e20 or gamma > 1e20:
            return count, delta, gamma, fi


def exact_solution(K_full, rhs, x0=None):
    """Solution of K p = rhs to well below double precision: SuperLU with
    partial pivoting, then iterative refinement with residuals accumulated in
    80-bit long double."""
    lu = spla.splu(K_full.tocsc())
    coo = K_full.tocoo()
    r_idx, c_idx = coo.row, coo.col
    v = coo.data.astype(np.longdouble)
    b = rhs.astype(np.longdouble)
    x = (lu.solve(rhs) if x0 is None else x0).astype(np.longdouble)
    for _ in range(8):
        r = b.copy()
        np.subtract.at(r, r_idx, v * x[c_idx])
        dx = lu.solve(np.asarray(r, dtype=np.float64))
        x = x + dx
        if np.abs(dx).max() <= 1e-19 * max(np.abs(x).max(), 1e-300):
            break
    return x, lu


def replay(name, N, iterations, ordering, compare_amd_factor=True,
           arithmetic=None):
    """Returns (rows, report): report[k] is a dict of the quantities listed in
    the module docstring for iteration k = 1 … iterations−1."""
    O, rows, (d_f, d_ce, d_ci) = oracle_rows(name, N, iterations)
    P = sb.Problem(name, N)
    D = P.open_device()
    D.set_scaling(d_f, d_ce, d_ci)
    if arithmetic is not None:
        D.set_factor_arithmetic(arithmetic)   # slpb_factor_arithmetic
    n, me, mi, dim = P.n, P.me, P.mi, P.n + P.me
    D.set_iterate(rows[0].x, rows[0].s, rows[0].y, rows[0].z)
    D.eval_current(1)
    st = D.analyze(ordering)
    perm = D.permutation()
    _, _, cp, ri = D.pattern(-1)
    diag_pos = np.array([cp[c] + np.searchsorted(ri[cp[c]:cp[c + 1]], c)
                         for c in range(dim)])
    report = []
    for k in range(1, len(rows)):
        prev, row = rows[k - 1], rows[k]
        if row.type != 0 or prev.type != 0:
            break
        mu = prev.mu
        tau = max(0.99, 1.0 - mu)
        D.set_iterate(prev.x, prev.s, prev.y, prev.z)
        D.eval_current(1)
        count, delta, gamma, fi = regularization_loop(D, n, me, prev.delta)
        step = D.solve(mu, tau)
        rep = dict(iteration=k, factorizations=(count, row.factorizations),
                   delta=(delta, row.delta), gamma=(gamma, row.gamma),
                   inertia=(fi.n_pos, fi.n_neg, fi.n_zero),
                   alpha_max=(step.alpha_max, row.alpha_max),
                   alpha_z=(step.alpha_z, row.alpha_z),
                   solves=row.solves, trials=row.trials, mu=mu)
        # ---- the linear system the device solved, and its exact solution -----
        kv = D.download(sb.ARR_KKT_VAL)
        kvr = kv.copy()
        kvr[diag_pos[:n]] += delta
        kvr[diag_pos[n:]] -= gamma
        rhs = D.download(sb.ARR_RHS)
        p_gpu = np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)])
        low = sp.csc_matrix((kvr, ri, cp), shape=(dim, dim))
        full = (low + sp.tril(low, -1).T).tocsc()
        p_star, lu = exact_solution(full, rhs)
        p_star64 = np.asarray(p_star, dtype=np.float64)
        scale = max(np.abs(p_star64).max(), 1e-300)
        rep["e_gpu"] = float(np.abs(p_gpu - p_star).max() / scale)
        rep["res_gpu"] = float(np.abs(full @ p_gpu - rhs).max() /
                               max(np.abs(rhs).max(), 1e-300))
        if compare_amd_factor:
            _, _, p_cpu, _ = ldlt(dim, cp, ri, kvr, rhs, None)   # oracle, AMD
            rep["e_cpu"] = float(np.abs(p_cpu - p_star).max() / scale)
            rep["res_cpu"] = float(np.abs(full @ p_cpu - rhs).max() /
                                   max(np.abs(rhs).max(), 1e-300))
            rep["gap_p"] = float(np.abs(p_gpu - p_cpu).max() / scale)
        # 1-norm condition estimate of the regularised system
        inv = spla.LinearOperator((dim, dim), matvec=lu.solve, rmatvec=lu.solve)
        rep["cond1"] = float(spla.onenormest(full) * spla.onenormest(inv))
        # ---- what the ORACLE's inputs were at this iterate: its own K and rhs
        # (interior_point.hpp:426-448 in numpy on the oracle's matrices). The
        # two sides' inputs differ in the last bits (libdevice vs glibc
        # sin/cos, summation order of the sweeps); pert = K⁻¹(δrhs − δK p*) is
        # the first-order effect of that difference on the Newton step.
        x0, s0, y0, z0 = prev.x, prev.s, prev.y, prev.z
        Hc, Aec, Aic = O.H(x0, y0, z0), O.A_e(x0), O.A_i(x0)
        Hs = sp.csc_matrix((Hc.val, Hc.rowidx, Hc.colptr), shape=(n, n))
        Aes = sp.csc_matrix((Aec.val, Aec.rowidx, Aec.colptr), shape=(me, n))
        Ais = sp.csc_matrix((Aic.val, Aic.rowidx, Aic.colptr), shape=(mi, n))
        sigma = z0 / s0
        TL = Hs + sp.tril(Ais.T @ sp.diags(sigma) @ Ais)
        low_c = sp.bmat([[TL, None], [Aes, sp.csc_matrix((me, me))]], format="csc")
        full_c = (low_c + sp.tril(low_c, -1).T +
                  sp.diags(np.concatenate([np.full(n, delta), np.full(me, -gamma)]))).tocsc()
        gc, cec, cic = O.g(x0), O.c_e(x0), O.c_i(x0)
        t = -sigma * cic + mu / s0 + z0
        rhs_c = np.concatenate([-gc + Aes.T @ y0 + Ais.T @ t, -cec])
        rep["input_gap"] = dict(K=float(abs(full - full_c).max() / max(abs(full_c).max(), 1e-300)),
                                rhs=rel(rhs, rhs_c))
        pert = lu.solve((rhs - rhs_c) - (full - full_c) @ p_star64)
        rep["pert"] = float(np.abs(pert).max() / scale)
        # ---- the step the oracle accepted, then the commit --------------------
        if row.solves == 1:
            D.trial(row.alpha, row.alpha_z)
            D.accept(mu)   # z is clamped with the μ of this iteration (:797-801)
            x, s, y, z = D.get_iterate()
            rep["next"] = dict(x=rel(x, row.x), s=rel(s, row.s),
                               y=rel(y, row.y), z=rel(z, row.z))
            # The step the ORACLE actually took, recovered from its trajectory,
            # against the exact Newton step p* of the system — its own distance
            # from the exact step, in the units of the iterate comparison
            # (α‖Δp‖∞/‖next iterate‖∞), beside the device's.
            cis = cic - s0
            ps_star = cis + Ais @ p_star64[:n]
            pz_star = mu / s0 - z0 - sigma * ps_star
            run = dict(x=(row.x - prev.x) / row.alpha, y=(row.y - prev.y) / row.alpha_z,
                       s=(row.s - prev.s) / row.alpha if mi else np.zeros(0),
                       z=(row.z - prev.z) / row.alpha_z if mi else np.zeros(0))
            dev = dict(x=D.download(sb.ARR_P_X), y=D.download(sb.ARR_P_Y),
                       s=D.download(sb.ARR_P_S), z=D.download(sb.ARR_P_Z))
            star = dict(x=p_star64[:n], y=-p_star64[n:], s=ps_star, z=pz_star)
            step = dict(x=row.alpha, y=row.alpha_z, s=row.alpha, z=row.alpha_z)
            nxt = dict(x=row.x, y=row.y, s=row.s, z=row.z)
            rep["err_run"], rep["err_dev"] = {}, {}
            for b in "xysz":
                if star[b].size == 0:
                    rep["err_run"][b] = rep["err_dev"][b] = 0.0
                    continue
                den = max(np.abs(nxt[b]).max(), 1e-300)
                rep["err_run"][b] = float(step[b] * np.abs(run[b] - star[b]).max() / den)
                rep["err_dev"][b] = float(step[b] * np.abs(dev[b] - star[b]).max() / den)
        report.append(rep)
    P.close_device()
    P.close()
    O.close()
    return rows, report, dict(nnz_l=st.nnz_l, n_levels=st.n_levels,
                              max_front=st.max_front, perm=perm)


def format_report(report):
    lines = ["it  fact(gpu/cpu)  delta      e_gpu     e_cpu     gap_p     cond1     "
             "next x     next y     next z     a_max rel"]
    for r in report:
        nx = r.get("next", {})
        lines.append(
            f"{r['iteration']:2d}  {r['factorizations'][0]}/{r['factorizations'][1]}"
            f"          {r['delta'][0]:.1e}  {r['e_gpu']:.2e}  "
            f"{r.get('e_cpu', float('nan')):.2e}  {r.get('gap_p', float('nan')):.2e}  "
            f"{r['cond1']:.2e}  {nx.get('x', float('nan')):.2e}   "
            f"{nx.get('y', float('nan')):.2e}   {nx.get('z', float('nan')):.2e}   "
            f"{abs(r['alpha_max'][0] - r['alpha_max'][1]) / max(abs(r['alpha_max'][1]), 1e-300):.1e}"
            + (f"   own error of the step: oracle x {r['err_run']['x']:.1e} y {r['err_run']['y']:.1e}"
               f" z {r['err_run']['z']:.1e} | device x {r['err_dev']['x']:.1e} y {r['err_dev']['y']:.1e}"
               f" z {r['err_dev']['z']:.1e} | pert {r['pert']:.0e}" if "err_run" in r else ""))
    return "\n".join(lines)
