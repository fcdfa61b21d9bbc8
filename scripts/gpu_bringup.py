"""First-contact script for a GPU box: small problems, kernel parity vs the
oracle, LDLT parity, a traced solve, and quick timings. Prints a log."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sleipnir_b200 as sb
from oracle.pyoracle import OracleProblem, EXIT_STATUS, ldlt as oracle_ldlt

def section(t): print("\n==== " + t + " ====", flush=True)

def small_problems():
    section("small known-answer problems")
    cases = [("lp_maximize", (375, 250)), ("quartic", (1,)), ("wachter_biegler", (1, 0, 0.5)),
             ("qp_inequality_2d", (3 + 1 / 3, 1 + 2 / 3)), ("conflicting_bounds", None),
             ("locally_infeasible_ineq", None), ("nonfinite_ineq", None), ("nonfinite_ineq_jacobian", None),
             ("rosenbrock_disk", (1, 1)), ("rosenbrock_cubic_line", None)]
    for name, expect in cases:
        try:
            P = sb.Problem(name, 0, -0.5, 1.2)
            st = P.solve()
            x, *_ = P.solution()
            O = OracleProblem(name, 0, -0.5, 1.2); so = O.solve(); xo, *_ = O.solution()
            print(f"{name:28s} gpu {sb.EXIT_STATUS[st]:24s} x={np.round(x,8)} iters {len(P.trace())} | oracle {EXIT_STATUS[so]} x={np.round(xo,8)} iters {len(O.trace())} expect {expect}", flush=True)
        except Exception as e:
            print(name, "EXCEPTION", repr(e)); traceback.print_exc()

def kernel_parity(name, N):
    section(f"kernel parity {name} N={N}")
    P = sb.Problem(name, N); O = OracleProblem(name, N); O.eval_setup()
    df, dce, dci = O.scaling()
    D = P.open_device()
    D.set_scaling(df, dce, dci)
    rng = np.random.default_rng(3)
    n, me, mi = P.n, P.me, P.mi
    x = O.initial_guess() + 0.02 * rng.standard_normal(n); y = 0.1 * rng.standard_normal(me)
    z = np.abs(rng.standard_normal(mi)) + 0.2; s = np.abs(rng.standard_normal(mi)) + 0.2
    D.set_iterate(x, s, y, z)
    info = D.eval_current(1)
    def cmp(nm, a, b):
        a = np.asarray(a); b = np.asarray(b)
        if a.shape != b.shape: print(f"  {nm}: SHAPE {a.shape} vs {b.shape}"); return
        if a.size == 0: print(f"  {nm}: empty"); return
        print(f"  {nm:6s} max|diff| {np.abs(a-b).max():.3e} rel {np.abs(a-b).max()/max(np.abs(b).max(),1e-300):.3e} bitexact {bool((a==b).all())}", flush=True)
    cmp("f", [info.f], [O.f(x)])
    cmp("c_e", D.download(sb.ARR_C_E), O.c_e(x)); cmp("c_i", D.download(sb.ARR_C_I), O.c_i(x))
    cmp("g", D.download(sb.ARR_G), O.g(x))
    for nm, arr, pat, M in [("A_e", sb.ARR_A_E_VAL, sb.OUT_A_E, O.A_e(x)), ("A_i", sb.ARR_A_I_VAL, sb.OUT_A_I, O.A_i(x)), ("H", sb.ARR_H_VAL, sb.OUT_H_C, O.H(x, y, z))]:
        r, c, cp, ri = D.pattern(pat)
        ok = np.array_equal(cp, M.colptr) and np.array_equal(ri, M.rowidx)
        print(f"  {nm} pattern equal {ok} nnz {len(ri)}")
        if ok: cmp(nm, D.download(arr), M.val)
    print("  point info:", info.f, info.ce_l1, info.cis_l1, info.log_s_sum, bin(info.finite), info.ci_all_positive)
    ce, ci = O.c_e(x), O.c_i(x)
    print("  expect    :", O.f(x), np.abs(ce).sum(), np.abs(ci - s).sum(), np.log(s).sum())
    # factor / solve
    st = D.analyze(sb.ORDER_NESTED_DISSECTION)
    print(f"  symbolic: dim {st.dim} nnzK {st.nnz_kkt} nnzL {st.nnz_l} stored {st.nnz_l_stored} supers {st.n_supernodes} levels {st.n_levels} maxfront {st.max_front} etree {st.etree_height}")
    perm = D.permutation()
    for delta, gamma in [(1e-4, 1e-10), (1.0, 1e-6)]:
        fi = D.factor(delta, gamma, True)
        r, c, cp, ri = D.pattern(-1)
        kv = D.download(sb.ARR_KKT_VAL)
        kvr = kv.copy(); dim = n + me
        reg = np.concatenate([np.full(n, delta), np.full(me, -gamma)])
        for col in range(dim):
            seg = ri[cp[col]:cp[col + 1]]; k = cp[col] + np.searchsorted(seg, col); kvr[k] += reg[col]
        mu, tau = 0.1 * df, 0.99
        si = D.solve(mu, tau)
        rhs = D.download(sb.ARR_RHS)
        nnzL, Do, xo, hgt = oracle_ldlt(dim, cp, ri, kvr, rhs, perm)
        Dg = D.download(sb.ARR_D)
        px, py = D.download(sb.ARR_P_X), D.download(sb.ARR_P_Y)
        sol = np.concatenate([px, -py])
        print(f"  δ={delta:g} γ={gamma:g}: inertia ({fi.n_pos},{fi.n_neg},{fi.n_zero}) zp {fi.zero_pivot} min|D| {fi.min_abs_d:.3e} | oracle ({(Do>0).sum()},{(Do<0).sum()}) nnzL {nnzL} | D rel diff {np.abs(Dg-Do).max()/np.abs(Do).max():.2e} | sol rel diff {np.abs(sol-xo).max()/np.abs(xo).max():.2e} | alpha_max {si.alpha_max:.6g} alpha_z {si.alpha_z:.6g}", flush=True)
    P.close_device()

def traced_solve(name, N, iters):
    section(f"traced solve {name} N={N} (first {iters} iterations), same permutation on both sides")
    P = sb.Problem(name, N)
    t = time.time(); st = P.solve(max_iterations=iters, keep_iterates=True); tg = time.time() - t
    tr = P.trace(); sym = P.symbolic_stats(); cnt = P.counters()
    print(f"gpu: {sb.EXIT_STATUS[st]} rows {len(tr)} total {tg:.3f}s loop {P.loop_seconds():.4f}s -> {len(tr)/max(P.loop_seconds(),1e-9):.1f} steps/s | {sym} | launches {cnt['kernel_launches']} fact {cnt['factorizations']} solves {cnt['solves']}")
    # need the permutation: re-open device session to fetch it
    D = P.open_device(); D.analyze(sb.ORDER_NESTED_DISSECTION); perm = D.permutation(); P.close_device()
    O = OracleProblem(name, N)
    so = O.solve(max_iterations=iters, perm=perm, force_sparse=1)
    to = O.trace()
    print(f"oracle: {EXIT_STATUS[so]} rows {len(to)} solve {O.solve_seconds():.3f}s -> {len(to)/max(O.solve_seconds(),1e-9):.2f} steps/s")
    for k in range(min(len(tr), len(to))):
        a, b = tr[k], to[k]
        dx = np.abs(a.x - b.x).max() / max(np.abs(b.x).max(), 1e-300)
        dy = np.abs(a.y - b.y).max() / max(np.abs(b.y).max(), 1e-300) if a.y.size else 0.0
        dz = np.abs(a.z - b.z).max() / max(np.abs(b.z).max(), 1e-300)
        if k < 12 or k % 10 == 0:
            print(f"  it {k:3d} dx {dx:.2e} dy {dy:.2e} dz {dz:.2e} | delta {a.delta:.3g}/{b.delta:.3g} fact {a.factorizations}/{b.factorizations} trials {a.trials}/{b.trials} alpha {a.alpha:.6g}/{b.alpha:.6g} err {a.error:.6e}/{b.error:.6e}", flush=True)

def timing(name, N, iters):
    section(f"timing {name} N={N}, {iters} iterations")
    P = sb.Problem(name, N)
    t = time.time(); st = P.solve(max_iterations=iters); tg = time.time() - t
    tr = P.trace(); sym = P.symbolic_stats(); cnt = P.counters()
    print(f"gpu: {sb.EXIT_STATUS[st]} rows {len(tr)} total {tg:.3f}s loop {P.loop_seconds():.4f}s -> {len(tr)/max(P.loop_seconds(),1e-9):.1f} steps/s")
    print(f"   {sym}\n   {cnt}")
    print(f"   fact/iter {sum(r.factorizations for r in tr)/len(tr):.2f} solves/iter {sum(r.solves for r in tr)/len(tr):.2f} trials/iter {sum(r.trials for r in tr)/len(tr):.2f}", flush=True)

if __name__ == "__main__":
    what = sys.argv[1:] or ["small", "parity", "trace", "timing"]
    for w in what:
        try:
            if w == "small": small_problems()
            elif w == "parity":
                kernel_parity("flywheel", 50); kernel_parity("cart_pole", 20); kernel_parity("cart_pole", 200)
            elif w == "trace":
                traced_solve("flywheel", 50, 100); traced_solve("cart_pole", 100, 60)
            elif w == "timing":
                timing("cart_pole", 1000, 30); timing("cart_pole", 5000, 30)
        except Exception as e:
            print("SECTION", w, "FAILED:", repr(e)); traceback.print_exc()
