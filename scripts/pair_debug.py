import sys
sys.path.insert(0, ".")
import numpy as np
import sleipnir_b200 as sb
name, N = "gfold", 20
P = sb.Problem(name, N)
D = P.open_device()
x = P.initial_guess(); s = np.ones(P.mi); y = np.zeros(P.me); z = np.ones(P.mi)
D.set_iterate(x, s, y, z); D.eval_current(1); st = D.analyze()
print("sym", st.dim, st.n_supernodes, st.n_levels, st.max_front)
mu = 0.1
for (d, g) in [(1e-4, 1e-10), (5e-5, 1e-10)]:
    f1 = D.factor(d, g, True); s1 = D.solve(mu, 0.99); px1 = D.download(sb.ARR_P_X); D1 = D.download(sb.ARR_D)
    fa, fb = D.factor_pair([0.0, d], [0.0, g], True); D.select_factor(1); s2 = D.solve(mu, 0.99); px2 = D.download(sb.ARR_P_X); D2 = D.download(sb.ARR_D)
    print("single", (f1.n_pos, f1.n_neg, f1.n_zero, f1.zero_pivot, f1.min_abs_d), "pair v0", (fa.n_pos, fa.n_neg, fa.n_zero, fa.zero_pivot, fa.min_abs_d), "v1", (fb.n_pos, fb.n_neg, fb.n_zero, fb.zero_pivot, fb.min_abs_d))
    print("  D equal", np.array_equal(D1, D2), "px equal", np.array_equal(px1, px2), "max|dpx|", np.nanmax(np.abs(px1 - px2)), "nan in px2", np.isnan(px2).sum(), "alpha", s1.alpha_max, s2.alpha_max)
