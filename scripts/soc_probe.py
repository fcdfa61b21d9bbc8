"""Development aid: first iteration of cart_pole_eq by hand through the ABI,
against dense numpy algebra on the oracle's matrices."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sleipnir_b200 as sb  # noqa: E402
from oracle.pyoracle import OracleProblem  # noqa: E402

name, N = sys.argv[1], int(sys.argv[2])
P, O = sb.Problem(name, N), OracleProblem(name, N)
O.eval_setup()
d_f, d_ce, d_ci = O.scaling()
n, me = P.n, P.me
D = P.open_device()
D.set_scaling(d_f, d_ce, d_ci)
x = O.initial_guess()
y = np.zeros(me)
e = np.zeros(0)
D.set_iterate(x, e, y, e)
pi = D.eval_current(1)
D.analyze()
delta, gamma = 1e-4, 1e-10
fi = D.factor(0.0, gamma, True)
print("inertia d=0:", fi.n_pos, fi.n_neg, fi.n_zero)
fi = D.factor(delta, gamma, True)
print("inertia d=1e-4:", fi.n_pos, fi.n_neg, fi.n_zero)
si = D.solve(0.1 * d_f, 0.99)
print("alpha_max", si.alpha_max, si.alpha_z)
px, py = D.download(sb.ARR_P_X), D.download(sb.ARR_P_Y)

H, Ae = O.H(x, y, e), O.A_e(x)
Hs = sp.csc_matrix((H.val, H.rowidx, H.colptr), shape=(n, n))
Hs = Hs + sp.tril(Hs, -1).T
Aes = sp.csc_matrix((Ae.val, Ae.rowidx, Ae.colptr), shape=(me, n))
K = sp.bmat([[Hs + delta * sp.eye(n), Aes.T], [Aes, -gamma * sp.eye(me)]], format="csc")
lu = spl.splu(K)
g, ce = O.g(x), O.c_e(x)
rhs = np.concatenate([-g + Aes.T @ y, -ce])
p = lu.solve(rhs)
rel = lambda a, b: np.abs(a - b).max() / max(1.0, np.abs(b).max())
print("px", rel(px, p[:n]), "py", rel(py, -p[n:]))
ti = D.trial(1.0, 1.0)
tce = O.c_e(x + p[:n])
print("viol cur", pi.ce_l1, np.abs(ce).sum(), "trial", ti.ce_l1, np.abs(tce).sum())
D.soc_begin()
ce_soc = ce.copy()
a_soc = 1.0
tce_dev = D.download(sb.ARR_TRIAL_C_E)
for it in range(3):
    ssi = D.soc_iterate(0.1 * d_f, 0.99, a_soc)
    ce_soc = a_soc * ce_soc + tce
    rhs[n:] = -ce_soc
    ps = lu.solve(rhs)
    spx = D.download(sb.ARR_P_X) if not hasattr(sb, "ARR_SOC_P_X") else D.download(sb.ARR_SOC_P_X)
    ti = D.trial(ssi.alpha_max, ssi.alpha_max, 1, 0)
    tce = O.c_e(x + ps[:n])
    print("soc", it, "alpha", ssi.alpha_max, "dev viol", ti.ce_l1, "ref viol", np.abs(tce).sum(),
          "f", ti.f, O.f(x + ps[:n]))
