import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sleipnir_b200 as sb
P = sb.Problem("cart_pole", 5000)
st = P.solve()
tr = P.trace()
f = np.array([r.factorizations for r in tr]); d = np.array([r.delta for r in tr])
print(sb.EXIT_STATUS[st], len(tr), "hist", np.bincount(f))
prev = np.concatenate([[0.0], d[:-1]])
launch = np.where(prev == 0, f, np.maximum(1, f - 1))
print("launches/iter", launch.mean(), "factorizations/iter", f.mean())
print("prev==0 & f>=2:", np.sum((prev == 0) & (f >= 2)), "prev!=0 & f>=3:", np.sum((prev != 0) & (f >= 3)), "f>=4", np.sum(f >= 4))
i3 = np.where((prev != 0) & (f >= 3))[0]
print("ratio delta/prev where f>=3:", np.unique(np.round(d[i3] / prev[i3], 3), return_counts=True))
print("always-pair", np.maximum(1, f - 1).mean(), "triple", np.maximum(1, f - 2).mean())
