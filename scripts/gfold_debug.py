import sys
sys.path.insert(0, ".")
import numpy as np
import sleipnir_b200 as sb
from oracle.pyoracle import OracleProblem, EXIT_STATUS
name, N = "gfold", 20
P = sb.Problem(name, N); st = P.solve(max_iterations=1500); tr = P.trace()
D = P.open_device(); D.analyze(); perm = D.permutation(); P.close_device()
O = OracleProblem(name, N); so = O.solve(max_iterations=1500, keep_iterates=False, perm=perm, force_sparse=1); to = O.trace()
O2 = OracleProblem(name, N); so2 = O2.solve(max_iterations=1500, keep_iterates=False); to2 = O2.trace()
fr = lambda t: sum(r.type == 1 for r in t)
print("gpu", sb.EXIT_STATUS[st], len(tr), fr(tr), "| oracle same perm", EXIT_STATUS[so], len(to), fr(to), "| oracle amd", EXIT_STATUS[so2], len(to2), fr(to2))
k0 = next((i for i, r in enumerate(tr) if r.type == 1), len(tr))
for k in list(range(0, 6)) + list(range(max(0, k0 - 6), min(len(tr), k0 + 6))):
    a = tr[k]; b = to[k] if k < len(to) else None
    line = f"{k:4d} g[t{a.type} err {a.error:10.3e} inf {a.infeasibility:9.2e} mu {a.mu:8.2e} d {a.delta:8.2e} a {a.alpha:8.2e} f{a.factorizations} t{a.trials}]"
    if b: line += f" o[t{b.type} err {b.error:10.3e} inf {b.infeasibility:9.2e} mu {b.mu:8.2e} d {b.delta:8.2e} a {b.alpha:8.2e} f{b.factorizations} t{b.trials}]"
    print(line)
