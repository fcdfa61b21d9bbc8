import sys
sys.path.insert(0, ".")
import numpy as np
import sleipnir_b200 as sb
from oracle.pyoracle import OracleProblem, EXIT_STATUS
name, N = "cart_pole", int(sys.argv[1]) if len(sys.argv) > 1 else 15
P = sb.Problem(name, N); st = P.solve(max_iterations=400, keep_iterates=True); tr = P.trace()
O = OracleProblem(name, N); so = O.solve(max_iterations=400, keep_iterates=True); to = O.trace()
print("gpu", sb.EXIT_STATUS[st], len(tr), "oracle", EXIT_STATUS[so], len(to))
def first_fr(t):
    for i, r in enumerate(t):
        if r.type == 1: return i
    return len(t)
ig, io = first_fr(tr), first_fr(to)
print("first restoration row: gpu", ig, "oracle", io)
for k in range(max(0, min(ig, io) - 3), min(len(tr), len(to), max(ig, io) + 25)):
    a, b = tr[k], to[k]
    print(f"{k:4d} g[t{a.type} err {a.error:10.3e} cost {a.cost:11.4e} inf {a.infeasibility:9.2e} mu {a.mu:8.2e} d {a.delta:8.2e} g {a.gamma:7.1e} a {a.alpha:8.2e} f{a.factorizations} t{a.trials}] "
          f"o[t{b.type} err {b.error:10.3e} cost {b.cost:11.4e} inf {b.infeasibility:9.2e} mu {b.mu:8.2e} d {b.delta:8.2e} g {b.gamma:7.1e} a {b.alpha:8.2e} f{b.factorizations} t{b.trials}]")
