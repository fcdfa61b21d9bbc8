"""Development aid: per-iteration comparison of the device SQP/Newton branches
with the oracle (same permutation). Usage: python scripts/sqp_probe.py name N"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sleipnir_b200 as sb  # noqa: E402
from oracle.pyoracle import EXIT_STATUS, OracleProblem  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max())) if len(b) else 0.0


name, N = sys.argv[1], int(sys.argv[2])
P = sb.Problem(name, N)
st = P.solve(keep_iterates=True)
tr = P.trace()
perm = None
if P.n + P.me > 1:
    D = P.open_device(); D.analyze(); perm = D.permutation(); P.close_device()
O = OracleProblem(name, N)
so = O.solve(perm=perm, force_sparse=1)
to = O.trace()
print(name, N, P.solver_kind(), sb.EXIT_STATUS[st], EXIT_STATUS[so], len(tr), len(to))
for a, b in list(zip(tr, to))[:60]:
    print(a.iteration, a.factorizations, b.factorizations, a.trials, b.trials,
          a.delta, b.delta, a.alpha, b.alpha, "%.3e %.3e" % (a.error, b.error),
          "dx %.2e dy %.2e" % (rel(a.x, b.x), rel(a.y, b.y)))
