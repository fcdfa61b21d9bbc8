"""Development aid: device path vs oracle on problems that enter feasibility
restoration."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import sleipnir_b200 as sb
from oracle.pyoracle import OracleProblem, EXIT_STATUS
cases = [("cart_pole", N) for N in (15, 30, 40, 50, 80, 20, 10, 200)]
for name, N in cases:
    P = sb.Problem(name, N)
    t = time.time(); st = P.solve(max_iterations=3000); tr = P.trace(); dt = time.time() - t
    O = OracleProblem(name, N); so = O.solve(max_iterations=3000, keep_iterates=False); to = O.trace()
    xg, xo = P.solution()[0], O.solution()[0]
    print(f"{name} N={N:4d} gpu {sb.EXIT_STATUS[st]:32s} iters {len(tr):5d} (restoration {sum(r.type == 1 for r in tr):4d}) {dt:6.2f}s | "
          f"oracle {EXIT_STATUS[so]:32s} iters {len(to):5d} (restoration {sum(r.type == 1 for r in to):4d}) | max|dx| {np.abs(xg - xo).max():.2e}", flush=True)
    P.close(); O.close()
