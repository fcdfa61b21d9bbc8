"""Device solves of the BASELINE configs to their exit status, with the default
(nested dissection) order and the reference's AMD order, beside the oracle's
goldens (tests/golden/converge_*.npz). Needs a B200."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sleipnir_b200 as sb  # noqa: E402

cases = [a.split(":") for a in sys.argv[1:]] or [["cart_pole", "300", "nd"]]
for case in cases:
    name, N, order = case[:3]
    T = float(case[3]) if len(case) > 3 else 0.0
    N = int(N)
    P = sb.Problem(name, N, T)
    t0 = time.perf_counter()
    st = P.solve(ordering=sb.ORDER_AMD if order == "amd" else sb.ORDER_NESTED_DISSECTION)
    dt = time.perf_counter() - t0
    tr = P.trace()
    rest = [i for i, r in enumerate(tr) if r.type == 1]
    g = os.path.join(ROOT, "tests", "golden", f"converge_{name}_{N}.npz")
    gold = ""
    if os.path.exists(g):
        G = np.load(g)
        gold = (f" | oracle: {sb.EXIT_STATUS[int(G['status'])]} after {int(G['iterations'])} "
                f"(restoration entry {int(G['restoration_entry'])}), cost {float(G['final_cost']):.9g}")
    print(f"{name} N={N} T={T or 'default'} {order}: {sb.EXIT_STATUS[st]} after {len(tr)} iterations "
          f"({len(rest)} in restoration, entry {rest[0] if rest else -1}), cost "
          f"{tr[-1].cost:.9g}, error {tr[-1].error:.3g}, {dt:.1f} s{gold}", flush=True)
    P.close()
