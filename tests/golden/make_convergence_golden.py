"""Generates tests/golden/converge_*.npz: the reference algorithm (oracle on the
reference's OWN autodiff core, AMD ordering like Eigen::SimplicialLDLT, default
Options) run to its exit status on the BASELINE.json configurations. Run in
the authoring container only (minutes per case):

    python tests/golden/make_convergence_golden.py [name:N[:T] ...]

Stored per case: exit status, iteration count, the iteration at which
feasibility restoration was entered (−1: never), the scalar trace (error, cost,
infeasibility, δ, α, μ, factorisations, iteration type) and the final primal
iterate. The GPU tests compare the driver-visible end-to-end solve with these.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import EXIT_STATUS, OracleProblem, have_reference  # noqa: E402

CASES = [("cart_pole", 300), ("cart_pole", 1000), ("cart_pole", 5000),
         ("gfold", 2000)]


def run(name, N, T=0.0, max_iterations=5000):
    backend = "reference" if have_reference() else "restated"
    P = OracleProblem(name, N, T, backend=backend)
    t0 = time.perf_counter()
    st = P.solve(max_iterations=max_iterations, keep_iterates=False)
    dt = time.perf_counter() - t0
    tr = P.trace()
    x, s, y, z = P.solution()
    types = np.array([r.type for r in tr], dtype=np.int8)
    rest = np.flatnonzero(types == 1)
    out = dict(status=st, iterations=len(tr),
               restoration_entry=int(rest[0]) if rest.size else -1,
               restoration_iterations=int(rest.size),
               x=x, final_cost=tr[-1].cost if tr else np.nan,
               final_error=tr[-1].error if tr else np.nan,
               final_infeasibility=tr[-1].infeasibility if tr else np.nan,
               error=np.array([r.error for r in tr]),
               cost=np.array([r.cost for r in tr]),
               infeasibility=np.array([r.infeasibility for r in tr]),
               alpha=np.array([r.alpha for r in tr]),
               delta=np.array([r.delta for r in tr]),
               mu=np.array([r.mu for r in tr]),
               factorizations=np.array([r.factorizations for r in tr], dtype=np.int16),
               type=types, backend=backend, seconds=dt)
    P.close()
    print(f"{name} N={N}: {EXIT_STATUS[st]} after {len(tr)} iterations "
          f"({rest.size} in restoration, entry {out['restoration_entry']}), "
          f"cost {out['final_cost']:.9g}, {dt:.1f} s", flush=True)
    return out


if __name__ == "__main__":
    cases = [c + (0.0,) for c in CASES]
    if len(sys.argv) > 1:   # name:N[:T]  (T: horizon in seconds, default the config's)
        cases = []
        for a in sys.argv[1:]:
            f = a.split(":")
            cases.append((f[0], int(f[1]), float(f[2]) if len(f) > 2 else 0.0))
    for name, N, T in cases:
        tag = f"{name}_{N}" + (f"_T{T:g}" if T else "")
        np.savez_compressed(os.path.join(HERE, f"converge_{tag}.npz"),
                            **run(name, N, T))
