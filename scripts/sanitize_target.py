"""Target of the compute-sanitizer runs (scripts/sanitize.sh): smoke(), then a
cart-pole N=300 solve that exercises the dependency-driven tree kernels
(k_factor_tree incl. the speculated pair of slpb_factor_pair, k_solve_tree) and
the TMA-streamed autodiff sweep for a few dozen Newton iterations, then two
minimum-time problems whose dense row gives fronts above order 32 / 158 (the
hybrid path and the global-memory workspace)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as g  # noqa: E402
import sleipnir_b200 as sb  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 25
g.smoke()
P = sb.Problem("cart_pole", 300)
st = P.solve(max_iterations=iters)
tr = P.trace()
c = P.counters()
print(f"[sanitize] cart-pole N=300: {sb.EXIT_STATUS[st]} after {len(tr)} iterations, "
      f"{c['factorizations']} factorisations, {c['solves']} solves, "
      f"{c['kernel_launches']} launches")
P.close()
# fronts above order 32 (hybrid tree + block kernels) and above one block's
# shared memory (global workspace)
for N in (50, 100):
    P = sb.Problem("differential_drive_ocp", N)
    st = P.solve(max_iterations=min(iters, 6))
    sym = P.symbolic_stats()
    print(f"[sanitize] differential-drive minimum time N={N}: {sb.EXIT_STATUS[st]} after "
          f"{len(P.trace())} iterations, max front {sym['max_front']}")
    P.close()
print("[sanitize] done")
