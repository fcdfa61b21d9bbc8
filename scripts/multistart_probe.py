"""Development aid: aggregate Newton steps/s of B concurrent solves on one GPU
(slp::multistart, one host thread + CUDA stream per start).
Usage: python scripts/multistart_probe.py [N] [B ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sleipnir_b200 as sb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
Bs = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
P = sb.Problem("cart_pole", N)
n = P.n
P.solve(max_iterations=20)      # warm the context, kernels, pools
P.close()
for B in Bs:
    r = sb.multistart("cart_pole", N, [5.0] * B, n_vars=n)
    its = sum(s[2] for s in r["starts"])
    print(f"N={N} B={B}: wall {r['wall_s']:.3f} s, {its} iterations, "
          f"{its / r['wall_s']:.0f} steps/s aggregate, per-start "
          f"{min(s[3] for s in r['starts']):.3f}..{max(s[3] for s in r['starts']):.3f} s, "
          f"status {sorted(set(s[0] for s in r['starts']))}", flush=True)
