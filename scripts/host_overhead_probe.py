"""Where the wall time of a Newton iteration goes besides the kernels: the same
200 iterations of cart-pole N=5000 with and without the per-phase device
timers (SLPB_NO_TIMERS=1), wall time per iteration against the device time of
the timed kernel groups."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sleipnir_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
for rep in range(3):
    P = sb.Problem("cart_pole", N)
    st = P.solve(max_iterations=205)
    tr = P.trace()
    tim = P.timers()
    dev = sum(v["mean_ms"] * v["launches"] for v in tim.values()) / max(len(tr), 1)
    cnt = P.counters()
    print(f"rep {rep}: {len(tr)} iterations, wall {1e3 * P.loop_seconds() / len(tr):.4f} ms/iteration, "
          f"timed device groups {dev:.4f} ms/iteration, launches/iteration {cnt['kernel_launches'] / len(tr):.1f}",
          flush=True)
    P.close()
