import sys, time, os
sys.path.insert(0, ".")
os.environ["SLPB_DESTROY_TIMING"] = "1"
import sleipnir_b200 as sb
for rep in range(4):
    P = sb.Problem("cart_pole", 5000)
    t1 = time.perf_counter(); st = P.solve(max_iterations=1200); t2 = time.perf_counter()
    ph = P.phase_seconds()
    print(f"rep {rep}: solve {t2-t1:.3f}s iters {len(P.trace())} " + " ".join(f"{k}={v:.3f}" for k, v in ph.items()), flush=True)
    P.close()
