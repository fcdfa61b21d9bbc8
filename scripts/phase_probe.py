import sys, time
sys.path.insert(0, ".")
import sleipnir_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
for rep in range(2):
    t0 = time.perf_counter(); P = sb.Problem("cart_pole", N); t1 = time.perf_counter()
    st = P.solve(max_iterations=50); t2 = time.perf_counter()
    ph = P.phase_seconds()
    print(f"rep {rep}: construct {t1-t0:.3f}s solve {t2-t1:.3f}s loop {P.loop_seconds():.3f}s phases " + " ".join(f"{k}={v:.3f}" for k, v in ph.items()), flush=True)
    t3 = time.perf_counter(); P.close(); print(f"   close {time.perf_counter()-t3:.3f}s")
