import sys, time
sys.path.insert(0, "/root/repo")
import sleipnir_b200 as sb
for name, N in (("differential_drive_ocp", 50), ("differential_drive_ocp", 70), ("differential_drive_ocp", 200), ("cart_pole_ocp", 100)):
    P = sb.Problem(name, N)
    st = P.solve(); tr = P.trace(); sym = P.symbolic_stats(); tim = P.timers()
    print(name, N, sb.EXIT_STATUS[st], len(tr), "iterations,", round(1e3 * P.loop_seconds() / len(tr), 4), "ms/iteration; max front", sym["max_front"], "levels", sym["n_levels"], "factor ms", round(tim["factor"]["mean_ms"], 4), "solve ms", round(tim["solve"]["mean_ms"], 4))
    P.close()
