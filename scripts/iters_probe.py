"""Exit status and iteration count of the cart-pole swing-up (T = 5 s) per
horizon length, default options — which instances the globalisation gets
through depends on rounding-level details of the factorisation."""
import sys, time; sys.path.insert(0,'/root/repo')
import sleipnir_b200 as sb
Ns = [int(a) for a in sys.argv[1:]] or [300, 1000, 2000, 3000, 5000]
for N in Ns:
    P = sb.Problem("cart_pole", N)
    t=time.time(); st = P.solve(max_iterations=3000); tr = P.trace()
    rest = sum(1 for r in tr if r.type == 1)
    print(N, sb.EXIT_STATUS[st], "iters", len(tr), "restoration", rest, "loop %.3fs total %.2fs"%(P.loop_seconds(), time.time()-t), "last err %.3e alpha %.3e delta %.3e"%(tr[-1].error, tr[-1].alpha, tr[-1].delta), flush=True)
    P.close()
