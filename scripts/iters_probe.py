import sys, time; sys.path.insert(0,'/root/repo')
import sleipnir_b200 as sb
for N in (1000, 5000):
    P = sb.Problem("cart_pole", N)
    t=time.time(); st = P.solve(max_iterations=3000); tr = P.trace()
    print(N, sb.EXIT_STATUS[st], "iters", len(tr), "loop %.3fs total %.2fs"%(P.loop_seconds(), time.time()-t), "last err %.3e alpha %.3e delta %.3e"%(tr[-1].error, tr[-1].alpha, tr[-1].delta), flush=True)
    P.close()
