import sys
sys.path.insert(0, ".")
import numpy as np
import sleipnir_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xs = []
for rep in range(3):
    P = sb.Problem("cart_pole", N)
    st = P.solve(max_iterations=int(sys.argv[2]) if len(sys.argv) > 2 else 5000)
    tr = P.trace()
    xs.append(P.solution()[0].copy())
    print(rep, sb.EXIT_STATUS[st], len(tr), "restoration", sum(r.type == 1 for r in tr), "fact", sum(r.factorizations for r in tr), "x checksum %.17g" % xs[-1].sum(), flush=True)
    P.close()
print("identical:", all(np.array_equal(xs[0], x) for x in xs))
