"""Worker of tests/test_gpu_parity.py::test_sharded_two_gpu (one process per
GPU under torch.distributed.run): both ranks solve the same cart-pole problem
with the re-linearisation sweep, the factorisation and the triangular solves
sharded over the ranks (own subtrees eliminated locally, one small
all-gather of the subtree roots, the top of the tree replicated, solution
pieces gathered) and compare with an unsharded solve on their own GPU — the
iterates must be bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sleipnir_b200 as sb  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 25

box = [sb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)

ref = sb.Problem("cart_pole", N)
ref.solve(max_iterations=iters, device=local, keep_iterates=True)
tr_ref = ref.trace()

P = sb.Problem("cart_pole", N)
P.set_comm(rank, world, box[0])
P.solve(max_iterations=iters, device=local, keep_iterates=True)
tr = P.trace()
assert len(tr) == len(tr_ref) == iters
for a, b in zip(tr, tr_ref):
    assert a.delta == b.delta and a.alpha == b.alpha
    assert np.array_equal(a.x, b.x) and np.array_equal(a.z, b.z)
# all ranks hold the same iterate
x = torch.from_numpy(P.solution()[0].copy())
xs = [torch.zeros_like(x) for _ in range(world)]
dist.all_gather(xs, x)
assert all(torch.equal(xs[0], t) for t in xs)
t_ref = tr_ref[-1].t_end - tr_ref[4].t_end
t_sh = tr[-1].t_end - tr[4].t_end
cs = P.comm_stats()
assert cs["subtree_roots"]["count"] > 0 and cs["solution"]["count"] > 0, cs
if rank == 0:
    per = {k: (v["bytes"] / max(v["count"], 1), 1e3 * v["total_ms"] / max(v["count"], 1))
           for k, v in cs.items()}
    print(f"SHARDED_OK N={N} world={world} iterations={iters} "
          f"ms/step single={1e3 * t_ref / (iters - 5):.3f} sharded={1e3 * t_sh / (iters - 5):.3f} "
          + " ".join(f"{k}: {b / 1024:.1f} KiB {us:.0f} us" for k, (b, us) in per.items()))
dist.barrier()
dist.destroy_process_group()
