"""World-size-2 gloo run of the multi-rank bookkeeping bench.py uses (one
process per GPU, max-over-ranks timing, whole-job value = units of all ranks /
slowest rank's time). Each rank runs an independent replica of the path: the
interior-point loop has no data-path collective in this round."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from bench import steady_rate
    from oracle.pyoracle import OracleProblem
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # each rank walks its own replica on the CPU checker (no GPU here)
    P = OracleProblem("cart_pole", 10 + 5 * rank)
    P.solve(max_iterations=8, keep_iterates=False)
    k, dt = steady_rate(P.trace(), 3, 5)
    t = torch.tensor([dt], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ks = torch.tensor([float(k)], dtype=torch.float64)
    dist.all_reduce(ks, op=dist.ReduceOp.SUM)
    assert t.item() >= dt and ks.item() == 5 * world
    if rank == 0:
        print("VALUE", ks.item() / t.item())
    dist.barrier()
    dist.destroy_process_group()
""") % ROOT


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
         str(port), str(script)],
        capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "VALUE" in res.stdout
