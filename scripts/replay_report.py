"""Prints the one-step replay table of tests/replay.py for one config:
    python scripts/replay_report.py cart_pole 5000 12 nd|amd
(needs a B200: run through gpurun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sleipnir_b200 as sb  # noqa: E402
from replay import format_report, replay  # noqa: E402

name, N, iters, order = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
t0 = time.perf_counter()
rows, report, sym = replay(name, N, iters,
                           sb.ORDER_AMD if order == "amd" else sb.ORDER_NESTED_DISSECTION)
print(f"== {name} N={N}, device ordering {order}: nnz(L)={sym['nnz_l']}, "
      f"{sym['n_levels']} levels, max front {sym['max_front']} "
      f"({time.perf_counter() - t0:.1f} s)")
print(format_report(report), flush=True)
