"""Times the batched LDLᵀ (slpb_batch_factor / slpb_batch_solve) for several
batch sizes on the cart-pole KKT system:
    python scripts/batch_bench.py [N=5000] [B ...]
(needs a B200: run through gpurun)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import os as _os
_os.environ.setdefault("SLPB_TIMER_EVERY", "1")  # time every launch
import sleipnir_b200 as sb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
sizes = [int(a) for a in sys.argv[2:]] or [32, 64, 128, 256]
P = sb.Problem("cart_pole", N)
D = P.open_device()
rng = np.random.default_rng(0)
x = P.initial_guess() + 0.01 * rng.standard_normal(P.n)
D.set_iterate(x, 0.5 + rng.random(P.mi), 0.1 * rng.standard_normal(P.me), 0.5 + rng.random(P.mi))
D.eval_current(1)
st = D.analyze()
fi = D.factor(1.0, 1e-6, True)
D.solve(0.1, 0.99)
kkt, rhs = D.download(sb.ARR_KKT_VAL), D.download(sb.ARR_RHS)
sol1 = np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)])
f1, s1 = D.last_device_ms(3), D.last_device_ms(4)
fac_bytes = 12 * st.nnz_kkt + 12 * st.nnz_l + 8 * st.dim
sol_bytes = 2 * 12 * st.nnz_l + 8 * st.dim + 4 * 8 * st.dim
peak = 6547.8
print(f"N={N} dim={st.dim} nnz(K)={st.nnz_kkt} nnz(L)={st.nnz_l}; algorithmic bytes "
      f"factor {fac_bytes / 1e6:.2f} MB, solve {sol_bytes / 1e6:.2f} MB per instance")
print(f"single instance: factor {f1:.3f} ms ({fac_bytes / f1 / 1e6:.1f} GB/s, "
      f"{fac_bytes / f1 / 1e6 / peak:.4f}), full solve {s1:.3f} ms")
for Bn in sizes:
    B = sb.Batch(D, Bn)
    for i in range(Bn):
        B.set_system(i, kkt * (1 + 1e-3 * rng.standard_normal(kkt.size)) if i else kkt, rhs)
    best_f = best_s = 1e9
    for rep in range(4):
        info = B.factor(1.0, 1e-6)
        B.solve()
        f, s = B.last_ms()
        best_f, best_s = min(best_f, f), min(best_s, s)
    ok = np.array_equal(B.get(0), sol1)
    pe, ue = B.stored_entries()
    gf = Bn * fac_bytes / best_f / 1e6
    gs = Bn * sol_bytes / best_s / 1e6
    gt = Bn * (fac_bytes + sol_bytes) / (best_f + best_s) / 1e6
    print(f"B={Bn:4d}: factor {best_f:8.3f} ms {gf:7.1f} GB/s ({gf / peak:.3f})  "
          f"solve {best_s:8.3f} ms {gs:7.1f} GB/s ({gs / peak:.3f})  "
          f"factor+solve {gt:7.1f} GB/s ({gt / peak:.3f})  per-instance "
          f"{(best_f + best_s) / Bn * 1e3:.1f} us  inertia ok "
          f"{all((i.n_pos, i.n_neg) == (fi.n_pos, fi.n_neg) for i in info)} "
          f"instance0 bit-identical {ok}", flush=True)
    B.close()
