"""Per-level latency breakdown of the dependency-driven factor kernel
(SLPB_TREE_DEBUG=1)."""
import ctypes as C, os, sys
import numpy as np
os.environ["SLPB_TREE_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import os as _os
_os.environ.setdefault("SLPB_TIMER_EVERY", "1")  # time every launch
import sleipnir_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
P = sb.Problem("cart_pole", N)
D = P.open_device()
rng = np.random.default_rng(0)
D.set_iterate(P.initial_guess(), np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
D.eval_current(1)
st = D.analyze()
for rep in range(3):
    D.factor(1.0, 1e-6, True)
print("factor ms", D.last_device_ms(3), "levels", st.n_levels, "supers", st.n_supernodes)
L = sb.device_lib()
ns = st.n_supernodes
stamps = np.zeros(4 * ns, dtype=np.uint64); level = np.zeros(ns, dtype=np.int32); parent = np.zeros(ns, dtype=np.int32)
L.slpb_debug_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
rc = L.slpb_debug_tree(D.raw, stamps.ctypes.data, level.ctypes.data, parent.ctypes.data)
assert rc == 0
t = stamps.reshape(ns, 4).astype(np.int64); t0 = t[:, 0].min(); t -= t0
print("total span us", t[:, 3].max() / 1e3)
child_done = np.zeros(ns, dtype=np.int64)
for s in range(ns):
    if parent[s] >= 0: child_done[parent[s]] = max(child_done[parent[s]], t[s, 3])
for lv in range(st.n_levels):
    m = level == lv
    sig = (t[m, 1] - np.maximum(child_done[m], t[m, 0]))
    print(f"level {lv:2d} n={m.sum():5d} ticket {t[m,0].min()/1e3:8.1f}..{t[m,0].max()/1e3:8.1f} us | ready-after-children median {np.median(sig)/1e3:6.2f} max {sig.max()/1e3:6.2f} us | extend-add median {np.median(t[m,2]-t[m,1])/1e3:6.2f} max {(t[m,2]-t[m,1]).max()/1e3:6.2f} us | elimination+write-out median {np.median(t[m,3]-t[m,2])/1e3:6.2f} max {(t[m,3]-t[m,2]).max()/1e3:6.2f} us | pre-wait work median {np.median(t[m,1]-t[m,0])/1e3:6.2f} | done at {t[m,3].max()/1e3:8.1f}")
