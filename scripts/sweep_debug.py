"""Phase timeline of one thread block of the full re-linearisation sweep
(k_ad_sweep, 512 threads). Needs a development build of the device library:

    make -C sleipnir_b200 -B EXTRA=-DSLPB_SWEEP_STAMPS

Prints cycles between the phase boundaries of block 0: tables → leaves → every
super-level (forward / value outputs / reverse) → adjoint outputs."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SLPB_SERIAL_SWEEPS"] = "1"  # no concurrent launch writes the stamps
import sleipnir_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
P = sb.Problem("cart_pole", N)
D = P.open_device()
D.set_iterate(P.initial_guess(), np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
for rep in range(5):
    D.eval_current(1)
print("eval_full ms", D.last_device_ms(0))
L = sb.device_lib()
out = np.zeros(128, dtype=np.int64)
L.slpb_debug_sweep_stamps.argtypes = [C.c_void_p]
rc = L.slpb_debug_sweep_stamps(out.ctypes.data)
assert rc == 0, "library was not built with -DSLPB_SWEEP_STAMPS"
for name, o in (("block 0 (shares its SM with a second task when tasks > SMs)", out[:64]),
                ("block 75 (alone on its SM)", out[64:])):
    if o[0] == 0:
        continue
    t = o - o[0]
    last = max(k for k in range(3, 62) if o[k] != 0)
    print(name)
    print(f"  tables in smem      {t[1]:8d} cycles")
    print(f"  leaves + constants  {t[2] - t[1]:8d}")
    prev = t[2]
    row = []
    for k in range(3, last + 1):
        row.append(f"{t[k] - prev}")
        prev = t[k]
    print("  super-levels        " + " ".join(row))
    print(f"  adjoint outputs     {t[62] - prev:8d}")
    print(f"  total               {t[62]:8d} cycles = {t[62] / 1.965e3:.1f} us at 1965 MHz")
