"""ctypes loader for tests/emu/libslpb_emu.so (host emulation of the device
library's logic; test infrastructure only)."""
import ctypes as C
import os

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                "emu", "libslpb_emu.so"))
        vp = C.c_void_p
        L.emu_create.restype = vp
        L.emu_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double]
        L.emu_destroy.argtypes = [vp]
        L.emu_set_shard_world.argtypes = [vp, C.c_int]
        L.emu_error.restype = C.c_char_p
        L.emu_error.argtypes = [vp]
        L.emu_dims.argtypes = [vp, _ip, _ip, _ip]
        L.emu_initial_guess.argtypes = [vp, _dp]
        L.emu_stats.argtypes = [vp, _lp]
        L.emu_pattern.argtypes = [vp, C.c_int, _ip, _ip, _lp, _ip, _ip]
        L.emu_eval.argtypes = [vp] + [_dp] * 3 + [C.c_double] + [_dp] * 9
        L.emu_kkt_build.restype = C.c_int64
        L.emu_kkt_build.argtypes = [vp]
        L.emu_kkt_pattern.argtypes = [vp, _ip, _ip]
        L.emu_kkt_assemble.argtypes = [vp, _dp, _dp]
        L.emu_analyze.restype = C.c_int
        L.emu_analyze.argtypes = [vp, C.c_int, _ip, _lp]
        L.emu_get_perm.argtypes = [vp, _ip]
        L.emu_fronts.argtypes = [vp, _ip, _ip, _ip, _ip]
        L.emu_tree_shard.restype = C.c_int
        L.emu_tree_shard.argtypes = [vp, C.c_int, _ip, _dp]
        L.emu_order_amd.argtypes = [C.c_int, _ip, _ip, _ip]
        L.emu_factor.restype = C.c_double
        L.emu_factor.argtypes = [vp, C.c_double, C.c_double, _ip, _dp]
        L.emu_solve.argtypes = [vp, _dp, _dp]
        L.emu_set_fused.argtypes = [vp, C.c_int]
        L.emu_set_kkt_values.argtypes = [vp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def order_amd(n, colptr, rowidx):
    """The product's AMD ordering (csrc/amd.cpp) of a lower-triangular pattern."""
    cp = np.ascontiguousarray(colptr, dtype=np.int32)
    ri = np.ascontiguousarray(rowidx, dtype=np.int32)
    p = np.zeros(n, dtype=np.int32)
    lib().emu_order_amd(n, _i(cp), _i(ri), _i(p))
    return p


class Emu:
    def __init__(self, name, N=0, p0=0.0, p1=0.0):
        self.L = lib()
        self.h = self.L.emu_create(name.encode(), N, p0, p1)
        assert self.h, f"unknown problem {name}"
        err = self.L.emu_error(self.h)
        assert not err, err
        n, me, mi = C.c_int(), C.c_int(), C.c_int()
        self.L.emu_dims(self.h, n, me, mi)
        self.n, self.me, self.mi = n.value, me.value, mi.value

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h)
            self.h = None

    def set_shard_world(self, world):
        """Emulate the derivative sweep as `world` ranks would run it."""
        self.L.emu_set_shard_world(self.h, world)

    def stats(self):
        out = (C.c_int64 * 12)()
        self.L.emu_stats(self.h, out)
        keys = ("tape_nodes", "value_clusters", "value_programs", "value_words",
                "deriv_clusters", "deriv_programs", "deriv_words", "max_smem",
                "instr", "visits", "contribs", "deriv_stage")
        return dict(zip(keys, out))

    def pattern(self, which):
        r, c, nnz = C.c_int(), C.c_int(), C.c_int64()
        self.L.emu_pattern(self.h, which, r, c, nnz, None, None)
        cp = np.zeros(c.value + 1, dtype=np.int32)
        ri = np.zeros(max(nnz.value, 1), dtype=np.int32)
        self.L.emu_pattern(self.h, which, r, c, nnz, _i(cp), _i(ri))
        return r.value, c.value, cp, ri[:nnz.value]

    def eval(self, x, y, z, d_f, d_ce, d_ci):
        n, me, mi = self.n, self.me, self.mi
        pad = lambda v, m: np.ascontiguousarray(v if m else np.zeros(1), dtype=np.float64)
        sizes = {w: len(self.pattern(w)[3]) for w in (5, 7, 3)}
        f = C.c_double()
        ce, ci, g = np.zeros(max(me, 1)), np.zeros(max(mi, 1)), np.zeros(n)
        ae, ai, hv = (np.zeros(max(sizes[5], 1)), np.zeros(max(sizes[7], 1)),
                      np.zeros(max(sizes[3], 1)))
        self.L.emu_eval(self.h, _d(np.ascontiguousarray(x)), _d(pad(y, me)),
                        _d(pad(z, mi)), d_f, _d(pad(d_ce, me)), _d(pad(d_ci, mi)),
                        C.byref(f), _d(ce), _d(ci), _d(g), _d(ae), _d(ai), _d(hv))
        return dict(f=f.value, c_e=ce[:me], c_i=ci[:mi], g=g, A_e=ae[:sizes[5]],
                    A_i=ai[:sizes[7]], H=hv[:sizes[3]])

    def kkt(self, sigma):
        nnz = self.L.emu_kkt_build(self.h)
        dim = self.n + self.me
        cp = np.zeros(dim + 1, dtype=np.int32)
        ri = np.zeros(nnz, dtype=np.int32)
        self.L.emu_kkt_pattern(self.h, _i(cp), _i(ri))
        kv = np.zeros(nnz)
        self.L.emu_kkt_assemble(self.h, _d(np.ascontiguousarray(sigma)), _d(kv))
        return cp, ri, kv

    def analyze(self, ordering=0, perm=None):
        out = (C.c_int64 * 8)()
        p = None if perm is None else _i(np.ascontiguousarray(perm, dtype=np.int32))
        rc = self.L.emu_analyze(self.h, ordering, p, out)
        assert rc == 0, self.L.emu_error(self.h)
        keys = ("dim", "nnz_l", "n_super", "n_levels", "max_front",
                "etree_height", "panel", "update")
        return dict(zip(keys, out))

    def fronts(self, n_super):
        """(F, np, level, parent) of every front of the last analysis."""
        a = [np.zeros(n_super, dtype=np.int32) for _ in range(4)]
        self.L.emu_fronts(self.h, *[_i(v) for v in a])
        return a

    def tree_shard(self, world, n_super):
        """(n_top, owner per front, work per rank + [top])."""
        owner = np.zeros(n_super, dtype=np.int32)
        work = np.zeros(world + 1)
        n_top = self.L.emu_tree_shard(self.h, world, _i(owner), _d(work))
        return n_top, owner, work

    def perm(self):
        p = np.zeros(self.n + self.me, dtype=np.int32)
        self.L.emu_get_perm(self.h, _i(p))
        return p

    def set_kkt_values(self, kv):
        """Overwrites the assembled lhs values (after kkt())."""
        self.L.emu_set_kkt_values(self.h, _d(np.ascontiguousarray(kv, dtype=np.float64)))

    def set_fused(self, fused=True):
        """Arithmetic mode of factor(): False = the reference's separate
        rounding, True = fused Schur updates (SLPB_ARITH_TENSOR)."""
        self.L.emu_set_fused(self.h, int(fused))

    def factor(self, delta, gamma):
        info = (C.c_int * 4)()
        D = np.zeros(self.n + self.me)
        mn = self.L.emu_factor(self.h, delta, gamma, info, _d(D))
        return tuple(info), mn, D

    def solve(self, rhs):
        x = np.zeros(self.n + self.me)
        self.L.emu_solve(self.h, _d(np.ascontiguousarray(rhs)), _d(x))
        return x
