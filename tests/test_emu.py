"""CPU-tier checks of the product's host logic and of the kernels' bodies run
on the host by tests/emu (one lane instead of a warp / thread block):
host DSL → tape → cluster programs → interpreter; KKT recipe; symbolic
analysis; multifrontal LDLT — each against the oracle and the golden vectors."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

import ctypes

from emu import Emu

emu_ip = ctypes.POINTER(ctypes.c_int)
from oracle.pyoracle import OracleProblem, ldlt

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PARAMS = {"rosenbrock_cubic_line": (0.3, 0.7), "rosenbrock_disk": (-0.5, 1.2)}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "eval_*.npz"))))
def test_interpreter_matches_reference_core_bit_for_bit(path):
    """Same libm on both sides here, so the product's DSL + compiler +
    interpreter must reproduce the reference core's numbers exactly."""
    g = np.load(path)
    name, N = os.path.basename(path)[5:-4].rsplit("_", 1)
    p0, p1 = PARAMS.get(name, (0, 0))
    E = Emu(name, int(N), p0, p1)
    out = E.eval(g["x"], g["y"], g["z"], float(g["d_f"]), g["d_ce"], g["d_ci"])
    assert out["f"] == g["f"]
    np.testing.assert_array_equal(out["c_e"], g["c_e"])
    np.testing.assert_array_equal(out["c_i"], g["c_i"])
    # One decision variable shared by every step (the OCPs' single variable time
    # step; the inputs of a single-shooting OCP) collects adjoint contributions
    # from all the per-step clusters: they are added cluster by cluster, not in
    # the reference's node order — a few entries differ in the last bit or two.
    shared = name in ("cart_pole_ocp", "differential_drive_ocp",
                      "flywheel_ocp_shooting")
    if shared:
        np.testing.assert_allclose(out["g"], g["g"], rtol=1e-14,
                                   atol=1e-15 * np.abs(g["g"]).max())
    elif name == "chained_rosenbrock":
        # A cost sum cut into terms adds each term's partial adjoint as a whole,
        # x + (c1 + c2), where the reference's sweep adds contribution by
        # contribution, (x + c1) + c2. The two agree bit for bit unless a leaf
        # collects SEVERAL contributions from a term the reference visits
        # later — here x₀, through pow(x₀, 2) = x₀·x₀ of the first term: 1 ulp.
        np.testing.assert_array_equal(out["g"][1:], g["g"][1:])
        np.testing.assert_allclose(out["g"][0], g["g"][0], rtol=2.3e-16)
    else:
        np.testing.assert_array_equal(out["g"], g["g"])
    for nm, which in (("A_e", 5), ("A_i", 7), ("H", 3)):
        _, _, cp, ri = E.pattern(which)
        np.testing.assert_array_equal(cp, g[nm + "_colptr"])
        np.testing.assert_array_equal(ri, g[nm + "_rowidx"])
        if shared:
            np.testing.assert_allclose(
                out[nm], g[nm + "_val"], rtol=1e-14,
                atol=1e-15 * np.abs(g[nm + "_val"]).max(initial=0.0))
        else:
            np.testing.assert_array_equal(out[nm], g[nm + "_val"])
    E.close()


def test_time_steps_share_one_program():
    """Every stage of a transcription compiles to the same position-independent
    program: the blob must not grow with the horizon."""
    a, b = Emu("cart_pole", 20), Emu("cart_pole", 40)  # both above the split threshold
    sa, sb_ = a.stats(), b.stats()
    assert sa["deriv_programs"] == sb_["deriv_programs"] <= 4
    assert sa["deriv_words"] == sb_["deriv_words"]
    assert sb_["deriv_clusters"] == 2 * 40  # dynamics + u_k² term per stage
    assert sb_["max_smem"] < 48 * 1024
    a.close(); b.close()


def test_long_cost_sum_is_split():
    """Σ_k (r − x_k)² with 301 terms: the row is split into independent terms
    and summed by the gather; values agree with the oracle to rounding."""
    N = 300
    E, O = Emu("flywheel", N), OracleProblem("flywheel", N)
    O.eval_setup()
    d_f, d_ce, d_ci = O.scaling()
    rng = np.random.default_rng(5)
    x = rng.standard_normal(E.n)
    y, z = rng.standard_normal(E.me), np.abs(rng.standard_normal(E.mi))
    out = E.eval(x, y, z, d_f, d_ce, d_ci)
    assert abs(out["f"] - O.f(x)) <= 1e-13 * abs(O.f(x))
    np.testing.assert_allclose(out["g"], O.g(x), rtol=1e-14, atol=0)
    np.testing.assert_array_equal(out["A_e"], O.A_e(x).val)
    np.testing.assert_array_equal(out["H"], O.H(x, y, z).val)
    assert E.stats()["value_clusters"] > N  # one cluster per term
    E.close(); O.close()


@pytest.mark.parametrize("name,N", [("cart_pole", 30), ("flywheel", 40),
                                    ("wachter_biegler", 0),
                                    # no inequality / no constraint blocks
                                    ("cart_pole_eq", 20), ("flywheel_eq", 30),
                                    ("chained_rosenbrock", 40)])
def test_kkt_assembly(name, N):
    E, O = Emu(name, N), OracleProblem(name, N)
    O.eval_setup()
    d_f, d_ce, d_ci = O.scaling()
    rng = np.random.default_rng(9)
    x = O.initial_guess() + 0.01 * rng.standard_normal(E.n)
    y = 0.1 * rng.standard_normal(E.me)
    z = 0.1 + np.abs(rng.standard_normal(E.mi))
    s = 0.1 + np.abs(rng.standard_normal(E.mi))
    E.eval(x, y, z, d_f, d_ce, d_ci)
    cp, ri, kv = E.kkt(z / s)
    n, me, mi = E.n, E.me, E.mi
    H, Ae, Ai = O.H(x, y, z), O.A_e(x), O.A_i(x)
    Hs = sp.csc_matrix((H.val, H.rowidx, H.colptr), shape=(n, n))
    Aes = sp.csc_matrix((Ae.val, Ae.rowidx, Ae.colptr), shape=(me, n))
    Ais = sp.csc_matrix((Ai.val, Ai.rowidx, Ai.colptr), shape=(mi, n))
    TL = Hs + sp.tril(Ais.T @ sp.diags(z / s) @ Ais)
    K = TL.tocsc() if me == 0 else \
        sp.bmat([[TL, None], [Aes, sp.csc_matrix((me, me))]], format="csc")
    Kemu = sp.csc_matrix((kv, ri, cp), shape=(n + me, n + me))
    assert abs(Kemu - K).max() <= 1e-12 * max(1.0, abs(K).max())
    # every diagonal entry is structurally present (pattern stability,
    # sparse_regularized_ldlt.hpp:65-67)
    for c in range(n + me):
        assert c in ri[cp[c]:cp[c + 1]]
    E.close(); O.close()


@pytest.mark.parametrize("ordering", [0, 2])
def test_multifrontal_ldlt_matches_oracle(ordering):
    E, O = Emu("cart_pole", 60), OracleProblem("cart_pole", 60)
    O.eval_setup()
    d_f, d_ce, d_ci = O.scaling()
    rng = np.random.default_rng(4)
    x = O.initial_guess() + 0.01 * rng.standard_normal(E.n)
    y = 0.1 * rng.standard_normal(E.me)
    z = 0.5 + np.abs(rng.standard_normal(E.mi))
    s = 0.5 + np.abs(rng.standard_normal(E.mi))
    E.eval(x, y, z, d_f, d_ce, d_ci)
    cp, ri, kv = E.kkt(z / s)
    st = E.analyze(ordering)
    perm = E.perm()
    assert sorted(perm) == list(range(E.n + E.me))
    dim = E.n + E.me
    delta, gamma = 1.0, 1e-6
    info, mn, D = E.factor(delta, gamma)
    assert info[:3] == (E.n, E.me, 0) and info[3] == 0
    kvr = kv.copy()
    for c in range(dim):
        k = cp[c] + np.searchsorted(ri[cp[c]:cp[c + 1]], c)
        kvr[k] += delta if c < E.n else -gamma
    rhs = rng.standard_normal(dim)
    nnzL, Do, xo, _ = ldlt(dim, cp, ri, kvr, rhs, perm)
    assert nnzL == st["nnz_l"]
    np.testing.assert_allclose(D, Do, rtol=1e-9)
    assert mn == pytest.approx(np.abs(Do).min(), rel=1e-9)
    xs = E.solve(rhs)
    np.testing.assert_allclose(xs, xo, rtol=0, atol=1e-9 * np.abs(xo).max())
    full = sp.csc_matrix((kvr, ri, cp), shape=(dim, dim))
    full = full + sp.tril(full, -1).T
    assert np.abs(full @ xs - rhs).max() < 1e-8
    E.close(); O.close()


@pytest.mark.parametrize("name,N", [("cart_pole", 40), ("cart_pole", 300),
                                    ("gfold", 12), ("flywheel", 50),
                                    ("arm_on_elevator", 30)])
def test_product_amd_equals_the_oracle_amd(name, N):
    """SLPB_ORDER_AMD (csrc/amd.cpp) returns the permutation of the oracle's
    restatement of Eigen's AMDOrdering (oracle/ldlt.hpp; reference call site
    sparse_regularized_ldlt.hpp:183) on the KKT patterns of the configs, and
    analysing with it reproduces the oracle's nnz(L) for that order."""
    from oracle.pyoracle import amd
    from emu import order_amd
    E, O = Emu(name, N), OracleProblem(name, N)
    E.eval(O.initial_guess(), np.zeros(E.me), np.ones(E.mi), 1.0,
           np.ones(E.me), np.ones(E.mi))
    O.close()
    cp, ri, kv = E.kkt(np.ones(E.mi))
    dim = E.n + E.me
    p_product = order_amd(dim, cp, ri)
    p_oracle = amd(dim, cp, ri)
    np.testing.assert_array_equal(p_product, p_oracle)
    st = E.analyze(1)   # SLPB_ORDER_AMD
    nnzL, *_ = ldlt(dim, cp, ri, np.where(ri == np.repeat(np.arange(dim), np.diff(cp)), 4.0, 0.01),
                    None, p_oracle)
    assert st["nnz_l"] == nnzL
    E.close()


@pytest.mark.parametrize("N,world", [(5000, 2), (5000, 4), (20000, 4), (20000, 8), (300, 8)])
def test_tree_partition_for_the_sharded_factorisation(N, world):
    """build_tree_shard (csrc/symbolic.cpp): a small replicated top, every
    subtree below it owned by exactly one rank, balanced work (SURVEY §8(e):
    local elimination to the interface, redundant interface solve)."""
    E = Emu("cart_pole", N)
    E.eval(np.zeros(E.n), np.zeros(E.me), np.ones(E.mi), 1.0, np.ones(E.me),
           np.ones(E.mi))
    E.kkt(np.ones(E.mi))
    st = E.analyze(0)
    n_top, owner, work = E.tree_shard(world, st["n_super"])
    assert n_top >= 1                      # (negative: a structural self-check failed)
    assert set(owner.tolist()) == set(range(-1, world))
    assert (owner < 0).sum() == n_top
    total = work.sum()
    assert work[world] <= (0.05 if N >= 5000 else 0.4) * total   # small replicated part
    assert work[:world].max() <= 1.25 * work[:world].mean()   # balanced
    F, npv, lvl, par = E.fronts(st["n_super"])
    assert lvl[owner < 0].min() >= lvl[owner >= 0].max() - 8
    E.close()


def test_nested_dissection_gives_a_shallow_tree():
    """The assembly tree's depth grows like log N (the reference's AMD order
    has an O(N) chain)."""
    depth = {}
    for N in (50, 400):
        E = Emu("cart_pole", N)
        E.eval(np.zeros(E.n), np.zeros(E.me), np.ones(E.mi), 1.0,
               np.ones(E.me), np.ones(E.mi))
        E.kkt(np.ones(E.mi))
        st = E.analyze(0)
        depth[N] = st["n_levels"]
        assert st["max_front"] <= 40
        E.close()
    assert depth[400] <= depth[50] + 4


def test_zero_pivot_flag():
    """δ = γ = 0 on the initial cart-pole KKT hits a structurally zero pivot
    (states without curvature): the factor must flag it, not crash."""
    E = Emu("cart_pole", 10)
    E.eval(np.zeros(E.n), np.zeros(E.me), np.ones(E.mi), 1.0, np.ones(E.me),
           np.ones(E.mi))
    E.kkt(np.ones(E.mi))
    E.analyze(2)
    info, mn, D = E.factor(0.0, 0.0)
    assert info[3] == 1 or info[2] > 0 or info[:2] != (E.n, E.me)
    E.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_sweep_equals_the_single_rank_sweep(world):
    """The multi-GPU shard plan (csrc/compile.cpp, build_shard_plan): every rank
    sweeps its share of the tasks, packs the stage slots it produced, and after
    the all-gather every slot is there exactly once — bit-identical derivatives
    for any number of ranks."""
    g = np.load(os.path.join(GOLDEN, "eval_cart_pole_40.npz"))
    E = Emu("cart_pole", 40)
    ref = E.eval(g["x"], g["y"], g["z"], float(g["d_f"]), g["d_ce"], g["d_ci"])
    E.set_shard_world(world)
    out = E.eval(g["x"], g["y"], g["z"], float(g["d_f"]), g["d_ce"], g["d_ci"])
    for k in ("g", "A_e", "A_i", "H", "c_e", "c_i"):
        np.testing.assert_array_equal(out[k], ref[k])
    assert out["f"] == ref["f"]
    E.close()


@pytest.mark.parametrize("name,N", [("cart_pole", 40), ("gfold", 12), ("flywheel", 30)])
def test_ordering_keeps_every_multiplier_behind_a_neighbour(name, N):
    """csrc/symbolic.cpp, defer_leading_multipliers: in the elimination order no
    multiplier (index ≥ n) precedes all of its neighbours — an unpivoted LDLᵀ
    would meet the bare −γ there (an exactly zero pivot for γ = 0). The order
    must still be a permutation."""
    E = Emu(name, N)
    n, me = E.n, E.me
    E.L.emu_kkt_build(E.h)
    E.analyze(0)
    perm = E.perm()
    dim = n + me
    assert sorted(perm) == list(range(dim))
    nnz = E.L.emu_kkt_build(E.h)
    cp = np.zeros(dim + 1, dtype=np.int32)
    ri = np.zeros(nnz, dtype=np.int32)
    E.L.emu_kkt_pattern(E.h, cp.ctypes.data_as(emu_ip), ri.ctypes.data_as(emu_ip))
    pos = np.empty(dim, dtype=np.int64)
    pos[perm] = np.arange(dim)
    nbrs = [[] for _ in range(dim)]
    for c in range(dim):
        for r in ri[cp[c]:cp[c + 1]]:
            if r != c:
                nbrs[r].append(c)
                nbrs[c].append(r)
    for d in range(n, dim):
        assert nbrs[d], "a multiplier without neighbours"
        assert min(pos[u] for u in nbrs[d]) < pos[d]
    E.close()


@pytest.mark.parametrize("name,N", [
    ("cart_pole", 100), ("cart_pole", 5000), ("flywheel", 200), ("gfold", 40),
    ("gfold", 2000), ("cart_pole_eq", 30), ("chained_rosenbrock", 2000),
    ("flywheel_eq", 400), ("flywheel_ocp", 100),
    ("flywheel_ocp_collocation", 100), ("flywheel_ocp_shooting", 40),
    ("flywheel_ocp_discrete", 100), ("cart_pole_ocp", 100),
    ("differential_drive_ocp", 50), ("double_integrator", 700),
    ("arm_on_elevator", 800), ("differential_drive", 20), ("all_ops", 0)])
def test_clusters_fit_one_thread_block(name, N, monkeypatch):
    """Every problem the GPU tier solves compiles into clusters that fit the
    224 KB of shared memory one thread block can have (kAdSmemMax in slpb.cu);
    unbounded cluster fusion once chained all steps of a collocation OCP into
    one 747 KB cluster."""
    monkeypatch.setenv("SLPB_EMU_SMEM_BUDGET", str(224 * 1024))
    E = Emu(name, N)
    assert E.n > 0
    E.close()


def test_concurrent_builds_are_independent():
    """Eight threads each build the cart-pole problem (thread-local expression
    pool; one of them also holds the process-wide fast slot), flatten it and
    compile it (the compiler starts threads of its own) at the same time — what
    slp::multistart does before its starts reach the device. Every one must
    reproduce the golden vector bit for bit."""
    import threading
    z = np.load(os.path.join(GOLDEN, "eval_cart_pole_40.npz"))
    g = {k: z[k] for k in z.files}          # (npz members load lazily)
    for _ in range(3):
        out = [None] * 8

        def work(i):
            E = Emu("cart_pole", 40)
            r = E.eval(g["x"], g["y"], g["z"], float(g["d_f"]), g["d_ce"], g["d_ci"])
            out[i] = (r["f"], r["g"].copy(), r["H"].copy(), r["A_e"].copy())
            E.close()
        threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for o in out:
            assert o is not None and o[0] == g["f"]
            np.testing.assert_array_equal(o[1], g["g"])
            np.testing.assert_array_equal(o[2], g["H_val"])
            np.testing.assert_array_equal(o[3], g["A_e_val"])


_SCHED_WORKER = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from emu import Emu
out = {}
for name, N in (("cart_pole", 40), ("all_ops", 0), ("flywheel", 60)):
    E = Emu(name, N)
    rng = np.random.default_rng(11)
    x = 0.3 + 0.5 * rng.random(E.n)
    y, z = rng.standard_normal(E.me), np.abs(rng.standard_normal(E.mi)) + 0.1
    r = E.eval(x, y, z, 0.7, np.full(max(E.me, 1), 1.3)[:E.me], np.full(max(E.mi, 1), 0.9)[:E.mi])
    for k, v in r.items():
        out[f"{name}.{k}"] = np.asarray(v)
    import ctypes as C
    hdr = (C.c_uint32 * 24)()
    n_prog = E.L.emu_program_header(E.h, 1, 0, hdr)
    best = None
    for p in range(n_prog):
        E.L.emu_program_header(E.h, 1, p, hdr)
        if best is None or hdr[14] > best[14]:
            best = list(hdr)
    out[f"{name}.header"] = np.array(best, dtype=np.int64)
    E.close()
np.savez(sys.argv[2], **out)
"""


def _run_schedule_variant(tmp_path, tag, env_extra):
    import subprocess
    import sys
    script = tmp_path / "sched_worker.py"
    script.write_text(_SCHED_WORKER)
    out = tmp_path / f"{tag}.npz"
    env = dict(os.environ)
    for k in ("SLPB_SCHED_LEGACY", "SLPB_NO_VALUE_REUSE", "SLPB_SCHED_WINDOW",
              "SLPB_SCHED_FLAT_COSTS", "SLPB_CHAIN_CAP"):
        env.pop(k, None)
    env.update(env_extra)
    subprocess.run([sys.executable, str(script), os.path.dirname(__file__), str(out)],
                   check=True, env=env, timeout=600)
    return np.load(out)


def test_schedule_and_value_reuse_do_not_change_a_single_bit(tmp_path):
    """The sweep compiler's list scheduling (which worker runs an item, in which
    super-level), its window, its cost model and the reuse of forward values as
    partials (d sin = the cos node, d exp = the node itself) only move work
    around: every output of the interpreter is bit-identical to the previous
    placement with every partial evaluated (the switches are read once per
    process, hence the subprocesses)."""
    ref = _run_schedule_variant(tmp_path, "default", {})
    variants = {
        "legacy": {"SLPB_SCHED_LEGACY": "1"},
        "no_reuse": {"SLPB_NO_VALUE_REUSE": "1"},
        "legacy_no_reuse": {"SLPB_SCHED_LEGACY": "1", "SLPB_NO_VALUE_REUSE": "1"},
        "window32": {"SLPB_SCHED_WINDOW": "32"},
        "flat": {"SLPB_SCHED_FLAT_COSTS": "1"},
    }
    for tag, env in variants.items():
        got = _run_schedule_variant(tmp_path, tag, env)
        for k in ref.files:
            if k.endswith(".header"):
                continue
            np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{tag}: {k}")
    # the switches did something: the legacy placement needs more barriers for
    # the cart-pole stage, and without reuse no partial is a negated product
    legacy = _run_schedule_variant(tmp_path, "legacy2", {"SLPB_SCHED_LEGACY": "1"})
    h_new, h_old = ref["cart_pole.header"], legacy["cart_pole.header"]
    assert h_new[12] + h_new[13] < h_old[12] + h_old[13]  # forward + reverse super-levels


def test_cart_pole_stage_fits_twice_on_an_sm():
    """The window of the list scheduler is chosen so that a 32-lane task of the
    cart-pole stage (scratch + tables + instruction ring) stays within 112 KB:
    two tasks per SM, as 157 tasks on 148 SMs need."""
    E = Emu("cart_pole", 64)
    hdr = (ctypes.c_uint32 * 24)()
    n_prog = E.L.emu_program_header(E.h, 1, 0, hdr)
    worst = 0
    for p in range(n_prog):
        E.L.emu_program_header(E.h, 1, p, hdr)
        scratch = (hdr[0] * 32 * 8 + 15) & ~15
        ring = (scratch + hdr[5] * 4 + 15) & ~15
        worst = max(worst, ring + hdr[22] * 4 + 3 * 8)
    assert 64 * 1024 < worst <= 112 * 1024
    E.close()
