"""The C-ABI libraries load, export every symbol include/slpb.h declares, and
refuse to run without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import sleipnir_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "slpb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slpb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_symbol_lists_agree():
    assert _declared_symbols() == sorted(sb.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = sb.device_lib()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"libslpb.so does not export {name}"


def test_host_library_loads():
    H = sb.host_lib()
    for name in ("slpbh_problem_create", "slpbh_solve", "slpbh_device_open",
                 "slpbh_trace_get", "slpbh_solution"):
        assert hasattr(H, name)


def test_problem_dimensions_match_the_survey_table():
    # cart-pole n = 5N+4, m_e = 4N+8, m_i = 4N+2; flywheel 2N+1, N+1, 2N
    P = sb.Problem("cart_pole", 100)
    assert (P.n, P.me, P.mi) == (504, 408, 402)
    assert P.types() == (3, 4, 2)
    P.close()
    P = sb.Problem("flywheel", 50)
    assert (P.n, P.me, P.mi) == (101, 51, 100)
    P.close()


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_fails_loudly_without_a_device():
    lib = sb.device_lib()
    h = C.c_void_p()
    rc = lib.slpb_create(0, C.byref(h))
    assert rc == -5  # SLPB_ERR_NO_DEVICE
    P = sb.Problem("quartic")
    with pytest.raises(sb.DeviceError, match="no CPU fallback"):
        P.solve()
    P.close()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_concurrent_graph_building_fails_loudly_too():
    """slp::multistart builds one problem per host thread (each with its own
    expression pool) before any start reaches the device: without a GPU all
    eight threads must get as far as the device call and report it — no crash
    in the thread-local pool machinery, no silent CPU path."""
    for _ in range(3):
        with pytest.raises(sb.DeviceError):
            sb.multistart("cart_pole", 40, [5.0] * 8)
    # the calling thread's DSL still works afterwards
    P = sb.Problem("flywheel", 20)
    assert (P.n, P.me, P.mi) == (41, 21, 40)
    P.close()


def test_trivial_problems_need_no_solver():
    """trivial_problem_test.cpp:14-66: an empty problem and one with neither
    cost nor constraints return SUCCESS before any solver (or device) is
    touched, and leave the variables alone."""
    P = sb.Problem("empty")
    assert (P.n, P.me, P.mi) == (0, 0, 0) and P.types() == (0, 0, 0)
    assert sb.EXIT_STATUS[P.solve()] == "SUCCESS"
    P.close()
    for value in (0.0, 1.0):
        P = sb.Problem("no_cost_unconstrained", 0, value)
        assert (P.n, P.me, P.mi) == (6, 0, 0) and P.types() == (0, 0, 0)
        assert sb.EXIT_STATUS[P.solve()] == "SUCCESS"
        assert list(P.initial_guess()) == [value] * 6
        P.close()


def test_ocp_front_end_shapes_and_types():
    """What the reference's OCP tests CHECK before solving
    (flywheel_ocp_test.cpp:66-68, cart_pole_ocp_test.cpp:83-85,
    differential_drive_ocp_test.cpp:59-61): expression types, and the decision
    variable layout U | dt | X."""
    QUADRATIC, LINEAR, NONLINEAR = 3, 2, 4
    for name, N, dims in (("flywheel_ocp", 50, (102, 51, 102)),
                          ("flywheel_ocp_collocation", 50, (102, 51, 102)),
                          ("flywheel_ocp_discrete", 50, (102, 51, 102)),
                          ("flywheel_ocp_shooting", 50, (51, 1, 102))):
        P = sb.Problem(name, N)
        assert (P.n, P.me, P.mi) == dims
        assert P.types() == (QUADRATIC, LINEAR, LINEAR)
        P.close()
    P = sb.Problem("cart_pole_ocp", 100)
    assert (P.n, P.me, P.mi) == (101 + 1 + 404, 400 + 8, 202 + 202)
    assert P.types() == (QUADRATIC, NONLINEAR, LINEAR)
    P.close()
    P = sb.Problem("differential_drive_ocp", 50)
    assert (P.n, P.me, P.mi) == (102 + 1 + 255, 250 + 10, 204 + 102)
    assert P.types() == (LINEAR, NONLINEAR, LINEAR)
    P.close()


def test_product_does_not_reference_the_oracle():
    """Nothing under sleipnir_b200/ (nor include/) may include, import or link
    anything under oracle/."""
    bad = []
    for base in ("sleipnir_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if not f.endswith((".hpp", ".h", ".cpp", ".cu", ".py", "Makefile")):
                    continue
                text = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r'#include\s+[<"].*oracle|from oracle|import oracle|'
                             r'liboracle|pyoracle', text):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_device_code_is_sm100a_with_tensor_core_and_tma_instructions():
    """What the cubin of libslpb.so contains (cuobjdump runs without a GPU):
    sm_100a code only; FP64 tensor-core products (DMMA.8x8x4: the rank-4
    updates of dense fronts, csrc/ldlt_dense.cuh) inside the factorisation
    kernel; 1-D TMA bulk copies + mbarriers (UBLKCP, SYNCS) inside the autodiff
    sweep (the instruction stream, slpb.cu TmaStream)."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump is not installed")
    lib = sb.LIB_DEVICE
    elf = subprocess.run([exe, "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf and "sm_80" not in elf
    sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True).stdout
    current, found = None, {}
    for line in sass.splitlines():
        if "Function :" in line:
            current = line.split("Function :")[1].strip()
        for mnemonic in ("DMMA", "UBLKCP", "SYNCS"):
            if mnemonic in line and current:
                found.setdefault(mnemonic, set()).add(current)
    assert any("k_factor_tree" in f for f in found.get("DMMA", ())), found.get("DMMA")
    assert any("k_ad_sweep" in f for f in found.get("UBLKCP", ()))
    assert any("k_ad_sweep" in f for f in found.get("SYNCS", ()))
