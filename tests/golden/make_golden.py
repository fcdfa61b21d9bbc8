"""Generates tests/golden/*.npz from the tier-A oracle (oracle/_ref: the
reference's OWN expression.hpp / expression_graph.hpp compiled where they lie
under /root/reference). Run in the authoring container only:

    python tests/golden/make_golden.py

The vectors pin (a) the autodiff outputs of the reference's graph walk at fixed
points and (b) the scalar trace of complete solves, so that the restated
oracle, the host emulation and the CUDA path can all be checked on a box that
has no /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import OracleProblem, have_reference  # noqa: E402

assert have_reference(), "build oracle/_ref first: make -C oracle ref"

EVAL_CASES = [("cart_pole", 8, 0.0, 0.0), ("cart_pole", 40, 0.0, 0.0),
              ("flywheel", 50, 0.0, 0.0), ("rosenbrock_cubic_line", 0, 0.3, 0.7),
              ("rosenbrock_disk", 0, -0.5, 1.2), ("wachter_biegler", 0, 0.0, 0.0),
              ("gfold", 12, 0.0, 0.0),
              # Newton / SQP branches (no inequality constraints)
              ("chained_rosenbrock", 30, 0.0, 0.0),
              ("cart_pole_eq", 12, 0.0, 0.0),
              ("min_distance_line", 0, 0.0, 0.0),
              # slp::OCP front end (optimization/ocp.hpp)
              ("all_ops", 0, 0.0, 0.0),
              ("double_integrator", 40, 0.0, 0.0),
              ("arm_on_elevator", 30, 0.0, 0.0),
              ("differential_drive", 20, 0.0, 0.0),
              ("flywheel_ocp", 30, 0.0, 0.0),
              ("flywheel_ocp_collocation", 30, 0.0, 0.0),
              ("flywheel_ocp_shooting", 12, 0.0, 0.0),
              ("flywheel_ocp_discrete", 30, 0.0, 0.0),
              ("cart_pole_ocp", 20, 0.0, 0.0),
              ("differential_drive_ocp", 20, 0.0, 0.0)]
SOLVE_CASES = [("flywheel", 50, 0.0, 0.0), ("cart_pole", 50, 0.0, 0.0),
               ("lp_maximize", 0, 0.0, 0.0), ("quartic", 0, 0.0, 0.0),
               ("qp_inequality_2d", 0, 0.0, 0.0),
               ("wachter_biegler", 0, 0.0, 0.0),
               ("rosenbrock_disk", 0, -0.5, 1.2),
               ("rosenbrock_cubic_line", 0, 0.3, 0.7),
               ("gfold", 20, 0.0, 0.0),
               ("chained_rosenbrock", 50, 0.0, 0.0),
               ("flywheel_eq", 50, 0.0, 0.0),
               ("cart_pole_eq", 30, 0.0, 0.0),
               ("eq_maximize_xy", 0, 0.0, 0.0),
               ("min_distance_line", 0, 0.0, 0.0),
               ("double_integrator", 700, 0.0, 0.0),
               ("arm_on_elevator", 800, 0.0, 0.0),
               ("differential_drive", 100, 0.0, 0.0),
               ("flywheel_ocp", 100, 0.0, 0.0),
               ("flywheel_ocp_collocation", 100, 0.0, 0.0),
               ("flywheel_ocp_shooting", 40, 0.0, 0.0),
               ("flywheel_ocp_discrete", 100, 0.0, 0.0),
               ("cart_pole_ocp", 100, 0.0, 0.0),
               ("differential_drive_ocp", 50, 0.0, 0.0)]


def eval_case(name, N, p0, p1):
    P = OracleProblem(name, N, p0, p1, backend="reference")
    P.eval_setup()
    d_f, d_ce, d_ci = P.scaling()
    rng = np.random.default_rng(1234)
    x = P.initial_guess() + 0.05 * rng.standard_normal(P.n)
    y = 0.3 * rng.standard_normal(P.me)
    z = 0.2 + np.abs(rng.standard_normal(P.mi))
    out = dict(x=x, y=y, z=z, d_f=d_f, d_ce=d_ce, d_ci=d_ci, f=P.f(x),
               c_e=P.c_e(x), c_i=P.c_i(x), g=P.g(x))
    for nm, M in (("A_e", P.A_e(x)), ("A_i", P.A_i(x)), ("H", P.H(x, y, z))):
        out[nm + "_colptr"] = M.colptr
        out[nm + "_rowidx"] = M.rowidx
        out[nm + "_val"] = M.val
    P.close()
    return out


def solve_case(name, N, p0, p1):
    P = OracleProblem(name, N, p0, p1, backend="reference")
    st = P.solve(keep_iterates=False)
    tr = P.trace()
    x, s, y, z = P.solution()
    out = dict(status=st, iterations=len(tr), x=x, s=s, y=y, z=z,
               error=np.array([r.error for r in tr]),
               cost=np.array([r.cost for r in tr]),
               alpha=np.array([r.alpha for r in tr]),
               delta=np.array([r.delta for r in tr]),
               mu=np.array([r.mu for r in tr]),
               factorizations=np.array([r.factorizations for r in tr]))
    P.close()
    return out


if __name__ == "__main__":
    only = set(sys.argv[1:])  # optional: regenerate only these problems
    if only:
        EVAL_CASES = [c for c in EVAL_CASES if c[0] in only]
        SOLVE_CASES = [c for c in SOLVE_CASES if c[0] in only]
    for c in EVAL_CASES:
        np.savez_compressed(os.path.join(HERE, f"eval_{c[0]}_{c[1]}.npz"),
                            **eval_case(*c))
    for c in SOLVE_CASES:
        np.savez_compressed(os.path.join(HERE, f"solve_{c[0]}_{c[1]}.npz"),
                            **solve_case(*c))
    print("golden vectors written to", HERE)
