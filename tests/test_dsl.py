"""CPU tier: the host DSL (sleipnir_b200/include/sleipnir) against the
reference's own DSL-level tests — constraints_test.cpp,
decision_variable_test.cpp, the expression-type CHECKs of the problem tests,
the pool-returned scope guard, multistart's selection rule — restated in
tests/dsl/dsl_tests.cpp, compiled here and run; no device call is made."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dsl_matches_the_reference_tests(tmp_path):
    exe = str(tmp_path / "dsl_tests")
    lib = os.path.join(ROOT, "sleipnir_b200", "lib")
    build = subprocess.run(
        [os.environ.get("CXX", "g++"), "-std=c++23", "-O1", "-pthread",
         "-I", os.path.join(ROOT, "sleipnir_b200", "include"),
         "-I", os.path.join(ROOT, "include"),
         os.path.join(ROOT, "tests", "dsl", "dsl_tests.cpp"), "-o", exe,
         "-L", lib, "-lslpb", f"-Wl,-rpath,{lib}"],
        capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-4000:]
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout[-4000:]
    assert run.stdout.startswith("ok ")
