"""Runs the UNMODIFIED reference package's own unit tests on this repo's casadi-free expression layer.

`tests/golden/ref_shim.py` installs `optas_b200.sym` as `casadi` and `optas_b200.urdf` as
`urdf_parser_py` (plus inert stubs for osqp / cvxopt / vtk / xacro), after which `/root/reference/optas`
imports as is.  Each reference test file is run in its own pytest process, as the reference's CI does
(.github/workflows/pytest.yaml:27-35; the class-level `SXContainer.is_discrete` dict otherwise leaks
between files).  Skipped where the reference checkout does not exist (the GPU box).
Not replayable here: tests/test_models.py (needs roboticstoolbox, pybullet), tests/test_examples.py
(pybullet, matplotlib), and the CasADi / OSQP / CVXOPT halves of tests/test_solver.py."""
import os
import shutil
import subprocess
import sys

import pytest

REF_TESTS = "/root/reference/tests"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILES = [
    ("test_sx_container.py", None, 5),
    ("test_optas_utils.py", None, 18),
    ("test_spatialmath.py", None, 66),
    ("test_optimization.py", None, 10),
    ("test_builder.py", None, 31),
    ("test_solver.py", "scipy", 1),   # 14 scipy.optimize methods on the Booth function through the reference's own ScipyMinimizeSolver
]


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference checkout not present on this machine")
@pytest.mark.parametrize("name,select,min_passed", FILES)
def test_reference_test_file_passes_on_the_shim(tmp_path, name, select, min_passed):
    shutil.copy(os.path.join(REF_TESTS, name), tmp_path / name)
    for extra in ("tester_robot.urdf", "tester_robot_revolute.urdf", "tester_robot_model.py"):
        if os.path.exists(os.path.join(REF_TESTS, extra)):
            shutil.copy(os.path.join(REF_TESTS, extra), tmp_path / extra)
    (tmp_path / "conftest.py").write_text(
        "import sys\n"
        f"sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})\n"
        "import ref_shim\n"
        "ref_shim.install()\n")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", name]
    if select:
        cmd += ["-k", select]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    n_passed = int(tail.split(" passed")[0].split()[-1])
    assert n_passed >= min_passed, tail
