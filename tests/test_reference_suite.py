"""Pins the CPU oracle with the REFERENCE'S OWN unit tests: the unmodified files under
/root/reference/tests/py are executed where they lie against tests/refshim/runko, a
stand-in `runko` package that binds the product's host logic (runko_b200.tiles) to the
oracle.  152 reference test cases cover fdtd2, the extended stencil, both binomial filters,
3 pushers x 2 interpolators, both zigzag depositers, sorting, the field setters/getters and
particle injection, the edge boundary conditions, the reflector wall, the moving injector and the antenna.

/root/reference exists only in the build container, so this file is skipped on the GPU box
(tests/test_kats.py restates the same known-answer cases for both backends there)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference/tests/py"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "test_emf_fdtd2.py": 6,
    "test_emf_stencil.py": 7,
    "test_emf_current_filter_binomial2.py": 4,
    "test_pic_particle_pusher.py": 54,
    "test_pic_current_depositer_zigzag_1st.py": 8,
    "test_pic_current_depositer_zigzag_1st_atomic.py": 8,
    "test_pic_particle_sorting.py": 2,
    "test_emf.py": 8,
    "test_pic.py": 6,
    "test_emf_edge_bc.py": 15,
    "test_pic_reflector_wall.py": 6,
    "test_pic_moving_injector.py": 13,
    "test_emf_antenna.py": 11,
    "test_emf_antenna_time_evolution.py": 4,
}


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("fname", sorted(FILES))
def test_reference_unit_tests_pass_on_the_oracle(fname):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "refshim"), os.path.dirname(HERE)])
    env["RUNKO_SHIM_BACKEND"] = "oracle"
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", os.path.join(REF, fname)],
                       cwd="/tmp", env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert f"{FILES[fname]} passed" in r.stdout, tail
