"""The REFERENCE's own python test modules (tests/simulator/**, unittest + ddt), run UNCHANGED against this implementation
through tests/reference_unittest_runner.py: pyphare drives `pybindlibs`, diagnostics are written in the reference's h5 layout
and read back by pyphare's own fromh5 reader.  They check, among others: B, density and bulk velocity as provided by the
user; particle number per cell; domain and level-ghost particles of refined levels equal to the split of the coarser
particles; overlapped patch fields equal to 5.5e-15 (with and without refined levels); fine fields coarsened onto the coarser
level through the sub-cycles; domain particles on refined levels after advancing.  MHD permutations are filtered out (another
solver).  CPU parity back end here; PHARE_B200_REFERENCE_TESTS_ON_GPU=1 runs the same modules on the CUDA back end."""
import os
import re
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests", "simulator")), reason="reference tree not mounted")

# module -> tests expected to run (the reference skips some itself)
MODULES = {
    "tests.simulator.data_wrangler": 1,
    "tests.simulator.initialize.test_fields_init_1d": 12,
    "tests.simulator.initialize.test_particles_init_1d": 23,
    "tests.simulator.advance.test_fields_advance_1d": 36,
    "tests.simulator.advance.test_particles_advance_1d": 6,
}


def run_module(module, tmp_path, backend):
    env = dict(os.environ, PHARE_B200_BACKEND=backend)
    r = subprocess.run([sys.executable, os.path.join(HERE, "reference_unittest_runner.py"), module, "-x", "MHD"],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=1500)
    m = re.search(r"RESULT \S+ run=(\d+) fail=(\d+) err=(\d+) skip=(\d+)", r.stdout)
    assert m, (r.stdout[-2000:], r.stderr[-4000:])
    return tuple(int(x) for x in m.groups()), r


@pytest.mark.parametrize("module", list(MODULES))
def test_reference_unittest_module_passes(cpu_oracle, cpu_ref, tmp_path, module):
    (run, fail, err, skip), r = run_module(module, tmp_path, "cpu")
    assert (fail, err) == (0, 0), r.stderr[-6000:]
    assert run == MODULES[module] and r.returncode == 0


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("PHARE_B200_REFERENCE_TESTS_ON_GPU") != "1", reason="opt-in")
@pytest.mark.parametrize("module", list(MODULES))
def test_reference_unittest_module_passes_on_the_gpu(tmp_path, module):
    (run, fail, err, skip), r = run_module(module, tmp_path, "gpu")
    assert (fail, err) == (0, 0), r.stderr[-6000:]
