"""The REFERENCE's own python test modules (tests/simulator/**, unittest + ddt), run UNCHANGED against this implementation
through tests/reference_unittest_runner.py: pyphare drives `pybindlibs`, diagnostics are written in the reference's h5 layout
and read back by pyphare's own fromh5 reader.  They check, among others: B, density and bulk velocity as provided by the
user; particle number per cell; domain and level-ghost particles of refined levels equal to the split of the coarser
particles; overlapped patch fields equal to 5.5e-15 (with and without refined levels); fine fields coarsened onto the coarser
level through the sub-cycles; domain particles on refined levels after advancing; restarts (every diagnostic of a resumed run
equal to the straight run's, time- and elapsed-triggered, modes, keep_last); the two-cell-jump exception and its emergency dump.  MHD permutations are filtered out (another
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
    "tests.simulator.test_restarts": 16,          # the reference's restart tests: write, resume, compare the diagnostics
    "tests.simulator.test_exceptions": 1,         # particle moved two cells -> exception + emergency dump
    "tests.simulator.test_validation": 60,
    "tests.simulator.test_simulation": 1,
    "tests.simulator.test_init_periodicity": 1,
    "tests.simulator.test_diagnostic_timestamps": 1,  # the dump cadence over 101 steps (the 30 000-step test is left out)
    # every diagnostic (incl. momentum tensors, particle_count, python attributes) present and readable at the dump times;
    # the 1-D permutations here, all 28 tests of the module (1/2/3-D, elapsed-time dumps) pass when run by hand (3 min)
    "tests.simulator.test_diagnostics": 9,
    # 2-D Harris sheet with refinement="tagging": runs restarted at several times have the layout and the data of the
    # continuous run
    "tests.simulator.refinement.test_regridding": 1,
}
FILTERS = {"tests.simulator.test_diagnostic_timestamps": ["-k", "test_hierarchy_timestamp_cadence"],
           "tests.simulator.test_diagnostics": ["-k", "'ndim': 1"]}


def start_module(module, cwd, backend):
    env = dict(os.environ, PHARE_B200_BACKEND=backend)
    os.makedirs(cwd, exist_ok=True)
    return subprocess.Popen([sys.executable, os.path.join(HERE, "reference_unittest_runner.py"), module, "-x", "MHD", *FILTERS.get(module, [])],
                            cwd=str(cwd), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def result_of(proc):
    out, err = proc.communicate(timeout=1500)
    m = re.search(r"RESULT \S+ run=(\d+) fail=(\d+) err=(\d+) skip=(\d+)", out)
    assert m, (out[-2000:], err[-4000:])
    return tuple(int(x) for x in m.groups()), proc.returncode, err


@pytest.fixture(scope="module")
def cpu_runs(tmp_path_factory, cpu_oracle, cpu_ref):
    """every module in its own process and directory, all started together (they are independent)"""
    base = tmp_path_factory.mktemp("reference_unittests")
    procs = {m: start_module(m, base / m.rsplit(".", 1)[1], "cpu") for m in MODULES}
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


@pytest.mark.parametrize("module", list(MODULES))
def test_reference_unittest_module_passes(cpu_runs, module):
    (run, fail, err, skip), rc, stderr = result_of(cpu_runs[module])
    assert (fail, err) == (0, 0), stderr[-6000:]
    assert run == MODULES[module] and rc == 0


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("PHARE_B200_REFERENCE_TESTS_ON_GPU") != "1", reason="opt-in")
@pytest.mark.parametrize("module", list(MODULES))
def test_reference_unittest_module_passes_on_the_gpu(tmp_path, module):
    (run, fail, err, skip), rc, stderr = result_of(start_module(module, tmp_path, "gpu"))
    assert (fail, err) == (0, 0), stderr[-6000:]
