"""Runs one of the REFERENCE's own unittest modules (/root/reference/tests/simulator/**), unchanged, against this
implementation: `pybindlibs` stands in for the reference's pybind modules, the compute back end is the CPU parity back end
(oracle/cpu_ops.py; the same files run on the GPU back end with PHARE_B200_BACKEND=gpu), diagnostics are written in the
reference's per-quantity h5 layout through phare_b200.h5lite (registered as `h5py`, absent here) and `ddt` is tests/shims/ddt.py.
Own process because it rebinds sys.modules.  Usage: python reference_unittest_runner.py <module> [-k substr] [-x substr]
Prints `RESULT <module> run=<n> fail=<n> err=<n> skip=<n>`; MHD permutations (another solver) are filtered out with -x MHD."""
import importlib
import os
import sys
import unittest
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PHARE_REFERENCE", "/root/reference")


def main(argv):
    module, keep, drop = argv[0], [], []
    it = iter(argv[1:])
    for a in it:
        (keep if a == "-k" else drop).append(next(it))
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections", "matplotlib.gridspec",
                 "matplotlib.animation", "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(name, mock.MagicMock(name=name))
    plt = sys.modules["matplotlib.pyplot"]
    if isinstance(plt, mock.MagicMock):  # no matplotlib here: `fig, ax = plt.subplots()` still has to unpack
        def subplots(nrows=1, ncols=1, **kw):
            ax = lambda: mock.MagicMock(name="Axes")
            if nrows * ncols == 1:
                return mock.MagicMock(name="Figure"), ax()
            if min(nrows, ncols) == 1:
                return mock.MagicMock(name="Figure"), tuple(ax() for _ in range(nrows * ncols))
            return mock.MagicMock(name="Figure"), tuple(tuple(ax() for _ in range(ncols)) for _ in range(nrows))
        plt.subplots = subplots
        sys.modules["matplotlib"].pyplot = plt
    sys.path[:0] = [os.path.dirname(HERE), os.path.join(REF, "pyphare"), REF, HERE, os.path.join(HERE, "shims")]
    os.environ["PHARE_B200_DIAG_FORMAT"] = "h5"
    from phare_b200 import h5lite
    h5lite.install()
    if os.environ.get("PHARE_B200_BACKEND", "cpu") == "cpu":
        import phare_b200.simulator as S
        from oracle.cpu_ops import CpuOps
        S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    m = importlib.import_module(module)

    def flatten(suite):
        for t in suite:
            if isinstance(t, unittest.TestSuite):
                yield from flatten(t)
            else:
                yield t
    def label(t):  # test id + the ddt datum it was generated from
        return t.id() + " " + repr(getattr(getattr(t, t._testMethodName, None), "ddt_value", ""))
    tests = [t for t in flatten(unittest.defaultTestLoader.loadTestsFromModule(m))
             if all(k in label(t) for k in keep) and not any(x in label(t) for x in drop)]
    r = unittest.TextTestRunner(verbosity=2).run(unittest.TestSuite(tests))
    print(f"RESULT {module} run={r.testsRun} fail={len(r.failures)} err={len(r.errors)} skip={len(r.skipped)}", flush=True)
    return 0 if r.wasSuccessful() else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
