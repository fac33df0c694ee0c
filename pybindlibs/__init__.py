"""Drop-in for the reference's `pybindlibs` extension modules, so that pyphare input scripts and
`pyphare.simulator.simulator.Simulator` run unchanged on the B200 back end:

    pybindlibs.dictator            src/initializer/dictator.cpp:41-62
    pybindlibs.cpp_etc             src/python3/cpp_etc.cpp:74-143
    pybindlibs.cpp_<dim>_<interp>_<nbRefinedPart>   src/python3/cpp_simulator.hpp:42-139 (one module per
                                   compile-time permutation in the reference; generated on import here)

Put the repository root on PYTHONPATH together with the reference's `pyphare/` directory.  The reference builds
these with pybind11 around its C++ Simulator; here the step driver above the C ABI is Python already
(phare_b200/solver.py), so the modules are plain Python over phare_b200.simulator.
"""
import importlib.abc
import importlib.machinery
import re
import sys
import types

_SIM_MODULE = re.compile(r"^pybindlibs\.cpp_(\d)_(\d)_(\d+)$")


class _SimulatorModuleFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """resolves pybindlibs.cpp_<dim>_<interp>_<nbRefinedPart> for every hybrid permutation (pyphare.cpp.simulator_id)"""

    def find_spec(self, fullname, path, target=None):
        if _SIM_MODULE.match(fullname):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return types.ModuleType(spec.name)

    def exec_module(self, module):
        from ._simulator_module import populate
        dim, interp, nref = (int(g) for g in _SIM_MODULE.match(module.__name__).groups())
        populate(module, dim, interp, nref)


if not any(isinstance(f, _SimulatorModuleFinder) for f in sys.meta_path):
    sys.meta_path.append(_SimulatorModuleFinder())
