"""contents of one pybindlibs.cpp_<dim>_<interp>_<nbRefinedPart> module (src/python3/cpp_simulator.hpp:42-139)"""
from phare_b200 import simulator as _sim


def populate(module, dim, interp, nref):
    if dim not in (1, 2, 3) or interp not in (1, 2, 3):
        raise ImportError(f"{module.__name__}: no such simulator permutation")

    class Simulator(_sim.Simulator):
        dims, interp_order, refined_particle_nbr = dim, interp, nref

        def __init__(self, hier):
            super().__init__(_sim.dict_instance(), hier, dim, interp, nref)

    def make_simulator(hier):
        return Simulator(hier)

    class Splitter:
        """amr::Splitter<dim, interp, nbRefinedPart> (src/amr/data/particles/refine/splitter.hpp): particle
        splitting belongs to level refinement, which this back end does not drive yet"""

        def __init__(self):
            raise NotImplementedError("particle splitting (refined levels) is not available in this back end")

    def split_pyarray_particles(*args, **kwargs):
        raise NotImplementedError("particle splitting (refined levels) is not available in this back end")

    module.Simulator = Simulator
    module.make_simulator = make_simulator
    module.DataWrangler = _sim.DataWrangler
    module.PatchLevel = _sim.PatchLevel
    module.Splitter = Splitter
    module.split_pyarray_particles = split_pyarray_particles
