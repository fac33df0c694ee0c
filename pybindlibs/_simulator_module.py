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
        """amr::Splitter<dim, interp, nbRefinedPart> (src/amr/data/particles/refine/splitter.hpp, split_{1,2,3}d.hpp):
        the flattened pattern table phb_split consumes (phare_b200/split.py)"""

        def __init__(self):
            from phare_b200.split import pattern
            self.deltas, self.weights, self.max_cell_distance = pattern(dim, interp, nref)
            self.nbRefinedPart = nref

    def split_pyarray_particles(particles):
        """splitPyArrayParticles (src/python3/particles.hpp): (iCell, delta, weight, charge, v) flat arrays of coarse
        particles -> the same five arrays for their refined particles, on the fine index space (runs phb_split)"""
        import numpy as np
        from phare_b200 import abi
        icell, delta, weight, charge, v = (np.asarray(a).reshape(-1) for a in particles)
        n = len(weight)
        ops = _sim.ops_factory(dim, interp)  # the CUDA back end; the CPU parity tests swap in the oracle here
        try:
            src, dst = ops.particles(max(n, 1)), ops.particles(max(n * nref, 1))
            ops.set_particles(src, icell.reshape(n, dim), delta.reshape(n, dim), weight, charge, v.reshape(n, 3))
            ops.set_count(dst, 0)
            big = 2 ** 30
            everywhere = abi.make_box([-big] * dim, [big] * dim)
            if ops.split(nref, src, 0, n, [everywhere], dst) is None:
                raise RuntimeError("split_pyarray_particles: destination store too small")
            ic, de, w, q, vv = ops.get_particles(dst)
        finally:
            ctx = getattr(ops, "ctx", None)  # the device context of this one call
            if ctx is not None:
                ctx.close()
        return ic.reshape(-1), de.reshape(-1), w, q, vv.reshape(-1)

    module.Simulator = Simulator
    module.make_simulator = make_simulator
    module.DataWrangler = _sim.DataWrangler
    module.PatchLevel = _sim.PatchLevel
    module.Splitter = Splitter
    module.split_pyarray_particles = split_pyarray_particles
