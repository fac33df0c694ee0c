"""Helpers for the front-end tests: fill the PHAREDict through pybindlibs.dictator exactly like
pyphare.pharein.populateDict does (initialize/general.py + hybrid.py), without needing pyphare."""
import numpy as np

XYZ = "xyz"


def populate(cells, dl, interp, pops, bfn, time_step=0.005, steps=4, eta=1e-3, nu=1e-3, Te=0.12, largest=None,
             diag_dir=None, diag_times=()):
    import pybindlibs.dictator as pp
    pp.stop()
    dim = len(cells)
    add_fn = getattr(pp, f"addInitFunction{dim}D")
    pp.add_string("simulation/name", "simulation_test")
    pp.add_int("simulation/dimension", dim)
    if largest is not None:
        pp.add_vector_int("simulation/AMR/largest_patch_size", [int(x) for x in largest])
    pp.add_string("simulation/grid/layout_type", "yee")
    for i in range(dim):
        pp.add_int(f"simulation/grid/nbr_cells/{XYZ[i]}", int(cells[i]))
        pp.add_double(f"simulation/grid/meshsize/{XYZ[i]}", float(dl[i]))
        pp.add_string(f"simulation/grid/boundary_type/{XYZ[i]}", "periodic")
    pp.add_int("simulation/interp_order", interp)
    pp.add_int("simulation/refined_particle_nbr", 2)
    pp.add_double("simulation/time_step", time_step)
    pp.add_int("simulation/time_step_nbr", steps)
    pp.add_double("simulation/final_time", time_step * steps)
    pp.add_int("simulation/AMR/max_nbr_levels", 1)
    pp.add_int("simulation/AMR/max_mhd_level", 0)
    pp.add_string("simulation/AMR/refinement/tagging/method", "none")
    pp.add_vector_string("simulation/models", ["HybridModel"])
    pp.add_string("simulation/algo/ion_updater/pusher/name", "modified_boris")
    pp.add_double("simulation/algo/ohm/resistivity", eta)
    pp.add_double("simulation/algo/ohm/hyper_resistivity", nu)
    pp.add_string("simulation/algo/ohm/hyper_mode", "constant")
    pp.add_size_t("simulation/ions/nbrPopulations", len(pops))
    for i, p in enumerate(pops):
        base = f"simulation/ions/pop{i}/"
        init = base + "particle_initializer/"
        pp.add_string(base + "name", p["name"])
        pp.add_double(base + "mass", p["mass"])
        pp.add_string(init + "name", "maxwellian")
        add_fn(init + "density", p["density"])
        for c, key in zip(XYZ, ("vx", "vy", "vz")):
            add_fn(init + f"bulk_velocity_{c}", p[key])
        for c, key in zip(XYZ, ("vthx", "vthy", "vthz")):
            add_fn(init + f"thermal_velocity_{c}", p[key])
        pp.add_double(init + "charge", p["charge"])
        pp.add_string(init + "basis", "cartesian")
        if p.get("seed") is not None:
            pp.add_optional_size_t(init + "init/seed", p["seed"])
        pp.add_int(init + "nbr_part_per_cell", p["ppc"])
        pp.add_double(init + "density_cut_off", p.get("cut", 1e-5))
    for c, f in zip(XYZ, bfn):
        add_fn(f"simulation/electromag/magnetic/initializer/{c}_component", f)
    pp.add_string("simulation/electrons/pressure_closure/name", "isothermal")
    pp.add_double("simulation/electrons/pressure_closure/Te", Te)
    if diag_dir is not None:
        for q in ("EM_B", "EM_E"):
            pp.add_string(f"simulation/diagnostics/electromag/{q}/type", "electromag")
            pp.add_string(f"simulation/diagnostics/electromag/{q}/quantity", "/" + q)
            pp.add_array_as_vector(f"simulation/diagnostics/electromag/{q}/write_timestamps", np.asarray(diag_times, float))
        for i, p in enumerate(pops):
            pp.add_string(f"simulation/diagnostics/fluid/flux{i}/type", "fluid")
            pp.add_string(f"simulation/diagnostics/fluid/flux{i}/quantity", f"/ions/pop/{p['name']}/flux")
            pp.add_array_as_vector(f"simulation/diagnostics/fluid/flux{i}/write_timestamps", np.asarray(diag_times[-1:], float))
            q = f"/ions/pop/{p['name']}/domain"
            pp.add_string(f"simulation/diagnostics/particle/part{i}/type", "particle")
            pp.add_string(f"simulation/diagnostics/particle/part{i}/quantity", q)
            pp.add_array_as_vector(f"simulation/diagnostics/particle/part{i}/write_timestamps", np.asarray(diag_times[-1:], float))
        pp.add_string("simulation/diagnostics/filePath", diag_dir)


def const(v):
    return lambda *x: np.full(len(x[0]), float(v))


def two_pop_1d(cells=64, dl=0.2, seed=77):
    Lx = cells * dl
    main = dict(name="protons", mass=1.0, charge=1.0, ppc=40, seed=seed,
                density=lambda x: 1.0 + 0.2 * np.sin(2 * np.pi * x / Lx), vx=const(0), vy=const(0), vz=const(0),
                vthx=const(0.3), vthy=const(0.3), vthz=const(0.3))
    beam = dict(name="beam", mass=2.0, charge=1.0, ppc=20, seed=seed + 1, density=const(0.1), vx=const(0.5), vy=const(0),
                vz=const(0), vthx=const(0.2), vthy=const(0.2), vthz=const(0.2))
    bfn = [const(1.0), lambda x: 0.1 * np.cos(2 * np.pi * x / Lx), lambda x: 0.05 * np.sin(2 * np.pi * x / Lx)]
    return [main, beam], bfn


def gather(sim, attr, comp):
    """physical + ghost arrays of every patch, keyed by patch id"""
    ops = sim.solver.ops
    out = {}
    for p in sim.solver.patches:
        h = getattr(p, attr)
        out[p.geom.id] = ops.get_field(h[comp] if comp is not None else h)
    return out
