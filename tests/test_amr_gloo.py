"""Refined hierarchies on two ranks (gloo, CPU back end): patches of both levels are dealt to the ranks, the
coarse -> fine gathers, the split particles and the coarsened data of the synchronisation cross ranks as packed
point-to-point messages.  The 2-rank result must equal the 1-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(ops, name, comm=None, steps=2):
    from amr_util import CASES, SOLVER_KW, B_init_nd
    from phare_b200.amr import build_hierarchy
    from solver_util import global_particles
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    dim = len(cells)
    gparts = global_particles(cells, interp, dx, ppc, 7)

    def particles_fn(i, L, pid):
        icell, delta, w, q, v = gparts[i]
        inside = np.ones(len(w), bool)
        for d in range(dim):
            inside &= (icell[:, d] >= L.amr_lower[d]) & (icell[:, d] < L.amr_lower[d] + L.ncells[d])
        return icell[inside], delta[inside], w[inside], q[inside], v[inside]

    h = build_hierarchy(ops, cells, grid, interp, dx, [dict(name="pop0", mass=1.0)], B_init_nd(cells, dx), particles_fn,
                        refinement_boxes=boxes, solver_kw=SOLVER_KW, comm=comm)
    for _ in range(steps):
        h.advance(0.004)
    return h


def _collect(h):
    from amr_util import level_fields
    res = {f"{k[0]}_{k[1]}_{k[2]}": v for k, v in level_fields(h).items()}
    for lvl in h.levels:
        for p in lvl.solver.patches:
            pop = p.pops[0]
            res[f"{lvl.number}_{p.geom.id}_counts"] = np.array(
                [h.ops.count(pop.domain)] + [h.ops.count(s) if s is not None else -1
                                             for s in (pop.level_ghost, pop.level_ghost_old, pop.level_ghost_new)])
    return res


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phare_b200.messenger import TorchComm
    from oracle.cpu_ops import CpuOps
    from amr_util import CASES
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    h = _run(CpuOps(len(cells), interp), name, comm=TorchComm(torch.device("cpu")))
    owners = {f"owner_{lvl.number}_{pg.id}": np.array([pg.owner]) for lvl in h.levels for pg in lvl.geom.patches}
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **_collect(h), **(owners if rank == 0 else {}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["1d_o1", "1d_o2_three_levels_two_root_patches", "1d_o1_at_the_periodic_boundary", "2d_o1",
                                  "3d_o1"])
def test_two_ranks_equal_one_process(cpu_ref, name, tmp_path):
    sys.path.insert(0, HERE)
    from oracle.cpu_ops import CpuOps
    from amr_util import CASES
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    want = _collect(_run(CpuOps(len(cells), interp), name))
    got = {}
    for r in range(2):
        got.update(np.load(os.path.join(str(tmp_path), f"rank{r}.npz")))
    owners = {k: int(v[0]) for k, v in got.items() if k.startswith("owner_")}
    # both ranks own root patches; the refined level is where the coarser patch under its corner is
    assert {o for k, o in owners.items() if k.startswith("owner_0_")} == {0, 1}
    assert set(want) == {k for k in got if not k.startswith("owner_")}
    for key, w in want.items():
        g = got[key]
        if key.endswith("_counts"):
            assert np.array_equal(g, w), key
            continue
        assert np.array_equal(np.isnan(g), np.isnan(w)), key
        ok = ~np.isnan(w) & np.isfinite(w)
        scale = np.max(np.abs(w[ok])) + 1e-30 if ok.any() else 1.0
        # identical plans and arithmetic; only the order of sums over particles that arrive from another rank differs
        assert np.max(np.abs(g[ok] - w[ok]), initial=0.0) <= 1e-11 * scale, key


def _sim_run(world_rank=None):
    """the dict-driven simulator with one refinement box, three steps; returns {key: array} of this rank's patches"""
    import phare_b200.simulator as S
    import pybindlibs.dictator as pp
    from oracle.cpu_ops import CpuOps
    from frontend_util import populate, two_pop_1d
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    pops, bfn = two_pop_1d(64)
    populate([64], [0.2], 1, pops, bfn, steps=3, largest=[16])
    pp.add_int("simulation/AMR/max_nbr_levels", 2)
    pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
    pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 2)
    for ib, (lo, hi) in enumerate(((10, 25), (26, 45))):   # two adjacent refined patches across root patch borders
        pp.add_int(f"simulation/AMR/refinement/boxes/L0/B{ib}/lower/x/", lo)
        pp.add_int(f"simulation/AMR/refinement/boxes/L0/B{ib}/upper/x/", hi)
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    for _ in range(3):
        sim.advance(sim.timeStep())
    res = {}
    for il, solver in enumerate(sim.level_solvers()):
        for p in solver.patches:
            for name, h in (("By", p.B[1]), ("Ez", p.E[2]), ("Ne", p.Ne), ("Vx", p.Vi[0])):
                res[f"{il}_{p.geom.id}_{name}"] = solver.ops.get_field(h)
            res[f"{il}_{p.geom.id}_n"] = np.array([solver.ops.count(pop.domain) for pop in p.pops])
    S.dict_instance().stop()
    return res


def _sim_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    np.savez(os.path.join(out_dir, f"sim_rank{rank}.npz"), **_sim_run())
    dist.barrier()
    dist.destroy_process_group()


def test_simulator_with_refinement_on_two_ranks_equals_one(cpu_ref, tmp_path):
    """pyphare-style dict -> Simulator on two ranks: 4 root patches and 2 refined patches dealt to the ranks"""
    sys.path.insert(0, HERE)
    import phare_b200.simulator as S
    old = S.ops_factory
    try:
        mp.spawn(_sim_worker, args=(2, 29500 + (os.getpid() % 2000) + 11, str(tmp_path)), nprocs=2, join=True)
        want = _sim_run()
    finally:
        S.ops_factory = old
    got = {}
    for r in range(2):
        got.update(np.load(os.path.join(str(tmp_path), f"sim_rank{r}.npz")))
    assert set(got) == set(want) and len([k for k in want if k.startswith("1_")]) == 2 * 5
    for k, w in want.items():
        if k.endswith("_n"):
            assert np.array_equal(got[k], w), k
        else:
            ok = np.isfinite(w)
            assert np.array_equal(np.isnan(got[k]), np.isnan(w)), k
            assert np.max(np.abs(got[k][ok] - w[ok]), initial=0.0) <= 1e-11 * (np.max(np.abs(w[ok])) + 1e-30), k


def _tagged_run():
    """a double tangential discontinuity advected at vx = 2 with refinement="tagging" (three levels): the hierarchy is
    built from tags and regridded after every root step (the finest level moves during the run)"""
    import phare_b200.simulator as S
    import pybindlibs.dictator as pp
    from oracle.cpu_ops import CpuOps
    from frontend_util import populate, const
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    cells, dl = 200, 1.0
    Lx = cells * dl
    Sx = lambda x, x0: 0.5 * (1 + np.tanh((x - x0) / 1.0))
    by = lambda x: -1 + 2 * (Sx(x, 0.25 * Lx) - Sx(x, 0.75 * Lx))
    T = lambda x: 1 - 0.5 * (by(x) ** 2 + 0.25)
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=30, seed=11, density=const(1.0), vx=const(2.0), vy=const(0),
               vz=const(0), vthx=T, vthy=T, vthz=T)
    populate([cells], [dl], 1, [pop], [const(0.0), by, const(0.5)], time_step=0.04, steps=12, nu=0.01, largest=[50])
    pp.add_int("simulation/AMR/max_nbr_levels", 3)
    pp.add_string("simulation/AMR/refinement/tagging/method", "auto")
    pp.add_double("simulation/AMR/refinement/tagging/threshold", 0.1)
    pp.add_int("simulation/AMR/tag_buffer", 1)
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    boxes0 = [[(int(p.box.lo[0]), int(p.box.hi[0])) for p in lvl.geom.patches] for lvl in sim.amr.levels]
    for _ in range(12):
        sim.advance(sim.timeStep())
    res = {"boxes0": np.array(sum(boxes0[1:], []))}
    res["boxes"] = np.array(sum([[(int(p.box.lo[0]), int(p.box.hi[0])) for p in lvl.geom.patches] for lvl in sim.amr.levels[1:]], []))
    for il, solver in enumerate(sim.level_solvers()):
        for p in solver.patches:
            for name, h in (("By", p.B[1]), ("Ez", p.E[2]), ("Ne", p.Ne)):
                res[f"{il}_{p.geom.id}_{name}"] = solver.ops.get_field(h)
            res[f"{il}_{p.geom.id}_n"] = np.array([solver.ops.count(pop.domain) for pop in p.pops])
    S.dict_instance().stop()
    return res


def _tagged_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    np.savez(os.path.join(out_dir, f"tag_rank{rank}.npz"), **_tagged_run())
    dist.barrier()
    dist.destroy_process_group()


def test_tagging_and_regridding_on_two_ranks_equal_one(cpu_ref, tmp_path):
    sys.path.insert(0, HERE)
    import phare_b200.simulator as S
    old = S.ops_factory
    try:
        mp.spawn(_tagged_worker, args=(2, 29500 + (os.getpid() % 2000) + 13, str(tmp_path)), nprocs=2, join=True)
        want = _tagged_run()
    finally:
        S.ops_factory = old
    got = {}
    for r in range(2):
        got.update(np.load(os.path.join(str(tmp_path), f"tag_rank{r}.npz")))
    assert len(want["boxes"]) == 4 and not np.array_equal(want["boxes"], want["boxes0"])   # the hierarchy moved
    assert set(got) == set(want)
    for k, w in want.items():
        if k.startswith("boxes") or k.endswith("_n"):
            assert np.array_equal(got[k], w), k
        else:
            ok = np.isfinite(w)
            assert np.array_equal(np.isnan(got[k]), np.isnan(w)), k
            assert np.max(np.abs(got[k][ok] - w[ok]), initial=0.0) <= 1e-10 * (np.max(np.abs(w[ok])) + 1e-30), k


def _restart_worker(rank, world, port, out_dir):
    """two ranks, refined level: 2 steps, restart written, 1 more step; then a new simulator resumed from the per-rank
    restart files takes that step again and must land on the same state bit for bit"""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import phare_b200.simulator as S
    import pybindlibs.dictator as pp
    from oracle.cpu_ops import CpuOps
    from frontend_util import populate, two_pop_1d
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    dt = 0.005

    def make(extra):
        S.dict_instance().stop()
        pops, bfn = two_pop_1d(64)
        populate([64], [0.2], 1, pops, bfn, steps=3, largest=[16])
        pp.add_int("simulation/AMR/max_nbr_levels", 2)
        pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
        pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 2)
        for ib, (lo, hi) in enumerate(((10, 25), (26, 45))):
            pp.add_int(f"simulation/AMR/refinement/boxes/L0/B{ib}/lower/x/", lo)
            pp.add_int(f"simulation/AMR/refinement/boxes/L0/B{ib}/upper/x/", hi)
        pp.add_string("simulation/restarts/filePath", out_dir)
        pp.add_string("simulation/restarts/serialized_simulation", "x")
        extra()
        sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
        sim.initialize()
        return sim

    def state(sim):
        ops = sim.solver.ops
        out = []
        for solver in sim.level_solvers():
            for p in solver.patches:
                out += [ops.get_field(f) for f in (*p.B, *p.E, p.Ne, *p.Vi)]
                for pop in p.pops:
                    out += list(ops.get_particles(pop.domain))
        return out

    sim = make(lambda: pp.add_array_as_vector("simulation/restarts/write_timestamps", np.array([2 * dt])))
    for _ in range(3):
        sim.dump_restarts(sim.currentTime(), dt)
        sim.advance(dt)
    want = state(sim)
    load = S.restart_path_for_time(out_dir, 2 * dt)
    assert sorted(os.listdir(load)) == ["restart_rank000000.npz", "restart_rank000001.npz"]

    def loading():
        pp.add_string("simulation/restarts/loadPath", load)
        pp.add_double("simulation/restarts/restart_time", 2 * dt)
    sim2 = make(loading)
    sim2.advance(dt)
    got = state(sim2)
    assert len(got) == len(want) and len(want) > 0
    for a, b in zip(want, got):
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_restart_on_two_ranks_resumes_bit_for_bit(cpu_ref, tmp_path):
    mp.spawn(_restart_worker, args=(2, 29500 + (os.getpid() % 2000) + 17, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _h5_worker(rank, world, port, out_dir):
    """two ranks write one quantity: rank 0 the file, rank 1 its `.rank1` piece; a reader sees the patches of both"""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PHARE_B200_DIAG_FORMAT"] = "h5"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import phare_b200.simulator as S
    from phare_b200 import h5lite
    from oracle.cpu_ops import CpuOps
    from frontend_util import populate, two_pop_1d
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    pops, bfn = two_pop_1d(64)
    populate([64], [0.2], 1, pops, bfn, steps=1, largest=[16], diag_dir=out_dir, diag_times=[0.0])
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    assert sim.dump_diagnostics(0.0, 0.005)
    sim.close_diagnostics()
    dist.barrier()
    if rank == 0:
        f = h5lite.File(os.path.join(out_dir, "EM_B.h5"))
        patches = f["t/0.0000000000/pl0"]
        assert len(patches.keys()) == 4 and {k[:2] for k in patches.keys()} == {"p0", "p1"}
        lowers = sorted(int(p.attrs["lower"][0]) for p in patches.values())
        assert lowers == [0, 16, 32, 48] and all(p["EM_B_y"].shape == (16 + 4,) for p in patches.values())
        assert {int(p.attrs["mpi_rank"]) for p in patches.values()} == {0, 1}
        with open(os.path.join(out_dir, "ok"), "w") as o:
            o.write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_h5_diagnostics_from_two_ranks_merge(cpu_ref, tmp_path):
    mp.spawn(_h5_worker, args=(2, 29500 + (os.getpid() % 2000) + 19, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
