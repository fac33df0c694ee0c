"""N>1 path on GPUs: one process per GPU (NCCL for the plumbing), field exchange phases through NVLink peer
memory (csrc/peer.cu + messenger.PeerArena).  The result must equal the same run staged through NCCL send/recv
(PHB_PEER_HALO=0) and the one-process multi-patch run.  Needs >= 2 GPUs: skipped on the single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, case, out_dir, peer):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PHB_PEER_HALO="1" if peer else "0",
                      PHB_PEER_ARENA_MB="64", PHB_PEER_TIMEOUT_S="10")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from phare_b200.messenger import TorchComm
    from phare_b200.solver import GpuOps
    from solver_util import global_particles, make_solver, FIELDS
    domain, grid, interp, dx, ppc, npop, steps = case
    dim = len(domain)
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    comm = TorchComm(torch.device(f"cuda:{rank}"))
    s = make_solver(GpuOps(dim, interp, f"cuda:{rank}"), domain, grid, interp, dx, gparts, comm=comm)
    assert (s.messenger.arena is not None) == peer
    for _ in range(steps):
        s.advance_level(0.005)
    res = {}
    for p in s.patches:
        for attr, comp, qty in FIELDS:
            h = getattr(p, attr)
            res[f"{p.geom.id}_{attr}_{comp}"] = s.ops.get_field(h[comp] if comp is not None else h)
        for i in range(npop):
            res[f"{p.geom.id}_n{i}"] = np.array([s.ops.count(p.pops[i].domain)])
    np.savez(os.path.join(out_dir, f"{'peer' if peer else 'nccl'}_rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("case", [
    ((48,), (4,), 1, (0.2,), 20, 1, 4),
    ((16, 12), (2, 2), 1, (0.4, 0.4), 8, 2, 3),
    ((8, 8, 8), (2, 1, 2), 1, (0.2, 0.2, 0.2), 4, 1, 3),
])
def test_peer_memory_halo_equals_nccl_and_single_process(case, tmp_path):
    sys.path.insert(0, HERE)
    from phare_b200.solver import GpuOps
    from solver_util import global_particles, make_solver, FIELDS
    domain, grid, interp, dx, ppc, npop, steps = case
    world = 2
    for k, peer in enumerate((True, False)):
        port = 29600 + (os.getpid() % 1000) + k
        mp.spawn(_worker, args=(world, port, case, str(tmp_path), peer), nprocs=world, join=True)
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    ref = make_solver(GpuOps(len(domain), interp, "cuda:0"), domain, grid, interp, dx, gparts)
    for _ in range(steps):
        ref.advance_level(0.005)
    got = {m: {} for m in ("peer", "nccl")}
    for m in got:
        for r in range(world):
            got[m].update(np.load(os.path.join(str(tmp_path), f"{m}_rank{r}.npz")))
    for p in ref.patches:
        for attr, comp, qty in FIELDS:
            h = getattr(p, attr)
            want = ref.ops.get_field(h[comp] if comp is not None else h)
            scale = np.nanmax(np.abs(want)) + 1e-30
            for m in got:  # identical plans and arithmetic; only the order of the atomic border sums may differ
                g = got[m][f"{p.geom.id}_{attr}_{comp}"]
                assert np.allclose(g, want, rtol=0, atol=1e-11 * scale, equal_nan=True), (m, p.geom.id, attr, comp)
        for i in range(npop):
            for m in got:
                assert int(got[m][f"{p.geom.id}_n{i}"][0]) == ref.ops.count(p.pops[i].domain)


def _sim_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), PHB_PEER_ARENA_MB="64")
    if world > 1:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import phare_b200.simulator as S
    from frontend_util import populate, const
    cells, dl = [32, 24], [0.25, 0.25]
    Lx = cells[0] * dl[0]
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=50, seed=None,  # unseeded -> the device (Philox) loader
               density=lambda x, y: 1.0 + 0.3 * np.sin(2 * np.pi * x / Lx), vx=const(0.1), vy=const(0), vz=const(0),
               vthx=const(0.3), vthy=const(0.3), vthz=const(0.3))
    populate(cells, dl, 1, [pop], [const(1.0), lambda x, y: 0.1 * np.cos(2 * np.pi * x / Lx), const(0.0)], largest=[16, 12])
    sim = S.make_simulator(S.make_hierarchy(), 2, 1, 2)
    sim.initialize()
    for _ in range(3):
        sim.advance(sim.timeStep())
    res = {}
    for p in sim.solver.patches:
        for name, h in (("Bz", p.B[2]), ("Ex", p.E[0]), ("Ne", p.Ne), ("Vx", p.Vi[0])):
            res[f"{p.geom.id}_{name}"] = sim.solver.ops.get_field(h)
        res[f"{p.geom.id}_n"] = np.array([sim.solver.ops.count(p.pops[0].domain)])
    np.savez(os.path.join(out_dir, f"sim_w{world}_rank{rank}.npz"), **res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_simulator_front_end_on_two_gpus_equals_one(tmp_path):
    """the dict-driven Simulator sharded over two ranks (device loader keyed by the global cell, peer-memory halo,
    NCCL migration) against the same dict on one GPU: same patches, same particles, fields equal to rounding"""
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_sim_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    mp.spawn(_sim_worker, args=(1, port + 1, str(tmp_path)), nprocs=1, join=True)
    one = dict(np.load(os.path.join(str(tmp_path), "sim_w1_rank0.npz")))
    two = {}
    for r in range(2):
        two.update(np.load(os.path.join(str(tmp_path), f"sim_w2_rank{r}.npz")))
    assert sorted(one) == sorted(two) and len(one) == 4 * 5
    for k, want in one.items():
        if k.endswith("_n"):
            assert int(two[k][0]) == int(want[0])
        else:
            scale = np.nanmax(np.abs(want)) + 1e-30
            assert np.allclose(two[k], want, rtol=0, atol=1e-11 * scale, equal_nan=True), k
