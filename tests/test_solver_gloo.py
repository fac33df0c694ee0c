"""N>1 path on CPU: two processes (gloo), patches dealt to ranks, halo phases and particle migration through
torch.distributed point-to-point.  The 2-rank result must equal the 1-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phare_b200.messenger import TorchComm
    from oracle.cpu_ops import CpuOps
    from solver_util import global_particles, make_solver, FIELDS
    domain, grid, interp, dx, ppc, npop, steps = case
    dim = len(domain)
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    comm = TorchComm(torch.device("cpu"))
    s = make_solver(CpuOps(dim, interp), domain, grid, interp, dx, gparts, comm=comm)
    for _ in range(steps):
        s.advance_level(0.005)
    res = {}
    for p in s.patches:
        for attr, comp, qty in FIELDS:
            h = getattr(p, attr)
            res[f"{p.geom.id}_{attr}_{comp}"] = s.ops.get_field(h[comp] if comp is not None else h)
        for i in range(npop):
            res[f"{p.geom.id}_n{i}"] = np.array([s.ops.count(p.pops[i].domain)])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", [
    ((48,), (4,), 1, (0.2,), 20, 1, 4),
    ((16, 12), (2, 2), 1, (0.4, 0.4), 8, 2, 3),
    ((8, 8, 8), (2, 1, 2), 1, (0.2, 0.2, 0.2), 4, 1, 2),
])
def test_two_ranks_equal_one_process(case, tmp_path):
    sys.path.insert(0, HERE)
    from oracle.cpu_ops import CpuOps
    from solver_util import global_particles, make_solver, FIELDS
    domain, grid, interp, dx, ppc, npop, steps = case
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    ref = make_solver(CpuOps(len(domain), interp), domain, grid, interp, dx, gparts)
    for _ in range(steps):
        ref.advance_level(0.005)
    got = {}
    for r in range(2):
        got.update(np.load(os.path.join(str(tmp_path), f"rank{r}.npz")))
    seen = 0
    for p in ref.patches:
        for attr, comp, qty in FIELDS:
            h = getattr(p, attr)
            want = ref.ops.get_field(h[comp] if comp is not None else h)
            g = got[f"{p.geom.id}_{attr}_{comp}"]
            # identical plans and arithmetic; only the order of the border sums may differ
            scale = np.nanmax(np.abs(want)) + 1e-30
            assert np.allclose(g, want, rtol=0, atol=1e-12 * scale, equal_nan=True), (p.geom.id, attr, comp)
            seen += 1
        for i in range(npop):
            assert int(got[f"{p.geom.id}_n{i}"][0]) == ref.ops.count(p.pops[i].domain)
    assert seen == len(ref.patches) * len(FIELDS)
