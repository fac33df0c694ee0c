"""Refined hierarchies on the GPU (phare_b200.amr over the C ABI: level operators of csrc/level.cu, phb_split, the
per-level PPC step) against the same orchestration on the CPU oracle, same initial particles and fields.
Tolerance as for the single-level step (north_star): moments and fields <= 1e-10 relative after N coarse steps."""
import numpy as np
import pytest

from amr_util import CASES, make_hierarchy, level_fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,pops", [("1d_o1", 1), ("1d_o2_td", 2), ("1d_o3_two_patches", 1), ("1d_o1_three_levels", 1),
                                       ("2d_o1", 2), ("2d_o2_L", 1), ("3d_o1", 1)])
def test_gpu_hierarchy_matches_cpu_oracle_hierarchy(cpu_ref, name, pops):
    from phare_b200.solver import GpuOps
    from oracle.cpu_ops import CpuOps
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    dim = len(cells)
    cpu = make_hierarchy(CpuOps(dim, interp), name, pops)
    gpu = make_hierarchy(GpuOps(dim, interp, "cuda:0"), name, pops)

    def compare(tag):
        a, b = level_fields(cpu), level_fields(gpu)
        assert a.keys() == b.keys()
        for key in a:
            x, y = a[key], b[key]
            assert np.array_equal(np.isnan(x), np.isnan(y)), (tag, key)
            ok = ~np.isnan(x) & np.isfinite(x)
            scale = np.max(np.abs(x[ok])) + 1e-30 if ok.any() else 1.0
            assert np.max(np.abs(x[ok] - y[ok]), initial=0.0) <= 1e-10 * scale + 1e-13, (tag, key)
        for lc, lg in zip(cpu.levels, gpu.levels):
            for pc, pg in zip(lc.solver.patches, lg.solver.patches):
                for i in range(pops):
                    for store in ("domain", "level_ghost", "level_ghost_old", "level_ghost_new"):
                        sc, sg = getattr(pc.pops[i], store), getattr(pg.pops[i], store)
                        if sc is not None:
                            assert cpu.ops.count(sc) == gpu.ops.count(sg), (tag, lc.number, pc.geom.id, i, store)

    compare("init")
    for step in range(2):
        cpu.advance(0.004)
        gpu.advance(0.004)
        compare(f"step {step}")


def _two_gpu_worker(rank, world, port, name, out_dir):
    import os
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PHB_PEER_ARENA_MB="64", PHB_PEER_TIMEOUT_S="10")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from phare_b200.messenger import TorchComm
    from phare_b200.solver import GpuOps
    from test_amr_gloo import _run, _collect
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    h = _run(GpuOps(len(cells), interp, f"cuda:{rank}"), name, comm=TorchComm(torch.device(f"cuda:{rank}")))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **_collect(h))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["1d_o1", "2d_o1", "3d_o1"])
def test_hierarchy_on_two_gpus_equals_one(cpu_ref, name, tmp_path):
    """patches of both levels dealt to two GPUs (root level: NVLink peer-memory halo; between levels and on the refined
    level: NCCL point-to-point) against the same hierarchy on one GPU.  Needs two GPUs: skipped on the single-GPU box."""
    import os
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from phare_b200.solver import GpuOps
    from test_amr_gloo import _run, _collect
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    mp.spawn(_two_gpu_worker, args=(2, 29800 + (os.getpid() % 1000), name, str(tmp_path)), nprocs=2, join=True)
    want = _collect(_run(GpuOps(len(cells), interp, "cuda:0"), name))
    got = {}
    for r in range(2):
        got.update(np.load(os.path.join(str(tmp_path), f"rank{r}.npz")))
    assert set(want) == set(got)
    for key, w in want.items():
        g = got[key]
        if key.endswith("_counts"):
            assert np.array_equal(g, w), key
            continue
        ok = ~np.isnan(w) & np.isfinite(w)
        scale = np.max(np.abs(w[ok])) + 1e-30 if ok.any() else 1.0
        assert np.array_equal(np.isnan(g), np.isnan(w)) and np.max(np.abs(g[ok] - w[ok]), initial=0.0) <= 1e-10 * scale, key
