"""The C++ level driver spread over two GPUs (include/phare_b200/solver_ppc.hpp: LevelMessenger::Distribution, NVLink peer
memory between the ranks, migration and the error vote in message headers; launched through libphare_b200_host.so, one
process per GPU) against the Python driver holding the whole level in one process: same patches, same particles (device
loader keyed by the global cell), fields equal to rounding, particle counts per patch and population equal.
Needs >= 2 GPUs: skipped on the single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu

CASES = [(1, (1024,), (4,), 4), (3, (64, 32), (2, 2), 3), (4, (32, 32), (2, 2), 2), (5, (16, 16, 16), (2, 2, 1), 3)]


def _cfg(k, cells, grid):
    from phare_b200 import configs
    cfg = configs.get(k).with_cells(cells, grid)
    cfg.pops = [dict(p, ppc=min(p["ppc"], 24)) for p in cfg.pops]
    return cfg


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PHB_PEER_ARENA_MB="256", PHB_PEER_TIMEOUT_S="20",
                      PHB_PEER_MIGRATION_CAP="65536")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from phare_b200 import abi, host_cpp
    from phare_b200.messenger import TorchComm
    k, cells, grid, steps = case
    cfg = _cfg(k, cells, grid)
    level = host_cpp.CppLevel(cfg, f"cuda:{rank}", comm=TorchComm(torch.device(f"cuda:{rank}")))
    level.initialize()
    level.advance(cfg.dt, steps)
    res = {}
    counts = level.counts()
    for ip, pid in enumerate(level.patch_ids):
        for which, name, q0 in ((host_cpp.B, "B", abi.BX), (host_cpp.E, "E", abi.EX), (host_cpp.VI, "Vi", abi.VX)):
            for c in range(3):
                res[f"{pid}_{name}{c}"] = level.get_field(ip, which, c, q0 + c)
        res[f"{pid}_Ni"] = level.get_field(ip, host_cpp.NI, 0, abi.RHO)
        res[f"{pid}_counts"] = np.array(counts[ip])
    np.savez(os.path.join(out_dir, f"cpp_rank{rank}.npz"), **res)
    dist.barrier()
    level.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("case", CASES, ids=[f"config{c[0]}" for c in CASES])
def test_cpp_level_on_two_gpus_equals_python_level_on_one(case, tmp_path):
    from phare_b200 import configs
    from phare_b200.messenger import LocalComm
    from phare_b200.solver import GpuOps
    world = 2
    port = 29800 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    k, cells, grid, steps = case
    cfg = _cfg(k, cells, grid)
    py = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
    for _ in range(steps):
        py.advance_level(cfg.dt)
    got = {}
    for r in range(world):
        got.update(np.load(os.path.join(str(tmp_path), f"cpp_rank{r}.npz")))
    assert len([key for key in got if key.endswith("_counts")]) == len(py.patches)
    for p in py.patches:
        pid = p.geom.id
        assert list(got[f"{pid}_counts"]) == [py.ops.count(pop.domain) for pop in p.pops]
        for name, attr in (("B", "B"), ("E", "E"), ("Vi", "Vi")):
            for c in range(3):
                want = py.ops.get_field(getattr(p, attr)[c])
                g = got[f"{pid}_{name}{c}"]
                ok = np.isfinite(want)
                assert np.array_equal(np.isfinite(g), ok)
                scale = np.max(np.abs(want[ok])) + 1e-300
                assert np.max(np.abs(g[ok] - want[ok])) <= 1e-11 * scale + 1e-14, (pid, name, c)
        want = py.ops.get_field(p.Ne)
        assert np.max(np.abs(got[f"{pid}_Ni"] - want)) <= 1e-11 * np.max(np.abs(want))
