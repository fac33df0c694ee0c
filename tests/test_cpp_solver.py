"""The C++ level driver (include/phare_b200/solver_ppc.hpp: SolverPPC + PeriodicMessenger over the operator mirror,
CUDA reached only through the C ABI): compiles everywhere; on a GPU box N steps of a one-patch periodic problem
must match the Python-driven step (same kernels and plans) and the CPU oracle step (north_star bound 1e-10)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from phare_b200 import abi
from solver_util import global_particles, make_solver, B_init, FIELDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_solver")


def build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp/test_solver.cpp"),
           "-o", EXE, "-L" + os.path.join(ROOT, "phare_b200/lib"), "-lphare_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "phare_b200/lib"), "-Wl,-rpath,/usr/local/cuda/lib64",
           "-L/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_solver_compiles_and_links():
    build()
    r = subprocess.run([EXE, "--compile-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "compiled" in r.stdout


def patch_layouts(domain, grid, interp, dx):
    from phare_b200.solver import make_level
    return make_level(domain, grid, interp, dx)[1]


def write_problem(path, domain, interp, dx, gparts, masses, steps, dt, eta, nu, Te, grid=None):
    from phare_b200.messenger import centering
    from phare_b200.setup import node_coords
    dim = len(domain)
    grid = grid or (1,) * dim
    with open(path, "wb") as f:
        f.write(struct.pack("<4i3I3I3d4d", dim, interp, steps, len(gparts), *(list(domain) + [0] * (3 - dim)),
                            *(list(grid) + [1] * (3 - dim)), *(list(dx) + [0.0] * (3 - dim)), dt, eta, nu, Te))
        bfn = B_init(domain, dx)
        for L in patch_layouts(domain, grid, interp, dx):
            for c in range(3):
                mesh = np.meshgrid(*node_coords(L, abi.BX + c, centering, domain), indexing="ij")
                f.write(np.ascontiguousarray(np.broadcast_to(bfn(c, *mesh), mesh[0].shape), dtype=np.float64).tobytes())
        rec = np.dtype(dict(names=["weight", "charge", "icell", "delta", "v"],
                            formats=["<f8", "<f8", ("<i4", (dim,)), ("<f8", (dim,)), ("<f8", (3,))],
                            offsets=[0, 8, 16, 16 + 4 * dim + (4 * dim) % 8, 16 + 4 * dim + (4 * dim) % 8 + 8 * dim],
                            itemsize={1: 56, 2: 64, 3: 80}[dim]))
        for (icell, delta, w, q, v), mass in zip(gparts, masses):
            a = np.zeros(len(w), rec)
            a["weight"], a["charge"], a["icell"], a["delta"], a["v"] = w, q, icell, delta, v
            f.write(struct.pack("<dQ", mass, len(w)))
            f.write(a.tobytes())


def read_result(path, domain, interp, npop, grid=None, dx=None):
    """per patch: {(attr, comp): array}, [particle count per population]"""
    import ctypes as C
    dim = len(domain)
    grid = grid or (1,) * dim
    lib = abi.load()
    raw, off, out = open(path, "rb").read(), 0, []
    for L in patch_layouts(domain, grid, interp, dx or [1.0] * dim):
        rec = {}
        for attr, comp, qty in [("B", 0, abi.BX), ("B", 1, abi.BY), ("B", 2, abi.BZ), ("E", 0, abi.EX), ("E", 1, abi.EY),
                                ("E", 2, abi.EZ), ("Ne", None, abi.RHO), ("Vi", 0, abi.VX), ("Vi", 1, abi.VY), ("Vi", 2, abi.VZ)]:
            s = (C.c_uint32 * 3)()
            n = lib.phb_field_shape(C.byref(L), qty, s)
            rec[(attr, comp)] = np.frombuffer(raw, np.float64, n, off).reshape([s[d] for d in range(dim)])
            off += 8 * n
        counts = np.frombuffer(raw, np.uint64, npop, off)
        off += 8 * npop
        out.append((rec, [int(c) for c in counts]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("domain,interp,dx,ppc,npop,steps", [
    ((64,), 1, (0.2,), 50, 2, 5),
    ((48,), 2, (0.25,), 40, 1, 4),
    ((24, 16), 1, (0.4, 0.4), 20, 2, 3),
    ((16, 12), 3, (0.2, 0.2), 12, 1, 2),
    ((12, 8, 8), 1, (0.2, 0.2, 0.2), 8, 1, 3),
])
def test_cpp_solver_matches_python_driver_and_oracle(domain, interp, dx, ppc, npop, steps, tmp_path):
    from phare_b200.solver import GpuOps
    from oracle.cpu_ops import CpuOps
    build()
    dim = len(domain)
    masses = (1.0, 2.0)[:npop]
    kw = dict(resistivity=1e-3, hyper_resistivity=1e-3, Te=0.12)
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    prob, res = str(tmp_path / "problem.bin"), str(tmp_path / "result.bin")
    write_problem(prob, domain, interp, dx, gparts, masses, steps, 0.005, kw["resistivity"], kw["hyper_resistivity"], kw["Te"])
    r = subprocess.run([EXE, prob, res], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    (got, counts), = read_result(res, domain, interp, npop)
    assert counts == [len(g[2]) for g in gparts]  # one periodic patch: the particle number is conserved
    for name, ops, tol in (("python driver", GpuOps(dim, interp, "cuda:0"), 1e-11), ("oracle", CpuOps(dim, interp), 1e-10)):
        s = make_solver(ops, domain, (1,) * dim, interp, dx, gparts, masses=masses, solver_kw=kw)
        for _ in range(steps):
            s.advance_level(0.005)
        p = s.patches[0]
        for attr, comp, qty in FIELDS:
            h = getattr(p, attr)
            want = s.ops.get_field(h[comp] if comp is not None else h)
            ok = np.isfinite(want)
            assert np.array_equal(np.isfinite(got[(attr, comp)]), ok), (name, attr, comp)
            scale = np.max(np.abs(want[ok])) + 1e-30
            assert np.max(np.abs(got[(attr, comp)][ok] - want[ok])) <= tol * scale + 1e-13, (name, attr, comp)


@pytest.mark.gpu
@pytest.mark.parametrize("domain,grid,interp,dx,ppc,npop,steps", [
    ((64,), (1,), 1, (0.2,), 50, 2, 5),
    ((64,), (4,), 1, (0.2,), 50, 2, 5),
    ((60,), (3,), 2, (0.25,), 40, 1, 4),
    ((24, 16), (1, 1), 1, (0.4, 0.4), 20, 2, 3),
    ((24, 16), (2, 2), 1, (0.4, 0.4), 20, 2, 4),
    ((16, 16), (2, 1), 3, (0.2, 0.2), 12, 2, 3),
    ((12, 8, 8), (1, 1, 1), 1, (0.2, 0.2, 0.2), 8, 1, 3),
    ((12, 8, 8), (2, 1, 2), 1, (0.2, 0.2, 0.2), 8, 1, 3),
])
def test_cpp_solver_matches_step_oracle(domain, grid, interp, dx, ppc, npop, steps, tmp_path):
    """the C++ SolverPPC (its own LevelMessenger box algebra in solver_ppc.hpp, CUDA through the C ABI), one patch and
    several, against the STEP ORACLE (the reference's own functors under an independent level loop,
    oracle/ref/ref_step.cpp): fields and moments <= 1e-10 per node, particle counts per patch exact"""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libphare_ref.so not built")
    from oracle import ref_step as rs
    from phare_b200.messenger import centering
    from phare_b200.setup import node_coords
    from step_oracle_util import node_errors, physical
    build()
    dim = len(domain)
    masses = (1.0, 2.0)[:npop]
    eta, nu, Te, dt = 1e-3, 1e-3, 0.12, 0.005
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    prob, res = str(tmp_path / "problem.bin"), str(tmp_path / "result.bin")
    write_problem(prob, domain, interp, dx, gparts, masses, steps, dt, eta, nu, Te, grid)
    r = subprocess.run([EXE, prob, res], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    result = read_result(res, domain, interp, npop, grid, dx)
    layouts = patch_layouts(domain, grid, interp, dx)
    boxes = [([L.amr_lower[d] for d in range(dim)], [L.amr_lower[d] + L.ncells[d] - 1 for d in range(dim)]) for L in layouts]
    ref = rs.RefStep(dim, interp, boxes, dx, [0.0] * dim, domain, masses, Te=Te, eta=eta, nu=nu)
    bfn = B_init(domain, dx)
    for ip, L in enumerate(layouts):
        B0 = []
        for c in range(3):
            mesh = np.meshgrid(*node_coords(L, abi.BX + c, centering, domain), indexing="ij")
            B0.append(np.ascontiguousarray(np.broadcast_to(bfn(c, *mesh), mesh[0].shape)))
        ref.set_vec(ip, rs.B, B0)
        for i, (icell, delta, w, q, v) in enumerate(gparts):
            inside = np.ones(len(w), bool)
            for d in range(dim):
                inside &= (icell[:, d] >= boxes[ip][0][d]) & (icell[:, d] <= boxes[ip][1][d])
            ref.set_particles(ip, i, icell[inside], delta[inside], w[inside], q[inside], v[inside])
    ref.initialize()
    for _ in range(steps):
        ref.advance(dt)
    for ip, (L, (got, counts)) in enumerate(zip(layouts, result)):
        assert counts == [ref.count(ip, i) for i in range(npop)]
        want = {("B", c): ref.vec(ip, rs.B)[c] for c in range(3)}
        want.update({("E", c): ref.vec(ip, rs.E)[c] for c in range(3)})
        want.update({("Vi", c): ref.vec(ip, rs.VI)[c] for c in range(3)})
        want[("Ne", None)] = ref.scalar(ip, rs.NI)
        for attr, comp, qty in FIELDS:
            a, b = got[(attr, comp)], want[(attr, comp)]
            if attr in ("Ne", "Vi"):
                a, b = physical(a, L, qty, centering), physical(b, L, qty, centering)
            assert node_errors(a, b) <= 1e-10, (ip, attr, comp, node_errors(a, b))
