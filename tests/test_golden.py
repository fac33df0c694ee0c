"""Known-answer tests from the reference's OWN golden generators (fixtures in tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference/tests/core/numerics/*/*.py).  The same assertions as the
reference's gtest files, at the same tolerances:
  faraday/test_main.cpp, ampere/test_main.cpp   1e-12 on physical nodes
  ohm/test_main.cpp                             1e-12 (1e-10 for the few hyper-resistive sums)
  interpolator/test_main.cpp:150-208            B-spline weights, EXPECT_DOUBLE_EQ (4 ulp)
  pusher/test_pusher.cpp:180-243                Boris trajectory vs scipy odeint, 1e-5
The CPU oracle is checked without a GPU; the CUDA kernels are checked with `-m gpu`."""
import os

import numpy as np
import pytest

from phare_b200 import abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = np.load(os.path.join(GOLD, "fields_golden.npz"))
CENT = {"Bx": "pdd", "By": "dpd", "Bz": "ddp", "Ex": "dpp", "Ey": "pdp", "Ez": "ppd", "Jx": "dpp", "Jy": "pdp", "Jz": "ppd"}
G = 2  # interp order 1


def layout_for(op, dim):
    nc = F[f"{op}_crop3d_ncells"] if dim == 3 else F[f"{op}_ncells"][:dim]
    return abi.make_layout(dim, 1, [int(x) for x in nc], [float(x) for x in F[f"{op}_dx"][:dim]])


def shape_of(L, name):
    return tuple(int(L.ncells[d]) + (1 if CENT.get(name, "ppp")[d] == "p" else 0) + 2 * G for d in range(L.dim))


def get(op, dim, name, L):
    k = f"{op}_{dim}d_{name}"
    return np.ascontiguousarray(F[k]) if k in F.files else np.zeros(shape_of(L, name))


def physical(L, name, a):
    sl = tuple(slice(G, G + int(L.ncells[d]) + (1 if CENT.get(name, "ppp")[d] == "p" else 0)) for d in range(L.dim))
    return a[sl]


def check_fields(run, tol=1e-12):
    for dim in (1, 2, 3):
        # Faraday
        L = layout_for("faraday", dim)
        E = [get("faraday", dim, n, L) for n in ("Ex", "Ey", "Ez")]
        B = [get("faraday", dim, n, L) for n in ("Bx", "By", "Bz")]
        got = run("faraday", L, dict(E=E, B=B, dt=float(F["faraday_dt"])))
        for c, n in enumerate(("Bx", "By", "Bz")):
            if f"faraday_{dim}d_{n}New" in F.files:
                want = F[f"faraday_{dim}d_{n}New"]
                assert np.max(np.abs(physical(L, n, got[c]) - physical(L, n, want))) <= tol, ("faraday", dim, n)
        # Ampere
        L = layout_for("ampere", dim)
        B = [get("ampere", dim, n, L) for n in ("Bx", "By", "Bz")]
        got = run("ampere", L, dict(B=B))
        for c, n in enumerate(("Jx", "Jy", "Jz")):
            if f"ampere_{dim}d_{n}" in F.files:
                want = F[f"ampere_{dim}d_{n}"]
                assert np.max(np.abs(physical(L, n, got[c]) - physical(L, n, want))) <= tol, ("ampere", dim, n)
        # Ohm (eta = 1, nu = 0.01, constant hyper-resistivity)
        L = layout_for("ohm", dim)
        args = dict(n=get("ohm", dim, "n", L), Pe=get("ohm", dim, "P", L),
                    Ve=[get("ohm", dim, n, L) for n in ("Vx", "Vy", "Vz")],
                    B=[get("ohm", dim, n, L) for n in ("Bx", "By", "Bz")],
                    J=[get("ohm", dim, n, L) for n in ("Jx", "Jy", "Jz")],
                    eta=float(F[f"ohm_{dim}d_eta"]), nu=float(F[f"ohm_{dim}d_nu"]))
        got = run("ohm", L, args)
        for c, n in enumerate(("Ex", "Ey", "Ez")):
            want = F[f"ohm_{dim}d_{n}New"]
            err = np.max(np.abs(physical(L, n, got[c]) - physical(L, n, want)))
            assert err <= 1e-10 and np.mean(np.abs(physical(L, n, got[c]) - physical(L, n, want)) <= 1e-12) > 0.99, \
                ("ohm", dim, n, err)


def test_field_solvers_oracle(cpu_oracle):
    def run(op, L, a):
        if op == "faraday":
            return cpu_oracle.faraday(L, a["B"], a["E"], a["dt"])
        if op == "ampere":
            return cpu_oracle.ampere(L, a["B"])
        return cpu_oracle.ohm(L, a["n"], a["Ve"], a["Pe"], a["B"], a["J"], a["eta"], a["nu"], 0)
    check_fields(run)


@pytest.mark.gpu
def test_field_solvers_gpu():
    from phare_b200.device import Context, DeviceArray, DeviceVec
    ctxs = {d: Context(d, 1) for d in (1, 2, 3)}

    def run(op, L, a):
        ctx = ctxs[L.dim]
        up = lambda q0, arrs: DeviceVec(ctx, L, q0, arrs)
        if op == "faraday":
            out = DeviceVec(ctx, L, abi.BX)
            ctx.faraday(L, up(abi.BX, a["B"]), up(abi.EX, a["E"]), out, a["dt"])
        elif op == "ampere":
            out = DeviceVec(ctx, L, abi.JX)
            ctx.ampere(L, up(abi.BX, a["B"]), out)
        else:
            out = DeviceVec(ctx, L, abi.EX)
            n = DeviceArray(ctx, a["n"].shape).upload(a["n"])
            pe = DeviceArray(ctx, a["Pe"].shape).upload(a["Pe"])
            ctx.ohm(L, n, up(abi.VX, a["Ve"]), pe, up(abi.BX, a["B"]), up(abi.JX, a["J"]), out, a["eta"], a["nu"], 0)
        return out.download()
    check_fields(run)
    for c in ctxs.values():
        c.close()


def test_bspline_weights_oracle(cpu_oracle):
    """interpolator/test_main.cpp:150-208: 10 positions x = 3 + 0.1 i, primal and dual, orders 1..3"""
    bs = np.load(os.path.join(GOLD, "bsplines_golden.npz"))
    for order in (1, 2, 3):
        for ci, centering in enumerate(("primal", "dual")):
            for ipos in range(10):
                delta = float(ipos) * 0.1
                start, w = cpu_oracle.weights(order, ci, 3, delta)
                assert start == int(bs[f"nodes_{order}_{centering}"][ipos][0])
                want = bs[f"weights_{order}_{centering}"][ipos]
                for a, b in zip(w, want):  # EXPECT_DOUBLE_EQ: within 4 ulp
                    assert abs(a - b) <= 4 * np.spacing(max(abs(a), abs(b), 1e-300)) or abs(a - b) < 1e-16, (order, centering, ipos)


def test_weights_sum_to_one(cpu_oracle):
    """interpolator/test_main.cpp:92-99"""
    rng = np.random.default_rng(0)
    for order in (1, 2, 3):
        for dual in (0, 1):
            for _ in range(2000):
                _, w = cpu_oracle.weights(order, dual, int(rng.integers(5, 100)), float(rng.random()))
                assert abs(sum(w) - 1.0) < 1e-10


def boris_trajectory(push_steps, nsteps):
    """pusher/test_pusher.cpp:180-243: uniform E = (0.01,-0.05,0.05), B = (1,1,1), dx = 0.05, dt = 1e-4,
    start x = 0.25, v = (0,10,0); x(t) and v(t) against the odeint solution"""
    g = np.load(os.path.join(GOLD, "pusher_golden.npz"))
    traj, stride, dt = g["trajectory"], int(g["stride"]), float(g["dt"])
    ncell, dx = 8000, 0.05
    L = abi.make_layout(1, 1, [ncell], [dx])
    shape = lambda q: (ncell + (1 if q in (abi.EY, abi.EZ, abi.BX) else 0) + 4,)
    E = [np.full(shape(abi.EX + c), v) for c, v in enumerate((0.01, -0.05, 0.05))]
    B = [np.full(shape(abi.BX + c), 1.0) for c in range(3)]
    x0 = 0.25 + ncell // 2 * dx
    icell, delta = int(x0 / dx), x0 / dx - int(x0 / dx)
    xs, vs = push_steps(L, E, B, icell, delta, [0.0, 10.0, 0.0], dt, nsteps, stride)
    nchk = nsteps // stride
    want = traj[1:nchk + 1]
    got_x = np.array(xs) * dx - ncell // 2 * dx
    assert np.max(np.abs(got_x - want[:, 0])) < 1e-5
    assert np.max(np.abs(np.array(vs) - want[:, 3:6])) < 2e-4  # the reference only asserts positions


def test_boris_trajectory_oracle(cpu_oracle):
    from oracle import HostParticles

    def steps(L, E, B, icell, delta, v, dt, n, stride):
        P = HostParticles.from_soa(np.array([[icell]], np.int32), np.array([[delta]]), np.ones(1), np.ones(1),
                                   np.array([v]))
        xs, vs = [], []
        for s in range(1, n + 1):
            rc, P = cpu_oracle.push(L, E, B, P, 1.0, dt, pout=P)
            assert rc == 0
            if s % stride == 0:
                xs.append(P.icell[0][0] + P.delta[0][0])
                vs.append([P.v[k][0] for k in range(3)])
        return xs, vs
    boris_trajectory(steps, 20000)


@pytest.mark.gpu
def test_boris_trajectory_gpu():
    from phare_b200.device import Context, DeviceVec, DeviceParticles

    def steps(L, E, B, icell, delta, v, dt, n, stride):
        ctx = Context(1, 1)
        dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
        P = DeviceParticles(ctx, 1).upload_soa(np.array([[icell]], np.int32), np.array([[delta]]), np.ones(1),
                                               np.ones(1), np.array([v]))
        xs, vs = [], []
        for s in range(1, n + 1):
            ctx.push(L, dE, dB, P, P, 1.0, dt)
            if s % stride == 0:
                ic, de, _, _, vv = P.download_soa()
                xs.append(ic[0, 0] + de[0, 0])
                vs.append(vv[0].tolist())
        ctx.poll_error()
        ctx.close()
        return xs, vs
    boris_trajectory(steps, 20000)
