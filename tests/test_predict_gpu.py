"""GPU parity of the predicted re-binning (csrc/predict.cu) through the C ABI against the CPU oracle.

phb_push_deposit_predict (the domain_only sweep + the plan) then phb_push_deposit_rebin (the all sweep in one pass) must
leave what the oracle's push -> deposit -> bin leaves: cell_start and class counts exact, per-cell multisets bit-exact,
moments of BOTH sweeps <= 1e-10.  The plan is made with the first sweep's fields and carried out with the second's:
 * same fields: every plan holds whatever eps is (0: nobody is risky; 0.6: everybody is);
 * slightly different fields (what a PPC step does): plans hold thanks to the risky list (misfiled == 0);
 * very different fields and eps = 0: plans fail, are COUNTED, and phb_bin of the result restores the exact order."""
import os

import numpy as np
import pytest

from phare_b200 import abi
from phare_b200.device import Context, DeviceArray, DeviceVec, DeviceParticles, PhbError
from oracle import HostParticles, canonical_rows
from util import random_vec, domain_box, grown, particle_ghosts
from test_tile_gpu import binned_store, layout_for, moments, close, SHAPES

pytestmark = pytest.mark.gpu
TILED = [(1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1)]


@pytest.fixture(scope="module")
def ctxs():
    cache = {}

    def get(dim, interp):
        if (dim, interp) not in cache:
            cache[(dim, interp)] = Context(dim, interp)
        return cache[(dim, interp)]

    yield get
    for c in cache.values():
        c.close()


class eps_env:
    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self.old = os.environ.get("PHB_PREDICT_EPS")
        if self.eps is not None:
            os.environ["PHB_PREDICT_EPS"] = str(self.eps)

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("PHB_PREDICT_EPS", None)
        else:
            os.environ["PHB_PREDICT_EPS"] = self.old


def problem(cpu, dim, interp, seed, tail=257, vth=1.0):
    rng = np.random.default_rng(seed + 10 * dim + interp)
    L = layout_for(dim, interp)
    ncell = int(np.prod(SHAPES[dim]))
    soa, n_sorted, cs = binned_store(cpu, rng, L, ncell * (12 if dim < 3 else 6), vth=vth, tail=tail)
    np.clip(soa[4], -3.4, 3.4, out=soa[4])  # at most one cell per half step (h <= 0.25 at dt = 0.1)
    # the particles of the outermost x layers leave the patch: through the upper face (new patch ghosts, kept) and through
    # the lower one (erased: the keep box below stops at the domain's lower x face)
    lo, hi = domain_box(L).lower[0], domain_box(L).upper[0]
    soa[4][soa[0][:, 0] >= hi - 1, 0] = 3.0
    soa[4][soa[0][:, 0] <= lo + 1, 0] = -3.0
    dom = domain_box(L)
    G = grown(dom, dim, particle_ghosts(interp))
    keep_lo = [G.lower[d] for d in range(dim)]
    keep_lo[0] = dom.lower[0]  # leavers through the lower x face are dropped (erased), the others are new patch ghosts
    keep = [abi.make_box(keep_lo, [G.upper[d] for d in range(dim)])]
    E = random_vec(rng, cpu.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu.field_shape, L, abi.BX, 0.3)
    return rng, L, soa, n_sorted, cs, dom, keep, E, B


def run_sweeps(ctx, cpu, L, soa, n_sorted, cs, dom, keep, EB1, EB2, dt, eps):
    """the two sweeps on the device; returns (moments sweep 1, moments sweep 2, cs_new, counts + misfiled, out, in after)"""
    n = len(soa[2])
    pin, pout = DeviceParticles(ctx, n + 100).upload_soa(*soa), DeviceParticles(ctx, n + 100)
    cs_old = DeviceArray(ctx, cs.shape, np.uint32).upload(cs)
    cs_new = DeviceArray(ctx, cs.shape, np.uint32)
    nbytes = ctx.predict_plan_bytes(L, dom, n + 100)
    plan = DeviceArray(ctx, ((nbytes + 3) // 4,), np.uint32)
    out = []
    with eps_env(eps):
        for k, (E, B) in enumerate((EB1, EB2)):
            (rn, rq), F = moments(ctx, L, cpu)
            dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
            if k == 0:
                ctx.push_deposit_predict(L, dE, dB, pin, n_sorted, 1.0, dt, rn, rq, F, 0.9, keep, dom, cs_old, keep,
                                         plan.ptr, nbytes)
            else:
                ctx.push_deposit_rebin(L, dE, dB, pin, n_sorted, 1.0, dt, rn, rq, F, 0.9, keep, dom, cs_old, keep, pout,
                                       cs_new, plan.ptr, nbytes)
            out.append([rn.download(), rq.download()] + F.download())
    counts = ctx.predict_counts(L, dom, cs_new, plan.ptr, pout)
    return out[0], out[1], cs_new, counts, pout, pin


def oracle_sweep(cpu, L, soa, E, B, dt, dom, keep):
    rc, pushed = cpu.push(L, E, B, HostParticles.from_soa(*soa), 1.0, dt)
    assert rc == 0
    return pushed, cpu.deposit(L, pushed, coef=0.9, sel=keep), cpu.bin(L, pushed, dom, keep)


def check_exact(ctx, L, dom, got_cs, counts, pout, want, want_cs, want_counts):
    assert counts[:3] == want_counts and pout.n == counts[0] + counts[1]
    assert np.array_equal(got_cs.download(), want_cs)
    got = pout.download_soa()  # [domain | new patch ghosts]: what phb_bin_counts / phb_predict_counts left in out->n
    assert np.array_equal(got[0], want.soa()[0])
    assert np.array_equal(canonical_rows(*got), canonical_rows(*want.soa()))


@pytest.fixture(params=["blocks_64_cells", "blocks_128_cells"])
def block_size(request):
    """the test patches are smaller than one wave of 128-cell blocks, so the kernels cut them into 64-cell blocks;
    PHB_TILE_BIG=1 (read at every call) keeps the 128-cell blocks the benchmark sizes run with"""
    old = os.environ.get("PHB_TILE_BIG")
    if request.param == "blocks_128_cells":
        os.environ["PHB_TILE_BIG"] = "1"
    else:
        os.environ.pop("PHB_TILE_BIG", None)
    yield request.param
    if old is None:
        os.environ.pop("PHB_TILE_BIG", None)
    else:
        os.environ["PHB_TILE_BIG"] = old


@pytest.mark.parametrize("eps", [None, 0, 0.6], ids=["eps_default", "eps_0", "all_risky"])
@pytest.mark.parametrize("dim,interp", TILED)
def test_same_fields_every_plan_holds(ctxs, cpu_oracle, dim, interp, eps, block_size):
    ctx = ctxs(dim, interp)
    rng, L, soa, n_sorted, cs, dom, keep, E, B = problem(cpu_oracle, dim, interp, 7000)
    dt = 0.1
    pushed, want_m, (want, want_cs, want_counts) = oracle_sweep(cpu_oracle, L, soa, E, B, dt, dom, keep)
    assert want_counts[1] > 0 and want_counts[2] > 0
    m1, m2, cs_new, counts, pout, pin = run_sweeps(ctx, cpu_oracle, L, soa, n_sorted, cs, dom, keep, (E, B), (E, B), dt, eps)
    ctx.poll_error()
    assert counts[3] == 0
    close(m1, want_m)
    close(m2, want_m)
    check_exact(ctx, L, dom, cs_new, counts, pout, want, want_cs, want_counts)
    for g, w in zip(pin.download_soa(), soa):  # neither sweep writes the input store
        assert np.array_equal(g, w)


@pytest.mark.parametrize("dim,interp", TILED)
def test_corrected_fields_plans_hold_through_the_risky_list(ctxs, cpu_oracle, dim, interp, block_size):
    """the second sweep sees fields that differ by 1e-5 (positions by ~1e-8 of a cell, as between the two sweeps of a PPC
    step): a few hundred of the 10^4..10^5 particles sit within 2^-12 of a face, and with eps = 0 some of them would be
    misfiled; with the default eps none is"""
    ctx = ctxs(dim, interp)
    rng, L, soa, n_sorted, cs, dom, keep, E, B = problem(cpu_oracle, dim, interp, 7100)
    # crowd the faces: a tenth of the particles gets a delta that the push leaves within 1e-7 of a face
    E2 = [e + 1e-5 * rng.standard_normal(e.shape) for e in E]
    B2 = [b + 1e-5 * rng.standard_normal(b.shape) for b in B]
    dt = 0.1
    _, want_m1, _ = oracle_sweep(cpu_oracle, L, soa, E, B, dt, dom, keep)
    pushed, want_m2, (want, want_cs, want_counts) = oracle_sweep(cpu_oracle, L, soa, E2, B2, dt, dom, keep)
    m1, m2, cs_new, counts, pout, _ = run_sweeps(ctx, cpu_oracle, L, soa, n_sorted, cs, dom, keep, (E, B), (E2, B2), dt, None)
    ctx.poll_error()
    close(m1, want_m1)
    close(m2, want_m2)
    assert counts[3] == 0
    check_exact(ctx, L, dom, cs_new, counts, pout, want, want_cs, want_counts)


@pytest.mark.parametrize("dim,interp", [(1, 1), (2, 1), (2, 3), (3, 1)])
def test_failed_plans_are_counted_and_phb_bin_restores_the_order(ctxs, cpu_oracle, dim, interp):
    ctx = ctxs(dim, interp)
    rng, L, soa, n_sorted, cs, dom, keep, E, B = problem(cpu_oracle, dim, interp, 7200)
    E2 = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)  # unrelated fields: many plans fail
    B2 = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    dt = 0.1
    pushed, want_m2, (want, want_cs, want_counts) = oracle_sweep(cpu_oracle, L, soa, E2, B2, dt, dom, keep)
    m1, m2, cs_new, counts, pout, _ = run_sweeps(ctx, cpu_oracle, L, soa, n_sorted, cs, dom, keep, (E, B), (E2, B2), dt, 0)
    ctx.poll_error()
    assert counts[3] > 0                      # plans failed and were counted ...
    close(m2, want_m2)                        # ... the moments are those of the particles' actual positions ...
    n = len(soa[2])
    assert sum(counts[:3]) == n               # ... nobody is lost ...
    pout.n = n
    got = pout.download_soa()
    assert np.array_equal(canonical_rows(*got), canonical_rows(*pushed.soa()))  # ... and every particle was pushed right
    # the caller's remedy (IonUpdater.maintain_arrays): phb_bin of the result
    fixed = DeviceParticles(ctx, n + 100)
    cs_fix = DeviceArray(ctx, cs.shape, np.uint32)
    c2 = ctx.bin(L, pout, fixed, dom, keep, cs_fix)
    assert c2 == want_counts
    assert np.array_equal(cs_fix.download(), want_cs)
    fixed.n = c2[0] + c2[1]
    assert np.array_equal(canonical_rows(*fixed.download_soa()), canonical_rows(*want.soa()))


def test_move_two_cells_is_reported_and_nobody_is_lost(ctxs, cpu_oracle):
    ctx = ctxs(2, 1)
    rng, L, soa, n_sorted, cs, dom, keep, E, B = problem(cpu_oracle, 2, 1, 7300, tail=50, vth=0.1)
    soa[4][1234, 0] = 1000.0  # one runaway in the ordered part
    soa[4][-3, 1] = -1000.0   # and one in the tail
    n = len(soa[2])
    m1, m2, cs_new, counts, pout, _ = run_sweeps(ctx, cpu_oracle, L, soa, n_sorted, cs, dom, keep, (E, B), (E, B), 0.1, None)
    with pytest.raises(PhbError) as e:
        ctx.poll_error()
    assert e.value.code == abi.PHB_ERR_MOVE_TWO_CELL
    assert sum(counts[:3]) == n and counts[3] == 0


@pytest.mark.parametrize("k,cells,grid", [(1, (512,), (2,)), (3, (64, 32), (2, 2)), (5, (16, 16, 16), (2, 1, 1))])
def test_step_with_forced_repair_equals_step_without_prediction(k, cells, grid, monkeypatch):
    """IonUpdater.maintain_arrays' remedy for failed plans (phb_bin of the re-binned store) driven from the solver: every second
    step the count of failed plans is reported as 1 although every plan held; the repaired run must equal the run without
    prediction (same particles in the same cells, fields to rounding) and the run that never repairs"""
    from phare_b200 import configs
    from phare_b200.messenger import LocalComm
    from phare_b200.solver import GpuOps

    def build(predict, force):
        cfg = configs.get(k).with_cells(cells, grid)
        cfg.pops = [dict(p, ppc=min(p["ppc"], 16)) for p in cfg.pops]
        s = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
        s.updater.predict = predict
        if force:
            real, calls = s.ops.predict_counts, [0]

            def lying(*a):
                c = real(*a)
                calls[0] += 1
                return c[:3] + ((1,) if calls[0] % 2 == 0 else (c[3],))
            monkeypatch.setattr(s.ops, "predict_counts", lying)
        return cfg, s

    runs = []
    for predict, force in ((True, True), (True, False), (False, False)):
        cfg, s = build(predict, force)
        for _ in range(4):
            s.advance_level(cfg.dt)
        runs.append(s)
    forced, plain, off = runs
    assert forced.updater.rebin_fallbacks > 0 and plain.updater.rebin_fallbacks == 0 and plain.updater.misfiled == 0
    # a failed plan widens the band of face-near particles that wait for the final fields (4x per failure)
    assert forced.ops.ctx.predict_eps() > plain.ops.ctx.predict_eps() == 2.0 ** -12
    for other in (plain, off):
        for pa, pb in zip(forced.patches, other.patches):
            for popa, popb in zip(pa.pops, pb.pops):
                assert forced.ops.count(popa.domain) == other.ops.count(popb.domain)
                a = forced.ops.get_particles(popa.domain)
                b = other.ops.get_particles(popb.domain)
                assert np.array_equal(np.sort(a[0], axis=0), np.sort(b[0], axis=0))  # the same cells are populated alike
                csa, csb = popa.cell_start.t.cpu().numpy(), popb.cell_start.t.cpu().numpy()
                assert np.array_equal(csa, csb)
            for name in ("B", "E", "Vi"):
                for c in range(3):
                    x = forced.ops.get_field(getattr(pa, name)[c])
                    y = other.ops.get_field(getattr(pb, name)[c])
                    ok = np.isfinite(y)
                    assert np.max(np.abs(x[ok] - y[ok])) <= 1e-11 * (np.max(np.abs(y[ok])) + 1e-300)


def test_small_patches_keep_the_two_pass_sweep():
    """below IonUpdater.predict_min_cells (a tile-kernel grid that would not fill the GPU) the all sweep stays two-pass"""
    from phare_b200 import configs
    from phare_b200.messenger import LocalComm
    from phare_b200.solver import GpuOps
    cfg = configs.get(3).with_cells((64, 32), (2, 2))
    cfg.pops = [dict(p, ppc=8) for p in cfg.pops]
    for min_cells, one_pass in ((128 * 2 * 148, False), (0, True)):
        s = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
        s.updater.predict_min_cells = min_cells
        s.ops.kernel_timing, s.ops.timed = True, {}
        s.advance_level(cfg.dt)
        assert ("move_all_rebin" in s.ops.timed) == one_pass
        assert ("deposit_scatter" in s.ops.timed) == (not one_pass)
