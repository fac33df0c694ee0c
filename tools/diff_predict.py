"""the step with and without the predicted re-binning on the same device-loaded config: largest field / moment difference"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phare_b200 import abi, configs
from phare_b200.messenger import LocalComm
from phare_b200.solver import GpuOps

CASES = {1: ((1024,), (4,), 4), 3: ((64, 32), (2, 2), 3), 5: ((16, 16, 16), (2, 2, 1), 3)}
for k in [int(a) for a in sys.argv[1:]] or [1, 3, 5]:
    cells, grid, steps = CASES[k]
    cfg = configs.get(k).with_cells(cells, grid)
    cfg.pops = [dict(p, ppc=min(p["ppc"], 24)) for p in cfg.pops]
    runs = []
    for predict in (True, False):
        s = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
        s.updater.predict = predict
        for _ in range(steps):
            s.advance_level(cfg.dt)
        runs.append(s)
    a, b = runs
    print("config", k, "misfiled", a.updater.misfiled, "fallbacks", a.updater.rebin_fallbacks)
    for pa, pb in zip(a.patches, b.patches):
        ca = [a.ops.count(p.domain) for p in pa.pops]
        cb = [b.ops.count(p.domain) for p in pb.pops]
        out = [f"counts {ca} {cb}"]
        for name in ("B", "E", "Vi"):
            for c in range(3):
                x, y = a.ops.get_field(getattr(pa, name)[c]), b.ops.get_field(getattr(pb, name)[c])
                ok = np.isfinite(y)
                out.append(f"{name}{c} {np.max(np.abs(x[ok] - y[ok])):.2e}/{np.max(np.abs(y[ok])):.2e}")
        x, y = a.ops.get_field(pa.Ne), b.ops.get_field(pb.Ne)
        out.append(f"Ne {np.max(np.abs(x - y)):.2e}/{np.max(np.abs(y)):.2e}")
        print("  patch", pa.geom.id, " ".join(out))
