"""Generates tests/golden/*.npz from the REFERENCE's own golden generators (run in the dev container,
where /root/reference exists; the GPU box only reads the committed .npz files).

Generators executed, unmodified, from where they lie:
  tests/core/numerics/faraday/test_faraday.py       -> dbxdt/dbydt/dbzdt_yee_{1,2,3}D_order1
  tests/core/numerics/ampere/test_ampere.py         -> jx/jy/jz_yee_{1,2,3}D_order1
  tests/core/numerics/ohm/test_ohm.py               -> ohmx/ohmy/ohmz_yee_{1,2,3}D_order1   (eta = 1, nu = 0.01)
  tests/core/numerics/interpolator/interpolator_test.py -> bsplines_{1,2,3}_{primal,dual}.dat
  tests/core/numerics/pusher/test_pusher.py         -> pusher_test_in.txt (scipy odeint trajectory)
The generators keep their analytic input fields in local variables; a profile hook copies every numpy local
at function return, so the fixtures hold the exact inputs AND the expected outputs, with no formula restated
here.  3-D arrays are kept on a stride-2 sub-lattice to keep the fixtures small.

usage: python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
CROP_OFF, CROP_N = (7, 5, 9), (12, 10, 8)  # first cell (array index) and cells of the 3-D window
CENTERING = {"Bx": "pdd", "By": "dpd", "Bz": "ddp", "Ex": "dpp", "Ey": "pdp", "Ez": "ppd", "Jx": "dpp", "Jy": "pdp",
             "Jz": "ppd", "n": "ppp", "Vx": "ppp", "Vy": "ppp", "Vz": "ppp", "P": "ppp"}


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_captured(mod, tmp):
    captured = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name.startswith("test_"):
            captured[frame.f_code.co_name] = {k: np.array(v, copy=True) for k, v in frame.f_locals.items()
                                              if isinstance(v, np.ndarray)}
            captured[frame.f_code.co_name].update({k: np.float64(v) for k, v in frame.f_locals.items()
                                                   if isinstance(v, float)})
    sys.setprofile(prof)
    try:
        mod.main(tmp)
    finally:
        sys.setprofile(None)
    return captured


def main():
    sys.path.insert(0, os.path.join(REF, "pyphare"))
    tmp = tempfile.mkdtemp()
    out = {}
    keep = {
        "faraday": ["Ex", "Ey", "Ez", "Bx", "By", "Bz", "BxNew", "ByNew", "BzNew"],
        "ampere": ["Bx", "By", "Bz", "Jx", "Jy", "Jz"],
        "ohm": ["n", "Vx", "Vy", "Vz", "P", "Bx", "By", "Bz", "Jx", "Jy", "Jz", "ExNew", "EyNew", "EzNew", "eta", "nu"],
    }
    for op in ("faraday", "ampere", "ohm"):
        mod = load(os.path.join(REF, "tests/core/numerics", op, f"test_{op}.py"), f"ref_{op}")
        cap = run_captured(mod, tmp)
        tv = mod.TestVariables()
        for fn, loc in cap.items():
            dim = int(fn[-2])
            for k in keep[op]:
                if k in loc:
                    a = loc[k]
                    if dim == 3 and a.ndim == 3:
                        # the stencils are local: keep a CROP_N-cell window (plus its ghost margin) of the
                        # 50x30x40 arrays; the window is a valid smaller layout whose physical nodes map onto
                        # physical nodes of the full one
                        sl = []
                        for d in range(3):
                            primal = CENTERING[k.replace("New", "")][d] == "p"
                            sl.append(slice(CROP_OFF[d], CROP_OFF[d] + CROP_N[d] + 2 * 2 + (1 if primal else 0)))
                        a = a[tuple(sl)].copy()
                    out[f"{op}_{dim}d_{k}"] = a
        out[f"{op}_crop3d_ncells"] = np.array(CROP_N)
        out[f"{op}_ncells"] = np.array(tv.nbrCells)
        out[f"{op}_dx"] = np.array(tv.meshSize)
        if hasattr(tv, "dt"):
            out[f"{op}_dt"] = np.float64(tv.dt)
    np.savez_compressed(os.path.join(HERE, "fields_golden.npz"), **out)
    print("fields_golden.npz:", {k: getattr(v, "shape", ()) for k, v in list(out.items())[:8]}, "...")

    # B-spline weights
    itp = load(os.path.join(REF, "tests/core/numerics/interpolator/interpolator_test.py"), "ref_itp")
    argv = sys.argv
    sys.argv = ["interpolator_test.py", tmp]
    itp.main()
    sys.argv = argv
    bs = {}
    for centering in ("primal", "dual"):
        for order in (1, 2, 3):
            raw = open(os.path.join(tmp, f"bsplines_{order}_{centering}.dat"), "rb").read()
            rec = np.dtype([("nodes", np.int32, order + 1), ("w", np.float64, order + 1)])
            a = np.frombuffer(raw, dtype=rec)
            bs[f"nodes_{order}_{centering}"] = a["nodes"].copy()
            bs[f"weights_{order}_{centering}"] = a["w"].copy()
    np.savez_compressed(os.path.join(HERE, "bsplines_golden.npz"), **bs)
    print("bsplines_golden.npz:", {k: v.shape for k, v in bs.items()})

    # Boris trajectory (odeint), every 1000th of the 100001 rows
    psh = load(os.path.join(REF, "tests/core/numerics/pusher/test_pusher.py"), "ref_pusher")
    sys.argv = ["test_pusher.py", tmp]
    psh.main()
    sys.argv = argv
    sol = np.loadtxt(os.path.join(tmp, "pusher_test_in.txt"))
    np.savez_compressed(os.path.join(HERE, "pusher_golden.npz"), trajectory=sol[::1000], stride=np.int64(1000),
                        dt=np.float64(1e-4), nrows=np.int64(len(sol)))
    print("pusher_golden.npz:", sol.shape, "->", sol[::1000].shape)


if __name__ == "__main__":
    main()
