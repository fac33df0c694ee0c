"""Generates tests/golden/gridlayout_golden.npz from the REFERENCE's own GridLayout golden generators, run unmodified from
where they lie (dev container only; the GPU box reads the committed .npz):

  tests/core/data/gridlayout/allocSizes.py                    -> allocSizes_{1,2,3}d_O{1,2,3}.txt
  tests/core/data/gridlayout/gridIndexing.py                  -> gridIndexing_{1,2,3}d_O{1,2,3}.txt
  tests/core/data/gridlayout/test_deriv.py                    -> d{x,y,z}{By,Ez}_interpOrder_{1,2,3}_{1,2,3}d.txt
  tests/core/data/gridlayout/test_laplacian.py                -> lapJ{x,y,z}_interpOrder_{1,2,3}_{1,2,3}d.txt
  tests/core/data/gridlayout/test_linear_combinations_yee.py  -> linear_coefs_yee_*.txt

The deriv / laplacian generators keep their analytic input fields (By, Ez, Jx, Jy, Jz) in local variables: np.savetxt is
hooked and the caller's locals are copied at the moment each expected array is written, so the fixture holds the exact
inputs AND the expected outputs with no formula restated here.  3-D arrays are cropped to a window (inputs one node wider
than outputs: the stencils reach +-1) to keep the fixture small; 1-D and 2-D arrays are whole.

usage: python tests/golden/make_gridlayout_golden.py"""
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
WIN_LO, WIN_N = (9, 7, 11), (10, 8, 9)  # 3-D window: first array index and extent of the kept OUTPUT block


def main():
    gl_dir = os.path.join(REF, "tests/core/data/gridlayout")
    sys.path.insert(0, os.path.join(REF, "pyphare"))
    sys.path.insert(0, gl_dir)
    tmp = tempfile.mkdtemp()
    sys.argv = [sys.argv[0]]
    import allocSizes
    import gridIndexing
    import test_deriv
    import test_laplacian
    import test_linear_combinations_yee

    out = {}
    real_savetxt = np.savetxt

    def crop(a, grow):
        if a.ndim < 3:
            return np.array(a, copy=True)
        sl = tuple(slice(WIN_LO[d] - grow, WIN_LO[d] + WIN_N[d] + grow) for d in range(3))
        return np.array(a[sl], copy=True)

    def hooked(fname, arr, *a, **k):
        name = os.path.basename(str(fname)).replace(".txt", "")
        frame = sys._getframe(1)
        arr = np.asarray(arr)
        nd = int(name[-2])  # ..._1d / _2d / _3d: the generators flatten for savetxt, find the original array
        if arr.ndim != nd:
            for v in frame.f_locals.values():
                if isinstance(v, np.ndarray) and v.ndim == nd and v.size == arr.size and np.array_equal(v.ravel(), arr.ravel()):
                    arr = v
                    break
            assert arr.ndim == nd, name
        out[f"{name}/expected"] = crop(arr, 0)
        out[f"{name}/shape"] = np.array(arr.shape)
        for key in ("By", "Ez", "Jx", "Jy", "Jz"):
            v = frame.f_locals.get(key)
            if isinstance(v, np.ndarray) and name[-2:] == "%dd" % v.ndim:
                # generators of one (dim, interp order) write several files: the inputs are stored once per order
                order = name.split("interpOrder_")[1][0]
                out[f"input_{key}_O{order}_{v.ndim}d"] = crop(v, 1)
                out[f"input_{key}_O{order}_{v.ndim}d/shape"] = np.array(v.shape)
        return real_savetxt(fname, arr.ravel(), *a, **k)

    np.savetxt = hooked
    try:
        test_deriv.main(tmp + "/")
        test_laplacian.main(tmp + "/")
    finally:
        np.savetxt = real_savetxt
    allocSizes.main(tmp + "/")
    gridIndexing.main(tmp + "/")
    test_linear_combinations_yee.main(tmp + "/")
    for f in sorted(os.listdir(tmp)):
        if f.startswith(("allocSizes", "gridIndexing", "linear_coefs")):
            out["text/" + f] = np.array(open(os.path.join(tmp, f)).read())
    out["window"] = np.array([WIN_LO, WIN_N])
    # interp orders 2 and 3 share their ghost widths, hence shapes and (for these generators) every value: keep 1 and 2
    for k in list(out):
        if ("interpOrder_3" in k or "_O3_" in k) and "3d" in k:
            del out[k]
    path = os.path.join(HERE, "gridlayout_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB,", len(out), "entries")


if __name__ == "__main__":
    main()
