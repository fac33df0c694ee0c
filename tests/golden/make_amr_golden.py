"""Generates tests/golden/amr_operators.npz from the REFERENCE's own Python restatements of its level operators
(run in the dev container only, /root/reference must exist):
  tests/amr/data/field/refine/test_refine_field.py   default_refine, refine_magnetic, refine_electric
  tests/simulator/utilities/field_coarsening.py      coarsen
These are the functions the reference's test-suite checks its C++ refiners / coarseners against
(tests/simulator/test_advance.py, test_initialization.py).  Each case stores the coarse (or fine) input array with
ghosts, the AMR cell box it belongs to and the reference output; tests/test_amr_operators.py replays them through the
oracle (CPU) and through the C ABI (GPU)."""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
# plotting / HDF5 packages the reference's pharesee package imports at module level are absent here and unused
from unittest import mock  # noqa: E402
for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections", "matplotlib.colors",
          "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
    sys.modules.setdefault(m, mock.MagicMock())
sys.path[:0] = [REF, os.path.join(REF, "pyphare")]
from pyphare.core import box as boxm                                   # noqa: E402
from pyphare.core.box import Box                                       # noqa: E402
from pyphare.core.gridlayout import GridLayout                         # noqa: E402
from pyphare.pharesee.hierarchy.patchdata import FieldData             # noqa: E402
from tests.amr.data.field.refine import test_refine_field as R         # noqa: E402
from tests.simulator.utilities.field_coarsening import coarsen         # noqa: E402

out = {}
rng = np.random.default_rng(20260101)


cases = []
for ndim, names in ((1, ["Bx", "By", "Ex", "Ey", "rho"]), (2, ["Bx", "By", "Bz", "Ex", "Ey", "Ez", "rho"])):
    for interp in (1, 2):
        for name in names:
            lower = [6, 10][:ndim]
            ncells = [12, 8][:ndim]
            box = Box(lower, [l + n - 1 for l, n in zip(lower, ncells)])
            layout = GridLayout(box, np.zeros(ndim), np.full(ndim, 0.4), interp_order=interp)
            # test_refine_field.fine_layout_from reads layout.options.interp_order (an attribute of an older GridLayout)
            layout.options = types.SimpleNamespace(interp_order=interp)
            probe = FieldData(layout, name, data=np.zeros(1))
            g = int(probe.ghosts_nbr[0])
            prim = probe.primal_directions()
            shape = [ncells[d] + 2 * g + int(prim[d]) for d in range(ndim)]
            data = rng.standard_normal(shape)
            fd = FieldData(layout, name, data=data)
            key = f"{ndim}d_o{interp}_{name}"
            out[key + "_coarse"] = data
            out[key + "_box"] = np.array([lower, [l + n - 1 for l, n in zip(lower, ncells)]])
            algos = {"default": R.default_refine}
            if name[0] == "B":
                algos["magnetic"] = R.refine_magnetic
            # refine_electric is only usable in 1-D for Ex: its 1-D primal branch fails on a shape mismatch and its 2-D
            # branch tests the centering of Ey while documenting (and being applied to) Ex -> returns zeros
            if name == "Ex" and ndim == 1:
                algos["electric"] = R.refine_electric
            for an, fn in algos.items():
                try:
                    fine = fn(fd, data=data)
                except ValueError as e:  # e.g. refine_electric on a 1-D primal component (shape mismatch upstream)
                    print("skipped", key, an, "-", e)
                    continue
                out[f"{key}_{an}_fine"] = np.asarray(fine.dataset[:])
                out[f"{key}_{an}_finebox"] = np.array([fine.box.lower, fine.box.upper])
            # coarsening: a random FINE field over the refined box, coarsened onto the coarse box
            if name[0] in "Er":
                flayout = GridLayout(boxm.refine(box, 2), np.zeros(ndim), np.full(ndim, 0.2), interp_order=interp)
                fshape = [2 * ncells[d] + 2 * g + int(prim[d]) for d in range(ndim)]
                fdata = rng.standard_normal(fshape)
                ffd = FieldData(flayout, name, data=fdata)
                cdata = np.zeros(shape)
                coarsen(name, fd, ffd, box, fdata, cdata)
                out[key + "_cz_fine"] = fdata
                out[key + "_cz_coarse"] = cdata

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amr_operators.npz")
np.savez_compressed(dst, **out)
print(dst, len(out), "arrays")
