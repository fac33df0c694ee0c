"""Reader of the `.npz` diagnostics of phare_b200.simulator into pyphare's own PatchHierarchy, so that the reference's
post-processing (pyphare.pharesee: flat_finest_field, finest_field, hierarchy arithmetic, plots) runs on the output of
this implementation.  The reference reads HDF5 files (pyphare/pharesee/hierarchy/fromh5.py); h5py is not
available in this image, the file layout (`t<time>/pl<level>/p<patch>/<dataset>` + the writer's attributes) is the same.
pyphare must be importable (the reference's pure-Python package)."""
import glob
import os

import numpy as np


def hierarchy_from_npz(diag_dir, time=None):
    """PatchHierarchy holding every field quantity dumped in `diag_dir` at `time` (default: the latest dump), all levels,
    all patches of all ranks.  Keys follow pyphare: Bx..Bz, Ex..Ez, Vx..Vz, rho (charge density), <pop>_rho, <pop>_Fx.."""
    from pyphare.core.box import Box
    from pyphare.core.gridlayout import GridLayout
    from pyphare.pharesee.hierarchy.hierarchy import PatchHierarchy
    from pyphare.pharesee.hierarchy.hierarchy_utils import field_qties
    from pyphare.pharesee.hierarchy.patch import Patch
    from pyphare.pharesee.hierarchy.patchdata import FieldData
    from pyphare.pharesee.hierarchy.patchlevel import PatchLevel

    files = sorted(glob.glob(os.path.join(diag_dir, "*.npz")))
    if not files:
        raise FileNotFoundError(f"no .npz diagnostics in {diag_dir}")
    stamp = lambda f: float(os.path.basename(f).rsplit("_rank", 1)[0].rsplit("_", 1)[1])
    if time is None:
        time = max(stamp(f) for f in files)
    files = [f for f in files if abs(stamp(f) - time) < 5e-6 and not os.path.basename(f).startswith("particle_")]
    if not files:
        raise FileNotFoundError(f"no field diagnostics at time {time} in {diag_dir}")
    patches = {}  # (level, patch id) -> {"meta": ..., "data": {name: array}}
    interp = cells = dl = None
    for f in files:
        z = np.load(f)
        interp, cells, dl = int(z["_meta/interp_order"]), z["_meta/domain_cells"], z["_meta/cell_width"]
        quantity = str(z["_meta/quantity"]).strip("/")           # e.g. ions/pop/protons/flux
        prefix = quantity.split("/")[2] + "_" if quantity.startswith("ions/pop/") else ""
        for key in z.files:
            if key.startswith("_meta/"):
                continue
            _, pl, pp, name = key.split("/", 3)
            entry = patches.setdefault((int(pl[2:]), int(pp[1:])), {"meta": {}, "data": {}})
            if name.startswith("_"):
                entry["meta"][name[1:]] = z[key]
            elif name in field_qties:
                entry["data"][prefix + field_qties[name]] = (field_qties[name], z[key])
    levels = {}
    for (il, pid), entry in sorted(patches.items()):
        m = entry["meta"]
        layout = GridLayout(Box(m["lower"], m["upper"]), m["origin"], np.asarray(dl) / 2 ** il, interp_order=interp)
        pdatas = {name: FieldData(layout, qty, data) for name, (qty, data) in entry["data"].items()}
        levels.setdefault(il, []).append(Patch(pdatas, patch_id=f"p{il}#{pid}"))
    patch_levels = {il: PatchLevel(il, ps) for il, ps in levels.items()}
    domain = Box([0] * len(cells), [int(c) - 1 for c in cells])
    return PatchHierarchy(patch_levels, domain, refinement_ratio=2, times=[time])
