"""Split patterns of the reference's particle Splitter (src/amr/data/particles/refine/split_{1,2,3}d.hpp): the
flattened (delta, weight) table per (dim, interp_order, nbRefinedPart), generated from the reference headers by
tools/make_split_patterns.py and re-checked against them by tests/test_split.py."""
import json
import os

import numpy as np

_TABLE = None


def table():
    global _TABLE
    if _TABLE is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "split_patterns.json")) as f:
            _TABLE = json.load(f)
    return _TABLE


def pattern(dim, interp, nref):
    """(deltas float32 (nref, dim), weights float32 (nref,), maxCellDistanceFromSplit)"""
    key = f"{dim}_{interp}_{nref}"
    if key not in table():
        raise KeyError(f"no Splitter<{dim}, {interp}, {nref}> (valid: meta_utilities.hpp:73-88)")
    e = table()[key]
    return (np.asarray(e["deltas"], np.float32).reshape(nref, dim), np.asarray(e["weights"], np.float32),
            int(e["max_cell_distance"]))


def permutations():
    return sorted(tuple(int(x) for x in k.split("_")) for k in table())
