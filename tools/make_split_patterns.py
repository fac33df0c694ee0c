"""Generates phare_b200/split_patterns.json — the (delta, weight) table of every particle Splitter permutation
(src/amr/data/particles/refine/split_{1,2,3}d.hpp; published at github.com/PHAREHUB/PHARE/wiki/SplitPattern) — by
asking the reference itself (oracle/_ref/libphare_ref.so::phr_split_pattern, built in place from the reference
headers).  Values are IEEE float32 like the reference's, stored as the exact double they convert to.
Run in the container that has /root/reference; tests/test_split.py re-checks the committed table against the
reference whenever the reference build is available."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PERMUTATIONS = {(1, 1): (2, 3), (1, 2): (2, 3, 4), (1, 3): (2, 3, 4, 5),
                (2, 1): (4, 5, 8, 9), (2, 2): (4, 5, 8, 9, 16), (2, 3): (4, 5, 8, 9, 25),
                (3, 1): (6, 12, 27), (3, 2): (6, 12, 27), (3, 3): (6, 12, 27)}


def reference_pattern(lib, dim, interp, nref):
    d = np.zeros((nref, dim), np.float32)
    w = np.zeros(nref, np.float32)
    m = C.c_int()
    rc = lib.phr_split_pattern(dim, interp, nref, d.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.byref(m))
    assert rc == 0, (dim, interp, nref)
    return d, w, m.value


def main():
    import oracle
    oracle.build(ref=True)
    lib = oracle.Cpu("ref").lib
    table = {}
    for (dim, interp), nrefs in PERMUTATIONS.items():
        for nref in nrefs:
            d, w, m = reference_pattern(lib, dim, interp, nref)
            table[f"{dim}_{interp}_{nref}"] = dict(deltas=[[float(x) for x in row] for row in d],
                                                   weights=[float(x) for x in w], max_cell_distance=m)
    with open(os.path.join(ROOT, "phare_b200", "split_patterns.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
    print(f"{len(table)} splitters written")


if __name__ == "__main__":
    main()
