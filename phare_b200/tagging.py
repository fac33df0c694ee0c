"""Tagging-driven refinement (`refinement="tagging"`): which cells of a level get refined, and the boxes of the next finer
level.

  default_tags   restates the reference's criterion exactly: DefaultTaggerStrategy<HybridModel>::tag
                 (src/amr/tagging/default_tagger_strategy.hpp:51-196), evaluated on the host from the patch's B arrays
                 (three small arrays per patch and regrid; the criterion is a few flops per cell).
  cluster        turns the tagged cells into boxes.  The reference hands the tags to SAMRAI's GriddingAlgorithm
                 (TileClustering / BergerRigoutsos, tag buffer, proper nesting, load balancer: src/amr/wrappers/
                 integrator.hpp:170-215), which cannot be built here and whose box choice is not part of PHARE: this is a
                 tile clustering of the same kind (fixed lattice of tiles, a tile holding a tag is refined, adjacent tiles
                 are coalesced, boxes chopped to the largest patch size) honouring tag_buffer, nesting_buffer and the patch
                 size limits.  The boxes are therefore NOT guaranteed identical to SAMRAI's; everything downstream of them
                 (phare_b200.amr) is the same as for user-given refinement boxes.
"""
import itertools

import numpy as np

from .boxes import Box


def default_tags(dim, interp, B, ncells, threshold=0.1):
    """B: the three host arrays of the patch (ghosts included); returns an int array of shape ncells (1 = refine)."""
    g = 2 if interp == 1 else 4
    Bx, By, Bz = B
    tags = np.zeros(tuple(int(n) for n in ncells), dtype=np.int32)
    if dim == 1:
        # :73-97: 5-point running means of By, Bz at ix and ix+1 (ix = g + iCell); the last cell is skipped when the
        # ghost layer is too thin for ix+3 (field_ghost_width <= 2)
        n = int(ncells[0]) if g > 2 else int(ncells[0]) - 1
        ix = g + np.arange(n)
        mean = lambda F, s: 0.2 * (F[ix - 2 + s] + F[ix - 1 + s] + F[ix + s] + F[ix + 1 + s] + F[ix + 2 + s])
        by0, bz0, by1, bz1 = mean(By, 0), mean(Bz, 0), mean(By, 1), mean(Bz, 1)
        cby = np.abs(by1 - by0) / (1 + np.abs(by0))
        cbz = np.abs(bz1 - bz0) / (1 + np.abs(bz0))
        tags[:n] = np.sqrt(cby * cby + cbz * cbz) > threshold
        return tags
    # :99-196: max over components and directions of |F(i+2) - F(i)| / (1 + |F(i+1) - F(i)|), every component indexed
    # from the dual physical start index (= g) whatever its own centering
    idx = np.meshgrid(*[g + np.arange(int(ncells[d])) for d in range(dim)], indexing="ij")
    crit = np.zeros(tags.shape)
    for F in (Bx, By, Bz):
        f0 = F[tuple(idx)]
        for d in range(dim):
            up1 = [i.copy() for i in idx]
            up2 = [i.copy() for i in idx]
            up1[d] = up1[d] + 1
            up2[d] = up2[d] + 2
            crit = np.maximum(crit, np.abs(F[tuple(up2)] - f0) / (1 + np.abs(F[tuple(up1)] - f0)))
    tags[...] = crit > threshold
    return tags


def level_tag_mask(level, ops, threshold, comm=None):
    """(mask, origin): boolean mask of the tagged cells of `level` over the bounding box of its patches; every rank
    tags the patches it owns, the masks are merged (element-wise max over the ranks)"""
    patches = level.solver.patches
    lo = np.min([p.box.lo for p in level.geom.patches], axis=0)
    hi = np.max([p.box.hi for p in level.geom.patches], axis=0)
    mask = np.zeros(tuple(int(x) for x in hi - lo + 1), dtype=bool)
    dim = len(lo)
    for p in patches:
        B = [ops.get_field(p.B[c]) for c in range(3)]
        t = default_tags(dim, p.layout.interp, B, [p.layout.ncells[d] for d in range(dim)], threshold)
        sl = tuple(slice(int(p.geom.box.lo[d] - lo[d]), int(p.geom.box.hi[d] - lo[d]) + 1) for d in range(dim))
        mask[sl] |= t.astype(bool)
    if comm is not None and comm.size > 1:
        mask = comm.allreduce_array_max(mask.astype(np.int32)).astype(bool)
    return mask, lo


def _dilate(mask, w):
    """box-shaped dilation by w cells (the tag buffer)"""
    out = mask.copy()
    for d in range(mask.ndim):
        acc = out.copy()
        for s in range(1, w + 1):
            a = [slice(None)] * mask.ndim
            b = [slice(None)] * mask.ndim
            a[d], b[d] = slice(s, None), slice(None, -s)
            acc[tuple(a)] |= out[tuple(b)]
            acc[tuple(b)] |= out[tuple(a)]
        out = acc
    return out


def cluster(mask, origin, allowed_boxes, tile, largest=None):
    """boxes (in the index space of the mask's level) covering the tagged cells: every tile of `tile` cells per
    direction (lattice anchored at index 0) that holds a tag AND lies entirely inside one of `allowed_boxes` is selected;
    selected tiles are coalesced into boxes (runs along the last direction first, then equal runs of adjacent rows), and
    boxes longer than `largest` cells are chopped at tile boundaries."""
    dim = mask.ndim
    tile = [int(t) for t in np.broadcast_to(tile, (dim,))]
    origin = np.asarray(origin, dtype=np.int64)
    upper = origin + np.asarray(mask.shape) - 1
    t_lo = [int(np.floor(origin[d] / tile[d])) for d in range(dim)]
    t_hi = [int(np.floor(upper[d] / tile[d])) for d in range(dim)]
    shape = [t_hi[d] - t_lo[d] + 1 for d in range(dim)]
    sel = np.zeros(shape, dtype=bool)
    for t in itertools.product(*[range(s) for s in shape]):
        lo = np.array([(t_lo[d] + t[d]) * tile[d] for d in range(dim)], dtype=np.int64)
        box = Box(lo, lo + np.asarray(tile) - 1)
        if not any(a.contains(box) for a in allowed_boxes):
            continue
        clip = Box(np.maximum(box.lo, origin), np.minimum(box.hi, upper))
        sl = tuple(slice(int(clip.lo[d] - origin[d]), int(clip.hi[d] - origin[d]) + 1) for d in range(dim))
        sel[t] = bool(mask[sl].any())
    # coalesce: runs along the last direction, then merge runs with the same extent in adjacent rows, direction by
    # direction towards the first one
    boxes = []
    for t in itertools.product(*[range(s) for s in shape[:-1]]):
        row = sel[t]
        k = 0
        while k < len(row):
            if not row[k]:
                k += 1
                continue
            e = k
            while e + 1 < len(row) and row[e + 1]:
                e += 1
            boxes.append(Box(list(t) + [k], list(t) + [e]))
            k = e + 1
    for d in range(dim - 2, -1, -1):
        merged = True
        while merged:
            merged = False
            for a, b in itertools.combinations(boxes, 2):
                other = [k for k in range(dim) if k != d]
                if all(a.lo[k] == b.lo[k] and a.hi[k] == b.hi[k] for k in other) and (a.hi[d] + 1 == b.lo[d] or b.hi[d] + 1 == a.lo[d]):
                    boxes.remove(a)
                    boxes.remove(b)
                    boxes.append(Box(np.minimum(a.lo, b.lo), np.maximum(a.hi, b.hi)))
                    merged = True
                    break
    # tiles -> cells, then chop
    out = []
    for b in boxes:
        lo = np.array([(t_lo[d] + int(b.lo[d])) * tile[d] for d in range(dim)], dtype=np.int64)
        hi = np.array([(t_lo[d] + int(b.hi[d]) + 1) * tile[d] - 1 for d in range(dim)], dtype=np.int64)
        pieces = [Box(lo, hi)]
        if largest is not None:
            big = [max(tile[d], (int(np.broadcast_to(largest, (dim,))[d]) // tile[d]) * tile[d]) for d in range(dim)]
            for d in range(dim):
                nxt = []
                for p in pieces:
                    start = int(p.lo[d])
                    while start <= p.hi[d]:
                        end = min(start + big[d] - 1, int(p.hi[d]))
                        l, h = p.lo.copy(), p.hi.copy()
                        l[d], h[d] = start, end
                        nxt.append(Box(l, h))
                        start = end + 1
                pieces = nxt
        out += pieces
    return sorted(out, key=lambda b: tuple(int(x) for x in b.lo))


class Tagger:
    """what the simulator needs of ConcreteTagger + GriddingAlgorithm: the boxes of level i+1 from the state of level i"""

    def __init__(self, threshold=0.1, tag_buffer=1, nesting_buffer=1, smallest_patch_size=None, largest_patch_size=None,
                 max_nbr_levels=2):
        self.threshold, self.tag_buffer, self.nesting_buffer = threshold, int(tag_buffer), nesting_buffer
        self.smallest, self.largest, self.max_nbr_levels = smallest_patch_size, largest_patch_size, int(max_nbr_levels)

    def boxes(self, hierarchy, il):
        """cell boxes of level il (its own index space) to be covered by level il + 1"""
        level = hierarchy.levels[il]
        dim = level.geom.dim
        mask, origin = level_tag_mask(level, hierarchy.ops, self.threshold, hierarchy.comm)
        if self.tag_buffer:
            mask = _dilate(mask, self.tag_buffer)
        # where level il+1 may live: anywhere on the root level; on a refined level inside its patches shrunk by the
        # nesting buffer (proper nesting)
        nest = np.broadcast_to(self.nesting_buffer, (dim,)).astype(np.int64)
        domain = Box([0] * dim, [int(s) - 1 for s in level.geom.domain_shape])
        if il == 0:
            allowed = [domain]
        else:
            allowed = [b for b in (Box(p.box.lo + nest, p.box.hi - nest) * domain for p in level.geom.patches) if b is not None]
            # adjacent patches: the seam between them is interior to the level, not a boundary to nest away from
            allowed += [b for b in (Box(np.minimum(p.box.lo, q.box.lo) + nest, np.maximum(p.box.hi, q.box.hi) - nest) * domain
                                    for p, q in itertools.combinations(level.geom.patches, 2)
                                    if Box(np.minimum(p.box.lo, q.box.lo), np.maximum(p.box.hi, q.box.hi)).volume()
                                    == p.box.volume() + q.box.volume()) if b is not None]
        smallest = np.broadcast_to(self.smallest if self.smallest is not None else 8, (dim,))
        tile = [max(2, int(s) // 2) for s in smallest]  # a tile refines into one smallest patch of level il + 1
        largest = None if self.largest is None else [max(int(x) // 2, t) for x, t in zip(np.broadcast_to(self.largest, (dim,)), tile)]
        return cluster(mask, origin, allowed, tile, largest)
