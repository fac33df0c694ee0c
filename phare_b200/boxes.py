"""Integer box algebra on the host (inclusive lower/upper corners), restating what the reference gets
from core::Box (src/core/utilities/box/box.hpp:28-202) and SAMRAI's BoxContainer: intersection (`*`),
grow, shift, and box removal (remove(), box.hpp:260-330) used by the field geometry
(src/amr/data/field/field_geometry.hpp:227-304)."""
import itertools

import numpy as np


class Box:
    __slots__ = ("lo", "hi")

    def __init__(self, lo, hi):
        self.lo = np.asarray(lo, dtype=np.int64).copy()
        self.hi = np.asarray(hi, dtype=np.int64).copy()

    @property
    def dim(self):
        return len(self.lo)

    def empty(self):
        return bool(np.any(self.hi < self.lo))

    def shape(self):
        return tuple(int(x) for x in (self.hi - self.lo + 1))

    def volume(self):
        return 0 if self.empty() else int(np.prod(self.hi - self.lo + 1))

    def shift(self, t):
        t = np.asarray(t, dtype=np.int64)
        return Box(self.lo + t, self.hi + t)

    def grow(self, w):
        w = np.asarray(w, dtype=np.int64)
        return Box(self.lo - w, self.hi + w)

    def __mul__(self, other):
        """intersection, None when empty (core::Box::operator*)"""
        b = Box(np.maximum(self.lo, other.lo), np.minimum(self.hi, other.hi))
        return None if b.empty() else b

    def contains(self, other):
        return bool(np.all(other.lo >= self.lo) and np.all(other.hi <= self.hi))

    def __eq__(self, other):
        return np.array_equal(self.lo, other.lo) and np.array_equal(self.hi, other.hi)

    def __repr__(self):
        return f"Box({self.lo.tolist()}, {self.hi.tolist()})"

    def minus(self, other):
        """self \\ other as a list of disjoint boxes"""
        inter = self * other
        if inter is None:
            return [Box(self.lo, self.hi)]
        out = []
        lo, hi = self.lo.copy(), self.hi.copy()
        for d in range(self.dim):
            if lo[d] < inter.lo[d]:
                h = hi.copy()
                h[d] = inter.lo[d] - 1
                out.append(Box(lo, h))
                lo = lo.copy()
                lo[d] = inter.lo[d]
            if hi[d] > inter.hi[d]:
                l = lo.copy()
                l[d] = inter.hi[d] + 1
                out.append(Box(l, hi))
                hi = hi.copy()
                hi[d] = inter.hi[d]
        return out


def periodic_shifts(domain_shape):
    """all shifts in {-1,0,1}^d x domain extent (the level is periodic in every direction,
    src/amr/wrappers/hierarchy.hpp:347-349)"""
    dim = len(domain_shape)
    return [np.array([s[d] * domain_shape[d] for d in range(dim)], dtype=np.int64)
            for s in itertools.product((-1, 0, 1), repeat=dim)]
