"""Host-side pieces of the spatial index that stay O(1) scalars.

The per-point work of the reference's RectangularMesh (digitize / argsort /
searchsorted) runs on the GPU (csrc/mesh.cu); what is left on the host is the cell
GEOMETRY - how many reference cells per dimension, their size, the search window -
which must be computed with exactly the reference's rules so that the GPU visits the
same pairs with the same periodic shifts:

  sample1_cell_size / sample2_cell_sizes   pair_counters/rectangular_mesh.py:25-82
  RectangularMesh num_divs / cell_size     pair_counters/rectangular_mesh.py:202-209
  RectangularDoubleMesh checks             pair_counters/rectangular_mesh.py:376-417
  _enclose_in_box / _enclose_in_square     pair_counters/mesh_helpers.py:17-110
  _set_approximate_cell_sizes / _2d_       pair_counters/mesh_helpers.py:113-180
  _cell1_parallelization_indices           pair_counters/mesh_helpers.py:183-221
  _enforce_maximum_search_length           pair_counters/mesh_helpers.py:224-255
"""
from copy import copy
from math import floor

import numpy as np

from .._lib import MeshGeom

__all__ = ("double_mesh_geometry", "_enclose_in_box", "_enclose_in_square", "_set_approximate_cell_sizes",
           "_set_approximate_2d_cell_sizes", "_cell1_parallelization_indices",
           "_enforce_maximum_search_length")

default_max_cells_per_dimension_cell1 = 50
default_max_cells_per_dimension_cell2 = 50


def _cell1_size(period, search_length, approx_cell_size, max_cells):
    period = float(period)
    if search_length > period / 3.0:
        raise ValueError("Input ``search_length`` cannot exceed period/3")
    ndivs = max(int(floor(period / float(approx_cell_size))), 1)
    ndivs = min(max_cells, ndivs)
    nsearch = max(int(floor(period / float(search_length))), 1)
    ndivs = max(3, min(ndivs, nsearch))
    return period / float(ndivs)


def _cell2_size(period, cell1_size, approx_cell_size, max_cells):
    period = float(period)
    n1 = int(np.round(period / cell1_size))
    per = int(np.round(cell1_size / float(approx_cell_size)))
    per = min(max_cells, max(1, per))
    n2 = n1 * per
    if n2 > max_cells:
        n2 = (max_cells // n1) * n1
    return period / float(n2)


def _mesh_divs(period, approx_cell_size):
    ndivs = max(int(np.round(period / approx_cell_size)), 1)
    return ndivs, period / float(ndivs)


class DoubleMeshGeometry(object):
    """The scalar attributes of RectangularDoubleMesh(2D) the engines read."""

    def __init__(self, ndim, approx_cell1, approx_cell2, search, period, PBCs,
                 max_cells1=default_max_cells_per_dimension_cell1,
                 max_cells2=default_max_cells_per_dimension_cell2):
        names = "xyz"
        self.ndim = ndim
        self.PBCs = bool(PBCs)
        self.period = [float(np.atleast_1d(period[d])[0]) if np.ndim(period[d]) else float(period[d])
                       for d in range(ndim)]
        self.search = [float(search[d]) for d in range(ndim)]
        for d in range(ndim):
            if not (self.search[d] <= self.period[d] / 3.0):
                raise ValueError(
                    "\n The maximum length over which you search for pairs of points \n"
                    "cannot be larger than Lbox/3 in any dimension. \n"
                    "You tried to search for pairs out to a length of search_%slength = %.2f,\n"
                    "but the size of your box in this dimension is %speriod = %.2f.\n"
                    "If you need to count pairs on these length scales, \n"
                    "you should use a larger simulation.\n"
                    % (names[d], self.search[d], names[d], self.period[d]))
        self.ndivs1, self.cell1_size, self.ndivs2, self.cell2_size, self.cover = [], [], [], [], []
        for d in range(ndim):
            a1 = _cell1_size(self.period[d], self.search[d], approx_cell1[d], max_cells1)
            n1, c1 = _mesh_divs(self.period[d], a1)
            a2 = _cell2_size(self.period[d], c1, approx_cell2[d], max_cells2)
            n2, c2 = _mesh_divs(self.period[d], a2)
            self.ndivs1.append(n1)
            self.cell1_size.append(c1)
            self.ndivs2.append(n2)
            self.cell2_size.append(c2)
            self.cover.append(int(np.ceil(self.search[d] / c2)))   # npairs_3d_engine.pyx:74-79
        self.ncells1 = int(np.prod(self.ndivs1))

    def as_struct(self):
        g = MeshGeom()
        g.ndim = self.ndim
        g.pbc = 1 if self.PBCs else 0
        for d in range(3):
            on = d < self.ndim
            g.ndivs1[d] = self.ndivs1[d] if on else 1
            g.ndivs2[d] = self.ndivs2[d] if on else 1
            g.cover[d] = self.cover[d] if on else 0
            g.period[d] = self.period[d] if on else 1.0
            g.cell1_size[d] = self.cell1_size[d] if on else 1.0
            g.cell2_size[d] = self.cell2_size[d] if on else 1.0
            g.search[d] = self.search[d] if on else 0.0
        return g


def double_mesh_geometry(ndim, approx_cell1, approx_cell2, search, period, PBCs):
    return DoubleMeshGeometry(ndim, approx_cell1, approx_cell2, search, period, PBCs)


def _enclose(cols1, cols2, min_size):
    lo = np.min([np.min(c) for c in cols1] + [np.min(c) for c in cols2])
    hi = np.max([np.max(c) for c in cols1] + [np.max(c) for c in cols2]) - lo
    cols1 = [c - lo for c in cols1]
    cols2 = [c - lo for c in cols2]
    Lbox = np.array([hi] * len(cols1))
    if min_size is not None:
        min_size = np.atleast_1d(min_size)
        if np.any(Lbox < min_size):
            Lbox[(Lbox < min_size)] = min_size[(Lbox < min_size)]
    return cols1, cols2, Lbox


def _enclose_in_box(x1, y1, z1, x2, y2, z2, min_size=None):
    """Shift both samples by the global minimum coordinate (ONE scalar for all dimensions) and
    return a cubic box spanning the largest extent, padded to ``min_size``."""
    (x1, y1, z1), (x2, y2, z2), Lbox = _enclose([x1, y1, z1], [x2, y2, z2], min_size)
    return x1, y1, z1, x2, y2, z2, Lbox


def _enclose_in_square(x1, y1, x2, y2, min_size=None):
    (x1, y1), (x2, y2), Lbox = _enclose([x1, y1], [x2, y2], min_size)
    return x1, y1, x2, y2, Lbox


def _approx_sizes(approx_cell1_size, approx_cell2_size, period, n):
    if approx_cell1_size is None:
        approx_cell1_size = period / 10.0
    else:
        approx_cell1_size = np.atleast_1d(approx_cell1_size)
        if not (len(approx_cell1_size) == n and type(approx_cell1_size) is np.ndarray
                and approx_cell1_size.ndim == 1):
            raise ValueError("Input ``approx_cell1_size`` must be a length-3 sequence")
    if approx_cell2_size is None:
        approx_cell2_size = copy(approx_cell1_size)
    else:
        approx_cell2_size = np.atleast_1d(approx_cell2_size)
        if not (len(approx_cell2_size) == n and type(approx_cell2_size) is np.ndarray
                and approx_cell2_size.ndim == 1):
            raise ValueError("Input ``approx_cell2_size`` must be a length-3 sequence")
    return approx_cell1_size, approx_cell2_size


def _set_approximate_cell_sizes(approx_cell1_size, approx_cell2_size, period):
    return _approx_sizes(approx_cell1_size, approx_cell2_size, period, 3)


def _set_approximate_2d_cell_sizes(approx_cell1_size, approx_cell2_size, period):
    return _approx_sizes(approx_cell1_size, approx_cell2_size, period, 2)


def _cell1_parallelization_indices(ncells, num_threads):
    """Contiguous (first, last) ranges of mesh1 cells: the reference's work split over processes,
    re-used here as the split over GPUs / ranks."""
    if num_threads == 1:
        return 1, [(0, ncells)]
    elif num_threads > ncells:
        return ncells, [(a, a + 1) for a in np.arange(ncells)]
    parts = [a for a in np.array_split(np.arange(ncells), num_threads) if len(a) > 0]
    return num_threads, [(x[0], x[0] + len(x)) for x in parts]


def _enforce_maximum_search_length(search_length, period=None):
    search_length = np.atleast_1d(search_length)
    if period is None:
        period = np.zeros_like(search_length) + np.inf
    period = np.atleast_1d(period)
    if not np.all(search_length < period / 3.0):
        max_search_fraction = np.max(period - search_length)
        msg = ("The search algorithm used by the function you called \n"
               "does not permit you to look for pairs separated by values \n"
               "exceeding a fraction of Lbox/3. in any dimension.\n"
               "Your function call would require searching for pairs separated by a distance of {0:.2f}*Lbox.\n"
               "Either decrease your search length or use a larger simulation.")
        raise ValueError(msg.format(max_search_fraction))
