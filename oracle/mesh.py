"""TEST INFRASTRUCTURE — numpy restatement of the reference's spatial index.

Follows /root/reference/halotools/mock_observables/pair_counters/rectangular_mesh.py
(3-D: digitized_position :19-22, sample1_cell_size :25-55, sample2_cell_sizes :58-82,
RectangularMesh.__init__ :118-222, RectangularDoubleMesh :228-374) and
rectangular_mesh_2d.py (:15-19, :22-64, :69-145, :147-250).  The cell assignment
uses numpy's own float floor-division, argsort and searchsorted because those
ARE the reference semantics.  Only tests/, smoke() and
bench.py's CPU-baseline legs may import this module.
"""
from math import floor

import numpy as np

MAX_CELLS_PER_DIM = 50  # rectangular_mesh.py:15-16


def digitized_position(p, cell_size, num_divs):
    ip = np.floor(p // cell_size).astype(int)
    return np.where(ip >= num_divs, num_divs - 1, ip)


def sample1_cell_size(period, search_length, approx_cell_size, max_cells=MAX_CELLS_PER_DIM):
    period = float(period)
    if search_length > period / 3.0:
        raise ValueError("Input ``search_length`` cannot exceed period/3")
    ndivs = int(floor(period / float(approx_cell_size)))
    ndivs = max(ndivs, 1)
    ndivs = min(max_cells, ndivs)
    nsearch = int(floor(period / float(search_length)))
    nsearch = max(nsearch, 1)
    ndivs = min(ndivs, nsearch)
    ndivs = max(3, ndivs)
    return period / float(ndivs)


def sample2_cell_size(period, cell1_size, approx_cell_size, max_cells=MAX_CELLS_PER_DIM):
    period = float(period)
    n1 = int(np.round(period / cell1_size))
    per = int(np.round(cell1_size / float(approx_cell_size)))
    per = max(1, per)
    per = min(max_cells, per)
    n2 = n1 * per
    if n2 > max_cells:
        n2 = (max_cells // n1) * n1
    return period / float(n2)


class Mesh(object):
    """One sample binned into cells (any dimensionality 2 or 3), last dim fastest."""

    def __init__(self, coords, periods, approx_cell_sizes):
        self.ndim = len(coords)
        self.npts = coords[0].shape[0]
        self.period = [float(p) for p in periods]
        self.num_divs = [max(int(np.round(p / a)), 1) for p, a in zip(self.period, approx_cell_sizes)]
        self.cell_size = [p / float(n) for p, n in zip(self.period, self.num_divs)]
        self.ncells = int(np.prod(self.num_divs))
        idx = [digitized_position(c, cs, n) for c, cs, n in zip(coords, self.cell_size, self.num_divs)]
        cell_ids = idx[0]
        for d in range(1, self.ndim):
            cell_ids = cell_ids * self.num_divs[d] + idx[d]
        self.cell_ids = cell_ids
        self.idx_sorted = np.ascontiguousarray(np.argsort(cell_ids))
        cii = np.searchsorted(cell_ids, np.arange(self.ncells), sorter=self.idx_sorted)
        self.cell_id_indices = np.ascontiguousarray(np.append(cii, self.npts)).astype(np.int64)


class DoubleMesh(object):
    """RectangularDoubleMesh / RectangularDoubleMesh2D restated."""

    def __init__(self, coords1, coords2, approx_cell1, approx_cell2, search, period, PBCs=True):
        ndim = len(coords1)
        self.ndim = ndim
        self.period = [float(p) for p in period]
        self.search = [float(s) for s in search]
        self.PBCs = bool(PBCs)
        names = "xyz"
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
        a1 = [sample1_cell_size(self.period[d], self.search[d], approx_cell1[d]) for d in range(ndim)]
        self.mesh1 = Mesh(coords1, self.period, a1)
        a2 = [sample2_cell_size(self.period[d], self.mesh1.cell_size[d], approx_cell2[d]) for d in range(ndim)]
        self.mesh2 = Mesh(coords2, self.period, a2)
        self.per = [self.mesh2.num_divs[d] // self.mesh1.num_divs[d] for d in range(ndim)]
        # npairs_3d_engine.pyx:74-79
        self.cover = [int(np.ceil(self.search[d] / self.mesh2.cell_size[d])) for d in range(ndim)]

    def visited_pairs(self, first=0, last=None):
        """W_ref: number of (i, j) pairs the reference loop nest evaluates."""
        m1, m2 = self.mesh1, self.mesh2
        n1 = np.diff(m1.cell_id_indices).reshape(m1.num_divs)
        n2 = np.diff(m2.cell_id_indices).reshape(m2.num_divs).astype(np.float64)
        # sum of n2 over each cell1's window, separable box filter with wrap
        acc = n2
        for d in range(self.ndim):
            per, c, nd2, nd1 = self.per[d], self.cover[d], m2.num_divs[d], m1.num_divs[d]
            acc = np.moveaxis(acc, d, 0)
            out = np.zeros((nd1,) + acc.shape[1:], dtype=np.float64)
            for i1 in range(nd1):
                ids = np.arange(i1 * per - c, (i1 + 1) * per + c) % nd2
                out[i1] = acc[ids].sum(axis=0)
            acc = np.moveaxis(out, 0, d)
        w = (n1 * acc).ravel()
        if last is None:
            last = m1.ncells
        return float(w[first:last].sum())
