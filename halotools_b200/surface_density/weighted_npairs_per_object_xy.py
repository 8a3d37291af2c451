"""Drop-ins for ``halotools.mock_observables.surface_density.weighted_npairs_per_object_xy`` and
``total_mass_enclosed_per_cylinder``
(/root/reference/halotools/mock_observables/surface_density/weighted_npairs_per_object_xy.py:21-220,
mass_in_cylinders.py:126-262)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import (enforce_sample_has_correct_shape, enforce_sample_respects_pbcs, get_num_threads,
                       get_period, get_separation_bins_array)
from ..pair_counters.mesh_helpers import _set_approximate_2d_cell_sizes, double_mesh_geometry
from ..pair_counters._args import sample_columns
from .weighted_npairs_xy import _weighted_npairs_xy_process_args, weighted_npairs_xy

__all__ = ("weighted_npairs_per_object_xy", "total_mass_enclosed_per_cylinder",
           "total_mass_enclosed_in_stack_of_cylinders", "surface_density_in_annulus", "surface_density_in_cylinder")


def weighted_npairs_per_object_xy(sample1, sample2, sample2_mass, rp_bins,
                                  period=None, num_threads=1,
                                  approx_cell1_size=None, approx_cell2_size=None):
    """Mass of ``sample2`` inside the z-aligned cylinder of every radius in ``rp_bins`` around every point of
    ``sample1``: float64 (Npts1, len(rp_bins)), cumulative in rp, rows in input order
    (surface_density/engines/weighted_npairs_per_object_xy_engine.pyx:150-190)."""
    # the reference validates with the same helper as weighted_npairs_xy (weighted_npairs_per_object_xy.py:155-220)
    result = _weighted_npairs_xy_process_args(sample1, sample2, sample2_mass,
                                              rp_bins, period, num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, x2in, y2in, w2in = result[0:5]
    rp_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[5:]

    rp_max = np.max(rp_bins)
    search = [rp_max, rp_max]
    approx_cell1_size, approx_cell2_size = _set_approximate_2d_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(2, approx_cell1_size, approx_cell2_size, search, period[:2], PBCs)

    c1, c2 = sample_columns([x1in, y1in], [x2in, y2in], host_only="weighted_npairs_per_object_xy")
    counts = np.zeros((c1.n, len(rp_bins)), dtype=np.float64)
    first, last = _dist.cell1_range(geom.ncells1)
    w2 = np.ascontiguousarray(w2in, dtype=np.float64)
    g = geom.as_struct()
    rb = np.ascontiguousarray(rp_bins, dtype=np.float64)
    _lib.run_engine(
        "htb_weighted_npairs_per_object_xy_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(w2), _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        _lib._dp(counts))
    return _dist.allreduce_sum(counts)           # a fresh array already (no second copy of a large table)


def total_mass_enclosed_per_cylinder(centers, particles,
                                     particle_masses, downsampling_factor, rp_bins, period,
                                     num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """Total mass enclosed in infinitely long z-aligned cylinders of radii ``rp_bins`` around every centre:
    float64 (num_cyl, len(rp_bins)) (mass_in_cylinders.py:126-231)."""
    (centers, particles, particle_masses, downsampling_factor,
     rp_bins, period, num_threads, PBCs) = _enclosed_mass_process_args(
        centers, particles, particle_masses, downsampling_factor, rp_bins, period, num_threads)

    mean_particle_mass = np.mean(particle_masses)
    normalized_particle_masses = particle_masses/mean_particle_mass

    total_mass_per_cylinder = weighted_npairs_per_object_xy(
        centers, particles, normalized_particle_masses, rp_bins,
        period=period[:2], num_threads=num_threads,
        approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size)

    total_mass_per_cylinder *= downsampling_factor*mean_particle_mass
    return total_mass_per_cylinder


def total_mass_enclosed_in_stack_of_cylinders(centers, particles,
                                              particle_masses, downsampling_factor, rp_bins, period,
                                              num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """Total mass enclosed by the STACK of cylinders around all centres: float64 (len(rp_bins),)
    (mass_in_cylinders.py:21-123)."""
    (centers, particles, particle_masses, downsampling_factor,
     rp_bins, period, num_threads, PBCs) = _enclosed_mass_process_args(
        centers, particles, particle_masses, downsampling_factor, rp_bins, period, num_threads)

    mean_particle_mass = np.mean(particle_masses)
    normalized_particle_masses = particle_masses/mean_particle_mass

    total_mass_in_stack_of_cylinders = weighted_npairs_xy(
        centers, particles, normalized_particle_masses, rp_bins,
        period=period[:2], num_threads=num_threads,
        approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size)

    total_mass_in_stack_of_cylinders *= downsampling_factor*mean_particle_mass
    return total_mass_in_stack_of_cylinders


def surface_density_in_annulus(centers, particles, particle_masses,
                               downsampling_factor, rp_bins, period,
                               num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """Average surface mass density in a stack of annuli (surface_density.py:17-35)."""
    total_mass_in_stack_of_cylinders = total_mass_enclosed_in_stack_of_cylinders(
        centers, particles, particle_masses, downsampling_factor, rp_bins, period,
        num_threads=num_threads, approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size)
    total_mass_in_stack_of_annuli = np.diff(total_mass_in_stack_of_cylinders)
    rp_sq = rp_bins * rp_bins
    area_annuli = np.pi * np.diff(rp_sq)
    num_annuli = float(centers.shape[0])
    return total_mass_in_stack_of_annuli / (area_annuli * num_annuli)


def surface_density_in_cylinder(centers, particles, particle_masses,
                                downsampling_factor, rp_bins, period,
                                num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """Average surface mass density in a stack of cylinders (surface_density.py:38-53)."""
    total_mass_in_stack_of_cylinders = total_mass_enclosed_in_stack_of_cylinders(
        centers, particles, particle_masses, downsampling_factor, rp_bins, period,
        num_threads=num_threads, approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size)
    area_cylinders = np.pi * rp_bins * rp_bins
    num_cylinders = float(centers.shape[0])
    return total_mass_in_stack_of_cylinders / (area_cylinders * num_cylinders)


def _enclosed_mass_process_args(centers, particles, masses, downsampling_factor, rp_bins, period, num_threads):
    """mass_in_cylinders.py:234-262."""
    period, PBCs = get_period(period)

    centers = enforce_sample_has_correct_shape(centers)
    particles = enforce_sample_has_correct_shape(particles)

    masses = np.atleast_1d(masses)
    if len(masses) == 1:
        masses = np.zeros(particles.shape[0]) + masses[0]
    else:
        msg = "Must have same number of ``particle_masses`` as particles"
        assert masses.shape[0] == particles.shape[0], msg

    msg = "downsampling_factor = {0} < 1, which is impossible".format(downsampling_factor)
    assert downsampling_factor >= 1, msg

    enforce_sample_respects_pbcs(centers[:, 0], centers[:, 1], centers[:, 2], period)
    enforce_sample_respects_pbcs(particles[:, 0], particles[:, 1], particles[:, 2], period)

    rp_bins = get_separation_bins_array(rp_bins)
    num_threads = get_num_threads(num_threads, enforce_max_cores=False)
    return centers, particles, masses, downsampling_factor, rp_bins, period, num_threads, PBCs
