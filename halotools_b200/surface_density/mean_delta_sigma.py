"""Drop-in for ``halotools.mock_observables.mean_delta_sigma``
(/root/reference/halotools/mock_observables/surface_density/mean_delta_sigma.py:26-330)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import (custom_len, enforce_sample_has_correct_shape, enforce_sample_respects_pbcs,
                       get_num_threads, get_period, get_separation_bins_array)
from ..pair_counters._args import sample_columns
from ..pair_counters.mesh_helpers import (_enclose_in_box, _enclose_in_square,
                                          _set_approximate_2d_cell_sizes, double_mesh_geometry)

__all__ = ("mean_delta_sigma",)


def mean_delta_sigma(galaxies, particles, effective_particle_masses,
                     rp_bins, period=None, verbose=False, num_threads=1,
                     approx_cell1_size=None, approx_cell2_size=None, per_object=False):
    """Excess surface density Delta Sigma(rp) of ``particles`` around ``galaxies`` (projection along
    z), in the bins defined by ``rp_bins``: float64 (len(rp_bins)-1,), or (Ngal, len(rp_bins)-1)
    rows in input order when ``per_object`` is True.  Semantics follow the reference engine
    (surface_density/engines/mean_delta_sigma_engine.pyx:162-185)."""
    # one mass for every particle (the reference broadcasts the scalar, mean_delta_sigma.py:279-281):
    # the engine then needs no per-particle mass array and no per-pair logarithm
    uniform_mass = len(np.atleast_1d(effective_particle_masses)) == 1
    use_scalar = uniform_mass and not _lib.default_flags & _lib.FLAG_GENERIC
    keep = None
    if (period is not None and use_scalar and _lib.uploadable(galaxies) and _lib.uploadable(particles)
            and (_dist.device_collective() or (len(particles) >= 1000000 and _lib.library_present()
                                               and _lib.load().htb_device_count() >= 1))):
        # multi-GPU: every rank sends 1/world of each sample across PCIe, an all-gather over NVLink completes the copies
        # (instead of every rank uploading all 1e8 particles: 8 x 1.6 GB through the host's PCIe root).
        # One GPU, large samples: start the copies NOW, asynchronously, and let the bounds checks of the argument
        # processing run on the device copy (one HBM-bound pass) instead of scanning 2.4 GB on the host before the first
        # byte moves (config 5 end to end: 359 -> 340 ms)
        import torch
        keep = (galaxies, particles)
        with torch.cuda.stream(_lib.engine_stream()):
            galaxies, particles = _dist.to_device(galaxies), _dist.to_device(particles)
    result = _mean_delta_sigma_process_args(
        galaxies, particles, effective_particle_masses, rp_bins,
        period, num_threads, approx_cell1_size, approx_cell2_size, _broadcast_scalar_mass=not use_scalar)
    x1in, y1in, x2in, y2in, w2in = result[0:5]
    rp_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[5:]
    rp_max = np.max(rp_bins)
    search = [rp_max, rp_max]

    approx_cell1_size, approx_cell2_size = _set_approximate_2d_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(2, approx_cell1_size, approx_cell2_size, search, period[:2], PBCs)

    n1 = len(x1in)
    nbin = len(rp_bins) - 1
    # per_object=False only needs the column sums: they are formed on the device (HTB_FLAG_COLUMN_SUM) and the
    # (Ngal, nbin) rows never cross PCIe
    delta_sigma = np.zeros((n1, nbin) if per_object else (nbin,), dtype=np.float64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in], [x2in, y2in])
    extra = (_lib.FLAG_UNIFORM_MASS if use_scalar else 0) | (0 if per_object else _lib.FLAG_COLUMN_SUM)
    if c1.device and not use_scalar:
        raise TypeError("device-resident samples need a scalar ``effective_particle_masses``")
    m2 = np.ascontiguousarray(w2in[:1] if use_scalar else w2in, dtype=np.float64)
    if c1.device:
        # HTB_FLAG_DEVICE_INPUT covers every sample pointer, the mass included
        import torch
        m2_dev = torch.from_numpy(m2).to(x2in.device)
        m2_ptr = ctypes.cast(ctypes.c_void_p(int(m2_dev.data_ptr())), ctypes.POINTER(ctypes.c_double))
    else:
        m2_ptr = _lib._dp(m2)
    g = geom.as_struct()
    rb = np.ascontiguousarray(rp_bins, dtype=np.float64)
    if not per_object and _dist.device_collective():
        # multi-GPU: the column sums stay on the device, are all-reduced there on the engine's stream, and cross PCIe once
        import torch
        with torch.cuda.stream(_lib.engine_stream()):
            sums = torch.zeros(nbin, dtype=torch.float64, device="cuda")
            _lib.run_engine(
                "htb_mean_delta_sigma_engine", ctypes.byref(g),
                c1.ptrs[0], c1.ptrs[1], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
                c2.ptrs[0], c2.ptrs[1], ctypes.c_int64(c2.stride), m2_ptr, ctypes.c_int64(c2.n),
                _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
                _lib.out_pointer(sums, delta_sigma, ctypes.c_double), extra_flags=extra, device=c1.device, out_device=True)
            colsum = _dist.allreduce_device(sums).cpu().numpy()
        return colsum / float(n1) if n1 > 0 else colsum * np.nan
    _lib.run_engine(
        "htb_mean_delta_sigma_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], ctypes.c_int64(c2.stride), m2_ptr, ctypes.c_int64(c2.n),
        _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        _lib._dp(delta_sigma), extra_flags=extra, device=c1.device)
    if per_object:
        return _dist.allreduce_sum(delta_sigma)
    # rows outside this rank's mesh1 cells are zero, so the mean is the all-reduced column sum / N
    # (np.mean(axis=0) of the reference, mean_delta_sigma.py:252-255)
    colsum = _dist.allreduce_sum(delta_sigma)
    return colsum / float(n1) if n1 > 0 else colsum * np.nan


def _mean_delta_sigma_process_args(
        galaxies, particles, effective_particle_masses, rp_bins,
        period, num_threads, approx_cell1_size, approx_cell2_size, _broadcast_scalar_mass=True):
    """Same processing as mean_delta_sigma.py:258-330, including the reference's habit of writing
    the shifted coordinates back into the caller's arrays when ``period`` is None (:263-273)."""
    period, PBCs = get_period(period)
    if PBCs is False and (getattr(galaxies, "is_cuda", False) or getattr(particles, "is_cuda", False)):
        raise ValueError("device-resident samples need an explicit ``period``")

    if PBCs is False:
        _x1, _y1, _z1, _x2, _y2, _z2, period = _enclose_in_box(
            galaxies[:, 0], galaxies[:, 1], galaxies[:, 2],
            particles[:, 0], particles[:, 1], particles[:, 2])
        galaxies[:, 0] = _x1
        galaxies[:, 1] = _y1
        galaxies[:, 2] = _z1
        particles[:, 0] = _x2
        particles[:, 1] = _y2
        particles[:, 2] = _z2

    galaxies = enforce_sample_has_correct_shape(galaxies)
    particles = enforce_sample_has_correct_shape(particles)

    effective_particle_masses = np.atleast_1d(effective_particle_masses)
    if len(effective_particle_masses) == 1:
        # the reference broadcasts the scalar to one mass per particle (mean_delta_sigma.py:279-281); the engine
        # takes the scalar itself (HTB_FLAG_UNIFORM_MASS), so the Npart-long array is only built on request
        if _broadcast_scalar_mass:
            effective_particle_masses = np.zeros(particles.shape[0]) + effective_particle_masses[0]
        else:
            effective_particle_masses = np.asarray(effective_particle_masses, dtype=np.float64)
    else:
        msg = "Must have same number of ``particle_masses`` as particles"
        assert effective_particle_masses.shape[0] == particles.shape[0], msg

    enforce_sample_respects_pbcs(galaxies[:, 0], galaxies[:, 1], galaxies[:, 2], period)
    enforce_sample_respects_pbcs(particles[:, 0], particles[:, 1], particles[:, 2], period)

    x1 = galaxies[:, 0]
    y1 = galaxies[:, 1]
    x2 = particles[:, 0]
    y2 = particles[:, 1]

    rp_bins = get_separation_bins_array(rp_bins)
    rp_max = np.max(rp_bins)

    if period is None:
        PBCs = False
        x1, y1, x2, y2, period = (
            _enclose_in_square(x1, y1, x2, y2, min_size=[rp_max*3.0, rp_max*3.0]))

    num_threads = get_num_threads(num_threads, enforce_max_cores=False)

    if approx_cell1_size is None:
        approx_cell1_size = [rp_max, rp_max]
    elif custom_len(approx_cell1_size) == 1:
        approx_cell1_size = [approx_cell1_size, approx_cell1_size]
    if approx_cell2_size is None:
        approx_cell2_size = [rp_max, rp_max]
    elif custom_len(approx_cell2_size) == 1:
        approx_cell2_size = [approx_cell2_size, approx_cell2_size]

    return (x1, y1, x2, y2, effective_particle_masses, rp_bins,
            period, num_threads, PBCs, approx_cell1_size, approx_cell2_size)
