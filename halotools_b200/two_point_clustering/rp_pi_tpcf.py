"""Drop-in for ``halotools.mock_observables.rp_pi_tpcf``
(/root/reference/halotools/mock_observables/two_point_clustering/rp_pi_tpcf.py:33-535)."""
from math import pi

import numpy as np

from ..helpers import (enforce_sample_has_correct_shape, get_line_of_sight_bins_array, get_num_threads,
                       get_period, get_separation_bins_array)
from ..pair_counters import npairs_xy_z
from ..pair_counters.mesh_helpers import _enforce_maximum_search_length
from .. import _lib
from .. import distributed as _dist
from . import _device, _driver
from .clustering_helpers import process_optional_input_sample2, verify_tpcf_estimator
from .tpcf_estimators import _TP_estimator_requirements

__all__ = ["rp_pi_tpcf"]


def rp_pi_tpcf(sample1, rp_bins, pi_bins, sample2=None, randoms=None, period=None,
               do_auto=True, do_cross=True, estimator='Natural', num_threads=1,
               approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None, seed=None):
    """Redshift-space correlation function xi(rp, pi): (len(rp_bins)-1, len(pi_bins)-1) arrays,
    z being the line of sight.  Same arguments / return structure / errors as the reference."""
    return _rp_pi_tpcf(sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
                       approx_cell1_size, approx_cell2_size, approx_cellran_size, seed, 0.0)


def _rp_pi_tpcf(sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
                approx_cell1_size, approx_cell2_size, approx_cellran_size, seed, wp_pi_max):
    """``wp_pi_max`` > 0 (from ``wp``, device path only): the estimator kernel also applies wp's 2 * xi[:, 0] * pi_max."""
    (sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs) = _rp_pi_tpcf_process_args(
        sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
        approx_cell1_size, approx_cell2_size, approx_cellran_size, seed)

    do_DD, do_DR, do_RR = _TP_estimator_requirements(estimator)
    N1, N2 = len(sample1), len(sample2)
    NR = len(randoms) if randoms is not None else N1

    def analytic():
        # annular cylinders of a periodic box at the mean density (rp_pi_tpcf.py:443-467)
        nr = len(sample1)
        v = pi * np.outer(rp_bins ** 2.0, 2.0 * pi_bins)
        dv = np.diff(np.diff(v, axis=0), axis=1)
        volume = period.prod()
        n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
        D1R = n1 * (dv * (n1 / volume))
        D2R = n2 * (dv * (n2 / volume))
        return D1R, D2R, dv * (nr ** 2 / volume)

    if _device.available(npairs_xy_z):
        # K3 on the device: counts stay in HBM, one all-reduce, estimator kernel, ONE host synchronisation
        stat = _device.DeviceStatistic((len(rp_bins), len(pi_bins)))
        if same:
            sample1, randoms = stat.inputs(PBCs, sample1, randoms)
            sample2 = sample1
        else:
            sample1, sample2, randoms = stat.inputs(PBCs, sample1, sample2, randoms)

        def dcount(a, b, cell_a, cell_b):
            return stat.count(npairs_xy_z.enqueue, a, b, rp_bins, pi_bins, period=period, num_threads=num_threads,
                              approx_cell1_size=cell_a, approx_cell2_size=cell_b)

        def danalytic():
            return tuple(stat.analytic(v) for v in analytic())

        with _lib.upload_cache():
            D1D1, D1D2, D2D2 = _driver.data_counts(dcount, sample1, sample2, same, do_auto, do_cross,
                                                   approx_cell1_size, approx_cell2_size, always_auto1=True)
            D1R, D2R, RR = _driver.random_counts(dcount, danalytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                                 approx_cell1_size, approx_cell2_size, approx_cellran_size)
            return _device.combine(stat, same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator,
                                   wp_pi_max=wp_pi_max, out_shape=(len(rp_bins) - 1,) if wp_pi_max > 0.0 else None)

    # host restatement of the same flow (the pair counter was replaced: CPU tests of the driver logic)
    def count(a, b, cell_a, cell_b):
        c = npairs_xy_z(a, b, rp_bins, pi_bins, period=period, num_threads=num_threads,
                        approx_cell1_size=cell_a, approx_cell2_size=cell_b)
        return partial.add(np.diff(np.diff(c, axis=0), axis=1))

    # the engine's upload cache: every sample crosses PCIe once for all the counts of this call; multi-GPU: the
    # ranks' partial counts of ALL these calls are combined by one all-reduce at the end of the block
    partial = _dist.local_counts()
    with _lib.upload_cache(), partial:
        D1D1, D1D2, D2D2 = _driver.data_counts(count, sample1, sample2, same, do_auto, do_cross,
                                               approx_cell1_size, approx_cell2_size, always_auto1=True)
        D1R, D2R, RR = _driver.random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                             approx_cell1_size, approx_cell2_size, approx_cellran_size)
    return _driver.combine(same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator)


def _rp_pi_tpcf_process_args(sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross,
                             estimator, num_threads, approx_cell1_size, approx_cell2_size,
                             approx_cellran_size, seed):
    """Validation in the reference's order (rp_pi_tpcf.py:470-535)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)
    if randoms is not None and not getattr(randoms, "is_cuda", False):
        randoms = np.atleast_1d(randoms)

    rp_bins = get_separation_bins_array(rp_bins)
    rp_max = np.amax(rp_bins)
    pi_bins = get_line_of_sight_bins_array(pi_bins)
    pi_max = np.amax(pi_bins)

    period, PBCs = get_period(period)
    _enforce_maximum_search_length([rp_max, rp_max, pi_max], period)

    if (randoms is None) & (PBCs is False):
        raise ValueError("If no PBCs are specified, randoms must be provided.\n")
    try:
        assert do_auto == bool(do_auto)
        assert do_cross == bool(do_cross)
    except Exception:
        raise ValueError("`do_auto` and `do_cross` keywords must be boolean-valued.")

    num_threads = get_num_threads(num_threads)
    verify_tpcf_estimator(estimator)
    return (sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, num_threads,
            same, PBCs)
