"""Drop-in for ``halotools.mock_observables.tpcf``
(/root/reference/halotools/mock_observables/two_point_clustering/tpcf.py:212-498)."""
from math import gamma
from warnings import warn

import numpy as np

from ..custom_exceptions import HalotoolsError
from ..helpers import (enforce_sample_has_correct_shape, get_num_threads, get_period,
                       get_separation_bins_array)
from ..pair_counters import npairs_3d
from ..pair_counters.mesh_helpers import _enforce_maximum_search_length
from .. import _lib
from .. import distributed as _dist
from . import _device, _driver
from .clustering_helpers import (process_optional_input_sample2, tpcf_estimator_dd_dr_rr_requirements,
                                 verify_tpcf_estimator)

__all__ = ["tpcf"]


def tpcf(sample1, rbins, sample2=None, randoms=None, period=None,
         do_auto=True, do_cross=True, estimator='Natural', num_threads=1,
         approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None,
         RR_precomputed=None, NR_precomputed=None, seed=None):
    """Real-space two-point correlation function xi(r) in the bins ``rbins`` (len(rbins)-1 values).

    Same arguments, return structure (xi_11 | (xi_11, xi_12, xi_22) | xi_12 | (xi_11, xi_22)) and
    errors as the reference.  With ``randoms=None`` in a periodic box the random counts are the
    analytic shell volumes; otherwise DR / RR are counted on the GPU."""
    (sample1, rbins, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs, RR_precomputed, NR_precomputed) = _tpcf_process_args(
        sample1, rbins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
        approx_cell1_size, approx_cell2_size, approx_cellran_size, RR_precomputed, NR_precomputed, seed)

    do_DD, do_DR, do_RR = tpcf_estimator_dd_dr_rr_requirements[estimator]
    if RR_precomputed is not None:
        do_RR = False

    N1, N2 = len(sample1), len(sample2)
    if randoms is not None:
        NR = len(randoms)
    else:
        NR = NR_precomputed if NR_precomputed is not None else N1

    if _device.available(npairs_3d):
        # K3 on the device: counts stay in HBM, one all-reduce, estimator kernel, ONE host synchronisation
        return _tpcf_device(sample1, rbins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
                            approx_cell1_size, approx_cell2_size, approx_cellran_size, RR_precomputed, same,
                            do_DR, do_RR, N1, N2, NR)

    # host restatement of the same flow (the pair counter was replaced: CPU tests of the driver logic)
    def count(a, b, cell_a, cell_b):
        return partial.add(np.diff(npairs_3d(a, b, rbins, period=period, num_threads=num_threads,
                                             approx_cell1_size=cell_a, approx_cell2_size=cell_b)))

    def analytic():
        return _analytic_randoms(sample1, sample2, rbins, period)

    # the engine's upload cache: every sample crosses PCIe once for all the counts of this call; multi-GPU: the
    # ranks' partial counts of ALL these calls are combined by one all-reduce at the end of the block
    partial = _dist.local_counts()
    with _lib.upload_cache(), partial:
        D1D1, D1D2, D2D2 = _driver.data_counts(count, sample1, sample2, same, do_auto, do_cross,
                                               approx_cell1_size, approx_cell2_size)
        D1R, D2R, RR = _driver.random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                             approx_cell1_size, approx_cell2_size, approx_cellran_size)
    if RR_precomputed is not None:
        RR = RR_precomputed
    return _driver.combine(same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator)


def _analytic_randoms(sample1, sample2, rbins, period):
    """shells of a periodic box populated at the mean density (tpcf.py:121-145)"""
    nr = len(sample1)
    dv = np.diff((np.pi ** 1.5 / gamma(2.5)) * rbins ** 3)
    volume = period.prod()
    D1R = nr * (dv * (np.shape(sample1)[0] / volume))
    D2R = nr * (dv * (np.shape(sample2)[0] / volume))
    return D1R, D2R, dv * ((nr ** 2) / volume)


def _tpcf_device(sample1, rbins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
                 approx_cell1_size, approx_cell2_size, approx_cellran_size, RR_precomputed, same,
                 do_DR, do_RR, N1, N2, NR):
    stat = _device.DeviceStatistic((len(rbins),))
    if same:
        sample1, randoms = stat.inputs(period is not None, sample1, randoms)
        sample2 = sample1
    else:
        sample1, sample2, randoms = stat.inputs(period is not None, sample1, sample2, randoms)

    def count(a, b, cell_a, cell_b):
        return stat.count(npairs_3d.enqueue, a, b, rbins, period=period, num_threads=num_threads,
                          approx_cell1_size=cell_a, approx_cell2_size=cell_b)

    def analytic():
        return tuple(stat.analytic(v) for v in _analytic_randoms(sample1, sample2, rbins, period))

    with _lib.upload_cache():
        D1D1, D1D2, D2D2 = _driver.data_counts(count, sample1, sample2, same, do_auto, do_cross,
                                               approx_cell1_size, approx_cell2_size)
        D1R, D2R, RR = _driver.random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                             approx_cell1_size, approx_cell2_size, approx_cellran_size)
        if RR_precomputed is not None:
            RR = stat.analytic(RR_precomputed)
        return _device.combine(stat, same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator)


def _tpcf_process_args(sample1, rbins, sample2, randoms, period,
                       do_auto, do_cross, estimator, num_threads,
                       approx_cell1_size, approx_cell2_size, approx_cellran_size,
                       RR_precomputed, NR_precomputed, seed):
    """Validation in the reference's order with the reference's messages (tpcf.py:501-600)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)
    if randoms is not None and not getattr(randoms, "is_cuda", False):
        randoms = np.atleast_1d(randoms)

    rbins = get_separation_bins_array(rbins)
    rmax = np.amax(rbins)
    period, PBCs = get_period(period)
    _enforce_maximum_search_length(rmax, period)

    if (randoms is None) & (PBCs is False):
        raise ValueError("If no PBCs are specified, randoms must be provided.\n")
    try:
        assert do_auto == bool(do_auto)
        assert do_cross == bool(do_cross)
    except Exception:
        raise ValueError("`do_auto` and `do_cross` keywords must be boolean-valued.")

    num_threads = get_num_threads(num_threads)
    verify_tpcf_estimator(estimator)

    if (RR_precomputed is not None) | (NR_precomputed is not None):
        if not ((RR_precomputed is not None) & (NR_precomputed is not None)):
            raise HalotoolsError("\nYou must either provide both "
                                 "``RR_precomputed`` and ``NR_precomputed`` arguments, or neither\n")
        if len(RR_precomputed) != len(rbins) - 1:
            raise HalotoolsError("\nLength of ``RR_precomputed`` must match length of ``rbins``\n")
        if np.any(RR_precomputed == 0):
            warn("RR_precomputed has radial bin(s) which contain no pairs. \n"
                 "Consider increasing the number of randoms, or using larger bins.")
        try:
            assert len(randoms) == NR_precomputed
        except AssertionError:
            raise HalotoolsError("If passing in randoms and also NR_precomputed, \n"
                                 "the value of NR_precomputed must agree with the number of randoms\n")

    assert np.all(rbins > 0.0), "All values of input ``rbins`` must be positive"
    return (sample1, rbins, sample2, randoms, period, do_auto, do_cross, num_threads,
            same, PBCs, RR_precomputed, NR_precomputed)
