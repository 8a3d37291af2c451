"""Drop-in for ``halotools.mock_observables.tpcf_one_two_halo_decomp``
(/root/reference/halotools/mock_observables/two_point_clustering/tpcf_one_two_halo_decomp.py:36-631): xi(r) split
into pairs that share a host halo (marking function 3) and pairs that do not (marking function 4)."""
from math import gamma

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..custom_exceptions import HalotoolsError
from ..helpers import (enforce_sample_has_correct_shape, get_num_threads, get_period,
                       get_separation_bins_array)
from ..pair_counters import marked_npairs_3d, npairs_3d
from ..pair_counters.mesh_helpers import _enforce_maximum_search_length
from . import _driver
from .clustering_helpers import process_optional_input_sample2, verify_tpcf_estimator
from .tpcf_estimators import _TP_estimator_requirements

__all__ = ("tpcf_one_two_halo_decomp",)

np.seterr(divide="ignore", invalid="ignore")  # as the reference module does (:33)


def tpcf_one_two_halo_decomp(sample1, sample1_host_halo_id, rbins,
                             sample2=None, sample2_host_halo_id=None, randoms=None, period=None,
                             do_auto=True, do_cross=True, estimator="Natural", num_threads=1,
                             approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None, seed=None):
    """One-halo and two-halo terms of xi(r): the reference's return structure (one_halo_11, two_halo_11 | the six /
    two / four arrays of the cross-correlation cases)."""
    (sample1, sample1_host_halo_id, rbins, sample2, sample2_host_halo_id, randoms, period,
     do_auto, do_cross, num_threads, same, PBCs) = _process_args(
        sample1, sample1_host_halo_id, rbins, sample2, sample2_host_halo_id, randoms, period,
        do_auto, do_cross, estimator, num_threads)

    do_DD, do_DR, do_RR = _TP_estimator_requirements(estimator)
    N1, N2 = len(sample1), len(sample2)
    NR = len(randoms) if randoms is not None else N1

    # halo id + a column of ones: the marking functions return w1[1] * w2[1] = 1 for the selected pairs (:486-489)
    marks = {id(sample1): np.vstack((sample1_host_halo_id, np.ones(len(sample1_host_halo_id)))).T}
    marks.setdefault(id(sample2), np.vstack((sample2_host_halo_id, np.ones(len(sample2_host_halo_id)))).T)

    def marked_count(wfunc):
        def count(a, b, cell_a, cell_b):
            return partial.add(np.diff(marked_npairs_3d(a, b, rbins, weights1=marks[id(a)], weights2=marks[id(b)],
                                                        weight_func_id=wfunc, period=period, num_threads=num_threads)))
        return count

    def count(a, b, cell_a, cell_b):
        return partial.add(np.diff(npairs_3d(a, b, rbins, period=period, num_threads=num_threads,
                                             approx_cell1_size=cell_a, approx_cell2_size=cell_b)))

    def analytic():
        # shells of a periodic box populated at the mean density (:440-466)
        nr = len(sample1)
        dv = np.diff((np.pi ** 1.5 / gamma(2.5)) * rbins ** 3)
        volume = period.prod()
        D1R = nr * (dv * (np.shape(sample1)[0] / volume))
        D2R = nr * (dv * (np.shape(sample2)[0] / volume))
        return D1R, D2R, dv * ((nr ** 2) / volume)

    partial = _dist.local_counts()
    with _lib.upload_cache(), partial:
        one = _driver.data_counts(marked_count(3), sample1, sample2, same, do_auto, do_cross, None, None)
        two = _driver.data_counts(marked_count(4), sample1, sample2, same, do_auto, do_cross, None, None)
        D1R, D2R, RR = _driver.random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                             approx_cell1_size, approx_cell2_size, approx_cellran_size)
    xi1 = _driver.combine(same, do_auto, do_cross, one[0], one[1], one[2], D1R, D2R, RR, N1, N2, NR, estimator)
    xi2 = _driver.combine(same, do_auto, do_cross, two[0], two[1], two[2], D1R, D2R, RR, N1, N2, NR, estimator)
    if not isinstance(xi1, tuple):
        return xi1, xi2
    # (one_11, two_11, one_12, two_12, one_22, two_22) etc.: the terms interleaved (:283-358)
    out = []
    for a, b in zip(xi1, xi2):
        out += [a, b]
    return tuple(out)


def _process_args(sample1, sample1_host_halo_id, rbins, sample2, sample2_host_halo_id, randoms, period,
                  do_auto, do_cross, estimator, num_threads):
    """Validation in the reference's order with the reference's messages (:542-631)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample1_host_halo_id = np.atleast_1d(sample1_host_halo_id).astype(int)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)
    if same is True:
        sample2_host_halo_id = sample1_host_halo_id
    else:
        if sample2_host_halo_id is None:
            raise ValueError("If passing an input ``sample2``, must also pass sample2_host_halo_id")
        sample2_host_halo_id = np.atleast_1d(sample2_host_halo_id).astype(int)
    if randoms is not None:
        randoms = np.atleast_1d(randoms)

    if np.shape(sample1_host_halo_id) != (len(sample1),):
        raise HalotoolsError("\n `sample1_host_halo_id` must be a 1-D \narray the same length as `sample1`.")
    if np.shape(sample2_host_halo_id) != (len(sample2),):
        raise HalotoolsError("\n `sample2_host_halo_id` must be a 1-D \narray the same length as `sample2`.")

    rbins = get_separation_bins_array(rbins)
    rmax = np.max(rbins)
    period, PBCs = get_period(period)
    _enforce_maximum_search_length(rmax, period)

    if (randoms is None) & (PBCs is False):
        raise HalotoolsError("\n If no PBCs are specified, randoms must be provided.")
    try:
        assert do_auto == bool(do_auto)
        assert do_cross == bool(do_cross)
    except Exception:
        raise ValueError("`do_auto` and `do_cross` keywords must be boolean-valued.")

    num_threads = get_num_threads(num_threads)
    verify_tpcf_estimator(estimator)
    return (sample1, sample1_host_halo_id, rbins, sample2, sample2_host_halo_id, randoms, period,
            do_auto, do_cross, num_threads, same, PBCs)
