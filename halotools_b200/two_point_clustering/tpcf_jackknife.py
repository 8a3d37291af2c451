"""Drop-ins for ``halotools.mock_observables.tpcf_jackknife`` and ``wp_jackknife``
(/root/reference/halotools/mock_observables/two_point_clustering/tpcf_jackknife.py:33-653,
wp_jackknife.py:36-633): the statistic of the full sample plus its jackknife covariance over cuboid sub-volumes.
One driver serves both; the pair counts come from the jackknife counters of the GPU engine."""
import numpy as np

from .. import _lib
from ..catalog_analysis_helpers import cuboid_subvolume_labels
from ..custom_exceptions import HalotoolsError
from ..helpers import (enforce_sample_has_correct_shape, get_line_of_sight_bins_array, get_num_threads,
                       get_period, get_separation_bins_array)
from ..pair_counters import npairs_jackknife_3d, npairs_jackknife_xy_z
from ..pair_counters.mesh_helpers import _enforce_maximum_search_length
from .clustering_helpers import process_optional_input_sample2, verify_tpcf_estimator
from .marked_tpcf import _SeededNumpyRNG
from .tpcf_estimators import _TP_estimator, _TP_estimator_crossx, _TP_estimator_requirements

__all__ = ("tpcf_jackknife", "wp_jackknife", "rp_pi_tpcf_jackknife")

np.seterr(divide="ignore", invalid="ignore")  # as the reference modules do (tpcf_jackknife.py:30)


def tpcf_jackknife(sample1, randoms, rbins, Nsub=[5, 5, 5], sample2=None, period=None,
                   do_auto=True, do_cross=True, estimator="Natural", num_threads=1, seed=None):
    """xi(r) of the full sample and the jackknife covariance matrix over ``prod(Nsub)`` cuboid sub-volumes; the
    reference's return structure: ``xi, cov`` | ``xi_11, xi_12, xi_22, cov_11, cov_12, cov_22`` | ..."""
    (sample1, rbins, Nsub, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs) = _tpcf_jackknife_process_args(sample1, randoms, rbins, Nsub, sample2, period,
                                                do_auto, do_cross, estimator, num_threads, seed)

    def count(a, b, ja, jb, nsub):
        c = npairs_jackknife_3d(a, b, rbins, period=period, jtags1=ja, jtags2=jb, N_samples=nsub,
                                num_threads=num_threads)
        return np.diff(c, axis=1)

    return _jackknife_statistic(count, lambda c: c, 1.0, sample1, sample2, randoms, Nsub, period, PBCs, same,
                                do_auto, do_cross, estimator)


def wp_jackknife(sample1, randoms, rp_bins, pi_max, Nsub=[5, 5, 5], sample2=None, period=None,
                 do_auto=True, do_cross=True, estimator="Natural", num_threads=1, seed=None,
                 approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None):
    """w_p(r_p) of the full sample (integration to ``pi_max``) and its jackknife covariance matrix."""
    pi_bins = np.array([0.0, float(pi_max)])
    (sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs) = _wp_jackknife_tpcf_process_args(sample1, rp_bins, pi_bins, sample2, randoms, period,
                                                   do_auto, do_cross, estimator, num_threads, seed)

    def count(a, b, ja, jb, nsub):
        c = npairs_jackknife_xy_z(a, b, rp_bins, pi_bins, period=period, jtags1=ja, jtags2=jb, N_samples=nsub,
                                  num_threads=num_threads)
        return np.diff(np.diff(c, axis=1), axis=2)

    # the single pi bin is dropped (wp_jackknife.py:296-345) and the estimator scaled by 2 pi_max (:395-403)
    return _jackknife_statistic(count, lambda c: c[:, :, 0], 2.0 * pi_max, sample1, sample2, randoms, Nsub, period,
                                PBCs, same, do_auto, do_cross, estimator)


def rp_pi_tpcf_jackknife(sample1, randoms, rp_bins, pi_bins, Nsub=[5, 5, 5], sample2=None, period=None,
                         do_auto=True, do_cross=True, estimator="Natural", num_threads=1, seed=None,
                         approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None):
    """xi(rp, pi) of the full sample ((len(rp_bins)-1, len(pi_bins)-1) values) and its jackknife covariance matrix over
    the row-major flattened bins (rp_pi_tpcf_jackknife.py:36-670)."""
    (sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs) = _wp_jackknife_tpcf_process_args(sample1, rp_bins, pi_bins, sample2, randoms, period,
                                                   do_auto, do_cross, estimator, num_threads, seed, error=KeyError)

    def count(a, b, ja, jb, nsub):
        c = npairs_jackknife_xy_z(a, b, rp_bins, pi_bins, period=period, jtags1=ja, jtags2=jb, N_samples=nsub,
                                  num_threads=num_threads)
        return np.diff(np.diff(c, axis=1), axis=2)

    return _jackknife_statistic(count, lambda c: c, 1.0, sample1, sample2, randoms, Nsub, period, PBCs, same,
                                do_auto, do_cross, estimator, flatten=True)


def _jackknife_statistic(count, squeeze, scale, sample1, sample2, randoms, Nsub, period, PBCs, same,
                         do_auto, do_cross, estimator, flatten=False):
    """tpcf_jackknife.py:260-396 / wp_jackknife.py:262-420."""
    if PBCs is False:
        sample1, sample2, randoms, Lbox = _enclose_in_box(sample1, sample2, randoms)
    else:
        Lbox = period

    do_DD, do_DR, do_RR = _TP_estimator_requirements(estimator)
    N1, N2, NR = len(sample1), len(sample2), len(randoms)

    j_index_1, N_sub_vol = cuboid_subvolume_labels(sample1, Nsub, Lbox)
    j_index_2, N_sub_vol = cuboid_subvolume_labels(sample2, Nsub, Lbox)
    j_index_random, N_sub_vol = cuboid_subvolume_labels(randoms, Nsub, Lbox)

    # points left in each jackknife sample
    NR_subs = NR - get_subvolume_numbers(j_index_random, N_sub_vol)
    N1_subs = N1 - get_subvolume_numbers(j_index_1, N_sub_vol)
    N2_subs = N2 - get_subvolume_numbers(j_index_2, N_sub_vol)

    def cov(xi_sub):
        # rows = jackknife samples; 2-d statistics are flattened row-major first (rp_pi_tpcf_jackknife.py:423-441)
        if flatten:
            xi_sub = np.reshape(xi_sub, (N_sub_vol, -1))
        return np.array(np.cov(xi_sub.T, bias=True)) * (N_sub_vol - 1.0)

    def full_sub(c):
        if c is None:
            return None, None
        c = squeeze(c)
        return c[0], c[1:]

    with _lib.upload_cache():
        # data pairs (jnpair_counts, tpcf_jackknife.py:469-533)
        D1D1 = count(sample1, sample1, j_index_1, j_index_1, N_sub_vol) if do_auto else None
        if same:
            D1D2 = D2D2 = D1D1
        else:
            D1D2 = count(sample1, sample2, j_index_1, j_index_2, N_sub_vol) if do_cross else None
            D2D2 = count(sample2, sample2, j_index_2, j_index_2, N_sub_vol) if do_auto else None
        # random pairs (jrandom_counts, :536-583)
        D1R = count(sample1, randoms, j_index_1, j_index_random, N_sub_vol) if do_DR else None
        RR = count(randoms, randoms, j_index_random, j_index_random, N_sub_vol) if do_RR else None
        if same:
            D2R = D1R
        else:
            D2R = count(sample2, randoms, j_index_2, j_index_random, N_sub_vol) if do_DR else None

    D1D1_full, D1D1_sub = full_sub(D1D1)
    D1D2_full, D1D2_sub = full_sub(D1D2)
    D2D2_full, D2D2_sub = full_sub(D2D2)
    D1R_full, D1R_sub = full_sub(D1R)
    D2R_full, D2R_sub = full_sub(D2R)
    RR_full, RR_sub = full_sub(RR)

    if do_auto is True:
        xi_11_full = scale * _TP_estimator(D1D1_full, D1R_full, RR_full, N1, N1, NR, NR, estimator)
        xi_22_full = scale * _TP_estimator(D2D2_full, D2R_full, RR_full, N2, N2, NR, NR, estimator)
        xi_11_sub = scale * _TP_estimator(D1D1_sub, D1R_sub, RR_sub, N1_subs, N1_subs, NR_subs, NR_subs, estimator)
        xi_22_sub = scale * _TP_estimator(D2D2_sub, D2R_sub, RR_sub, N2_subs, N2_subs, NR_subs, NR_subs, estimator)
        xi_11_cov = cov(xi_11_sub)
        xi_22_cov = cov(xi_22_sub)
    if do_cross is True:
        xi_12_full = scale * _TP_estimator_crossx(D1D2_full, D1R_full, D2R_full, RR_full, N1, N2, NR, NR, estimator)
        xi_12_sub = scale * _TP_estimator_crossx(D1D2_sub, D1R_sub, D2R_sub, RR_sub,
                                                 N1_subs, N2_subs, NR_subs, NR_subs, estimator)
        xi_12_cov = cov(xi_12_sub)

    if same:
        return xi_11_full, xi_11_cov
    if (do_auto is True) & (do_cross is True):
        return xi_11_full, xi_12_full, xi_22_full, xi_11_cov, xi_12_cov, xi_22_cov
    elif do_auto is True:
        return xi_11_full, xi_22_full, xi_11_cov, xi_22_cov
    elif do_cross is True:
        return xi_12_full, xi_12_cov


def _enclose_in_box(data1, data2, data3):
    """Shift the three samples so that the smallest coordinate is 0; cube side = largest extent
    (tpcf_jackknife.py:399-437)."""
    lo = min(np.min(d[:, :3]) for d in (data1, data2, data3))
    hi = max(np.max(d[:, :3]) for d in (data1, data2, data3)) - lo
    out = [np.vstack((d[:, 0] - lo, d[:, 1] - lo, d[:, 2] - lo)).T for d in (data1, data2, data3)]
    return out[0], out[1], out[2], np.array([hi, hi, hi])


def get_subvolume_numbers(j_index, N_sub_vol):
    """Points per sub-volume, empty sub-volumes included (tpcf_jackknife.py:440-454)."""
    temp = np.hstack((j_index, np.arange(1, N_sub_vol + 1, 1)))
    labels, N = np.unique(temp, return_counts=True)
    return N - 1


def _process_randoms(randoms, period, PBCs, seed, error):
    """``randoms = [N]`` asks for N uniform randoms in the periodic box (tpcf_jackknife.py:601-613)."""
    if np.shape(randoms) == (1,):
        N_randoms = randoms[0]
        if PBCs is True:
            with _SeededNumpyRNG(seed):
                randoms = np.random.random((N_randoms, 3)) * period
        else:
            msg = ("\n When no `period` parameter is passed, \n"
                   "the user must provide true randoms, and \n"
                   "not just the number of randoms desired.")
            raise error(msg)
    return randoms


def _tpcf_jackknife_process_args(sample1, randoms, rbins, Nsub, sample2, period,
                                 do_auto, do_cross, estimator, num_threads, seed):
    """Validation in the reference's order with the reference's messages (tpcf_jackknife.py:586-653)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)
    period, PBCs = get_period(period)
    randoms = _process_randoms(randoms, period, PBCs, seed, HalotoolsError)

    rbins = get_separation_bins_array(rbins)
    rmax = np.amax(rbins)

    Nsub = np.atleast_1d(Nsub)
    if len(Nsub) == 1:
        Nsub = np.array([Nsub[0]] * 3)
    try:
        assert np.all(Nsub < np.inf)
        assert np.all(Nsub > 0)
    except AssertionError:
        raise HalotoolsError("\n Input `Nsub` must be a bounded positive number in all dimensions")

    _enforce_maximum_search_length(rmax, period)

    try:
        assert do_auto == bool(do_auto)
        assert do_cross == bool(do_cross)
    except Exception:
        raise ValueError("`do_auto` and `do_cross` keywords must be boolean-valued.")

    num_threads = get_num_threads(num_threads)
    verify_tpcf_estimator(estimator)
    return (sample1, rbins, Nsub, sample2, randoms, period, do_auto, do_cross, num_threads, same, PBCs)


def _wp_jackknife_tpcf_process_args(sample1, rp_bins, pi_bins, sample2, randoms, period,
                                    do_auto, do_cross, estimator, num_threads, seed, error=ValueError):
    """wp_jackknife.py:557-633 (rp_pi_tpcf_jackknife.py:575-660 differs only in raising KeyError for randoms given by
    number without a period)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)
    if randoms is not None:
        randoms = np.atleast_1d(randoms)

    rp_bins = get_separation_bins_array(rp_bins)
    rp_max = np.amax(rp_bins)
    pi_bins = get_line_of_sight_bins_array(pi_bins)
    pi_max = np.amax(pi_bins)

    period, PBCs = get_period(period)
    randoms = _process_randoms(randoms, period, PBCs, seed, error)

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
    return (sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, num_threads, same, PBCs)
