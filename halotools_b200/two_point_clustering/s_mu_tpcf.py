"""Drop-ins for ``halotools.mock_observables.s_mu_tpcf`` and ``tpcf_multipole``
(/root/reference/halotools/mock_observables/two_point_clustering/s_mu_tpcf.py:33-586, tpcf_multipole.py:15-88)."""
import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import (enforce_sample_has_correct_shape, get_line_of_sight_bins_array, get_num_threads,
                       get_period, get_separation_bins_array)
from ..pair_counters import npairs_s_mu
from ..pair_counters.mesh_helpers import _enforce_maximum_search_length
from . import _driver
from .clustering_helpers import process_optional_input_sample2, verify_tpcf_estimator
from .tpcf_estimators import _TP_estimator_requirements

__all__ = ("s_mu_tpcf", "tpcf_multipole")

np.seterr(divide="ignore", invalid="ignore")  # as the reference module does (s_mu_tpcf.py:30)


def s_mu_tpcf(sample1, s_bins, mu_bins, sample2=None, randoms=None, period=None,
              do_auto=True, do_cross=True, estimator="Natural", num_threads=1,
              approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None, seed=None):
    """Redshift-space correlation function xi(s, mu) in the bins ``s_bins`` x ``mu_bins``
    ((len(s_bins)-1, len(mu_bins)-1) values per returned array; mu = cos of the angle to the line of sight, z).
    Same arguments, return structure and errors as the reference."""
    (sample1, s_bins, mu_bins, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs) = _s_mu_tpcf_process_args(sample1, s_bins, mu_bins, sample2, randoms, period,
                                           do_auto, do_cross, estimator, num_threads)
    do_DD, do_DR, do_RR = _TP_estimator_requirements(estimator)
    N1, N2 = len(sample1), len(sample2)
    NR = len(randoms) if randoms is not None else N1

    def count(a, b, cell_a, cell_b):
        c = npairs_s_mu(a, b, s_bins, mu_bins, period=period, num_threads=num_threads,
                        approx_cell1_size=cell_a, approx_cell2_size=cell_b)
        return partial.add(np.diff(np.diff(c, axis=0), axis=1))

    def analytic():
        # spherical wedge sectors of a periodic box at the mean density (s_mu_tpcf.py:324-334,412-440)
        nr = len(sample1)
        mu_rev = np.sort(mu_bins)[::-1]
        theta = np.arccos(mu_rev)
        vol = (2.0 * np.pi / 3.0) * np.outer((s_bins ** 3.0), (1.0 - np.cos(theta))) * 2.0
        dv = np.diff(np.diff(vol, axis=1), axis=0)
        volume = period.prod()
        n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
        D1R = (n1 - 1.0) * (dv * (n1 / volume))
        D2R = (n2 - 1.0) * (dv * (n2 / volume))
        return D1R, D2R, dv * (nr ** 2 / volume)

    partial = _dist.local_counts()
    with _lib.upload_cache(), partial:
        D1D1, D1D2, D2D2 = _driver.data_counts(count, sample1, sample2, same, do_auto, do_cross,
                                               approx_cell1_size, approx_cell2_size)
        D1R, D2R, RR = _driver.random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                             approx_cell1_size, approx_cell2_size, approx_cellran_size)
    xi = _driver.combine(same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator)
    # the counts run in order of increasing theta_LOS, i.e. decreasing mu: reverse the mu axis (s_mu_tpcf.py:299-321)
    if isinstance(xi, tuple):
        return tuple(x[:, ::-1] for x in xi)
    return xi[:, ::-1]


def tpcf_multipole(s_mu_tcpf_result, mu_bins, order=0):
    """Multipole of order ``order`` of xi(s, mu): numerical integration over the mu bins (tpcf_multipole.py:72-88)."""
    from scipy.special import legendre
    s_mu_tcpf_result = np.atleast_1d(s_mu_tcpf_result)
    mu_bins = np.atleast_1d(mu_bins)
    order = int(order)
    mu_bin_centers = (mu_bins[:-1] + mu_bins[1:]) / 2.0
    Ln = legendre(order)
    return (2.0 * order + 1.0) / 2.0 * np.sum(
        s_mu_tcpf_result * np.diff(mu_bins) * (Ln(mu_bin_centers) + Ln(-1.0 * mu_bin_centers)), axis=1)


def _s_mu_tpcf_process_args(sample1, s_bins, mu_bins, sample2, randoms, period,
                            do_auto, do_cross, estimator, num_threads):
    """Validation in the reference's order with the reference's messages (s_mu_tpcf.py:514-586)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)
    if randoms is not None:
        randoms = np.atleast_1d(randoms)

    s_bins = get_separation_bins_array(s_bins)
    s_max = np.max(s_bins)
    mu_bins = get_line_of_sight_bins_array(mu_bins)
    if (np.min(mu_bins) < 0.0) | (np.max(mu_bins) > 1.0):
        raise ValueError("`mu_bins` must be in the range [0,1].")

    period, PBCs = get_period(period)
    _enforce_maximum_search_length(s_max, period)

    if (randoms is None) & (PBCs is False):
        raise ValueError("\n If no PBCs are specified, randoms must be provided.\n")
    try:
        assert do_auto == bool(do_auto)
        assert do_cross == bool(do_cross)
    except Exception:
        raise ValueError("`do_auto` and `do_cross` keywords must be boolean-valued.")

    num_threads = get_num_threads(num_threads)
    verify_tpcf_estimator(estimator)
    return (sample1, s_bins, mu_bins, sample2, randoms, period, do_auto, do_cross, num_threads, same, PBCs)
