"""Shared control flow of tpcf / rp_pi_tpcf: which of D1D1, D1D2, D2D2, D1R, D2R, RR are counted
and how they are combined.  One table-driven driver serves the 3-d and the (rp, pi) statistic; the
decisions follow /root/reference/halotools/mock_observables/two_point_clustering/tpcf.py:39-205,
428-498 and rp_pi_tpcf.py:228-293,296-467 (including rp_pi_tpcf always counting D1D1, :312-322).
"""
import numpy as np

from .tpcf_estimators import _TP_estimator, _TP_estimator_crossx

__all__ = ("data_counts", "random_counts", "combine")


def data_counts(count, sample1, sample2, same, do_auto, do_cross, cell1, cell2, always_auto1=False):
    """(D1D1, D1D2, D2D2) differential counts; ``count(a, b, cell_a, cell_b)`` returns them."""
    D1D1 = count(sample1, sample1, cell1, cell1) if (do_auto or always_auto1) else None
    if same:
        return D1D1, D1D1, D1D1
    D1D2 = count(sample1, sample2, cell1, cell2) if do_cross else None
    D2D2 = count(sample2, sample2, cell2, cell2) if do_auto else None
    return D1D1, D1D2, D2D2


def random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR, cell1, cell2, cellran):
    """(D1R, D2R, RR): counted against ``randoms`` when given, else analytic (periodic box)."""
    if randoms is None:
        return analytic()
    RR = count(randoms, randoms, cellran, cellran) if do_RR else None
    D1R = count(sample1, randoms, cell1, cellran) if do_DR else None
    D2R = count(sample2, randoms, cell2, cellran) if (do_DR and not same) else None
    return D1R, D2R, RR


def combine(same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator):
    if same:
        return _TP_estimator(D1D1, D1R, RR, N1, N1, NR, NR, estimator)
    if (do_auto is True) & (do_cross is True):
        xi_11 = _TP_estimator(D1D1, D1R, RR, N1, N1, NR, NR, estimator)
        xi_12 = _TP_estimator_crossx(D1D2, D1R, D2R, RR, N1, N2, NR, NR, estimator)
        xi_22 = _TP_estimator(D2D2, D2R, RR, N2, N2, NR, NR, estimator)
        return xi_11, xi_12, xi_22
    elif do_cross is True:
        return _TP_estimator_crossx(D1D2, D1R, D2R, RR, N1, N2, NR, NR, estimator)
    elif do_auto is True:
        xi_11 = _TP_estimator(D1D1, D1R, RR, N1, N1, NR, NR, estimator)
        xi_22 = _TP_estimator(D2D2, D2R, RR, N2, N2, NR, NR, estimator)
        return xi_11, xi_22
