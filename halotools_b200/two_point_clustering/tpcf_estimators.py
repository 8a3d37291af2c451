"""Two-point estimators over differential pair counts.

Behaviour (formulae, which counts each estimator needs, the zero-division ValueError text the
reference's tests look for) follows
/root/reference/halotools/mock_observables/two_point_clustering/tpcf_estimators.py:14-183 and the
requirement table of clustering_helpers.py:18-24.
"""
import numpy as np

from ..custom_exceptions import HalotoolsError

__all__ = ["_TP_estimator", "_TP_estimator_crossx", "_list_estimators", "_TP_estimator_requirements",
           "tpcf_estimator_dd_dr_rr_requirements"]

# estimator -> (do_DD, do_DR, do_RR)
tpcf_estimator_dd_dr_rr_requirements = {
    'Natural': (True, False, True),
    'Davis-Peebles': (True, True, False),
    'Hewett': (True, True, True),
    'Hamilton': (True, True, True),
    'Landy-Szalay': (True, True, True),
}

_ZERO_MSG = ("When calculating the two-point function, there was at least one \n"
             "separation bin with zero {0} pairs. Since the ``{1}`` estimator you chose \n"
             "divides by {0}, you will have at least one NaN returned value.\n"
             "Most likely, the innermost separation bin is the problem.\n"
             "Try increasing the number of randoms and/or using broader bins.\n"
             "To estimate the number of required randoms, the following expression \n"
             "for the expected number of pairs inside a sphere of radius ``r`` may be useful:\n\n"
             "<Npairs> = (Nran_tot)*(4pi/3)*(r/Lbox)^3 \n\n")


def _list_estimators():
    return ["Natural", "Davis-Peebles", "Hewett", "Hamilton", "Landy-Szalay"]


def _TP_estimator_requirements(estimator):
    """(do_DD, do_DR, do_RR) for ``estimator``; HalotoolsError for an unknown name."""
    if estimator not in tpcf_estimator_dd_dr_rr_requirements:
        raise HalotoolsError("Input `estimator` must be one of the following:{0}".format(_list_estimators()))
    return tpcf_estimator_dd_dr_rr_requirements[estimator]


def _test_for_zero_division(DD, DR, RR, ND1, ND2, NR1, NR2, estimator):
    if (estimator in ("Natural", "Davis-Peebles", "Hewett", "Landy-Szalay")) & (np.any(RR == 0)):
        raise ValueError(_ZERO_MSG.format("RR", estimator))
    if (estimator in ("Hamilton",)) & (np.any(DR == 0)):
        raise ValueError(_ZERO_MSG.format("DR", estimator))


def _norm(*ns):
    """Counts normalisations as 1-d arrays + the broadcasting product used by the jackknife
    callers (rows = subsamples) or plain multiplication otherwise."""
    arrs = [np.atleast_1d(n) for n in ns]
    if any(len(a) > 1 for a in arrs):
        def mult(x, y):
            return (x * y.T).T
    else:
        def mult(x, y):
            return x * y
    return arrs, mult


def _unwrap(xi):
    return xi[0] if np.shape(xi)[0] == 1 else xi


def _TP_estimator(DD, DR, RR, ND1, ND2, NR1, NR2, estimator):
    """xi from auto-correlation counts."""
    (ND1, ND2, NR1, NR2), mult = _norm(ND1, ND2, NR1, NR2)
    _test_for_zero_division(DD, DR, RR, ND1, ND2, NR1, NR2, estimator)
    with np.errstate(divide="ignore", invalid="ignore"):
        if estimator == "Natural":               # DD/RR - 1
            factor = ND1 * ND2 / (NR1 * NR2)
            xi = mult(1.0 / factor, DD / RR) - 1.0
        elif estimator == "Davis-Peebles":       # DD/DR - 1
            factor = ND1 * ND2 / (ND1 * NR2)
            xi = mult(1.0 / factor, DD / DR) - 1.0
        elif estimator == "Hewett":              # (DD - DR)/RR
            factor1 = ND1 * ND2 / (NR1 * NR2)
            factor2 = ND1 * NR2 / (NR1 * NR2)
            xi = mult(1.0 / factor1, DD / RR) - mult(1.0 / factor2, DR / RR)
        elif estimator == "Hamilton":            # DD RR / DR^2 - 1
            xi = (DD * RR) / (DR * DR) - 1.0
        elif estimator == "Landy-Szalay":        # (DD - 2 DR + RR)/RR
            factor1 = ND1 * ND2 / (NR1 * NR2)
            factor2 = ND1 * NR2 / (NR1 * NR2)
            xi = mult(1.0 / factor1, DD / RR) - mult(1.0 / factor2, 2.0 * DR / RR) + 1.0
        else:
            raise ValueError("unsupported estimator!")
    return _unwrap(xi)


def _TP_estimator_crossx(DD, D1R, D2R, RR, ND1, ND2, NR1, NR2, estimator):
    """xi from cross-correlation counts (Natural, Hamilton and Landy-Szalay only)."""
    (ND1, ND2, NR1, NR2), mult = _norm(ND1, ND2, NR1, NR2)
    _test_for_zero_division(DD, D1R, RR, ND1, ND2, NR1, NR2, estimator)
    _test_for_zero_division(DD, D2R, RR, ND1, ND2, NR1, NR2, estimator)
    with np.errstate(divide="ignore", invalid="ignore"):
        if estimator == "Natural":
            factor = ND1 * ND2 / (NR1 * NR2)
            xi = mult(1.0 / factor, DD / RR) - 1.0
        elif estimator == "Hamilton":
            xi = (DD * RR) / (D1R * D2R) - 1.0
        elif estimator == "Landy-Szalay":
            factor1 = ND1 * ND2 / (NR1 * NR2)
            factor2 = ND1 * NR2 / (NR1 * NR2)
            xi = (mult(1.0 / factor1, DD / RR) - mult(1.0 / factor2, D1R / RR)
                  - mult(1.0 / factor2, D2R / RR) + 1.0)
        else:
            raise ValueError("{0} estimator is not supported for cross-correlations".format(estimator))
    return _unwrap(xi)
