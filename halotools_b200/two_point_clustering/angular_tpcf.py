"""Drop-in for ``halotools.mock_observables.angular_tpcf``
(/root/reference/halotools/mock_observables/two_point_clustering/angular_tpcf.py:28-392): w(theta) from chord-length
pair counts of points on the unit sphere (non-periodic ``npairs_3d``)."""
import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..custom_exceptions import HalotoolsError
from ..helpers import array_is_monotonic, get_num_threads
from ..pair_counters import npairs_3d
from . import _driver
from .clustering_helpers import process_optional_input_sample2, verify_tpcf_estimator
from .tpcf_estimators import _TP_estimator_requirements

__all__ = ("angular_tpcf",)

np.seterr(divide="ignore", invalid="ignore")  # as the reference module does (:25)


def spherical_to_cartesian(ra, dec):
    """(ra, dec) in degrees -> unit vectors (/root/reference/halotools/utils/spherical_geometry.py:12-43)."""
    rar = np.radians(ra)
    decr = np.radians(dec)
    return np.cos(rar) * np.cos(decr), np.sin(rar) * np.cos(decr), np.sin(decr)


def chord_to_cartesian(theta, radians=True):
    """Chord length on the unit sphere of the angle theta (spherical_geometry.py:46-79)."""
    theta = np.asarray(theta, dtype=float)
    if radians is False:
        theta = np.radians(theta)
    return 2.0 * np.sin(theta / 2.0)


def angular_tpcf(sample1, theta_bins, sample2=None, randoms=None,
                 do_auto=True, do_cross=True, estimator="Natural", num_threads=1):
    """Angular two-point correlation function w(theta) of (ra, dec) samples in degrees, in the bins ``theta_bins``
    (degrees).  Same arguments, return structure and errors as the reference."""
    (sample1, theta_bins, sample2, randoms, do_auto, do_cross, num_threads,
     same) = _angular_tpcf_process_args(sample1, theta_bins, sample2, randoms, do_auto, do_cross,
                                        estimator, num_threads)

    chord_bins = chord_to_cartesian(theta_bins, radians=False)
    x, y, z = spherical_to_cartesian(sample1[:, 0], sample1[:, 1])
    sample1 = np.vstack((x, y, z)).T
    if same:
        sample2 = sample1
    else:
        x, y, z = spherical_to_cartesian(sample2[:, 0], sample2[:, 1])
        sample2 = np.vstack((x, y, z)).T
    if randoms is not None:
        x, y, z = spherical_to_cartesian(randoms[:, 0], randoms[:, 1])
        randoms = np.vstack((x, y, z)).T

    do_DD, do_DR, do_RR = _TP_estimator_requirements(estimator)
    N1, N2 = len(sample1), len(sample2)
    NR = len(randoms) if randoms is not None else N1

    def count(a, b, cell_a, cell_b):
        return partial.add(np.diff(npairs_3d(a, b, chord_bins, num_threads=num_threads)))

    def analytic():
        # spherical caps of the unit sphere at the mean density (:202-232)
        nr = len(sample1)
        h = 1.0 - np.sqrt(1.0 - chord_bins ** 2)
        da = np.diff(np.pi * (chord_bins ** 2 + h ** 2))
        area = 4.0 * np.pi
        n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
        D1R = n1 * (da * (n1 / area))
        D2R = n2 * (da * (n2 / area))
        return D1R, D2R, da * (nr ** 2 / area)

    partial = _dist.local_counts()
    with _lib.upload_cache(), partial:
        D1D1, D1D2, D2D2 = _driver.data_counts(count, sample1, sample2, same, do_auto, do_cross, None, None)
        D1R, D2R, RR = _driver.random_counts(count, analytic, sample1, sample2, randoms, same, do_RR, do_DR,
                                             None, None, None)
    return _driver.combine(same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator)


def _angular_tpcf_process_args(sample1, theta_bins, sample2, randoms, do_auto, do_cross, estimator, num_threads):
    """Validation in the reference's order with the reference's messages (:336-392)."""
    sample1 = np.atleast_1d(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross, ndim=2)
    if randoms is not None:
        randoms = np.atleast_1d(randoms)

    theta_bins = np.atleast_1d(theta_bins)
    theta_max = np.max(theta_bins)
    try:
        assert theta_bins.ndim == 1
        assert len(theta_bins) > 1
        if len(theta_bins) > 2:
            assert array_is_monotonic(theta_bins, strict=True) == 1
    except AssertionError:
        raise HalotoolsError("\n Input `theta_bins` must be a monotonically increasing 1-D \n"
                             "array with at least two entries.")
    if theta_max >= 180.0:
        raise HalotoolsError("\n The maximum length over which you search for pairs of points \n"
                             "cannot be larger than 180.0 deg. \n")
    if (type(do_auto) is not bool) | (type(do_cross) is not bool):
        raise HalotoolsError("\n `do_auto` and `do_cross` keywords must be of type boolean.")
    num_threads = get_num_threads(num_threads)
    verify_tpcf_estimator(estimator)
    return sample1, theta_bins, sample2, randoms, do_auto, do_cross, num_threads, same
