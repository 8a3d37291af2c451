"""Drop-in for ``halotools.mock_observables.wp``
(/root/reference/halotools/mock_observables/two_point_clustering/wp.py:20-240)."""
import numpy as np

import importlib

from . import _device
from .rp_pi_tpcf import _rp_pi_tpcf_process_args, rp_pi_tpcf

__all__ = ['wp']

# the MODULE (the package re-exports the function under the same name)
_rp_pi = importlib.import_module(__package__ + ".rp_pi_tpcf")


def wp(sample1, rp_bins, pi_max, sample2=None, randoms=None, period=None,
       do_auto=True, do_cross=True, estimator='Natural', num_threads=1,
       approx_cell1_size=None, approx_cell2_size=None, approx_cellran_size=None, seed=None):
    """Projected correlation function wp(rp) = 2 * pi_max * xi(rp, [0, pi_max]) (one line-of-sight
    bin, wp.py:196,219-240); same arguments and return structure as the reference."""
    pi_max = float(pi_max)
    pi_bins = np.array([0.0, pi_max])
    (sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, num_threads,
     same, PBCs) = _rp_pi_tpcf_process_args(
        sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, estimator, num_threads,
        approx_cell1_size, approx_cell2_size, approx_cellran_size, seed)
    if same:
        sample2 = None

    if _device.available(_rp_pi.npairs_xy_z):
        # the pi integration 2 * xi[:, 0] * pi_max (wp.py:219-221) is applied by the estimator kernel
        return _rp_pi._rp_pi_tpcf(sample1, rp_bins, pi_bins, sample2, randoms, period, do_auto, do_cross, estimator,
                                  num_threads, approx_cell1_size, approx_cell2_size, approx_cellran_size, None, pi_max)

    result = rp_pi_tpcf(sample1, rp_bins=rp_bins, pi_bins=pi_bins, sample2=sample2, randoms=randoms,
                        period=period, do_auto=do_auto, do_cross=do_cross, estimator=estimator,
                        num_threads=num_threads, approx_cell1_size=approx_cell1_size,
                        approx_cell2_size=approx_cell2_size, approx_cellran_size=approx_cellran_size)

    def integrate(xi):
        return 2.0 * xi[:, 0] * pi_max

    if same or ((do_auto is False) & (do_cross is True)):
        return integrate(result)
    if (do_auto is True) & (do_cross is True):
        return integrate(result[0]), integrate(result[1]), integrate(result[2])
    if (do_auto is True) & (do_cross is False):
        return integrate(result[0]), integrate(result[1])
