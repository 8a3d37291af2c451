"""Argument helpers of the clustering functions
(/root/reference/halotools/mock_observables/two_point_clustering/clustering_helpers.py:18-95)."""
from warnings import warn

import numpy as np

from ..helpers import enforce_sample_has_correct_shape
from .tpcf_estimators import tpcf_estimator_dd_dr_rr_requirements

__all__ = ('verify_tpcf_estimator', 'process_optional_input_sample2', 'tpcf_estimator_dd_dr_rr_requirements')


def verify_tpcf_estimator(estimator):
    available_estimators = list(tpcf_estimator_dd_dr_rr_requirements.keys())
    if estimator in available_estimators:
        return estimator
    msg = (u"Your estimator ``{0}`` \n"
           "is not in the list of available estimators:\n {1}".format(estimator, available_estimators))
    raise ValueError(msg)


def process_optional_input_sample2(sample1, sample2, do_cross, ndim=3):
    """sample2=None (or equal to sample1) means auto-correlation only; returns
    (sample2, _sample1_is_sample2, do_cross)."""
    if sample2 is None:
        return sample1, True, do_cross
    sample2 = enforce_sample_has_correct_shape(sample2, ndim=ndim)
    same = (sample1.shape == sample2.shape) and bool(np.all(sample1 == sample2))
    if same and do_cross:
        warn(u"\n `sample1` and `sample2` are exactly the same, \n"
             "only the auto-correlation will be returned.\n")
        do_cross = False
    return sample2, same, do_cross
