"""Drop-in for ``halotools.mock_observables.marked_tpcf``
(/root/reference/halotools/mock_observables/two_point_clustering/marked_tpcf.py:28-597)."""
import numpy as np

from ..custom_exceptions import HalotoolsError
from ..helpers import (enforce_sample_has_correct_shape, get_num_threads, get_period,
                       get_separation_bins_array)
from ..pair_counters import marked_npairs_3d, npairs_3d
from ..pair_counters.mesh_helpers import _enforce_maximum_search_length
from .clustering_helpers import process_optional_input_sample2

__all__ = ['marked_tpcf']


class _SeededNumpyRNG(object):
    """Seed numpy's global RNG inside the block and restore its state afterwards (the behaviour of
    astropy.utils.misc.NumpyRNGContext the reference relies on, marked_tpcf.py:339-353)."""

    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        self._state = np.random.get_state()
        np.random.seed(self.seed)

    def __exit__(self, exc_type, exc_value, traceback):
        np.random.set_state(self._state)


def marked_tpcf(sample1, rbins, sample2=None,
                marks1=None, marks2=None, period=None, do_auto=True, do_cross=True,
                num_threads=1, weight_func_id=1,
                normalize_by='random_marks', iterations=1, randomize_marks=None, seed=None):
    """Marked correlation function M(r) = WW(r) / RR(r): weighted pair sums normalised either by
    the same sums over shuffled marks ('random_marks') or by plain pair counts ('number_counts').
    Arguments, return structure and errors are the reference's."""
    (sample1, rbins, sample2, marks1, marks2, period, do_auto, do_cross, num_threads,
     weight_func_id, normalize_by, same, PBCs, randomize_marks) = _marked_tpcf_process_args(
        sample1, rbins, sample2, marks1, marks2, period, do_auto, do_cross, num_threads,
        weight_func_id, normalize_by, iterations, randomize_marks, seed)

    def wcount(a, b, wa, wb):
        return np.diff(marked_npairs_3d(a, b, rbins, weights1=wa, weights2=wb,
                                        weight_func_id=weight_func_id, period=period,
                                        num_threads=num_threads))

    def triple(count, w1a, w1b, w2a, w2b, x1, x2):
        """(11, 12, 22) counts with the reference's do_auto / do_cross / same-sample rules."""
        c11 = count(sample1, sample1, w1a, w1b) if do_auto else None
        if same:
            return c11, c11, c11
        c12 = count(sample1, sample2, x1, x2) if do_cross else None
        c22 = count(sample2, sample2, w2a, w2b) if do_auto else None
        return c11, c12, c22

    W1W1, W1W2, W2W2 = triple(wcount, marks1, marks1, marks2, marks2, marks1, marks2)

    if normalize_by == 'number_counts':
        def ncount(a, b, wa, wb):
            return np.diff(npairs_3d(a, b, rbins, period=period, num_threads=num_threads))
        R1R1, R1R2, R2R2 = triple(ncount, None, None, None, None, None, None)
    else:
        def shuffled_counts():
            with _SeededNumpyRNG(seed):
                permutate1 = np.random.permutation(np.arange(0, len(sample1)))
                permutate2 = np.random.permutation(np.arange(0, len(sample2)))
            # the reference shuffles the mark arrays IN PLACE through an alias
            # (marked_tpcf.py:418-425), so both weight arguments of each count are the shuffled marks
            for i in range(marks1.shape[1]):
                if randomize_marks[i]:
                    marks1[:, i] = marks1[permutate1, i]
            for i in range(marks2.shape[1]):
                if randomize_marks[i]:
                    marks2[:, i] = marks2[permutate2, i]
            return triple(wcount, marks1, marks1, marks2, marks2, marks1, marks2)

        if iterations > 1:
            nb = len(rbins) - 1
            R1R1, R1R2, R2R2 = (np.zeros((iterations, nb)), np.zeros((iterations, nb)),
                                np.zeros((iterations, nb)))
            for i in range(iterations):
                R1R1[i, :], R1R2[i, :], R2R2[i, :] = shuffled_counts()
            R1R1, R1R2, R2R2 = (np.median(R1R1, axis=0), np.median(R1R2, axis=0),
                                np.median(R2R2, axis=0))
        else:
            R1R1, R1R2, R2R2 = shuffled_counts()

    with np.errstate(divide="ignore", invalid="ignore"):
        if same:
            return W1W1 / R1R1
        if (do_auto is True) & (do_cross is True):
            return W1W1 / R1R1, W1W2 / R1R2, W2W2 / R2R2
        elif do_cross is True:
            return W1W2 / R1R2
        elif do_auto is True:
            return W1W1 / R1R1, W2W2 / R2R2


def _as_marks(marks, n, which):
    marks = np.atleast_1d(marks).astype(float) if marks is not None else np.ones(n).astype(float)
    if marks.ndim == 1:
        marks = marks.reshape((len(marks), 1))
    elif marks.ndim != 2:
        msg = ("\n You must either pass in a 1-D or 2-D array \n"
               "for the input `marks%i`. \n"
               "The `pair_counters._wnpairs_process_weights` function received \n"
               "a `marks%i` array of dimension %i")
        raise HalotoolsError(msg % (which, which, marks.ndim))
    return marks


def _marked_tpcf_process_args(sample1, rbins, sample2, marks1, marks2,
                              period, do_auto, do_cross, num_threads,
                              wfunc, normalize_by, iterations, randomize_marks, seed):
    """Validation in the reference's order with its messages (marked_tpcf.py:497-597)."""
    sample1 = enforce_sample_has_correct_shape(sample1)
    sample2, same, do_cross = process_optional_input_sample2(sample1, sample2, do_cross)

    try:
        int(wfunc) == wfunc
    except Exception:
        raise ValueError("\n `wfunc` parameter must be an integer ID of the desired function.")
    if normalize_by not in ['random_marks', 'number_counts']:
        raise ValueError("\n `normalize_by` parameter not recognized.")

    marks1 = _as_marks(marks1, len(sample1), 1)
    marks2 = _as_marks(marks2, len(sample2), 2)
    if len(marks1) != len(sample1):
        raise HalotoolsError("\n `marks1` must have same length as `sample1`.")
    if len(marks2) != len(sample2):
        raise HalotoolsError("\n `marks2` must have same length as `sample2`.")

    if randomize_marks is not None:
        randomize_marks = np.atleast_1d(randomize_marks)
    else:
        randomize_marks = np.array([True]*marks1.shape[1])
    if randomize_marks.ndim == 1:
        if len(randomize_marks) != marks1.shape[1]:
            raise HalotoolsError("\n `randomize_marks` must have same length \n"
                                 " as the number of weights per point.")
    else:
        raise HalotoolsError("\n `randomize_marks` must be one dimensional.")

    rbins = get_separation_bins_array(rbins)
    rmax = np.amax(rbins)
    period, PBCs = get_period(period)
    _enforce_maximum_search_length(rmax, period)

    try:
        assert do_auto == bool(do_auto)
        assert do_cross == bool(do_cross)
    except Exception:
        raise ValueError("`do_auto` and `do_cross` keywords must be boolean-valued.")

    num_threads = get_num_threads(num_threads)
    return (sample1, rbins, sample2, marks1, marks2, period, do_auto, do_cross,
            num_threads, wfunc, normalize_by, same, PBCs, randomize_marks)
