"""Parity gates for float outputs, shared by the GPU tests.

delta-sigma is a cancelling difference, element (i, k) = (2 dlog_k M_i(<rp_k) - sum_inbin m (1 - ln(rp_{k+1}^2/d^2))) / norm:
the reference against itself under a re-ordering of the particles moves it by 1.8e-12 RELATIVE but only 5e-16 of
A_ik = the sum of the absolute values of the accumulated terms.  Gate: |gpu - ref| <= 1e-12 |ref| + 1e-12 A_ik
(per object), and the same with the column means of A for per_object=False.  A comes from the C oracle
(``oracle.mean_delta_sigma(..., return_abs=True)``).  With HTB_RECORD_ERRORS=<file> every call appends the achieved
errors (in units of the gate, relative, and relative to A) as one JSON line.
"""
import json
import os

import numpy as np


def delta_sigma_gate(got, want, A, label=""):
    got, want, A = np.asarray(got), np.asarray(want), np.asarray(A)
    assert got.shape == want.shape == A.shape, (got.shape, want.shape, A.shape)
    err = np.abs(got - want)
    gate = 1e-12 * np.abs(want) + 1e-12 * A
    with np.errstate(divide="ignore", invalid="ignore"):
        units = np.where(err > 0, err / gate, 0.0)
        rel = np.where(err > 0, err / np.abs(want), 0.0)
        rel_a = np.where(err > 0, err / A, 0.0)
    path = os.environ.get("HTB_RECORD_ERRORS")
    if path:
        with open(path, "a") as fh:
            fh.write(json.dumps({"label": label, "shape": list(got.shape), "max_err_over_gate": float(np.max(units, initial=0.0)),
                                 "max_rel": float(np.max(rel, initial=0.0)), "max_err_over_A": float(np.max(rel_a, initial=0.0))}) + "\n")
    assert np.all(err <= gate), "delta-sigma outside 1e-12 |ref| + 1e-12 A: %.3g gate units (%s)" % (np.max(units), label)


def oracle_delta_sigma(oracle, args, kwargs, num_threads=4):
    """(want, A) of a mean_delta_sigma call from the C oracle; the non-periodic path shifts its inputs in place, so
    it gets copies."""
    kw = dict(kwargs)
    kw.pop("num_threads", None)
    kw.pop("verbose", None)
    a = list(args)
    if kw.get("period") is None:
        a[0], a[1] = np.array(a[0], copy=True), np.array(a[1], copy=True)
    return oracle.mean_delta_sigma(*a, num_threads=num_threads, return_abs=True, **kw)
