"""Generates tests/golden/golden.npz by running the UNMODIFIED reference (astropy/halotools at
/root/reference, built into /tmp/htb_ref_build by oracle/build_ref.py) through its PUBLIC functions
on the seeded inputs of tests/golden/cases.py.  Run only in the build container:

    python oracle/build_ref.py && python tests/golden/make_golden.py

The GPU box never runs this (it has no /root/reference); it only reads golden.npz.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/tmp/htb_ref_build/src")

from tests.golden import cases  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    import halotools.mock_observables as mo
    assert "/tmp/htb_ref_build" in mo.__file__, mo.__file__
    out = {}
    for name in cases.names():
        fn, args, kwargs = cases.get(name)
        res = cases.flatten(getattr(mo, fn)(*args, **kwargs))
        out[name + "/n"] = np.array(len(res))
        for i, r in enumerate(res):
            out["%s/%d" % (name, i)] = r
        print("%-32s %s" % (name, " ".join(str(r.shape) for r in res)))
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), len(cases.names()), "cases")


if __name__ == "__main__":
    main()
