"""CPU tests (no GPU) of the HOST layer of the statistics: drivers, estimators, analytic randoms, jackknife algebra,
mark shuffling.  The pair counters the statistics call are replaced by the CPU oracle (test infrastructure), so what is
compared with the golden vectors of the unmodified reference is everything above the engine boundary."""
import importlib

import numpy as np
import pytest

import halotools_b200 as hb


def _mod(name):
    # the packages re-export the functions under the module names: fetch the MODULES
    return importlib.import_module("halotools_b200." + name)


tpcf, rp_pi_tpcf, marked_tpcf = (_mod("two_point_clustering." + n) for n in ("tpcf", "rp_pi_tpcf", "marked_tpcf"))
tpcf_jackknife, s_mu_tpcf = _mod("two_point_clustering.tpcf_jackknife"), _mod("two_point_clustering.s_mu_tpcf")
tpcf_one_two_halo_decomp = _mod("two_point_clustering.tpcf_one_two_halo_decomp")
angular_tpcf = _mod("two_point_clustering.angular_tpcf")
wpo_module = _mod("surface_density.weighted_npairs_per_object_xy")
from oracle import oracle
from tests.golden import cases

STAT_FUNCS = ("tpcf", "wp", "rp_pi_tpcf", "marked_tpcf", "tpcf_jackknife", "wp_jackknife", "rp_pi_tpcf_jackknife",
              "s_mu_tpcf", "tpcf_one_two_halo_decomp", "angular_tpcf", "total_mass_enclosed_in_stack_of_cylinders",
              "surface_density_in_annulus", "surface_density_in_cylinder", "total_mass_enclosed_per_cylinder",
              "return_xyz_formatted_array", "apply_zspace_distortion")
STAT_CASES = [n for n in cases.names() if cases._cases()[n][0] in STAT_FUNCS]


def _drop(fn, *names):
    def call(*args, **kwargs):
        for n in names:
            kwargs.pop(n, None)
        return fn(*args, **kwargs)
    return call


@pytest.fixture
def oracle_counters(monkeypatch):
    """every pair counter the statistic modules imported, served by the CPU oracle"""
    n3d = oracle.npairs_3d
    nxyz = oracle.npairs_xy_z
    nsmu = oracle.npairs_s_mu
    m3d = oracle.marked_npairs_3d
    jk3d = _drop(oracle.npairs_jackknife_3d, "num_threads")
    jkxyz = _drop(oracle.npairs_jackknife_xy_z, "num_threads")
    for mod, names in ((tpcf, {"npairs_3d": n3d}), (rp_pi_tpcf, {"npairs_xy_z": nxyz}),
                       (marked_tpcf, {"npairs_3d": n3d, "marked_npairs_3d": m3d}),
                       (tpcf_jackknife, {"npairs_jackknife_3d": jk3d, "npairs_jackknife_xy_z": jkxyz}),
                       (s_mu_tpcf, {"npairs_s_mu": nsmu}),
                       (tpcf_one_two_halo_decomp, {"npairs_3d": n3d, "marked_npairs_3d": m3d}),
                       (angular_tpcf, {"npairs_3d": n3d}),
                       (wpo_module, {"weighted_npairs_xy": _drop(oracle.weighted_npairs_xy, "num_threads"),
                                     "weighted_npairs_per_object_xy": _drop(oracle.weighted_npairs_per_object_xy, "num_threads")})):
        for name, fn in names.items():
            assert hasattr(mod, name), (mod.__name__, name)
            monkeypatch.setattr(mod, name, fn)


@pytest.mark.parametrize("name", STAT_CASES)
def test_statistic_host_logic_matches_reference_golden(name, golden, oracle_counters):
    fn, args, kwargs = cases.get(name)
    got = cases.flatten(getattr(hb, fn)(*args, **kwargs))
    want = golden(name)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape
        scale = np.max(np.abs(w[np.isfinite(w)])) if np.any(np.isfinite(w)) else 1.0
        assert np.allclose(g, w, rtol=1e-8, atol=1e-10 * scale, equal_nan=True), (name, np.nanmax(np.abs(g - w)))
