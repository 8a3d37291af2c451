"""Synthetic inputs of the BASELINE.json configurations (there is no network for real catalogs).

``fakesim_zheng07_mock`` is a numpy restatement of "zheng07 HOD populated on FakeSim-style halos"
(config 2): halo recipe after /root/reference/halotools/sim_manager/fake_sim.py:63,110-138
(Lbox 250, 10 log-spaced masses 1e10..1e16, uniform positions, rvir = 10^((log10 M - 15)/3),
conc ~ U(4, 15)); occupation after
/root/reference/halotools/empirical_models/occupation_models/zheng07_components.py:190-201,514-516
(threshold -20: logMmin 12.02, sigma_logM 0.26, logM0 11.38, logM1 13.31, alpha 1.06): centrals
Bernoulli(0.5 (1 + erf((logM - logMmin)/sigma))) at the halo centre, satellites
Poisson(((M - M0)/M1)^alpha) drawn from an isotropic NFW profile truncated at rvir, wrapped into
the box.  It only has to be the same SHAPE of input (extremely clustered, ~5e5 points); it is
not used for parity with the reference's own populate_mock().
"""
from math import erf

import numpy as np

__all__ = ("uniform_points", "fakesim_zheng07_mock", "config_rbins")


def uniform_points(seed, n, Lbox, dim=3):
    return np.random.RandomState(seed).uniform(0, Lbox, (int(n), dim))


def config_rbins():
    """15 log bins 0.1 - 20 Mpc/h (configs 1, 2, 4)."""
    return np.logspace(-1, np.log10(20), 15)


def _nfw_radii(rng, conc, n):
    """Inverse-transform sample of r/rvir for an NFW profile truncated at rvir."""
    def m(x):
        return np.log(1.0 + x) - x / (1.0 + x)
    u = rng.uniform(0, 1, n) * m(conc)
    # Newton iterations on m(x) = u, x in (0, conc)
    x = conc * rng.uniform(0.1, 0.9, n)
    for _ in range(40):
        f = m(x) - u
        fp = x / (1.0 + x) ** 2
        x = np.clip(x - f / np.maximum(fp, 1e-12), 1e-8, conc)
    return x / conc


def fakesim_zheng07_mock(num_halos_per_massbin=560, Lbox=250.0, seed=43):
    rng = np.random.RandomState(seed)
    massbins = np.logspace(10, 16, 10)
    mvir = np.repeat(massbins, num_halos_per_massbin)
    nh = len(mvir)
    rvir = 10.0 ** ((np.log10(mvir) - 15.0) / 3.0)
    conc = rng.uniform(4, 15, nh)
    pos = rng.uniform(0, Lbox, (nh, 3))
    logMmin, sigma, logM0, logM1, alpha = 12.02, 0.26, 11.38, 13.31, 1.06
    logm = np.log10(mvir)
    pcen = np.array([0.5 * (1.0 + erf((lm - logMmin) / sigma)) for lm in logm])
    has_cen = rng.uniform(0, 1, nh) < pcen
    M0, M1 = 10.0 ** logM0, 10.0 ** logM1
    mean_sat = np.where(mvir > M0, ((np.maximum(mvir - M0, 0.0)) / M1) ** alpha, 0.0)
    nsat = rng.poisson(mean_sat)
    host = np.repeat(np.arange(nh), nsat)
    ntot = len(host)
    r = _nfw_radii(rng, conc[host], ntot) * rvir[host]
    cost = rng.uniform(-1, 1, ntot)
    phi = rng.uniform(0, 2 * np.pi, ntot)
    sint = np.sqrt(1.0 - cost ** 2)
    sat = pos[host] + np.vstack([r * sint * np.cos(phi), r * sint * np.sin(phi), r * cost]).T
    gal = np.vstack([pos[has_cen], sat])
    gal = np.mod(gal, Lbox)
    gal[gal >= Lbox] = 0.0
    return np.ascontiguousarray(gal)
