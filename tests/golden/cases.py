"""Seeded input recipes shared by tests/golden/make_golden.py (which runs the UNMODIFIED reference
on them in the build container) and by the parity tests (which run the oracle and the CUDA path on
the same inputs).  Nothing here touches /root/reference.

Each case is ``name -> (function name, args tuple, kwargs dict)``; arrays are regenerated from
fixed ``np.random.RandomState`` seeds, so only the reference OUTPUTS are stored in golden.npz.
Fixture shapes follow the reference's own tests
(/root/reference/halotools/mock_observables/pair_counters/test_pair_counters/test_npairs_3d.py:33-267,
test_npairs_xy_z.py, test_npairs_s_mu.py, test_marked_npairs_3d.py, test_non_cubic_volumes.py,
surface_density/tests/test_mean_delta_sigma.py, two_point_clustering/tests/*).
"""
import numpy as np

NUM_WEIGHTS = {1: 1, 2: 1, 3: 2, 4: 2, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2, 10: 2, 11: 2,
               12: 4, 13: 4, 14: 3, 15: 3, 16: 5, 17: 5}


class FlatLCDM(object):
    """minimal cosmology object for the redshift-space-distortion cases (only ``efunc`` is used)"""

    def efunc(self, z):
        return np.sqrt(0.3 * (1.0 + z) ** 3 + 0.7)


def pts(seed, n, L=1.0, dim=3):
    return np.random.RandomState(seed).uniform(0, L, (n, dim))


def locus(seed, n, center, eps=0.001):
    """n points within +-eps of ``center`` (cf. cf_helpers.generate_locus_of_3d_points)."""
    rng = np.random.RandomState(seed)
    return np.asarray(center)[None, :] + rng.uniform(-eps, eps, (n, 3))


def grid(n_per_dim, L):
    """Regular grid with one point per cell centre offset (cf. cf_helpers.generate_3d_regular_mesh)."""
    edges = np.linspace(0, L, n_per_dim + 1)
    c = (edges[:-1] + edges[1:]) / 2.0
    x, y, z = np.meshgrid(c, c, c, indexing="ij")
    return np.vstack([x.ravel(), y.ravel(), z.ravel()]).T


def clustered(seed, nhalo, nper, L, sigma):
    """Clustered sample: Gaussian blobs around random centres, wrapped into the box."""
    rng = np.random.RandomState(seed)
    cen = rng.uniform(0, L, (nhalo, 3))
    p = cen[rng.randint(0, nhalo, nhalo * nper)] + rng.normal(0, sigma, (nhalo * nper, 3))
    return np.mod(p, L)


def weights(seed, n, wid):
    rng = np.random.RandomState(seed)
    nw = NUM_WEIGHTS[wid]
    w = rng.uniform(0.5, 1.5, (n, nw))
    if wid in (3, 4, 16, 17):
        # equality-type functions need repeated labels to be non-trivial
        col = 0 if wid in (3, 4) else 4
        w[:, col] = rng.randint(0, 3, n).astype(float)
    return w


def _cases():
    C = {}
    rb = np.array([0.001, 0.1, 0.2, 0.3])
    s1, s2 = pts(43, 1000), pts(44, 1000)
    C["n3d_periodic"] = ("npairs_3d", (s1, s2, rb), dict(period=1.0))
    C["n3d_nonperiodic"] = ("npairs_3d", (s1, s2, rb), dict(period=None))
    C["n3d_auto"] = ("npairs_3d", (s1, s1, rb), dict(period=1.0))
    C["n3d_rbins_from_zero"] = ("npairs_3d", (s1, s1, np.array([0.0, 0.05, 0.25])), dict(period=1.0))
    C["n3d_two_bins"] = ("npairs_3d", (s1, s2, np.array([0.1, 0.3])), dict(period=1.0))
    C["n3d_linear_30bins"] = ("npairs_3d", (s1, s2, np.linspace(0.01, 0.3, 30)), dict(period=1.0))
    scale = np.array([1.0, 2.0, 3.0])
    C["n3d_noncubic"] = ("npairs_3d", (s1 * scale, s2 * scale, rb), dict(period=[1.0, 2.0, 3.0]))
    C["n3d_cellsizes_a"] = ("npairs_3d", (s1, s2, rb),
                            dict(period=1.0, approx_cell1_size=[0.2, 0.2, 0.2], approx_cell2_size=[0.15, 0.15, 0.15]))
    C["n3d_cellsizes_b"] = ("npairs_3d", (s1, s2, rb),
                            dict(period=1.0, approx_cell1_size=0.1, approx_cell2_size=[0.05, 0.1, 0.3]))
    C["n3d_cellsizes_c"] = ("npairs_3d", (s1, s2, np.array([0.01, 0.05, 0.1])),
                            dict(period=1.0, approx_cell1_size=0.02, approx_cell2_size=0.01))
    C["n3d_search_third"] = ("npairs_3d", (s1, s2, np.array([0.1, 1.0 / 3.0])), dict(period=1.0))
    l1, l2 = locus(43, 100, (0.1, 0.1, 0.1)), locus(44, 100, (0.1, 0.1, 0.2))
    C["n3d_locus"] = ("npairs_3d", (l1, l2, np.array([0.001, 0.05, 0.15, 0.3])), dict(period=1.0))
    l3, l4 = locus(43, 100, (0.5, 0.5, 0.05)), locus(44, 100, (0.5, 0.5, 0.95))
    C["n3d_locus_wrap"] = ("npairs_3d", (l3, l4, np.array([0.001, 0.05, 0.15, 0.3])), dict(period=1.0))
    C["n3d_locus_nowrap"] = ("npairs_3d", (l3, l4, np.array([0.001, 0.05, 0.15, 0.3])), dict(period=None))
    g10 = grid(10, 1.0)
    C["n3d_grid"] = ("npairs_3d", (g10, g10, np.array([0.001, 0.101, 0.1 * np.sqrt(2) + 0.001, 0.1 * np.sqrt(3) + 0.001])),
                     dict(period=1.0))
    # points sitting EXACTLY on cell boundaries (numpy floor-division quirk)
    k = np.arange(0, 13)
    b = np.array([(250.0 / 12) * i for i in k])
    onb = np.vstack([np.repeat(b, 13), np.tile(b, 13), np.full(169, 125.0)]).T
    onb = np.clip(onb, 0, 250.0)
    C["n3d_on_boundaries"] = ("npairs_3d", (onb, pts(45, 2000, 250.0), np.logspace(-1, np.log10(20), 15)),
                              dict(period=250.0))
    s5 = pts(43, 5000, 250.0)
    C["n3d_c1_small"] = ("npairs_3d", (s5, s5, np.logspace(-1, np.log10(20), 15)), dict(period=250.0))
    cl = clustered(46, 40, 100, 250.0, 1.5)
    C["n3d_clustered"] = ("npairs_3d", (cl, cl, np.logspace(-1, np.log10(20), 15)), dict(period=250.0))
    C["n3d_clustered_x_random"] = ("npairs_3d", (cl, pts(47, 6000, 250.0), np.logspace(-1, np.log10(20), 15)),
                                   dict(period=250.0))
    C["n3d_empty_cells"] = ("npairs_3d", (locus(43, 50, (0.3, 0.3, 0.3), 0.05), locus(44, 70, (0.32, 0.3, 0.31), 0.05),
                                          np.array([0.01, 0.03, 0.06])), dict(period=1.0))
    C["n3d_c1_full"] = ("npairs_3d", (("pts", 43, 100000, 250.0), "same", np.logspace(-1, np.log10(20), 15)),
                        dict(period=250.0))

    rp, pi = np.array([0.001, 0.1, 0.2, 0.3]), np.array([0.001, 0.1, 0.2, 0.3])
    C["xyz_periodic"] = ("npairs_xy_z", (s1, s2, rp, pi), dict(period=1.0))
    C["xyz_nonperiodic"] = ("npairs_xy_z", (s1, s2, rp, pi), dict(period=None))
    C["xyz_pi_from_zero"] = ("npairs_xy_z", (s1, s1, rp, np.array([0.0, 0.15])), dict(period=1.0))
    C["xyz_noncubic"] = ("npairs_xy_z", (s1 * scale, s2 * scale, rp, pi), dict(period=[1.0, 2.0, 3.0]))
    C["xyz_wp_like"] = ("npairs_xy_z", (pts(43, 4000, 1000.0), "same", np.logspace(-1, np.log10(30), 15),
                                        np.array([0.0, 60.0])), dict(period=1000.0))
    C["xyz_cellsizes"] = ("npairs_xy_z", (s1, s2, rp, pi),
                          dict(period=1.0, approx_cell1_size=[0.2, 0.2, 0.2], approx_cell2_size=[0.15, 0.15, 0.15]))

    sb, mb = np.linspace(0.01, 0.3, 8), np.linspace(0, 1.0, 7)
    C["smu_periodic"] = ("npairs_s_mu", (s1, s2, sb, mb), dict(period=1.0))
    C["smu_nonperiodic"] = ("npairs_s_mu", (s1, s2, sb, mb), dict(period=None))
    C["smu_auto"] = ("npairs_s_mu", (s1, s1, np.array([0.0, 0.1, 0.2]), np.linspace(0, 1.0, 4)), dict(period=1.0))

    for wid in range(1, 18):
        C["marked_id%02d" % wid] = ("marked_npairs_3d", (s1, s2, rb, wid),
                                    dict(period=1.0, weights1=weights(50 + wid, 1000, wid),
                                         weights2=weights(80 + wid, 1000, wid)))
    C["marked_nonperiodic"] = ("marked_npairs_3d", (s1, s2, rb, 1),
                               dict(period=None, weights1=weights(51, 1000, 1), weights2=weights(81, 1000, 1)))
    wg = np.random.RandomState(7).randint(1, 4, (1000, 1)).astype(float)
    C["marked_grid_integer"] = ("marked_npairs_3d", (g10, g10, np.array([0.001, 0.101, 0.15, 0.18]), 1),
                                dict(period=1.0, weights1=wg, weights2=wg))
    C["marked_default_weights"] = ("marked_npairs_3d", (s1, s2, rb, 1), dict(period=1.0))
    C["marked_logbins"] = ("marked_npairs_3d", (s5, "same", np.logspace(-1, np.log10(20), 15), 1),
                           dict(period=250.0, weights1=weights(90, 5000, 1), weights2="same"))

    gal, ptcl = pts(43, 300, 1.0), pts(44, 20000, 1.0)
    rpb = np.logspace(np.log10(0.02), np.log10(0.25), 8)
    C["ds_periodic_per_object"] = ("mean_delta_sigma", (gal, ptcl, 1.0, rpb), dict(period=1.0, per_object=True))
    C["ds_periodic_mean"] = ("mean_delta_sigma", (gal, ptcl, 1.0, rpb), dict(period=1.0))
    C["ds_masses"] = ("mean_delta_sigma", (gal, ptcl, np.random.RandomState(48).uniform(0.5, 2.0, 20000), rpb),
                      dict(period=1.0, per_object=True))
    C["ds_nonperiodic"] = ("mean_delta_sigma", (gal, ptcl, 2.5, np.logspace(np.log10(0.02), np.log10(0.2), 6)),
                           dict(period=None, per_object=True))
    C["ds_cellsizes"] = ("mean_delta_sigma", (gal, ptcl, 1.0, rpb),
                         dict(period=1.0, per_object=True, approx_cell1_size=0.1, approx_cell2_size=0.05))

    rb2 = np.logspace(-2, -0.7, 8)
    ran = pts(49, 3000)
    C["tpcf_natural_analytic"] = ("tpcf", (s1, rb2), dict(period=1.0))
    for est in ("Natural", "Davis-Peebles", "Hewett", "Hamilton", "Landy-Szalay"):
        C["tpcf_randoms_" + est] = ("tpcf", (s1, rb2), dict(randoms=ran, period=1.0, estimator=est))
    C["tpcf_cross_ls"] = ("tpcf", (s1, rb2), dict(sample2=s2, randoms=ran, period=1.0, estimator="Landy-Szalay"))
    C["tpcf_cross_only"] = ("tpcf", (s1, rb2), dict(sample2=s2, period=1.0, do_auto=False))
    C["tpcf_auto_only"] = ("tpcf", (s1, rb2), dict(sample2=s2, period=1.0, do_cross=False))
    C["tpcf_nonperiodic"] = ("tpcf", (s1, rb2), dict(randoms=ran, period=None, estimator="Landy-Szalay"))
    rpw = np.logspace(-2, -0.8, 7)
    C["wp_auto"] = ("wp", (s1, rpw, 0.2), dict(period=1.0))
    C["wp_cross_randoms"] = ("wp", (s1, rpw, 0.2), dict(sample2=s2, randoms=ran, period=1.0, estimator="Landy-Szalay"))
    C["rp_pi_auto"] = ("rp_pi_tpcf", (s1, rpw, np.linspace(0, 0.25, 5)), dict(period=1.0))
    C["rp_pi_cross"] = ("rp_pi_tpcf", (s1, rpw, np.linspace(0, 0.25, 5)), dict(sample2=s2, period=1.0))
    m1 = np.random.RandomState(60).uniform(0.5, 1.5, 1000)
    m2 = np.random.RandomState(61).uniform(0.5, 1.5, 1000)
    C["marked_tpcf_random_marks"] = ("marked_tpcf", (s1, rb2), dict(marks1=m1, period=1.0, seed=43))
    C["marked_tpcf_number_counts"] = ("marked_tpcf", (s1, rb2),
                                      dict(marks1=m1, period=1.0, normalize_by="number_counts"))
    C["marked_tpcf_cross"] = ("marked_tpcf", (s1, rb2),
                              dict(sample2=s2, marks1=m1, marks2=m2, period=1.0, seed=43, iterations=2))
    # ---- npairs_projected, npairs_per_object_3d, marked_npairs_xy_z, weighted_npairs_xy
    # (fixture shapes: test_pair_counters/test_npairs_projected.py, test_npairs_per_object_3d.py,
    # test_marked_npairs_xy_z.py, surface_density/tests/test_weighted_npairs_xy.py)
    C["proj_periodic"] = ("npairs_projected", (s1, s2, rp, 0.2), dict(period=1.0))
    C["proj_nonperiodic"] = ("npairs_projected", (s1, s2, rp, 0.2), dict(period=None))
    C["proj_auto_logbins"] = ("npairs_projected", (s5, "same", np.logspace(-1, np.log10(20), 15), 40.0),
                              dict(period=250.0))
    C["proj_noncubic"] = ("npairs_projected", (s1 * scale, s2 * scale, rp, 0.25), dict(period=[1.0, 2.0, 3.0]))
    C["proj_cellsizes"] = ("npairs_projected", (s1, s2, rp, 0.1),
                           dict(period=1.0, approx_cell1_size=[0.2, 0.2, 0.2], approx_cell2_size=[0.15, 0.15, 0.15]))
    C["perobj_periodic"] = ("npairs_per_object_3d", (s1, s2, rb), dict(period=1.0))
    C["perobj_nonperiodic"] = ("npairs_per_object_3d", (s1, s2, rb), dict(period=None))
    C["perobj_auto_logbins"] = ("npairs_per_object_3d", (s5, "same", np.logspace(-1, np.log10(20), 15)),
                                dict(period=250.0))
    C["perobj_clustered"] = ("npairs_per_object_3d", (cl, pts(47, 6000, 250.0), np.logspace(-1, np.log10(20), 15)),
                             dict(period=250.0))
    C["perobj_cellsizes"] = ("npairs_per_object_3d", (s1, s2, np.array([0.01, 0.05, 0.1])),
                             dict(period=1.0, approx_cell1_size=0.1, approx_cell2_size=[0.05, 0.1, 0.3]))
    for wid in range(1, 16):
        C["mxyz_id%02d" % wid] = ("marked_npairs_xy_z", (s1, s2, rp, pi),
                                  dict(period=1.0, weights1=weights(50 + wid, 1000, wid),
                                       weights2=weights(80 + wid, 1000, wid), weight_func_id=wid))
    C["mxyz_nonperiodic"] = ("marked_npairs_xy_z", (s1, s2, rp, pi),
                             dict(period=None, weights1=weights(51, 1000, 1), weights2=weights(81, 1000, 1),
                                  weight_func_id=1))
    C["mxyz_auto_many_pi"] = ("marked_npairs_xy_z", (s5, "same", np.logspace(-1, np.log10(20), 10),
                                                     np.linspace(0.0, 40.0, 21)),
                              dict(period=250.0, weights1=weights(90, 5000, 1), weights2="same", weight_func_id=1))
    C["mxyz_default_weights"] = ("marked_npairs_xy_z", (s1, s2, rp, pi), dict(period=1.0, weight_func_id=2))
    g2a, g2b = pts(43, 1500, 450.0, 2), pts(44, 20000, 450.0, 2)
    mass2 = np.random.RandomState(48).uniform(0.0, 1.0, 20000)
    C["wxy_periodic"] = ("weighted_npairs_xy", (g2a, g2b, mass2, np.logspace(-1, 1.5, 15)), dict(period=[450.0, 450.0]))
    C["wxy_nonperiodic"] = ("weighted_npairs_xy", (g2a, g2b, mass2, np.logspace(-1, 1.5, 15)), dict(period=None))
    C["wxy_scalar_period_cellsizes"] = ("weighted_npairs_xy", (g2a, g2b, mass2, np.logspace(-0.5, 1.3, 9)),
                                        dict(period=450.0, approx_cell1_size=30.0, approx_cell2_size=[15.0, 10.0]))
    # ---- jackknife pair counters (fixture shapes: test_pair_counters/test_npairs_jackknife_3d.py)
    def jt(seed, n, ns):
        return np.random.RandomState(seed).randint(1, ns + 1, n)
    w1j, w2j = np.random.RandomState(70).uniform(0.5, 1.5, 1000), np.random.RandomState(71).uniform(0.5, 1.5, 1000)
    C["jk3d_periodic"] = ("npairs_jackknife_3d", (s1, s2, rb, jt(72, 1000, 10), jt(73, 1000, 10), 10), dict(period=1.0))
    C["jk3d_weights"] = ("npairs_jackknife_3d", (s1, s2, rb, jt(72, 1000, 10), jt(73, 1000, 10), 10),
                         dict(period=1.0, weights1=w1j, weights2=w2j))
    C["jk3d_nonperiodic"] = ("npairs_jackknife_3d", (s1, s2, rb, jt(72, 1000, 7), jt(73, 1000, 7), 7),
                             dict(period=None, weights1=w1j, weights2=w2j))
    C["jk3d_auto_logbins"] = ("npairs_jackknife_3d", (s5, "same", np.logspace(-1, np.log10(20), 15),
                                                      jt(74, 5000, 27), jt(74, 5000, 27), 27), dict(period=250.0))
    C["jk3d_one_sample"] = ("npairs_jackknife_3d", (s1, s2, rb, np.ones(1000, dtype=int), np.ones(1000, dtype=int), 1),
                            dict(period=1.0))
    C["jkxyz_periodic"] = ("npairs_jackknife_xy_z", (s1, s2, rp, pi, jt(72, 1000, 10), jt(73, 1000, 10), 10),
                           dict(period=1.0, weights1=w1j, weights2=w2j))
    C["jkxyz_wp_like"] = ("npairs_jackknife_xy_z", (s5, "same", np.logspace(-1, np.log10(20), 12), np.array([0.0, 40.0]),
                                                    jt(74, 5000, 8), jt(74, 5000, 8), 8), dict(period=250.0))
    # jackknife statistics (fixture shapes: two_point_clustering/tests/test_tpcf_jackknife.py, test_wp_jackknife.py)
    ranj = pts(75, 4000)
    C["tpcf_jk_auto"] = ("tpcf_jackknife", (s1, ranj, rb2), dict(Nsub=3, period=1.0))
    C["tpcf_jk_cross_ls"] = ("tpcf_jackknife", (s1, ranj, rb2),
                             dict(Nsub=[2, 3, 2], sample2=s2, period=1.0, estimator="Landy-Szalay"))
    C["tpcf_jk_nonperiodic"] = ("tpcf_jackknife", (s1, ranj, rb2), dict(Nsub=2, period=None, estimator="Landy-Szalay"))
    C["tpcf_jk_randoms_by_number"] = ("tpcf_jackknife", (s1, [3000], rb2), dict(Nsub=2, period=1.0, seed=43))
    C["wp_jk_auto"] = ("wp_jackknife", (s1, ranj, rpw, 0.2), dict(Nsub=3, period=1.0))
    C["wp_jk_cross"] = ("wp_jackknife", (s1, ranj, rpw, 0.15),
                        dict(Nsub=2, sample2=s2, period=1.0, estimator="Landy-Szalay", do_auto=False))
    C["wpoxy_periodic"] = ("weighted_npairs_per_object_xy", (g2a, g2b, mass2, np.logspace(-1, 1.5, 15)),
                           dict(period=[450.0, 450.0]))
    C["wpoxy_nonperiodic"] = ("weighted_npairs_per_object_xy", (g2a, g2b, mass2, np.logspace(-0.5, 1.3, 9)), dict(period=None))
    cen3, ptc3 = pts(43, 800, 250.0), pts(44, 30000, 250.0)
    C["mass_per_cylinder"] = ("total_mass_enclosed_per_cylinder",
                              (cen3, ptc3, np.random.RandomState(48).uniform(0.5, 2.0, 30000), 2.5,
                               np.logspace(-1, 1.2, 10), 250.0), dict())
    C["mass_per_cylinder_scalar"] = ("total_mass_enclosed_per_cylinder",
                                     (cen3, ptc3, 3.0e9, 1.0, np.logspace(-1, 1.2, 10), [250.0, 250.0, 250.0]), dict())
    # xi(s, mu) (fixture shapes: two_point_clustering/tests/test_s_mu_tpcf.py)
    sbj, mbj = np.linspace(0.01, 0.25, 7), np.linspace(0.0, 1.0, 6)
    C["smu_tpcf_auto_analytic"] = ("s_mu_tpcf", (s1, sbj, mbj), dict(period=1.0))
    C["smu_tpcf_randoms_ls"] = ("s_mu_tpcf", (s1, sbj, mbj), dict(randoms=ran, period=1.0, estimator="Landy-Szalay"))
    C["smu_tpcf_cross"] = ("s_mu_tpcf", (s1, sbj, mbj), dict(sample2=s2, period=1.0))
    C["smu_tpcf_nonperiodic"] = ("s_mu_tpcf", (s1, sbj, mbj), dict(randoms=ran, period=None, estimator="Landy-Szalay"))
    # one- / two-halo decomposition (fixture shapes: two_point_clustering/tests/test_tpcf_one_two_halo.py)
    hid1 = np.random.RandomState(76).randint(0, 60, 1000)
    hid2 = np.random.RandomState(77).randint(0, 60, 1000)
    C["tpcf_12h_auto"] = ("tpcf_one_two_halo_decomp", (s1, hid1, rb2), dict(period=1.0))
    C["tpcf_12h_cross_ls"] = ("tpcf_one_two_halo_decomp", (s1, hid1, rb2),
                              dict(sample2=s2, sample2_host_halo_id=hid2, randoms=ran, period=1.0, estimator="Landy-Szalay"))
    C["tpcf_12h_cross_only"] = ("tpcf_one_two_halo_decomp", (s1, hid1, rb2),
                                dict(sample2=s2, sample2_host_halo_id=hid2, period=1.0, do_auto=False))
    # w(theta) (fixture shapes: two_point_clustering/tests/test_angular_tpcf.py)
    def sky(seed, n):
        r = np.random.RandomState(seed)
        return np.vstack([r.uniform(0, 360.0, n), np.degrees(np.arcsin(r.uniform(-1, 1, n)))]).T
    tb = np.logspace(-1, 1.2, 9)
    C["ang_auto_analytic"] = ("angular_tpcf", (sky(78, 2000), tb), dict())
    C["ang_cross_randoms_ls"] = ("angular_tpcf", (sky(78, 2000), tb),
                                 dict(sample2=sky(79, 1500), randoms=sky(80, 4000), estimator="Landy-Szalay"))
    mcyl = np.random.RandomState(48).uniform(0.5, 2.0, 30000)
    rpc = np.logspace(-1, 1.2, 10)
    C["mass_stack_of_cylinders"] = ("total_mass_enclosed_in_stack_of_cylinders", (cen3, ptc3, mcyl, 2.5, rpc, 250.0), dict())
    C["sigma_in_annulus"] = ("surface_density_in_annulus", (cen3, ptc3, mcyl, 1.0, rpc, 250.0), dict())
    C["sigma_in_cylinder"] = ("surface_density_in_cylinder", (cen3, ptc3, 2.0e9, 3.0, rpc, [250.0, 250.0, 250.0]), dict())
    # wide jackknife tables (more than 48 cells per point row) and the statistic on them
    rpj, pij = np.logspace(-2, -0.8, 9), np.linspace(0.0, 0.28, 8)
    C["jkxyz_wide"] = ("npairs_jackknife_xy_z", (s1, s2, rpj, pij, jt(72, 1000, 10), jt(73, 1000, 10), 10),
                       dict(period=1.0, weights1=w1j, weights2=w2j))
    C["rp_pi_jk_auto"] = ("rp_pi_tpcf_jackknife", (s1, ranj, rpj, pij), dict(Nsub=2, period=1.0))
    C["rp_pi_jk_cross_ls"] = ("rp_pi_tpcf_jackknife", (s1, ranj, rpw, np.linspace(0, 0.25, 4)),
                              dict(Nsub=[2, 2, 3], sample2=s2, period=1.0, estimator="Landy-Szalay"))
    # input formatting step: positions + redshift-space distortions
    rngx = np.random.RandomState(81)
    xx, yy, zz = rngx.uniform(-50, 300, (3, 1000))
    vv = rngx.normal(0, 300, 1000)
    C["xyz_plain"] = ("return_xyz_formatted_array", (xx, yy, zz), dict(period=250.0))
    C["xyz_rsd_z0"] = ("return_xyz_formatted_array", (xx, yy, zz),
                       dict(period=[250.0, 200.0, 100.0], velocity=vv, velocity_distortion_dimension="z", cosmology=FlatLCDM()))
    C["xyz_rsd_z07_mask"] = ("return_xyz_formatted_array", (xx, yy, zz),
                             dict(period=250.0, velocity=vv, velocity_distortion_dimension="x", redshift=0.7,
                                  cosmology=FlatLCDM(), mask=rngx.rand(1000) > 0.5))
    C["zspace_distortion"] = ("apply_zspace_distortion", (zz, vv, 0.5, FlatLCDM(), 250.0), dict())
    return C


_CACHE = {}


def names():
    if not _CACHE:
        _CACHE.update(_cases())
    return sorted(_CACHE.keys())


def _materialise(a, first=None):
    if isinstance(a, tuple) and len(a) == 4 and a[0] == "pts":
        return pts(a[1], a[2], a[3])
    if isinstance(a, str) and a == "same":
        return first
    return a


def get(name, copy=True):
    """(function name, args, kwargs) with fresh array copies (some reference functions mutate inputs)."""
    if not _CACHE:
        _CACHE.update(_cases())
    fn, args, kwargs = _CACHE[name]
    args = list(args)
    args[0] = _materialise(args[0])
    for i in range(1, len(args)):
        args[i] = _materialise(args[i], args[0])
    kwargs = dict(kwargs)
    for k in list(kwargs):
        if isinstance(kwargs[k], str) and kwargs[k] == "same":
            kwargs[k] = kwargs["weights1"]
    if copy:
        same01 = len(args) > 1 and args[1] is args[0]
        args = [np.array(a, copy=True) if isinstance(a, np.ndarray) else a for a in args]
        if same01:
            args[1] = args[0]
        wsame = kwargs.get("weights2") is kwargs.get("weights1") and kwargs.get("weights1") is not None
        kwargs = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in kwargs.items()}
        if wsame:
            kwargs["weights2"] = kwargs["weights1"]
    return fn, tuple(args), kwargs


def flatten(result):
    """Reference results are arrays or tuples of arrays; store them as a list of arrays."""
    if isinstance(result, tuple):
        return [np.asarray(r) for r in result]
    return [np.asarray(result)]
