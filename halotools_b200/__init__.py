"""halotools_b200 - B200-native (sm_100a) pair-counting engine with the call signatures of
``halotools.mock_observables``: ``npairs_3d``, ``npairs_xy_z``, ``npairs_s_mu``, ``marked_npairs_3d``,
``mean_delta_sigma`` and the statistics built on them (``tpcf``, ``wp``, ``rp_pi_tpcf``,
``marked_tpcf``).  Host code is Python; the mesh sort and the pair loops are hand-written CUDA
behind the C ABI of ``include/halotools_b200.h`` (``libhalotools_b200.so``).  No CPU fallback."""
from .custom_exceptions import HalotoolsError
from .catalog_analysis_helpers import (cuboid_subvolume_labels, return_xyz_formatted_array,
                                        apply_zspace_distortion)
from .pair_counters import (npairs_3d, npairs_xy_z, npairs_s_mu, marked_npairs_3d, marked_npairs_xy_z,
                            npairs_projected, npairs_per_object_3d, npairs_jackknife_3d, npairs_jackknife_xy_z)
from .surface_density import (mean_delta_sigma, weighted_npairs_xy, weighted_npairs_per_object_xy,
                              total_mass_enclosed_per_cylinder, total_mass_enclosed_in_stack_of_cylinders,
                              surface_density_in_annulus, surface_density_in_cylinder)
from .two_point_clustering import tpcf, wp, rp_pi_tpcf, marked_tpcf, tpcf_jackknife, wp_jackknife, rp_pi_tpcf_jackknife, s_mu_tpcf, tpcf_multipole, tpcf_one_two_halo_decomp, angular_tpcf

__version__ = "0.1.0"
__all__ = ("HalotoolsError", "cuboid_subvolume_labels", "return_xyz_formatted_array", "apply_zspace_distortion", "npairs_3d", "npairs_xy_z", "npairs_s_mu", "marked_npairs_3d",
           "marked_npairs_xy_z", "npairs_projected", "npairs_per_object_3d", "npairs_jackknife_3d", "npairs_jackknife_xy_z",
           "mean_delta_sigma", "weighted_npairs_xy", "weighted_npairs_per_object_xy",
           "total_mass_enclosed_per_cylinder", "total_mass_enclosed_in_stack_of_cylinders",
           "surface_density_in_annulus", "surface_density_in_cylinder", "tpcf", "wp", "rp_pi_tpcf", "marked_tpcf", "tpcf_jackknife", "wp_jackknife", "rp_pi_tpcf_jackknife", "s_mu_tpcf", "tpcf_multipole", "tpcf_one_two_halo_decomp", "angular_tpcf")
