from .tpcf import tpcf
from .wp import wp
from .rp_pi_tpcf import rp_pi_tpcf
from .marked_tpcf import marked_tpcf
from .tpcf_jackknife import tpcf_jackknife, wp_jackknife, rp_pi_tpcf_jackknife
from .s_mu_tpcf import s_mu_tpcf, tpcf_multipole
from .tpcf_one_two_halo_decomp import tpcf_one_two_halo_decomp
from .angular_tpcf import angular_tpcf

__all__ = ("tpcf", "wp", "rp_pi_tpcf", "marked_tpcf", "tpcf_jackknife", "wp_jackknife", "rp_pi_tpcf_jackknife", "s_mu_tpcf", "tpcf_multipole",
           "tpcf_one_two_halo_decomp", "angular_tpcf")
