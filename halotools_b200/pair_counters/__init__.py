"""Pair counters with the reference's names and signatures
(/root/reference/halotools/mock_observables/pair_counters/__init__.py)."""
from .npairs_3d import npairs_3d
from .npairs_xy_z import npairs_xy_z
from .npairs_s_mu import npairs_s_mu
from .marked_npairs_3d import marked_npairs_3d
from .marked_npairs_xy_z import marked_npairs_xy_z
from .npairs_projected import npairs_projected
from .npairs_per_object_3d import npairs_per_object_3d
from .npairs_jackknife_3d import npairs_jackknife_3d, npairs_jackknife_xy_z

__all__ = ("npairs_3d", "npairs_xy_z", "npairs_s_mu", "marked_npairs_3d", "marked_npairs_xy_z",
           "npairs_projected", "npairs_per_object_3d", "npairs_jackknife_3d", "npairs_jackknife_xy_z")
