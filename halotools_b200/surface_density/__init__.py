from .mean_delta_sigma import mean_delta_sigma
from .weighted_npairs_xy import weighted_npairs_xy

__all__ = ("mean_delta_sigma", "weighted_npairs_xy")
