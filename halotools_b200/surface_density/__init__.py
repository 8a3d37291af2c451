from .mean_delta_sigma import mean_delta_sigma
from .weighted_npairs_xy import weighted_npairs_xy
from .weighted_npairs_per_object_xy import (weighted_npairs_per_object_xy, total_mass_enclosed_per_cylinder,
                                            total_mass_enclosed_in_stack_of_cylinders, surface_density_in_annulus,
                                            surface_density_in_cylinder)

__all__ = ("mean_delta_sigma", "weighted_npairs_xy", "weighted_npairs_per_object_xy", "total_mass_enclosed_per_cylinder",
           "total_mass_enclosed_in_stack_of_cylinders", "surface_density_in_annulus", "surface_density_in_cylinder")
