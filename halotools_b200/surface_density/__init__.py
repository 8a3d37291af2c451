from .mean_delta_sigma import mean_delta_sigma

__all__ = ("mean_delta_sigma",)
