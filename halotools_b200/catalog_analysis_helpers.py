"""Catalogue helpers on either side of the pair-counting path: the input formatting step
(``return_xyz_formatted_array`` / ``apply_zspace_distortion``,
/root/reference/halotools/mock_observables/catalog_analysis_helpers.py:108-327) and the sub-volume labels of the
jackknife statistics (``cuboid_subvolume_labels``, :330-421).  Host numpy, as in the reference."""
import numpy as np

__all__ = ("cuboid_subvolume_labels", "return_xyz_formatted_array", "apply_zspace_distortion")


def _efunc(cosmology, redshift):
    """E(z) = H(z) / H0 of the caller's cosmology object (anything with an ``efunc`` method, e.g. an astropy
    cosmology).  The reference defaults to astropy's Planck15 (sim_defaults.py), which is not available here: without
    a cosmology only z = 0 (E = 1 exactly, the reference's default redshift) can be served."""
    if cosmology is not None:
        return cosmology.efunc(redshift)
    if np.all(np.asarray(redshift) == 0.0):
        return 1.0
    raise ValueError("redshift-space distortions at redshift != 0 need a ``cosmology`` object with an ``efunc`` method")


def return_xyz_formatted_array(x, y, z, period=np.inf, cosmology=None, redshift=0.0, **kwargs):
    """(Npts, 3) array of positions in the format of the pair counters, optionally with redshift-space distortions
    from ``velocity`` along ``velocity_distortion_dimension`` and wrapped into the periodic box; ``mask`` selects
    rows (catalog_analysis_helpers.py:204-265)."""
    period = np.atleast_1d(period)
    if len(period) == 1:
        period = np.repeat(period, 3)
    elif len(period) == 3:
        pass
    else:
        raise ValueError("Input ``period`` must be a single float or a 3-element sequence")

    x = np.mod(x, period[0])
    y = np.mod(y, period[1])
    z = np.mod(z, period[2])

    posdict = {"x": np.copy(x), "y": np.copy(y), "z": np.copy(z)}
    period_dict = {"x": period[0], "y": period[1], "z": period[2]}

    a = "velocity_distortion_dimension" in kwargs
    b = "velocity" in kwargs
    if a or b:
        if not (a and b):
            raise KeyError("You must either both or none of the following keyword arguments: "
                           "``velocity_distortion_dimension`` and ``velocity``\n")
        vel_dist_dim = kwargs["velocity_distortion_dimension"]
        velocity = np.copy(kwargs["velocity"])
        if vel_dist_dim not in ("x", "y", "z"):
            raise KeyError("\nInput ``velocity_distortion_dimension`` must be either \n"
                           "``'x'``, ``'y'`` or ``'z'``.")
        spatial_distortion = (1.0 + redshift) * np.copy(velocity) / 100.0 / _efunc(cosmology, redshift)
        posdict[vel_dist_dim] = np.copy(posdict[vel_dist_dim]) + spatial_distortion
        Lbox = period_dict[vel_dist_dim]
        if Lbox != np.inf:
            posdict[vel_dist_dim] = posdict[vel_dist_dim] % Lbox       # enforce_periodicity_of_box (model_helpers.py:164-169)

    pos = np.vstack([np.copy(posdict["x"]), np.copy(posdict["y"]), np.copy(posdict["z"])]).T
    try:
        return pos[kwargs["mask"]]
    except KeyError:
        return pos


def apply_zspace_distortion(true_pos, peculiar_velocity, redshift, cosmology, Lbox=None):
    """s = s_true + (1 + z) v_pec / H(z), optionally wrapped into the box (catalog_analysis_helpers.py:319-327)."""
    scale_factor = 1.0 / (1.0 + redshift)
    pos_err = peculiar_velocity / 100.0 / _efunc(cosmology, redshift) / scale_factor
    zspace_pos = true_pos + pos_err
    if Lbox is not None:
        zspace_pos = zspace_pos % Lbox
    return zspace_pos


def cuboid_subvolume_labels(sample, Nsub, Lbox):
    """Integer labels in [1, prod(Nsub)] of the cuboid sub-volume each point of ``sample`` lies in, and the number
    of sub-volumes.  Same rule as the reference: ``floor(sample / (Lbox / Nsub))``, a point exactly on the upper
    face belongs to the last sub-volume, labels count with the LAST dimension fastest."""
    sample = np.atleast_1d(sample).astype("f8")
    try:
        assert sample.ndim == 2
        assert sample.shape[1] == 3
    except AssertionError:
        raise TypeError("Input ``sample`` must have shape (Npts, 3)")

    Nsub = np.atleast_1d(Nsub).astype("i4")
    if len(Nsub) == 1:
        Nsub = np.array([Nsub[0], Nsub[0], Nsub[0]])
    elif len(Nsub) != 3:
        raise TypeError("Input ``Nsub`` must be a scalar or length-3 sequence")

    Lbox = np.atleast_1d(Lbox).astype("f8")
    if len(Lbox) == 1:
        Lbox = np.array([Lbox[0]] * 3)
    elif len(Lbox) != 3:
        raise TypeError("Input ``Lbox`` must be a scalar or length-3 sequence")

    dL = Lbox / Nsub
    N_sub_vol = int(np.prod(Nsub))
    inds = np.arange(1, N_sub_vol + 1).reshape(Nsub[0], Nsub[1], Nsub[2])
    index = np.floor(sample / dL).astype(int)
    for i in range(3):
        index[:, i] = np.where(index[:, i] == Nsub[i], Nsub[i] - 1, index[:, i])
    index = inds[index[:, 0], index[:, 1], index[:, 2]].astype(int)
    return index, int(N_sub_vol)
