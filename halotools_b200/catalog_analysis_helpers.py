"""Catalogue helpers on either side of the pair-counting path: the input formatting step
(``return_xyz_formatted_array`` / ``apply_zspace_distortion``,
/root/reference/halotools/mock_observables/catalog_analysis_helpers.py:108-327) and the sub-volume labels of the
jackknife statistics (``cuboid_subvolume_labels``, :330-421).  Host arrays go through numpy, as in the reference;
torch CUDA tensors go through one elementwise kernel of the library (``htb_return_xyz_formatted_array`` /
``htb_apply_zspace_distortion``) and come back as a device-resident (Npts, 3) sample the pair counters take as it is:
a mock that lives on the device never crosses PCIe."""
import ctypes

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

    if getattr(x, "is_cuda", False):
        return _xyz_formatted_device(x, y, z, period, cosmology, redshift, kwargs)

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


def _dev_ptr(t):
    return ctypes.cast(ctypes.c_void_p(int(t.data_ptr())), ctypes.POINTER(ctypes.c_double))


def _dev_f64(t, n, name):
    import torch
    if not (getattr(t, "is_cuda", False) and t.dtype == torch.float64 and t.dim() == 1 and int(t.shape[0]) == n):
        raise TypeError("device input ``%s`` must be a float64 CUDA tensor of shape (Npts,)" % name)
    return t.contiguous()


def _xyz_formatted_device(x, y, z, period, cosmology, redshift, kwargs):
    """The same function for torch CUDA tensors: one kernel, (Npts, 3) float64 CUDA tensor out, no host copy."""
    import torch
    from . import _lib
    n = int(x.shape[0])
    x, y, z = (_dev_f64(t, n, nm) for t, nm in ((x, "x"), (y, "y"), (z, "z")))
    a = "velocity_distortion_dimension" in kwargs
    b = "velocity" in kwargs
    dim, vel, efunc = -1, None, 1.0
    if a or b:
        if not (a and b):
            raise KeyError("You must either both or none of the following keyword arguments: "
                           "``velocity_distortion_dimension`` and ``velocity``\n")
        vel_dist_dim = kwargs["velocity_distortion_dimension"]
        if vel_dist_dim not in ("x", "y", "z"):
            raise KeyError("\nInput ``velocity_distortion_dimension`` must be either \n"
                           "``'x'``, ``'y'`` or ``'z'``.")
        dim = "xyz".index(vel_dist_dim)
        vel = _dev_f64(kwargs["velocity"], n, "velocity")
        if np.ndim(redshift) != 0:
            raise TypeError("device samples take a scalar ``redshift``")
        efunc = float(_efunc(cosmology, redshift))
    per = (ctypes.c_double * 3)(*[float(p) for p in period])
    lib = _lib.require_gpu()
    with torch.cuda.stream(_lib.engine_stream()):
        out = torch.empty((n, 3), dtype=torch.float64, device=x.device)
        _lib.check(lib.htb_return_xyz_formatted_array(
            _dev_ptr(x), _dev_ptr(y), _dev_ptr(z), ctypes.c_int64(n), per,
            _dev_ptr(vel) if vel is not None else None, ctypes.c_int32(dim), ctypes.c_double(float(redshift)),
            ctypes.c_double(efunc), _dev_ptr(out)))
        if "mask" in kwargs:
            out = out[kwargs["mask"]]
    # the sample is used by later engine calls on the same stream; other torch streams see it after this
    torch.cuda.current_stream().wait_stream(_lib.engine_stream())
    return out


def apply_zspace_distortion(true_pos, peculiar_velocity, redshift, cosmology, Lbox=None):
    """s = s_true + (1 + z) v_pec / H(z), optionally wrapped into the box (catalog_analysis_helpers.py:319-327)."""
    if getattr(true_pos, "is_cuda", False):
        import torch
        from . import _lib
        n = int(true_pos.shape[0])
        pos, vel = _dev_f64(true_pos, n, "true_pos"), _dev_f64(peculiar_velocity, n, "peculiar_velocity")
        if np.ndim(redshift) != 0:
            raise TypeError("device samples take a scalar ``redshift``")
        lib = _lib.require_gpu()
        with torch.cuda.stream(_lib.engine_stream()):
            out = torch.empty(n, dtype=torch.float64, device=pos.device)
            _lib.check(lib.htb_apply_zspace_distortion(
                _dev_ptr(pos), _dev_ptr(vel), ctypes.c_int64(n), ctypes.c_double(float(redshift)),
                ctypes.c_double(float(_efunc(cosmology, redshift))), ctypes.c_double(float(Lbox) if Lbox is not None else 0.0),
                ctypes.c_int32(0 if Lbox is None else 1), _dev_ptr(out)))
        torch.cuda.current_stream().wait_stream(_lib.engine_stream())
        return out
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
