"""The one catalogue helper the jackknife statistics need
(/root/reference/halotools/mock_observables/catalog_analysis_helpers.py:330-421)."""
import numpy as np

__all__ = ("cuboid_subvolume_labels",)


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
