"""Regular grids of sightlines (host-side mirror of the reference's griddedspectra.py)."""
import numpy as np

from . import abstractsnapshot as absn
from . import spectra


def grid_axes_and_cofm(box, nspec, axis):
    """Sightline directions and positions on an ``nspec x nspec`` grid perpendicular to ``axis``
    (1-based); ``axis < 0`` gives all three directions, 3*nspec^2 lines.  Spacing box/nspec, first
    line on coordinate 0.  Same layout as reference griddedspectra.py:59-88 (row-major over the two
    perpendicular coordinates, axis 1 block first)."""
    n, m = np.meshgrid(np.arange(nspec), np.arange(nspec), indexing="ij")
    n, m, zero = n.ravel(), m.ravel(), np.zeros(nspec * nspec, dtype=np.int64)
    blocks = {1: np.stack([zero, n, m], axis=1), 2: np.stack([n, zero, m], axis=1), 3: np.stack([n, m, zero], axis=1)}
    if axis < 0:
        grid_id = np.concatenate([blocks[1], blocks[2], blocks[3]])
        grid_axes = np.repeat(np.array([1., 2., 3.]), nspec * nspec)
    elif axis in blocks:
        grid_id = blocks[axis]
        grid_axes = axis * np.ones(nspec * nspec)
    else:
        raise ValueError('wrong axis number {}'.format(axis))
    dx = box / (1. * nspec)
    return grid_axes, dx * grid_id


class GriddedSpectra(spectra.Spectra):
    """Regular ``nspec x nspec`` grid of sightlines along ``axis`` (all three axes if ``axis < 0``);
    constructor arguments as the reference's griddedspectra.py:12-56."""

    def __init__(self, num, base, nspec=200, MPI=None, res=None, nbins=None, savefile="gridded_spectra.hdf5",
                 reload_file=True, grid_axes=None, grid_cofm=None, axis=1, **kwargs):
        if reload_file is False:
            grid_cofm = None
            grid_axes = None
        else:
            f = absn.AbstractSnapshotFactory(num, base)
            self.box = f.get_header_attr("BoxSize")
            del f
            if grid_cofm is None:
                grid_axes, grid_cofm = self.get_axes_and_cofm(nspec, axis)
        spectra.Spectra.__init__(self, num, base, cofm=grid_cofm, axis=grid_axes, MPI=MPI, res=res, nbins=nbins,
                                 savefile=savefile, reload_file=reload_file, **kwargs)

    def get_axes_and_cofm(self, nspec, axis):
        """Position of the skewers in the grid."""
        return grid_axes_and_cofm(self.box, nspec, axis)
