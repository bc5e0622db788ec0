"""Regular grids of sightlines (host-side mirror of the reference's griddedspectra.py)."""
import numpy as np


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
