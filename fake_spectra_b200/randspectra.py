"""Spectra along randomly placed sightlines: the host-side counterpart of the reference's RandSpectra
(randspectra.py:10-37), same constructor arguments and the same sightlines for a given seed."""
import numpy as np

from . import abstractsnapshot
from .spectra import Spectra


def _box_of(num, base):
    """Box side (kpc/h) from the snapshot header."""
    snap = abstractsnapshot.AbstractSnapshotFactory(num, base)
    try:
        return snap.get_header_attr("BoxSize")
    finally:
        del snap


class RandSpectra(Spectra):
    """``numlos`` sightlines parallel to the x axis through uniformly random points of the box.

    The points come from numpy's global generator after ``np.random.seed(seed)``, which is what makes a run
    repeatable and identical to the reference's for the same seed.  With a positive ``thresh`` (a column
    density, or a [low, high] pair) sightlines are redrawn until ``ndla`` of them pass it
    (``Spectra.replace_not_DLA`` for ``elem`` / ``ion``)."""

    def __init__(self, num, base, MPI=None, seed=23, ndla=1000, numlos=5000, thresh=10 ** 20.3,
                 savefile="rand_spectra_DLA.hdf5", elem="H", ion=1, **kwargs):
        self.box = _box_of(num, base)
        self.NumLos = numlos
        np.random.seed(seed)
        points = self.get_cofm()
        x_axis = np.full(numlos, 1.0)  # axis ids are 1-based: 1 = x
        super().__init__(num, base, points, x_axis, MPI, savefile=savefile, reload_file=True, load_halo=False, **kwargs)
        wants_filter = np.size(thresh) > 1 or thresh > 0
        if wants_filter:
            self.replace_not_DLA(ndla, thresh, elem=elem, ion=ion)
            print("Found objects over threshold")

    def get_cofm(self, num=None):
        """``num`` (default: NumLos) further points, uniform in the box, from numpy's global generator."""
        count = self.NumLos if num is None else num
        return np.random.random_sample((count, 3)) * self.box
