"""Sightlines at random positions (host-side mirror of the reference's randspectra.py:10-37)."""
import numpy as np

from . import abstractsnapshot as absn
from . import spectra


class RandSpectra(spectra.Spectra):
    """``numlos`` sightlines along the x axis at positions drawn with ``np.random.seed(seed)``;
    with ``thresh > 0`` sightlines are redrawn until ``ndla`` exceed the column-density threshold."""

    def __init__(self, num, base, MPI=None, seed=23, ndla=1000, numlos=5000, thresh=10 ** 20.3,
                 savefile="rand_spectra_DLA.hdf5", elem="H", ion=1, **kwargs):
        f = absn.AbstractSnapshotFactory(num, base)
        self.box = f.get_header_attr("BoxSize")
        del f
        self.NumLos = numlos
        axis = np.ones(self.NumLos)  # 1 for x, 2 for y, 3 for z
        np.random.seed(seed)
        cofm = self.get_cofm()
        spectra.Spectra.__init__(self, num, base, cofm, axis, MPI, savefile=savefile, reload_file=True, load_halo=False,
                                 **kwargs)
        if np.size(thresh) > 1 or thresh > 0:
            self.replace_not_DLA(ndla, thresh, elem=elem, ion=ion)
            print("Found objects over threshold")

    def get_cofm(self, num=None):
        """More sightlines at uniformly random positions in the box."""
        if num is None:
            num = self.NumLos
        return self.box * np.random.random_sample((num, 3))
