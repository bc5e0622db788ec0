"""Small utilities on finished spectra used by the host classes (reference spec_utils.py)."""
import numpy as np


def res_corr(flux, dvbin, fwhm=8):
    """Convolve spectra (last axis, periodic) with the Gaussian of a spectrograph of the given FWHM in km/s; ``dvbin``
    is the pixel width in km/s.  Same discrete filter as the reference (spec_utils.py:5-25 calls
    scipy.ndimage.gaussian_filter1d(mode='wrap')): the Gaussian sampled at the pixel centres, cut at 4 sigma
    (radius int(4 sigma + 0.5) pixels) and renormalised; for sigma around a pixel this differs from the continuous
    Gaussian by tens of per cent, and the maxima of the smoothed spectra pick the line get_observer_tau returns."""
    sigma = (fwhm / dvbin) / (2 * np.sqrt(2 * np.log(2)))
    flux = np.asarray(flux, dtype=np.float64)
    if sigma <= 0:
        return np.array(flux)
    radius = int(4.0 * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    w /= w.sum()
    out = np.zeros_like(flux)
    for k, wk in zip(x, w):  # correlation with a symmetric kernel, periodic
        out += wk * np.roll(flux, -int(k), axis=-1)
    return out


def get_rolled_spectra(tau):
    """Cycle every spectrum so that its peak sits in the middle: (roll offsets, rolled array)."""
    tau = np.asarray(tau)
    mid = int(tau.shape[1] / 2)
    roll = mid - np.argmax(tau, axis=1)
    return roll, np.array([np.roll(t, r) for t, r in zip(tau, roll)])
