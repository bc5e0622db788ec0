"""Small utilities on finished spectra used by the host classes (reference spec_utils.py)."""
import numpy as np


def res_corr(flux, dvbin, fwhm=8):
    """Convolve spectra (last axis, periodic) with the Gaussian of a spectrograph of the given FWHM
    in km/s; ``dvbin`` is the pixel width in km/s."""
    sigma = (fwhm / dvbin) / (2 * np.sqrt(2 * np.log(2)))
    if sigma <= 0:
        return np.array(flux)
    n = np.shape(flux)[-1]
    # periodic Gaussian kernel, normalised, applied in Fourier space
    k = np.fft.rfftfreq(n)
    window = np.exp(-2 * (np.pi * k * sigma) ** 2)
    return np.fft.irfft(np.fft.rfft(flux, axis=-1) * window, n=n, axis=-1)


def get_rolled_spectra(tau):
    """Cycle every spectrum so that its peak sits in the middle: (roll offsets, rolled array)."""
    tau = np.asarray(tau)
    mid = int(tau.shape[1] / 2)
    roll = mid - np.argmax(tau, axis=1)
    return roll, np.array([np.roll(t, r) for t, r in zip(tau, roll)])
