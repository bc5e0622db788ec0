"""numpy restatement of the reference's flux statistics (test infrastructure: the checker for
fake_spectra_b200.fluxstatistics), following fluxstatistics.py line by line."""
import math

import numpy as np


def flux_pdf_np(tau, nbins=20, scale=1.):
    """fluxstatistics.py:43-52 (mean-flux rescaling applied by the caller through ``scale``)."""
    flux = np.exp(-scale * tau)
    bins = np.arange(nbins + 1) / (1. * nbins)
    fpdf, _ = np.histogram(flux, bins=bins, density=True)
    return (bins[1:] + bins[:-1]) / 2., fpdf


def powerspectrum_np(inarray, axis=-1):
    """fluxstatistics.py:54-61."""
    rfftd = np.fft.rfft(inarray, axis=axis)
    return np.abs(rfftd) ** 2 / np.shape(inarray)[axis] ** 2


def window_function_np(k, *, R, dv):
    """fluxstatistics.py:63-72."""
    sigma = R / (2 * np.sqrt(2 * np.log(2)))
    return np.exp(-0.5 * (k * sigma) ** 2) * np.sinc(k * dv / 2 / math.pi)


def flux_power_bins_np(vmax, npix):
    """fluxstatistics.py:197-215."""
    return np.fft.rfftfreq(npix) * 2.0 * math.pi * npix / vmax


def flux_power_np(tau, vmax, spec_res=8, scale=1., mean_flux_desired=None, window=False):
    """fluxstatistics.py:74-108 (``scale`` = the mean-flux rescaling, found by the caller)."""
    if mean_flux_desired is None:
        mean_flux_desired = np.mean(np.exp(-tau))
    nspec, npix = np.shape(tau)
    mean_flux_power = np.zeros(npix // 2 + 1, dtype=tau.dtype)
    for i in range(10):
        end = min((i + 1) * nspec // 10, nspec)
        dflux = np.exp(-scale * tau[i * nspec // 10:end]) / mean_flux_desired - 1.
        mean_flux_power += vmax * np.sum(powerspectrum_np(dflux, axis=1), axis=0)
    mean_flux_power /= nspec
    kf = flux_power_bins_np(vmax, npix)
    if window and spec_res > 0:
        mean_flux_power /= window_function_np(kf, R=spec_res, dv=vmax / npix) ** 2
    return kf, mean_flux_power
