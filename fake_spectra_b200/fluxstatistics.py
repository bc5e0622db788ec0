"""Flux statistics of optical depths: mean-flux rescaling, flux PDF and 1-D flux power spectrum
(mirror of the reference's fluxstatistics.py:20-108, 197-215 on the device).

Same function names, arguments and return values as the reference.  ``tau`` may be a numpy array
(uploaded once) or a CUDA tensor that is already resident, e.g. what
``native.CandidateIndex.compute_tau`` returns: nothing is copied back but the small results.  The
reductions, the histogram and the Fourier transform behind the power spectrum are this library's kernels
(fsb_stats.cu; no FFT library).  There is no CPU fallback.

The 3-D flux power (nbodykit) is outside the hot path and not provided.
"""
import ctypes as C
import math

import numpy as np

from . import _lib


def _device_tau(tau):
    """tau as a contiguous float64 CUDA tensor (reference: tau.astype(np.float64), fluxstatistics.py:41)."""
    import torch
    if isinstance(tau, torch.Tensor):
        if not tau.is_cuda:
            tau = tau.cuda()
        return tau.to(torch.float64).contiguous()
    if not torch.cuda.is_available():
        raise RuntimeError("fake_spectra_b200.fluxstatistics needs a CUDA device (no CPU fallback)")
    return torch.from_numpy(np.ascontiguousarray(tau, dtype=np.float64)).cuda()


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def obs_mean_tau(redshift):
    """Effective optical depth 0.0023 (1+z)^3.65 of 0711.1862 (fluxstatistics.py:20-23)."""
    return 0.0023 * (1.0 + redshift) ** 3.65


def mean_flux(tau, mean_flux_desired, tol=1e-5, thresh=1e30):
    """Scale factor s with <exp(-s tau)> = mean_flux_desired over the pixels with tau <= thresh
    (fluxstatistics.py:25-41 -> get_mean_flux_scale, py_module.cpp:235-262)."""
    if np.size(tau) == 0 if not hasattr(tau, "numel") else tau.numel() == 0:
        return 0
    import torch
    t = _device_tau(tau)
    scale = C.c_double(0)
    with torch.cuda.device(t.device):
        rc = _lib.load().fsb_rescale_mean_flux(C.c_void_p(t.data_ptr()), t.numel(), float(mean_flux_desired), float(tol),
                                               float(thresh), C.byref(scale), None, _stream())
    _lib.check(rc, "fsb_rescale_mean_flux")
    return scale.value


def flux_pdf(tau, nbins=20, mean_flux_desired=None):
    """Normalised histogram of the flux exp(-tau) on nbins equal bins of [0, 1]
    (fluxstatistics.py:43-52): returns (bin centres, density)."""
    import torch
    t = _device_tau(tau)
    scale = 1.
    if mean_flux_desired is not None:
        scale = mean_flux(t, mean_flux_desired)
    counts = torch.zeros(int(nbins), dtype=torch.int64, device=t.device)
    with torch.cuda.device(t.device):
        rc = _lib.load().fsb_flux_pdf(C.c_void_p(t.data_ptr()), t.numel(), float(scale), int(nbins),
                                      C.c_void_p(counts.data_ptr()), _stream())
    _lib.check(rc, "fsb_flux_pdf")
    counts = counts.cpu().numpy()
    bins = np.arange(nbins + 1) / (1. * nbins)
    # numpy.histogram(density=True): counts / total / bin width
    fpdf = counts / np.diff(bins) / counts.sum()
    cbins = (bins[1:] + bins[:-1]) / 2.
    return cbins, fpdf


def _powerspectrum(inarray, axis=-1):
    """|rfft|^2 / n^2 along ``axis`` (fluxstatistics.py:54-61) by this library's own transform (fsb_flux_power: a
    two-level direct Fourier sum in shared memory, any length); device tensors stay on the device."""
    import torch
    on_device = isinstance(inarray, torch.Tensor)
    t = _device_tau(inarray)
    t = t.movedim(axis, -1).contiguous()
    lead, n = t.shape[:-1], t.shape[-1]
    rows = t.reshape(-1, n)
    out = torch.empty((rows.shape[0], n // 2 + 1), dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        rc = _lib.load().fsb_flux_power(C.c_void_p(rows.data_ptr()), rows.shape[0], n, 1, 1.0, 1.0, 1.0, None,
                                        C.c_void_p(out.data_ptr()), _stream())
    _lib.check(rc, "fsb_flux_power")
    out = out.reshape(lead + (n // 2 + 1,)).movedim(-1, axis)
    return out if on_device else out.cpu().numpy()


def _window_function(k, *, R, dv):
    """Spectrograph response: Gaussian of FWHM R times the pixel sinc (fluxstatistics.py:63-72)."""
    sigma = R / (2 * np.sqrt(2 * np.log(2)))
    return np.exp(-0.5 * (k * sigma) ** 2) * np.sinc(k * dv / 2 / math.pi)


def _flux_power_bins(vmax, npix):
    """k of the rfft modes in s/km (fluxstatistics.py:197-215)."""
    kf = np.fft.rfftfreq(npix)
    return kf * 2.0 * math.pi * npix / vmax


def flux_power(tau, vmax, spec_res=8, mean_flux_desired=None, window=False):
    """Mean 1-D flux power spectrum of delta_F = exp(-tau)/<F> - 1 over the sightlines
    (fluxstatistics.py:74-108): returns (k [s/km], P_F [km/s]), both of length npix//2 + 1.
    One pass over the resident optical depths: the flux contrast is formed inside the transform kernel
    (fsb_flux_power), nothing of the size of tau is written."""
    import torch
    t = _device_tau(tau)
    if t.dim() != 2:
        raise ValueError("tau must have shape (NumLos, npix)")
    nspec, npix = t.shape
    scale = 1.
    if mean_flux_desired is not None:
        scale = mean_flux(t, mean_flux_desired)
    else:
        mean_flux_desired = _mean_exp(t)  # np.mean(np.exp(-tau)), fluxstatistics.py:96
    nk = npix // 2 + 1
    power = torch.zeros(nk, dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        # vmax * |F|^2 / npix^2, averaged over all sightlines
        rc = _lib.load().fsb_flux_power(C.c_void_p(t.data_ptr()), nspec, npix, 0, float(scale), float(mean_flux_desired),
                                        vmax / (1. * npix * npix) / nspec, C.c_void_p(power.data_ptr()), None, _stream())
    _lib.check(rc, "fsb_flux_power")
    mean_flux_power = power.cpu().numpy()
    kf = _flux_power_bins(vmax, npix)
    if window and spec_res > 0:
        mean_flux_power /= _window_function(kf, R=spec_res, dv=vmax / npix) ** 2
    return kf, mean_flux_power


def flux_sums(tau, scale=1.0, thresh=1e30):
    """(sum exp(-scale tau), sum tau exp(-scale tau), pixels used) over tau <= thresh: one pass of the
    reduction behind mean_flux.  sum / used at scale 1 is the mean flux (spectra.py:1276)."""
    import torch
    t = _device_tau(tau)
    sf, stf, used = C.c_double(0), C.c_double(0), C.c_int64(0)
    with torch.cuda.device(t.device):
        rc = _lib.load().fsb_flux_sums(C.c_void_p(t.data_ptr()), t.numel(), float(scale), float(thresh), C.byref(sf),
                                       C.byref(stf), C.byref(used), _stream())
    _lib.check(rc, "fsb_flux_sums")
    return sf.value, stf.value, used.value


def row_max(tau):
    """Maximum optical depth of every sightline (device reduction, fsb_row_max)."""
    import torch
    t = _device_tau(tau)
    if t.dim() != 2:
        raise ValueError("tau must have shape (NumLos, npix)")
    out = torch.empty(t.shape[0], dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        rc = _lib.load().fsb_row_max(C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], C.c_void_p(out.data_ptr()), _stream())
    _lib.check(rc, "fsb_row_max")
    return out.cpu().numpy()


def mask_damped_region(tt, taueff, tau_thresh=1e6, thresh2=0.25):
    """One spectrum with a damped absorber (spectra.py:1219-1252, after Chabanier et al. 2019): around every maximum above
    ``tau_thresh`` the pixels are set to ``taueff`` outwards in both directions (periodic) for as long as they exceed
    ``taueff + thresh2``.  Alters ``tt``; returns (tt, a pixel count as the reference tallies it)."""
    n = tt.shape[0]
    limit = taueff + thresh2
    tot = 0
    while tt.max() > tau_thresh:
        peak = int(tt.argmax())
        for direction, first in ((-1, 0), (1, 1)):
            j = first
            while tt[(peak + direction * j) % n] > limit:
                tt[(peak + direction * j) % n] = taueff
                j += 1
            # the reference's upward counter runs modulo the pixel count once it wraps (spectra.py:1243-1250)
            tot += j if direction < 0 or peak + j < n else j - n
    return tt, tot


def filter_tau(tau, tau_thresh):
    """Spectra._filter_tau (spectra.py:1254-1270): sightlines whose maximum optical depth exceeds ``tau_thresh`` have
    their damped regions replaced by the effective optical depth -log <exp(-tau)> of the whole sample.  The two
    reductions over the sample (mean flux, row maxima) run on the device; the few affected rows are edited on the host.
    Alters ``tau`` (a host array) in place like the reference."""
    if tau_thresh is None:
        return tau
    t = _device_tau(tau)
    taueff = -math.log(_mean_exp(t))
    rows = np.where(row_max(t) > tau_thresh)[0]
    for i in rows:
        tau[i], _ = mask_damped_region(tau[i], taueff, tau_thresh=tau_thresh)
    if rows.size:
        assert max(tau[i].max() for i in rows) < tau_thresh * 1.01
    return tau


def _mean_exp(t):
    sf, _, used = flux_sums(t)
    return sf / used if used else float("nan")
