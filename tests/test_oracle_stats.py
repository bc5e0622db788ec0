"""Flux statistics, CPU side: the oracle's restatement of get_mean_flux_scale
(py_module.cpp:235-262) and a numpy restatement of flux_pdf / flux_power (fluxstatistics.py:43-108),
pinned against the known answers of the reference's own tests (fake_spectra/tests/test_statistics.py:8-68,
values restated below)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import statcases  # noqa: E402


def test_mean_flux_known_answers(oracle):
    """test_statistics.py:8-19: tau = log n gives mean flux <n^-s>, so the scale for <n^-s> is s."""
    tol = 1e-4
    nn = np.arange(1, 101)
    tau = np.log(nn)
    for s in (1.0, 2.0, 0.5):
        mf = np.mean(nn ** (-s))
        assert abs(oracle.mean_flux_scale(tau, mf, tol) - s) < tol


def test_mean_flux_thresh_and_empty(oracle):
    rng = np.random.default_rng(3)
    tau = rng.exponential(0.6, 5000)
    tau[::50] = 1e4  # saturated pixels, excluded by thresh
    s = oracle.mean_flux_scale(tau, 0.7, 1e-8, thresh=100.)
    keep = tau <= 100.
    assert abs(np.mean(np.exp(-s * tau[keep])) - 0.7) < 1e-7
    assert oracle.mean_flux_scale(np.zeros(0), 0.5) == 0.0
    # the clamp of py_module.cpp:257-259: a target above 1 drives the scale to the 1e-10 floor
    assert oracle.mean_flux_scale(tau[keep], 1.5, 1e-3) == 1e-10


def test_flux_pdf_known_answers():
    """test_statistics.py:21-37."""
    nn = np.arange(1, 101, dtype=np.double)
    bins, hist = statcases.flux_pdf_np(np.log(nn), 20)
    assert bins[0] == 0. + 1 / 40. and bins[-1] == 1. - 1. / 40.
    assert np.min(hist) == 0. and np.max(hist) > 1.
    expected = np.array([16., 2.2, 0.6, 0.2, 0.2, 0.2, 0.2, 0., 0., 0., 0.2, 0., 0., 0., 0., 0., 0., 0., 0., 0.2])
    assert np.abs(np.sum(expected) - np.sum(hist)) < 1e-3
    assert np.all(np.abs(hist[3:] - expected[3:]) < 1e-3)
    assert np.abs(hist[0] - expected[0]) < 1e-2


def test_flux_power_known_answers():
    """test_statistics.py:39-68: Parseval, a delta function at the input frequency, and the window."""
    xx = np.tile(np.arange(0, 1, 0.01) ** 2, (10, 1))
    fpk = statcases.powerspectrum_np(np.exp(-xx), axis=1) * (2 * math.pi)
    for ff in fpk:
        dpower = np.sum(ff) + np.sum(ff[1:])
        assert abs(dpower - 2 * math.pi * np.sum(np.exp(-xx[0, :]) ** 2) / np.shape(xx)[1]) < 0.1
    inn = np.sin(2 * math.pi * np.linspace(1, 51, 200))
    ff = statcases.powerspectrum_np(inn)
    assert np.where(np.max(ff) == ff)[0][0] == 50
    for bb in (200, 201):
        xx = np.linspace(0, 51, bb)
        inn = np.sin(2 * math.pi * xx) + 1.5
        ff = statcases.powerspectrum_np(np.exp(-inn) - 1)
        taus = np.vstack([inn, ] * 10)
        bins, power = statcases.flux_power_np(taus, vmax=1., spec_res=0.01, window=True)
        power /= 12.5569
        wind = statcases.window_function_np(bins[1:], R=0.01, dv=1 / np.size(xx))
        assert np.all(np.abs(power[1:] * wind ** 2 - ff[1:]) < 0.01 * ff[1:])
        assert power[0] < 1e-20
