"""Flux statistics on the device (fsb_stats.cu through fake_spectra_b200.fluxstatistics) against the
CPU oracle / the numpy restatement of the reference (tests/statcases.py), plus the reference's own
known answers (fake_spectra/tests/test_statistics.py) run through the device path.
Tolerances: histogram counts exact; scale factors agree to 100x the Newton tolerance's rounding
(the iteration is the same, the summation order is not); power spectra 1e-10 relative."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import statcases  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fstat():
    import torch
    assert torch.cuda.is_available()
    from fake_spectra_b200 import fluxstatistics
    return fluxstatistics


def forest(nspec=300, npix=1115, seed=5):
    """Log-normal-ish optical depths with a few saturated pixels and exact zeros."""
    rng = np.random.default_rng(seed)
    tau = np.exp(rng.normal(-1.0, 1.3, (nspec, npix)))
    tau[rng.random((nspec, npix)) < 0.02] = 0.0
    tau[rng.random((nspec, npix)) < 0.001] *= 1e4
    return tau


def test_mean_flux_known_answers_on_device(fstat):
    tol = 1e-4
    nn = np.arange(1, 101)
    tau = np.log(nn)
    for s in (1.0, 2.0, 0.5):
        assert abs(fstat.mean_flux(tau, np.mean(nn ** (-s)), tol) - s) < tol
    assert fstat.mean_flux(np.zeros(0), 0.5) == 0


@pytest.mark.parametrize("n", [1, 2, 3, 255, 257, 1000003])
def test_mean_flux_matches_oracle(fstat, oracle, n):
    rng = np.random.default_rng(n)
    tau = rng.exponential(0.8, n)
    if n > 10:
        tau[::37] = 1e5
    for target, thresh in ((0.75, 1e30), (0.6, 50.0)):
        want, it_want = oracle.mean_flux_scale(tau, target, 1e-10, thresh, return_iterations=True)
        got = fstat.mean_flux(tau, target, 1e-10, thresh)
        assert abs(got - want) <= 1e-9 * want, (n, target, got, want)


def test_mean_flux_of_unaligned_device_view(fstat, oracle):
    """A device tensor that does not start on a 16-byte boundary takes the scalar-load path."""
    import torch
    tau = forest(7, 333)
    t = torch.from_numpy(tau.ravel()).cuda()[1:]
    assert t.data_ptr() % 16 == 8
    want = oracle.mean_flux_scale(tau.ravel()[1:], 0.7, 1e-10)
    assert abs(fstat.mean_flux(t, 0.7, 1e-10) - want) <= 1e-9 * want
    sf, stf, used = fstat.flux_sums(t)
    assert used == t.numel() and abs(sf - np.exp(-tau.ravel()[1:]).sum()) <= 1e-12 * sf


def test_rescale_drop_in(oracle):
    from fake_spectra_b200 import _spectra_priv as priv
    tau = forest(20, 200)
    want = oracle.mean_flux_scale(tau.ravel(), 0.8, 1e-8)
    assert abs(priv._rescale_mean_flux(tau, 0.8, tau.size, 1e-8, 1e30) - want) <= 1e-8 * want
    with pytest.raises(TypeError):
        priv._rescale_mean_flux(tau.astype(np.float32), 0.8, tau.size, 1e-8, 1e30)


def test_flux_pdf_known_answers_on_device(fstat):
    nn = np.arange(1, 101, dtype=np.double)
    bins, hist = fstat.flux_pdf(np.log(nn), 20)
    want_bins, want = statcases.flux_pdf_np(np.log(nn), 20)
    assert np.array_equal(bins, want_bins) and np.array_equal(hist, want)
    assert bins[0] == 1 / 40. and bins[-1] == 1. - 1. / 40.


@pytest.mark.parametrize("nbins", [1, 7, 20, 1000])
def test_flux_pdf_matches_numpy_histogram(fstat, oracle, nbins):
    tau = forest()
    bins, hist = fstat.flux_pdf(tau, nbins)
    wb, want = statcases.flux_pdf_np(tau, nbins)
    assert np.array_equal(bins, wb)
    assert np.allclose(hist, want, rtol=0, atol=1e-12 * want.max())
    # pixels whose flux sits on a bin edge up to the last bit of exp(): each may land on either side (the device's
    # exp is within an ulp of libm's; the reference's own test allows the same, test_statistics.py:35), nothing else moves
    edges = np.arange(1, nbins + 1) / nbins
    tau.ravel()[:edges.size] = -np.log(edges)
    _, hist = fstat.flux_pdf(tau, nbins)
    _, want = statcases.flux_pdf_np(tau, nbins)
    to_counts = tau.size / nbins
    assert np.abs(hist - want).sum() * to_counts <= 2 * edges.size + 1e-6
    # exact edges (flux 0 and 1 included) follow numpy's rule: tau = 0 -> flux 1.0 -> last bin
    _, hist = fstat.flux_pdf(np.zeros(10), nbins)
    assert np.array_equal(hist, statcases.flux_pdf_np(np.zeros(10), nbins)[1]) and hist[:-1].sum() == 0 and hist[-1] > 0
    # with rescaling: the histogram of the rescaled flux, scale from the same Newton iteration
    scale = oracle.mean_flux_scale(tau.ravel(), 0.7, 1e-5)
    _, hist = fstat.flux_pdf(tau, nbins, mean_flux_desired=0.7)
    _, want = statcases.flux_pdf_np(tau, nbins, scale=scale)
    assert np.abs(hist - want).sum() <= 1e-4 * want.sum()   # a pixel may change bin with the last digits of the scale


@pytest.mark.parametrize("npix", [200, 201, 1115])
def test_flux_power_matches_numpy(fstat, oracle, npix):
    tau = forest(123, npix)
    kf, power = fstat.flux_power(tau, vmax=1234.5)
    wk, want = statcases.flux_power_np(tau, vmax=1234.5)
    assert np.array_equal(kf, wk)
    assert np.max(np.abs(power - want)) <= 1e-10 * np.max(want)
    scale = oracle.mean_flux_scale(tau.ravel(), 0.66, 1e-5)
    kf, power = fstat.flux_power(tau, vmax=1234.5, mean_flux_desired=0.66, spec_res=8, window=True)
    wk, want = statcases.flux_power_np(tau, vmax=1234.5, scale=scale, mean_flux_desired=0.66, spec_res=8, window=True)
    assert np.max(np.abs(power / want - 1)) <= 1e-6   # the two Newton runs stop within tol = 1e-5 of each other


@pytest.mark.parametrize("npix", [1, 2, 3, 17, 64, 223, 1115, 4460, 8921, 9973])
def test_own_transform_matches_numpy_rfft(fstat, npix):
    """|rfft|^2 / n^2 by the two-level direct Fourier sum (fsb_flux_power) against numpy for composite, prime, even,
    odd and degenerate lengths: C1 (1115 = 5 223), C2 (4460 = 2^2 5 223), C3 (8921 = 11 811), a prime close to the
    shared-memory limit (9973)."""
    rng = np.random.default_rng(npix)
    x = rng.normal(0, 1, (5, npix)) + 0.3
    want = np.abs(np.fft.rfft(x, axis=1)) ** 2 / npix ** 2
    got = fstat._powerspectrum(x, axis=1)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 2e-12 * np.max(want)
    if npix > 1:
        assert np.allclose(fstat._powerspectrum(x.T.copy(), axis=0), want.T, rtol=0, atol=2e-12 * np.max(want))


def test_flux_power_known_answers_on_device(fstat):
    """test_statistics.py:55-68 through the device path."""
    for bb in (200, 201):
        xx = np.linspace(0, 51, bb)
        inn = np.sin(2 * math.pi * xx) + 1.5
        ff = statcases.powerspectrum_np(np.exp(-inn) - 1)
        taus = np.vstack([inn, ] * 10)
        bins, power = fstat.flux_power(taus, vmax=1., spec_res=0.01, window=True)
        power /= 12.5569
        wind = fstat._window_function(bins[1:], R=0.01, dv=1 / np.size(xx))
        assert np.all(np.abs(power[1:] * wind ** 2 - ff[1:]) < 0.01 * ff[1:])
        assert power[0] < 1e-20
    assert np.allclose(fstat._powerspectrum(np.exp(-taus), axis=1), statcases.powerspectrum_np(np.exp(-taus), axis=1), rtol=1e-12, atol=1e-18)


def test_spectra_class_statistics(oracle):
    """Spectra.get_mean_flux / get_flux_pdf / get_flux_power_1D on a synthetic snapshot."""
    import hostcases
    from fake_spectra_b200 import randspectra
    rs = randspectra.RandSpectra(0, hostcases.snapshot(12), numlos=20, thresh=0., res=1.5, quiet=True)
    tau = rs.get_tau("H", 1, 1215)
    assert abs(rs.get_mean_flux() - np.mean(np.exp(-tau))) < 1e-13
    _, pdf = rs.get_flux_pdf(nbins=10)
    assert np.allclose(pdf, statcases.flux_pdf_np(tau, 10)[1], atol=1e-12)
    kf, pk = rs.get_flux_power_1D()
    wk, want = statcases.flux_power_np(tau, rs.vmax)
    assert np.array_equal(kf, wk[1:]) and np.max(np.abs(pk - want[1:])) <= 1e-10 * np.max(want)
    assert abs(rs.get_mean_flux(tau_thresh=1e6) - np.mean(np.exp(-tau))) < 1e-13  # nothing is that thick here


def test_damped_absorbers_are_masked(fstat):
    """tau_thresh (spectra.py:1254-1270): rows whose maximum exceeds the threshold have the damped region set to the
    sample's effective optical depth out to taueff + 0.25 on both sides, periodically; other rows are untouched."""
    tau = forest(64, 500, seed=9) * 0.3
    tau[tau > 50] = 1.0
    ref = tau.copy()
    x = np.arange(500)
    tau[7] += 2e6 * np.exp(-((x - 120) / 4.0) ** 2)           # a damped absorber in the middle of row 7
    tau[30] += 5e6 * np.exp(-((((x - 2) + 250) % 500 - 250) / 3.0) ** 2)  # one that straddles the periodic edge of row 30
    dirty = tau.copy()
    taueff = -math.log(np.mean(np.exp(-dirty)))
    assert np.allclose(fstat.row_max(dirty), dirty.max(axis=1), rtol=0, atol=0)
    out = fstat.filter_tau(tau, 1e6)
    assert out is tau
    keep = np.ones(64, dtype=bool)
    keep[[7, 30]] = False
    assert np.array_equal(tau[keep], dirty[keep])
    for row, centre in ((7, 120), (30, 2)):
        masked = tau[row] != dirty[row]
        assert masked[centre] and np.all(np.abs(tau[row][masked] - taueff) < 1e-12) and np.unique(tau[row][masked]).size == 1
        assert tau[row].max() < 1e6 and masked.sum() > 10
        # the masked region is one periodic run around the peak, bounded by the first pixels at or below taueff + 0.25
        idx = (np.where(masked)[0] - centre + 250) % 500 - 250
        lo, hi = idx.min(), idx.max()
        assert np.array_equal(np.sort(idx), np.arange(lo, hi + 1))
        assert dirty[row][(centre + lo - 1) % 500] <= taueff + 0.25 and dirty[row][(centre + hi + 1) % 500] <= taueff + 0.25
        assert np.all(dirty[row][(centre + np.arange(lo, hi + 1)) % 500] > taueff + 0.25)
    assert ref.shape == tau.shape
