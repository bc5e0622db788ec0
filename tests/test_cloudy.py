"""Cloudy table host side against the library calls the reference makes:
scipy.interpolate.interp1d along redshift (convert_cloudy.py:152-153) and scipy.ndimage.map_coordinates
(:196-199).  Also pins the restatement the device kernel evaluates (csrc/fsb_prep.cu ion_fraction): 4 x 4 cubic B-spline
weights on scipy's own prefiltered, edge-padded coefficients reproduce map_coordinates(mode="nearest")."""
import numpy as np
import pytest
import scipy.interpolate as intp
from scipy.ndimage import map_coordinates

from fake_spectra_b200 import cloudy


def make_table(seed=3, nred=4):
    """A smooth synthetic table with the reference's grid: [redshift, 55 densities, 112 temperatures, 9 species, 17 ions]."""
    rng = np.random.default_rng(seed)
    dens, temp = np.arange(-7, 4, 0.2), np.arange(3, 8.6, 0.05)
    d, t = np.meshgrid(dens, temp, indexing="ij")
    table = np.empty((nred, dens.size, temp.size, 9, cloudy.NIONS))
    for z in range(nred):
        for s in range(9):
            for i in range(cloudy.NIONS):
                a, b, c = rng.uniform(0.2, 1.5, 3)
                table[z, :, :, s, i] = -3 * (1 + np.sin(a * d + 0.3 * z) * np.cos(b * t + i)) - c * 0.1 * (t - 5) ** 2
    table[:, :10, :20, 3, 4] = -30.  # N V: Cloudy's log(0) plateau with a sharp edge
    return table, np.array([0., 2., 3., 5.])[:nred]


def bspline_eval(tb, coef, c0, c1):
    """The device kernel's evaluation, in numpy."""
    pad = cloudy.SPLINE_PAD
    nd, nt = tb.dens.size, tb.temp.size
    c0 = np.clip(c0, -pad, nd - 1 + pad) + pad
    c1 = np.clip(c1, -pad, nt - 1 + pad) + pad

    def w(t):
        u = 1 - t
        return [u ** 3 / 6, (3 * t ** 3 - 6 * t ** 2 + 4) / 6, (3 * u ** 3 - 6 * u ** 2 + 4) / 6, t ** 3 / 6]
    i0, i1 = np.floor(c0).astype(int), np.floor(c1).astype(int)
    w0, w1 = w(c0 - i0), w(c1 - i1)
    out = np.zeros_like(c0)
    for a in range(4):
        for b in range(4):
            out += w0[a] * w1[b] * coef[np.clip(i0 - 1 + a, 0, coef.shape[0] - 1), np.clip(i1 - 1 + b, 0, coef.shape[1] - 1)]
    return out


def test_redshift_interpolation_matches_interp1d():
    table, reds = make_table()
    for z in (0., 1.3, 2., 2.999, 4.05, 5.):
        tb = cloudy.CloudyTable(z, table=table, reds=reds)
        want = intp.interp1d(reds, table, axis=0)(4.0 if 4.0 < z < 4.1 else z)
        assert np.allclose(tb.red_table, want, rtol=0, atol=1e-13)
    with pytest.raises(ValueError):
        cloudy.CloudyTable(5.5, table=table, reds=reds)


def test_ion_matches_the_reference_formula():
    table, reds = make_table()
    tb = cloudy.CloudyTable(2.4, table=table, reds=reds)
    rng = np.random.default_rng(0)
    rho = (10 ** rng.uniform(-6.9, 3.7, 4000)).astype(np.float32)
    temp = (10 ** rng.uniform(3.0, 8.5, 4000)).astype(np.float32)
    got = tb.ion("C", 4, np.array(rho), temp)
    scaled = rho * 0.774132
    crho = (np.log10(scaled) - tb.dens[0]) * (tb.dens.size - 1) / (tb.dens[-1] - tb.dens[0])
    ctemp = (np.log10(temp) - tb.temp[0]) * (tb.temp.size - 1) / (tb.temp[-1] - tb.temp[0])
    want = 10 ** map_coordinates(tb.red_table[:, :, 2, 3], np.vstack((crho, ctemp)), mode="nearest")
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        tb.ion("C", 4, np.array([1e5], dtype=np.float32), np.array([1e4], dtype=np.float32))
    assert tb.get_temp_bounds() == (10 ** 3.0, 10 ** np.max(tb.temp))


@pytest.mark.parametrize("species,ion", [("C", 4), ("Mg", 2), ("H", 1), ("N", 5)])
def test_device_restatement_matches_map_coordinates(species, ion):
    table, reds = make_table()
    tb = cloudy.CloudyTable(3., table=table, reds=reds)
    coef = tb.spline_coefficients(species, ion)
    assert coef.shape == (tb.dens.size + 24, tb.temp.size + 24)
    rng = np.random.default_rng(5)
    n = 20000
    # inside the grid and up to 5 cells outside (the reference refuses more than 0.2 dex = 1 density / 4 temperature cells)
    c0, c1 = rng.uniform(-5, tb.dens.size + 4, n), rng.uniform(-5, tb.temp.size + 4, n)
    want = map_coordinates(tb.red_table[:, :, tb.species.index(species), ion - 1], np.vstack((c0, c1)), mode="nearest")
    got = bspline_eval(tb, coef, c0, c1)
    assert np.max(np.abs(got - want)) < 1e-12


def test_table_file_roundtrip(tmp_path, monkeypatch):
    table, reds = make_table(nred=2)
    np.savez(tmp_path / "cloudy_table.npz", table=table)  # the reference's cache file: redshifts 0, 1, ...
    tb = cloudy.CloudyTable(0.5, str(tmp_path))
    assert np.allclose(tb.red_table, 0.5 * (table[0] + table[1]))
    monkeypatch.setenv("FAKE_SPECTRA_CLOUDY_DIR", str(tmp_path))
    assert cloudy.CloudyTable(1.0).red_table.shape == table.shape[1:]
    monkeypatch.delenv("FAKE_SPECTRA_CLOUDY_DIR")
    with pytest.raises(IOError):
        cloudy.CloudyTable(1.0)
