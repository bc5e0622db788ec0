"""Atomic line data for the transitions the hot path is benchmarked and tested on.

The reference parses VPFIT's atom.dat (line_data.py:59-81) into ``LineData[(elem, ion)][int(lambda)]``
objects with ``lambda_X`` (Angstrom), ``fosc_X`` and ``gamma_X`` (1/s).  This module keeps the same
lookup interface.  ``LineData()`` loads the shipped table of every transition of the nine species (data/
lines_9species.dat: the same 171 lines of 28 ions the reference holds, so that get_observer_tau chooses among the same
lines); a compact built-in table of the common IGM/CGM lines stands in when the data file is missing, and
:func:`read_vpfit` loads a user-supplied atom.dat.
"""
import os
import re

MASSES = {'H': 1.00794, 'He': 4.002602, 'C': 12.011, 'N': 14.00674, 'O': 15.9994, 'Ne': 20.18,
          'Mg': 24.3050, 'Si': 28.0855, 'Fe': 55.847}   # amu, as reference line_data.py:21

# (element, ion stage) -> [(lambda / Angstrom, f_osc, Gamma / s^-1), ...]
_BUILTIN = {
    ('H', 1): [(1215.6701, 0.416400, 6.265e8), (1025.7223, 0.079120, 1.897e8), (972.5368, 0.029000, 8.127e7),
               (949.7431, 0.013940, 4.204e7), (937.8035, 0.007799, 2.450e7)],
    ('He', 2): [(303.7822, 0.416, 6.270e8), (256.317, 0.0790, 1.897e8)],
    ('C', 2): [(1334.5323, 0.127800, 2.880e8), (1036.3367, 0.118000, 2.200e9)],
    ('C', 3): [(977.0201, 0.757000, 1.760e9)],
    ('C', 4): [(1548.2049, 0.189900, 2.642e8), (1550.77845, 0.094750, 2.628e8)],
    ('N', 5): [(1238.821, 0.1560, 3.391e8), (1242.804, 0.0770, 3.356e8)],
    ('O', 1): [(1302.1685, 0.048000, 5.650e8), (1039.2304, 0.00907, 1.87e8)],
    ('O', 6): [(1031.9261, 0.13250, 4.149e8), (1037.6167, 0.06580, 4.076e8)],
    ('Ne', 8): [(780.324, 0.050500, 1.000e8), (770.409, 0.103000, 1.000e8)],
    ('Mg', 1): [(2852.96328, 1.830000, 5.000e8)],
    ('Mg', 2): [(2796.3542699, 0.6155, 2.68e8), (2803.5314853, 0.3058, 2.66e8)],
    ('Si', 2): [(1526.70698, 0.133, 1.13e9), (1260.4221, 1.18, 2.95e9), (1193.2897, 0.582, 4.07e9),
                (1190.4158, 0.292, 4.08e9)],
    ('Si', 3): [(1206.500, 1.63, 2.48e9)],
    ('Si', 4): [(1393.76018, 0.513, 8.80e8), (1402.77291, 0.254, 8.62e8)],
    ('Fe', 2): [(2382.7641781, 0.320, 3.13e8), (2600.1724835, 0.2394, 2.70e8), (2344.2129601, 0.1142, 2.68e8),
                (1608.45085, 0.0577, 2.74e8)],
}

_ROMAN = {'I': 1, 'V': 5, 'X': 10}


class Line:
    """One transition: lambda_X (Angstrom), fosc_X, gamma_X (1/s).  Same attribute names as the
    reference (line_data.py:48-57) because ``Spectra._do_interpolation_work`` reads them."""

    def __init__(self, lambda_X, fosc_X, gamma_X):
        self.lambda_X = lambda_X
        self.fosc_X = fosc_X
        self.gamma_X = gamma_X

    def __repr__(self):
        return "Line(%g A, f=%g, Gamma=%g)" % (self.lambda_X, self.fosc_X, self.gamma_X)


def _roman(s):
    total = 0
    vals = [_ROMAN[c] for c in s]
    for i, v in enumerate(vals):
        total += -v if i + 1 < len(vals) and vals[i + 1] > v else v
    return total


TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lines_9species.dat")


def read_table(path):
    """The shipped table (data/lines_9species.dat): ``element ion lambda f_osc Gamma`` per row, every transition of the nine
    species that the reference's LineData holds (171 lines of 28 ions), keyed like it by the integer wavelength."""
    lines = {}
    with open(path) as fh:
        for raw in fh:
            tok = raw.split()
            if not tok or tok[0].startswith("#"):
                continue
            lam, fosc, gam = float(tok[2]), float(tok[3]), float(tok[4])
            lines.setdefault((tok[0], int(tok[1])), {})[int(lam)] = Line(lam, fosc, gam)
    return lines


def read_vpfit(path, species=tuple(MASSES)):
    """Parse a VPFIT atom.dat: species = leading letters, ion = roman numeral, then the first three
    floats are lambda, f, Gamma."""
    lines = {}
    with open(path) as fh:
        for raw in fh:
            m = re.match(r"([A-Z]\s*[a-z]?)([IVX]+)[\s*]", raw)
            if m is None:
                continue
            elem = re.sub(r"\s", "", m.group(1))
            if elem not in species:
                continue
            vals = []
            for tok in raw[m.end():].split():
                try:
                    vals.append(float(tok))
                except ValueError:
                    continue
                if len(vals) == 3:
                    break
            if len(vals) < 3:
                continue
            lines.setdefault((elem, _roman(m.group(2))), {})[int(vals[0])] = Line(*vals)
    return lines


class LineData:
    """``LineData()[(elem, ion)][int(lambda)] -> Line`` and ``get_mass(elem)``."""

    def __init__(self, vpdat=None):
        self.species = tuple(MASSES)
        self.masses = dict(MASSES)
        if vpdat is None:
            self.lines = read_table(TABLE) if os.path.exists(TABLE) else \
                {k: {int(l[0]): Line(*l) for l in v} for k, v in _BUILTIN.items()}
        else:
            self.lines = read_vpfit(vpdat, self.species)

    def __getitem__(self, specion):
        return self.lines[specion]

    def __len__(self):
        return len(self.lines)

    def get_mass(self, specie):
        return self.masses[specie]
