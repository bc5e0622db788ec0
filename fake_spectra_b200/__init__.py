"""fake_spectra_b200 — B200-native sightline interpolation (the hot path of sbird/fake_spectra).

Host classes keep the reference API (Spectra / RandSpectra / GriddedSpectra, get_tau,
get_col_density, ...); the native work runs in hand-written sm_100a CUDA kernels behind the C-ABI
declared in include/fsb200.h.  There is no CPU fallback: calls raise if the CUDA library or a
device is missing.
"""
__version__ = "0.1.0"
