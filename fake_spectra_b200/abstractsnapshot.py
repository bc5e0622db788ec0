"""Snapshot access for the host classes.

The reference's ``abstractsnapshot.AbstractSnapshotFactory(num, base, comm)`` opens Gadget/Arepo HDF5
or MP-Gadget BigFile output (abstractsnapshot.py:16-27).  Reading files is outside the hot path
(SURVEY section 8 marks the readers "next", row f3) and neither h5py nor bigfile exists in this
image, so the factory here accepts any in-memory object with the same duck-type
(``get_header_attr, get_data, get_n_segments, get_smooth_length, get_peculiar_velocity, get_temp,
get_kernel, get_npart, get_omega_baryon, get_units, get_blocklen``) — e.g.
:class:`fake_spectra_b200.synthetic.SyntheticSnapshot` — passed as ``base``; a path raises IOError
unless a reader is importable.
"""

_REQUIRED = ("get_header_attr", "get_data", "get_n_segments", "get_smooth_length", "get_peculiar_velocity",
             "get_temp", "get_kernel", "get_npart", "get_omega_baryon", "get_units")


def is_snapshot(obj):
    return all(hasattr(obj, name) for name in _REQUIRED)


def AbstractSnapshotFactory(num, base, comm=None):
    """Same call as the reference's factory; ``base`` may be a snapshot object."""
    _ = (num, comm)
    if is_snapshot(base):
        return base
    try:
        import h5py  # noqa: F401
    except ImportError as exc:
        raise IOError("cannot open snapshot %r: h5py is not available in this environment; pass an in-memory "
                      "snapshot object (e.g. fake_spectra_b200.synthetic.SyntheticSnapshot) as `base`" % (base,)) from exc
    raise IOError("HDF5 snapshot reading is not part of the B200 hot path (SURVEY 8f, row f3); pass an in-memory "
                  "snapshot object as `base`")
