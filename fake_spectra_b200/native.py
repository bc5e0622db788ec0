"""Device-resident interface to libfsb200.so: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is used only for device buffers and streams; all compute is in the library's own
sm_100a kernels.  This layer lets callers keep particles and results resident in HBM and reuse
one candidate index for several quantities (tau of several lines, column density, weighted
fields), which the one-shot boundary in :mod:`_spectra_priv` cannot.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _dptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA tensor of dtype %s" % (name, dtype))
    return t


class CandidateIndex:
    """Per-sightline candidate particle lists (replaces IndexTable + get_near_particles)."""

    def __init__(self, box, cofm, axis, pos, h, counts=None):
        """counts: optional int32 CUDA tensor [nlos] of list sizes from :func:`count_pairs` for these sightlines and
        these particles (skips the counting pass)."""
        self.lib = _lib.load()
        self.box = float(box)
        _need(cofm, torch.float64, "cofm"), _need(axis, torch.int32, "axis")
        _need(pos, torch.float32, "pos"), _need(h, torch.float32, "h")
        self.nlos = cofm.shape[0]
        self.npart = pos.shape[0]
        self.device = pos.device
        handle = C.c_void_p()
        if counts is not None:
            _need(counts, torch.int32, "counts")
            if counts.shape[0] != self.nlos:
                raise ValueError("counts must have one entry per sightline")
        with torch.cuda.device(self.device):
            rc = self.lib.fsb_index_build_counted(self.box, _dptr(cofm), _dptr(axis), self.nlos, _dptr(pos), _dptr(h),
                                                  self.npart, _dptr(counts), _stream(), C.byref(handle))
        _lib.check(rc, "fsb_index_build")
        self.handle = handle
        nlos, npairs, mx = C.c_int32(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.fsb_index_sizes(self.handle, C.byref(nlos), C.byref(npairs), C.byref(mx)), "fsb_index_sizes")
        self.npairs, self.max_list = npairs.value, mx.value

    def export(self):
        """(offsets int64[nlos+1], particle int32[npairs], dr2 float64[npairs]) as CUDA tensors."""
        off = torch.empty(self.nlos + 1, dtype=torch.int64, device=self.device)
        part = torch.empty(max(self.npairs, 1), dtype=torch.int32, device=self.device)
        dr2 = torch.empty(max(self.npairs, 1), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsb_index_export(self.handle, _dptr(off), _dptr(part), _dptr(dr2), _stream()),
                       "fsb_index_export")
        return off, part[:self.npairs], dr2[:self.npairs]

    def free(self):
        if getattr(self, "handle", None):
            with torch.cuda.device(self.device):
                self.lib.fsb_index_free(self.handle, _stream())
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- accumulation --------------------------------------------------------------------------
    def compute_tau(self, params, pos, vel, dens, temp, h, out=None, counters=None, push=None, lines=None):
        """params: one _lib.Params or a list of them (fused lines of one ion).
        Returns float64 [nlos, nbins] (or [nlines, nlos, nbins]); accumulates into ``out`` if given.
        push: a _lib.Push (see PeerRows.push_spec): finished rows are also stored into the listed full arrays
        (this rank's and its peers') from inside the kernel.
        lines: (begin, end): only these sightlines of the index are processed (rows outside stay untouched)."""
        plist = params if isinstance(params, (list, tuple)) else [params]
        nbins = plist[0].nbins
        shape = (len(plist), self.nlos, nbins)
        if out is None:
            out = torch.zeros(shape, dtype=torch.float64, device=self.device)
        arr = (_lib.Params * len(plist))(*plist)
        with torch.cuda.device(self.device):
            if lines is not None:
                if counters is not None or push is not None:
                    raise ValueError("a sightline range excludes counters and push")
                rc = self.lib.fsb_compute_tau_multi_range(self.handle, arr, len(plist), int(lines[0]), int(lines[1]), _dptr(pos),
                                                          _dptr(vel), _dptr(dens), _dptr(temp), _dptr(h), _dptr(out), _stream())
            elif push is not None:
                if counters is not None:
                    raise ValueError("counters and push are exclusive")
                rc = self.lib.fsb_compute_tau_multi_push(self.handle, arr, len(plist), _dptr(pos), _dptr(vel), _dptr(dens),
                                                         _dptr(temp), _dptr(h), _dptr(out), C.byref(push), _stream())
            else:
                rc = self.lib.fsb_compute_tau_multi(self.handle, arr, len(plist), _dptr(pos), _dptr(vel), _dptr(dens),
                                                    _dptr(temp), _dptr(h), _dptr(out), _dptr(counters), _stream())
        _lib.check(rc, "fsb_compute_tau")
        return out.view(shape) if isinstance(params, (list, tuple)) else out.view(self.nlos, nbins)

    def compute_colden(self, params, pos, dens, h, out=None, counters=None):
        """dens: float32 [npart] or [nweights, npart].  Returns float64 [nlos, nbins] or
        [nweights, nlos, nbins]."""
        nw = 1 if dens.dim() == 1 else dens.shape[0]
        shape = (nw, self.nlos, params.nbins)
        if out is None:
            out = torch.zeros(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.fsb_compute_colden(self.handle, C.byref(params), _dptr(pos), _dptr(dens), nw, _dptr(h),
                                             _dptr(out), _dptr(counters), _stream())
        _lib.check(rc, "fsb_compute_colden")
        return out.view(shape) if dens.dim() == 2 else out.view(self.nlos, params.nbins)

    def assign_cells(self, cofm, axis, pos):
        """Voronoi cell extents, float32 [npairs, 2] (replaces IndexTable::assign_cells)."""
        cells = torch.empty((max(self.npairs, 1), 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.fsb_assign_cells(self.handle, self.box, _dptr(cofm), _dptr(axis), _dptr(pos), _dptr(cells),
                                           _stream())
        _lib.check(rc, "fsb_assign_cells")
        return cells[:self.npairs]


MAX_PAIRS_PER_INDEX = 1 << 30  # one candidate index holds fewer than 2^31 pairs (32-bit list positions); half as a margin


def sightline_blocks(counts, max_pairs=MAX_PAIRS_PER_INDEX):
    """Contiguous sightline blocks [(begin, end), ...] whose candidate-pair totals stay at or below ``max_pairs``
    (``counts``: pairs per sightline, a host array; a single sightline above the bound gets a block of its own)."""
    counts = np.asarray(counts, dtype=np.int64)
    edges, begin, total = [], 0, 0
    for i, c in enumerate(counts):
        if i > begin and total + c > max_pairs:
            edges.append((begin, i))
            begin, total = i, 0
        total += int(c)
    edges.append((begin, len(counts)))
    return edges


class BlockedIndex:
    """A candidate index for any number of sightlines: one CandidateIndex when the pairs fit (fewer than 2^31),
    otherwise one per contiguous sightline block found from a count pass (fsb_count_pairs), built and used one after the
    other so that only one block's lists live in HBM at a time.  Same accumulation calls as CandidateIndex."""

    def __init__(self, box, cofm, axis, pos, h, max_pairs=MAX_PAIRS_PER_INDEX):
        self.box, self.cofm, self.axis, self.pos, self.h = float(box), cofm, axis, pos, h
        self.nlos, self.device = cofm.shape[0], pos.device
        counts = count_pairs(box, pos, h, axis, cofm)
        host = counts.cpu().numpy()
        self.npairs = int(host.astype(np.int64).sum())
        self.blocks = sightline_blocks(host, max_pairs)
        self._counts = counts
        self._single = None
        if len(self.blocks) == 1:
            self._single = CandidateIndex(box, cofm, axis, pos, h, counts=counts)

    def _each(self):
        if self._single is not None:
            yield 0, self.nlos, self._single
            return
        for b0, b1 in self.blocks:
            idx = CandidateIndex(self.box, self.cofm[b0:b1].contiguous(), self.axis[b0:b1].contiguous(), self.pos, self.h,
                                 counts=self._counts[b0:b1].contiguous())
            try:
                yield b0, b1, idx
            finally:
                idx.free()

    def compute_tau(self, params, pos, vel, dens, temp, h, out=None, counters=None, push=None, lines=None):
        if self._single is not None:
            return self._single.compute_tau(params, pos, vel, dens, temp, h, out=out, counters=counters, push=push, lines=lines)
        if push is not None or lines is not None or counters is not None:
            raise ValueError("push, lines and counters need a single index (fewer than 2^31 candidate pairs)")
        plist = params if isinstance(params, (list, tuple)) else [params]
        shape = (len(plist), self.nlos, plist[0].nbins)
        full = out.view(shape) if out is not None else torch.zeros(shape, dtype=torch.float64, device=self.device)
        for b0, b1, idx in self._each():
            if len(plist) == 1:
                idx.compute_tau(plist, pos, vel, dens, temp, h, out=full[:, b0:b1])  # one line: the block's rows are contiguous
            else:
                full[:, b0:b1] += idx.compute_tau(plist, pos, vel, dens, temp, h)
        return full if isinstance(params, (list, tuple)) else full.view(self.nlos, plist[0].nbins)

    def compute_colden(self, params, pos, dens, h, out=None, counters=None):
        if self._single is not None:
            return self._single.compute_colden(params, pos, dens, h, out=out, counters=counters)
        nw = 1 if dens.dim() == 1 else dens.shape[0]
        shape = (nw, self.nlos, params.nbins)
        full = out.view(shape) if out is not None else torch.zeros(shape, dtype=torch.float64, device=self.device)
        for b0, b1, idx in self._each():
            if nw == 1:
                idx.compute_colden(params, pos, dens, h, out=full[:, b0:b1])
            else:
                full[:, b0:b1] += idx.compute_colden(params, pos, dens, h).view(nw, b1 - b0, params.nbins)
        return full if dens.dim() == 2 else full.view(self.nlos, params.nbins)

    def free(self):
        if self._single is not None:
            self._single.free()
            self._single = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier: lets torch wrap device memory this library allocated."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerRows:
    """A full result array [K, numlos, nbins] (float64) on every rank of a sightline-sharded job, each rank's copy
    mapped into all the others (CUDA IPC over NVLink / NVSwitch).  The tau kernel of a rank stores every row it
    finishes into all copies (fsb_compute_tau_multi_push), so after a barrier every rank holds the complete array
    without a gather.  One process per GPU on ONE node; `group` as in torch.distributed."""

    def __init__(self, K, numlos, nbins, group=None):
        import torch.distributed as dist
        self.lib = _lib.load()
        self.K, self.numlos, self.nbins = int(K), int(numlos), int(nbins)
        self.group = group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        if self.size > _lib.MAX_PEERS:
            raise ValueError("at most %d ranks" % _lib.MAX_PEERS)
        self.device = torch.device("cuda", torch.cuda.current_device())
        nbytes = 8 * self.K * self.numlos * self.nbins
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(self.lib.fsb_peer_alloc(nbytes, C.byref(ptr), handle), "fsb_peer_alloc")
        self.ptr = ptr.value
        handles = [None] * self.size
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_ptr = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                self.peer_ptr.append(self.ptr)
                continue
            q = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(hb)
            _lib.check(self.lib.fsb_peer_open(buf, C.byref(q)), "fsb_peer_open")
            self.peer_ptr.append(q.value)
        self._carriers = [_RawCuda(q, (self.K, self.numlos, self.nbins)) for q in self.peer_ptr]
        self.views = [torch.as_tensor(c, device=self.device) for c in self._carriers]  # every rank's array, as seen from here
        self.full = self.views[self.rank]

    def zero_block(self, first_sightline, nrows):
        """Zero this rank's block of rows in every copy (a rank that launches no kernel for its block)."""
        for v in self.views:
            v[:, first_sightline:first_sightline + nrows].zero_()

    def push_spec(self, first_line, first_sightline):
        """Destinations for a compute_tau call whose first line is `first_line` of K and whose sightline block starts
        at `first_sightline` of numlos."""
        sp = _lib.Push()
        sp.npeers = self.size
        sp.line_stride = self.numlos * self.nbins
        off = 8 * ((int(first_line) * self.numlos + int(first_sightline)) * self.nbins)
        for r in range(self.size):
            sp.dest[r] = self.peer_ptr[r] + off
        return sp

    def barrier(self):
        """All ranks' kernels have finished (stream-ordered on each rank) => every copy is complete."""
        import torch.distributed as dist
        dist.barrier(group=self.group)

    def close(self):
        if getattr(self, "ptr", None) is None:
            return
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody is still writing into anybody's array
        for r, q in enumerate(self.peer_ptr):
            if r != self.rank:
                self.lib.fsb_peer_close(C.c_void_p(q))
        dist.barrier(group=self.group)
        self.full = None
        self.views = []
        self.lib.fsb_peer_free(C.c_void_p(self.ptr))
        self.ptr = None


def particle_interpolate(compute_tau, params, pos, vel, dens, temp, h, axis, cofm, out=None):
    """One-shot device call (index build + accumulation), device tensors in and out."""
    lib = _lib.load()
    nlos = cofm.shape[0]
    if out is None:
        out = torch.zeros((nlos, params.nbins), dtype=torch.float64, device=pos.device)
    with torch.cuda.device(pos.device):
        rc = lib.fsb_particle_interpolate(1 if compute_tau else 0, C.byref(params), _dptr(pos), _dptr(vel), _dptr(dens),
                                          _dptr(temp), _dptr(h), pos.shape[0], _dptr(axis), _dptr(cofm), nlos,
                                          _dptr(out), _stream())
    _lib.check(rc, "fsb_particle_interpolate")
    return out


def near_lines(box, pos, h, axis, cofm):
    """Ascending int32 indices (CUDA tensor) of particles near at least one sightline."""
    lib = _lib.load()
    out = torch.empty(max(pos.shape[0], 1), dtype=torch.int32, device=pos.device)
    count = C.c_int64(0)
    with torch.cuda.device(pos.device):
        rc = lib.fsb_near_lines(float(box), _dptr(pos), _dptr(h), pos.shape[0], _dptr(axis), _dptr(cofm), cofm.shape[0],
                                _dptr(out), C.byref(count), _stream())
    _lib.check(rc, "fsb_near_lines")
    return out[:count.value]


def count_pairs(box, pos, h, axis, cofm):
    """Candidate particles per sightline (int32 CUDA tensor [nlos]) without building the lists: the count pass
    that balances sightline blocks across GPUs (sharding.balanced_blocks)."""
    lib = _lib.load()
    counts = torch.empty(max(cofm.shape[0], 1), dtype=torch.int32, device=pos.device)
    with torch.cuda.device(pos.device):
        rc = lib.fsb_count_pairs(float(box), _dptr(pos), _dptr(h), pos.shape[0], _dptr(axis), _dptr(cofm),
                                 cofm.shape[0], _dptr(counts), _stream())
    _lib.check(rc, "fsb_count_pairs")
    return counts[:cofm.shape[0]]


def prepare_particles(cfg, index, position, velocity, density, ienergy, nelec, nh0, smoothing, mass_frac=None,
                      want_vel=True, want_temp=True, ion_table=None):
    """Snapshot fields of one segment (CUDA float32 tensors) -> (pos, vel, elem_den, temp, hh) of the particles in
    ``index`` (int32 CUDA tensor, or None for all): fsb_prepare_particles.  ``smoothing``: support radii of all
    particles (smoothing_lengths); ``mass_frac``: a 1-D view of one column of
    the metal table (any stride) or None; ``ion_table``: an ``_lib.IonTable`` (metal ions) or None."""
    lib = _lib.load()
    dev = position.device
    m = int(index.shape[0]) if index is not None else int(position.shape[0])
    f32 = dict(dtype=torch.float32, device=dev)
    pos, hh, elem_den = torch.empty((m, 3), **f32), torch.empty(m, **f32), torch.empty(m, **f32)
    vel = torch.empty((m, 3), **f32) if want_vel else None
    temp = torch.empty(m, **f32) if want_temp else None
    stride = _mass_frac_stride(mass_frac)
    with torch.cuda.device(dev):
        rc = lib.fsb_prepare_particles(C.byref(cfg), _dptr(index), m, _dptr(position), _dptr(velocity), _dptr(density),
                                       _dptr(ienergy), _dptr(nelec), _dptr(nh0), _dptr(smoothing), _dptr(mass_frac), stride,
                                       C.byref(ion_table) if ion_table is not None else None,
                                       _dptr(pos), _dptr(vel), _dptr(elem_den), _dptr(temp), _dptr(hh), _stream())
    _lib.check(rc, "fsb_prepare_particles")
    return pos, vel, elem_den, temp, hh


def smoothing_lengths(a, b=None, mode=0):
    """Kernel support radii from SmoothingLength (mode 0), Volume (1) or Masses / Density (2): fsb_smoothing_lengths."""
    lib = _lib.load()
    hh = torch.empty_like(a)
    with torch.cuda.device(a.device):
        rc = lib.fsb_smoothing_lengths(_dptr(a), _dptr(b), int(a.shape[0]), int(mode), _dptr(hh), _stream())
    _lib.check(rc, "fsb_smoothing_lengths")
    return hh


def _mass_frac_stride(mass_frac):
    if mass_frac is None:
        return 0
    if mass_frac.dtype != torch.float32 or mass_frac.dim() != 1:
        raise TypeError("mass_frac must be a 1-D float32 view")
    return int(mass_frac.stride(0))


def select_particles(cfg, index, density, mass_frac=None):
    """The entries of ``index`` (int32 CUDA tensor; None = all particles) with mass in the element, in order:
    fsb_prepare_select (the reference's _filter_particles, spectra.py:600)."""
    lib = _lib.load()
    dev = density.device
    m = int(index.shape[0]) if index is not None else int(density.shape[0])
    out = torch.empty(m, dtype=torch.int32, device=dev)
    count = C.c_int64(0)
    with torch.cuda.device(dev):
        rc = lib.fsb_prepare_select(C.byref(cfg), _dptr(index), m, _dptr(density), _dptr(mass_frac), _mass_frac_stride(mass_frac),
                                    _dptr(out), C.byref(count), _stream())
    _lib.check(rc, "fsb_prepare_select")
    return out[:count.value]


def voigt_profile(x, y, voigt=_lib.VOIGT_FAST):
    """Re w(x + i y) on the device (test hook)."""
    lib = _lib.load()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.fsb_voigt_profile(_dptr(x), _dptr(y), _dptr(out), x.numel(), int(voigt), _stream())
    _lib.check(rc, "fsb_voigt_profile")
    return out


def device_info():
    lib = _lib.load()
    v = [C.c_int32() for _ in range(4)]
    _lib.check(lib.fsb_device_info(*[C.byref(x) for x in v]), "fsb_device_info")
    return {"sm_count": v[0].value, "clock_khz": v[1].value, "cc": (v[2].value, v[3].value)}


def measure_fma_peak(fp64=True):
    """Measured FMA throughput (TFLOP/s) of the current device; the tau kernel's roofline peak."""
    lib = _lib.load()
    v = C.c_double(0)
    _lib.check(lib.fsb_measure_fma_peak(1 if fp64 else 0, C.byref(v), _stream()), "fsb_measure_fma_peak")
    return v.value
