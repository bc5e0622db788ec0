"""Multi-GPU partitioning of the interpolation (one process per GPU, torch.distributed).

Two modes:

* ``"sightlines"`` (default): every rank holds the whole particle set and computes a contiguous
  block of sightlines; blocks are disjoint rows of the result, so the data path needs NO
  collective — one all-gather at the end hands every rank the full array (what the reference's
  Allreduce leaves behind, spectra.py:828-830).
* ``"particles"``: every rank holds a slice of the particles and computes ALL sightlines for it;
  the partial [NumLos, nbins] arrays are summed with one all-reduce — the reference's MPI mode
  (abstractsnapshot.py:195-207, spectra.py:825-831), but in float64 instead of its float32.

The collectives run on whatever device the tensors live on: NCCL over NVLink for CUDA tensors, gloo
for the CPU tests of this host logic.
"""
import numpy as np
import torch
import torch.distributed as dist


def even_blocks(n, world):
    """Edges [world+1] of contiguous blocks of near-equal size."""
    return np.linspace(0, n, world + 1).astype(np.int64)


def balanced_blocks(weights, world):
    """Edges [world+1] of contiguous blocks with near-equal total weight (e.g. candidate pairs per
    sightline): block r ends at the first index whose running weight reaches (r+1)/world of the total."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    if n == 0 or w.sum() <= 0:
        return even_blocks(n, world)
    csum = np.cumsum(w)
    targets = csum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(csum, targets, side="left") + 1
    edges = np.concatenate([[0], np.minimum(inner, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(edges)


class Sharder:
    """Partition + recombination for one Spectra object.

    group : a torch.distributed process group, or None for the default group when
            torch.distributed is initialised (world of 1 otherwise).
    mode  : "sightlines" or "particles".
    """

    def __init__(self, mode="sightlines", group=None):
        if mode not in ("sightlines", "particles"):
            raise ValueError("shard must be 'sightlines' or 'particles', not %r" % (mode,))
        self.mode = mode
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.size = dist.get_world_size(group)
        else:
            self.rank, self.size = 0, 1
        self.edges = None

    # -- partition -----------------------------------------------------------------------------------
    def set_sightlines(self, numlos, weights=None):
        """Fix the sightline blocks (only used in "sightlines" mode)."""
        if weights is not None:
            self.edges = balanced_blocks(weights, self.size)
        else:
            self.edges = even_blocks(numlos, self.size)
        return self.edges

    def my_sightlines(self, numlos):
        """slice of the sightlines this rank computes."""
        if self.mode != "sightlines" or self.size == 1:
            return slice(0, numlos)
        if self.edges is None or self.edges[-1] != numlos:
            self.set_sightlines(numlos)
        return slice(int(self.edges[self.rank]), int(self.edges[self.rank + 1]))

    def my_particles(self, npart):
        """slice of an (already filtered) particle list this rank interpolates."""
        if self.mode != "particles" or self.size == 1:
            return slice(0, npart)
        e = even_blocks(npart, self.size)
        return slice(int(e[self.rank]), int(e[self.rank + 1]))

    def reduce_blocks(self, numlos, nblocks=4):
        """Particle-sharded mode: sightline blocks [(begin, end), ...] for computing and summing block by block: the
        all-reduce of one block's rows runs while the next block is being computed ("issued as blocks
        complete")."""
        e = even_blocks(numlos, max(1, min(int(nblocks), max(numlos, 1))))
        return [(int(e[i]), int(e[i + 1])) for i in range(len(e) - 1) if e[i + 1] > e[i]]

    def sum_block_async(self, full, begin, end):
        """Starts the FP64 sum over the ranks of rows [begin, end) of ``full`` ([K, numlos, ...] or [numlos, ...], K == 1
        for an in-place contiguous block) on the collective's own stream; returns a handle to wait() on."""
        if self.size == 1:
            return None
        blk = full[:, begin:end] if full.dim() == 3 else full[begin:end]
        if not blk.is_contiguous():
            raise ValueError("block sums need one line per array (rows of a block must be contiguous)")
        return dist.all_reduce(blk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    # -- recombination -------------------------------------------------------------------------------
    def combine(self, local, numlos, dim=0):
        """local: this rank's result (torch tensor, any device); dimension ``dim`` (0 or 1) runs over its
        sightline block ("sightlines") or over all sightlines ("particles").  Returns the full tensor (``numlos``
        along ``dim``) on every rank.

        Sightline mode moves every block exactly once: a block of rows of a row-major array is contiguous, so
        rank r's block is received straight into its place in the full array (one all-gather when the blocks are
        equal, one broadcast per block otherwise: no padding, no staging copies)."""
        if self.size == 1:
            return local
        if self.mode == "particles":
            local = local.contiguous()
            dist.all_reduce(local, op=dist.ReduceOp.SUM, group=self.group)
            return local
        if dim not in (0, 1):
            raise ValueError("dim must be 0 or 1")
        if self.edges is None or self.edges[-1] != numlos:
            self.set_sightlines(numlos)
        local = local.contiguous()
        lead = (1,) if dim == 0 else (local.shape[0],)
        tail = tuple(local.shape[dim + 1:])
        loc = local.view(lead + (local.shape[dim],) + tail)
        full = torch.empty(lead + (numlos,) + tail, dtype=local.dtype, device=local.device)
        sizes = np.diff(self.edges)
        for k in range(lead[0]):
            if np.all(sizes == sizes[0]):
                dist.all_gather_into_tensor(full[k], loc[k], group=self.group)
            else:
                for r in range(self.size):
                    blk = full[k, int(self.edges[r]):int(self.edges[r + 1])]
                    if r == self.rank:
                        blk.copy_(loc[k])
                    if blk.numel():
                        dist.broadcast(blk, src=dist.get_global_rank(self.group, r) if self.group is not None else r,
                                       group=self.group)
        return full[0] if dim == 0 else full
