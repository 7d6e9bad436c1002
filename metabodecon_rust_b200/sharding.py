"""Multi-GPU partitioning of the hot path: one process per GPU, no data-path collective.

The reference parallelises over independent spectra (`par_deconvolute_spectra`,
deconvoluter.rs:699-710) and over independent grid points (`par_superposition_vec`,
lorentzian.rs:656-663).  Both shard without any exchange: every rank deconvolutes a contiguous
block of the spectra list (or evaluates a contiguous slice of the grid with a replicated parameter
table) on its own GPU.  `torch.distributed` is used only to bring per-rank results back to the
caller (host objects) -- the kernels never communicate.

Batch error semantics are those of `collect::<Result<Vec<_>>>` (deconvoluter.rs:655-658): the
error of the failing spectrum with the LOWEST GLOBAL INDEX wins, whichever rank owned it.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} for world size {world}")
    if n_items < 0:
        raise ValueError("n_items must be non-negative")
    return rank * n_items // world, (rank + 1) * n_items // world


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_bounds(n_items, r, world) for r in range(world))]


def _dist():
    import torch.distributed as dist
    return dist


def deconvolute_spectra_sharded(spectra: Sequence, compute: Optional[Callable] = None, deconvoluter=None,
                                dst: Optional[int] = None):
    """Deconvolute `spectra` (the same list on every rank) with the work sharded over the ranks
    of the default process group.

    compute(shard) -> list of results, or raises; defaults to `deconvoluter.par_deconvolute_spectra`
    on this rank's current CUDA device.  Returns the full, ordered result list on rank `dst`
    (on every rank when dst is None); other ranks get None.  If any spectrum fails, the exception
    of the failing spectrum with the lowest global index is raised on every rank.
    """
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(len(spectra), rank, world)
    if compute is None:
        if deconvoluter is None:
            from .deconvoluter import Deconvoluter
            deconvoluter = Deconvoluter()
        compute = deconvoluter.par_deconvolute_spectra
    shard = list(spectra[lo:hi])
    results, failure = None, None
    try:
        results = list(compute(shard)) if shard else []
        if len(results) != len(shard):
            raise RuntimeError(f"compute returned {len(results)} results for {len(shard)} spectra")
    except Exception as err:  # noqa: BLE001 -- forwarded to every rank below
        # a batch call reports its first failing spectrum; replay one by one to learn its global
        # index so that the globally first failure can be chosen across ranks
        index = lo
        for i, sp in enumerate(shard):
            try:
                compute([sp])
            except Exception as inner:  # noqa: BLE001
                index, err = lo + i, inner
                break
        failure = (index, err)
    if world == 1:
        if failure is not None:
            raise failure[1]
        return results
    gathered: List = [None] * world
    dist.all_gather_object(gathered, (results, failure))
    failures = [f for _, f in gathered if f is not None]
    if failures:
        raise min(failures, key=lambda f: f[0])[1]
    if dst is not None and rank != dst:
        return None
    out: List = []
    for part, _ in gathered:
        out.extend(part)
    return out


def superposition_vec_sharded(x, params, compute: Optional[Callable] = None, dst: Optional[int] = None):
    """`Lorentzian::par_superposition_vec` with the grid `x` split into contiguous slices over the
    ranks; the (P, 3) parameter table is replicated.  compute(x_slice, params) -> values."""
    import numpy as np
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lo, hi = shard_bounds(x.size, rank, world)
    if compute is None:
        from .lorentzian import superposition_vec_array
        compute = superposition_vec_array
    part = np.ascontiguousarray(compute(x[lo:hi], params), dtype=np.float64)
    if part.size != hi - lo:
        raise RuntimeError("compute returned a slice of the wrong length")
    if world == 1:
        return part
    gathered: List = [None] * world
    dist.all_gather_object(gathered, part)
    if dst is not None and rank != dst:
        return None
    return np.concatenate(gathered)
