"""Host side of ``Fruit.transform`` for ordinary (pageable) numpy arrays.

``cudaMemcpy`` from pageable memory is staged by the driver at a few GB/s and
blocks the calling thread, so an array a user passes in (the reference's
calling convention: plain numpy in, plain numpy out) would reach a tenth of
the throughput of pinned buffers.  This module keeps a small ring of pinned
staging buffers per process (allocated once: page-locking is slow) and copies
between user memory and the ring with several threads (numpy releases the GIL
in its copy loops), so that the copies overlap the DMA transfers and the
kernels of the neighbouring chunks.  Nothing here touches the numerics.  Like
the reference API the ring is not re-entrant: one ``transform`` at a time per
process.
"""
import os
from concurrent.futures import ThreadPoolExecutor, wait

import torch

WORKERS = max(2, min(8, os.cpu_count() or 2))
_POOL = None
_PINNED: dict = {}


def pool() -> ThreadPoolExecutor:
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=WORKERS, thread_name_prefix="fruits-b200-copy")
    return _POOL


def pinned(slot: str, shape: tuple) -> torch.Tensor:
    """Pinned float64 buffer ``shape`` for ``slot`` (grown, never shrunk)."""
    need = 1
    for s in shape:
        need *= int(s)
    buf = _PINNED.get(slot)
    if buf is None or buf.numel() < need:
        _PINNED[slot] = None           # free the old one first
        buf = torch.empty((need,), dtype=torch.float64, pin_memory=True)
        _PINNED[slot] = buf
    return buf[:need].view(shape)


def copy_rows(dst, src) -> list:
    """``dst[...] = src`` (same shape, first axis = rows) split over the copy
    threads; returns the futures."""
    n = dst.shape[0]
    step = max(1, -(-n // WORKERS))
    d, s = _as_numpy(dst), _as_numpy(src)
    return [pool().submit(_copy, d[lo:lo + step], s[lo:lo + step]) for lo in range(0, n, step)]


def _as_numpy(a):
    return a.numpy() if isinstance(a, torch.Tensor) else a


def _copy(d, s) -> None:
    import numpy as np
    np.copyto(d, s)


def finish(futures: list) -> None:
    if futures:
        wait(futures)
        for f in futures:
            f.result()             # re-raise
