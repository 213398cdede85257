"""Series-sharded execution on several GPUs of one box.

The hot path is embarrassingly parallel over series (every numba kernel of
the reference is a ``prange`` over axis 0, e.g. fruits/iss/semiring.py:184),
so the multi-GPU layout is: one process per GPU (``torch.distributed``, NCCL),
rank ``r`` owns the contiguous rows ``[r*S, (r+1)*S)`` of the batch, the plan
and the fitted thresholds (a few kB) are replicated, and the only exchange of
``transform`` is the assembly of the ``[world*S, F]`` feature matrix on every
rank.  ``PeerGather`` fuses it into the feature kernels: their epilogue stores
with ``multimem.st`` through the NVSwitch multicast mapping of a symmetric
allocation; the fallbacks push row chunks with the copy engines or all-gather
them with NCCL on a side stream.

Fit (``fit_sharded``): every rank draws the same sample positions with the
global numpy RNG (same draws as the reference, fruits/fruit.py:430-438) and
keeps only the sampled rows it owns.  A threshold is a quantile over the WHOLE
sample of one iterated sum (fruits/sieving/segment.py:66-75): the radix
selections run in phases over the local rows and all-reduce their histograms
(``RowShard``), so all ranks end with the same thresholds and nothing but
histograms crosses NVLink.  ``shard="nodes"`` is the alternative for slices
that cannot be fitted row by row: the sample is gathered (bit-exactly) and the
iterated sums are split over the ranks instead.
"""
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(n: int, world: int, rank: int):
    """Contiguous row block ``[lo, hi)`` of rank ``rank`` (first ranks get the
    remainder rows)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_cost(n: int, costs, world: int, rank: int):
    """Contiguous block ``[lo, hi)`` of ``n`` items for ``rank`` such that the
    blocks carry about the same total cost (``costs[i]`` per item; None = equal
    costs, i.e. ``shard_rows``)."""
    if costs is None:
        return shard_rows(n, world, rank)
    csum = np.concatenate([[0.0], np.cumsum(np.asarray(costs, dtype=np.float64))])
    bounds = [int(np.searchsorted(csum, csum[-1] * r / world, side="left")) for r in range(world)]
    bounds.append(n)
    bounds = [min(max(b, 0), n) for b in bounds]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds[rank], bounds[rank + 1]


def sync_numpy_rng(group=None) -> None:
    """Give every rank rank 0's global numpy RNG state, so that all ranks draw
    the same fit sample and the same PPV subsamples (the reference consumes
    ``np.random`` in fit: fruits/fruit.py:434-437, fruits/sieving/implicit.py:104)."""
    state = [np.random.get_state() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(state, 0, group=group)
    np.random.set_state(state[0])


def gather_fit_sample(X_local: torch.Tensor, n_total: int, fit_sample_size, group=None):
    """Rows of the global batch that ``FruitSlice._select_fit_sample`` draws
    (fruits/fruit.py:430-438), assembled on every rank.  ``X_local`` holds the
    rows ``shard_rows(n_total, world, rank)``; the ranks must share one RNG
    state (``sync_numpy_rng``), every rank makes the same draw."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if isinstance(fit_sample_size, int) and fit_sample_size == 1:
        idx = np.array([np.random.randint(0, n_total)], dtype=np.int64)
    else:
        s = max(int(fit_sample_size * n_total), 1)
        idx = np.random.choice(n_total, size=s, replace=False).astype(np.int64)
    idx_t = torch.from_numpy(idx).to(X_local.device)
    lo, hi = shard_rows(n_total, world, rank)
    mine = (idx_t >= lo) & (idx_t < hi)
    sample = torch.zeros((idx_t.numel(),) + tuple(X_local.shape[1:]), dtype=X_local.dtype,
                         device=X_local.device)
    sample[mine] = X_local.index_select(0, idx_t[mine] - lo)
    # every sampled row is owned by exactly one rank and the others hold the bit
    # pattern 0: an INTEGER sum of the bit patterns assembles the rows exactly
    # (a floating point sum would turn -0.0 into +0.0 and drop NaN payloads, and
    # 1 / -0.0 = -inf in a letter with a negative exponent)
    assert sample.dtype == torch.float64
    dist.all_reduce(sample.view(torch.int64), op=dist.ReduceOp.SUM, group=group)
    return sample


def exchange_rows(T_local: torch.Tensor, n_total: int, need, group=None) -> torch.Tensor:
    """Rows ``need`` (global row numbers, any order, repeats allowed) of a
    row-sharded tensor: rank r holds the rows ``shard_rows(n_total, world, r)``
    in ``T_local``.  One all-to-all of the requests, one of the rows."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = T_local.device
    need = torch.as_tensor(np.asarray(need, dtype=np.int64), device=dev)
    bounds = [shard_rows(n_total, world, r) for r in range(world)]
    starts = torch.tensor([b[0] for b in bounds] + [n_total], dtype=torch.int64, device=dev)
    owner = torch.searchsorted(starts, need, right=True) - 1
    order = torch.argsort(owner, stable=True)
    req = need[order].contiguous()
    send_counts = torch.bincount(owner, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    asked = torch.empty((sum(rc),), dtype=torch.int64, device=dev)
    dist.all_to_all_single(asked, req, output_split_sizes=rc, input_split_sizes=sc, group=group)
    lo = bounds[rank][0]
    rows = T_local.index_select(0, asked - lo).contiguous()
    got = torch.empty((sum(sc),) + tuple(T_local.shape[1:]), dtype=T_local.dtype, device=dev)
    dist.all_to_all_single(got, rows, output_split_sizes=sc, input_split_sizes=rc, group=group)
    out = torch.empty_like(got)
    out[order] = got
    return out


class RowShard:
    """Fit on a row-sharded sample: every rank keeps only the
    sample rows it owns; a threshold -- a quantile over the WHOLE sample of one
    iterated sum (fruits/sieving/segment.py:66-75) -- comes from a radix select
    whose histograms are summed over the ranks between the passes
    (``fb_order_stats_dist``).  All ranks end with the same thresholds; nothing
    but histograms (a few kB per iterated sum) crosses NVLink."""

    def __init__(self, positions: np.ndarray, n_sample: int, group=None) -> None:
        self.positions = np.asarray(positions, dtype=np.int64)     # sample positions held here
        self.n_sample = int(n_sample)                              # rows of the whole sample
        self.group = group
        self.world = dist.get_world_size(group)
        self.n_local = len(self.positions)                         # (0: placeholder row only)
        self._n_max = None

    def n_local_max(self, dev) -> int:
        """Largest local row count over the ranks: chunk sizes are derived from
        it, so every rank runs the same number of chunks (and collectives)."""
        if self._n_max is None:
            t = torch.tensor([len(self.positions)], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self._n_max = max(int(t.item()), 1)
        return self._n_max

    def local_rows_of(self, sel: np.ndarray) -> np.ndarray:
        """Local row numbers of the sample positions ``sel`` held by this rank."""
        idx = np.searchsorted(self.positions, sel)
        idx = np.clip(idx, 0, max(len(self.positions) - 1, 0))
        hit = (len(self.positions) > 0) & (self.positions[idx] == sel) if len(self.positions) else \
            np.zeros(len(sel), dtype=bool)
        return idx[hit]

    # -- distributed order statistics -------------------------------------------
    def _reduce(self, work: torch.Tensor, off: int, count: int, dtype, op) -> None:
        width = 4 if dtype == torch.int32 else 8
        region = work[off:off + count * width].view(dtype)
        dist.all_reduce(region, op=op, group=self.group)

    def quantile_multi(self, V: torch.Tensor, t: int, pairs: list, n_rows: int = None) -> dict:
        """``sieving.abstract.quantile_multi`` over the rows of all ranks:
        ``V[P, n_local * t]`` holds this rank's rows of every problem."""
        import ctypes

        from . import _backend as be
        from .sieving.abstract import _lerp, _virtual_index
        V = V.contiguous()
        P, m_local = V.shape
        M = (self.n_sample if n_rows is None else n_rows) * t
        out = {}
        L = be.lib()
        for g in range(0, len(pairs), 4):
            grp = pairs[g:g + 4]
            S = len(grp)
            kg = [_virtual_index(M, q) for _, q in grp]
            incs = (ctypes.c_int32 * S)(*[int(i) for i, _ in grp])
            ks = (ctypes.c_int64 * S)(*[int(k) for k, _ in kg])
            lay = (ctypes.c_int64 * 5)()
            be.check(L.fb_order_stats_dist_layout(P, S, lay))
            work = torch.zeros((lay[4],), dtype=torch.uint8, device=V.device)
            lo, hi = be.empty((P, S)), be.empty((P, S))
            done = be.empty((P, S), dtype=torch.int32)
            ns = P * S
            for phase in range(10):
                be.check(L.fb_order_stats_dist(phase, V.data_ptr(), m_local, P, m_local, M, int(t),
                                               S, incs, ks, lo.data_ptr(), hi.data_ptr(),
                                               done.data_ptr(), work.data_ptr(), be.stream_ptr()))
                if phase in (0, 1):
                    self._reduce(work, lay[0], ns * 4096, torch.int32, dist.ReduceOp.SUM)
                elif phase in (2, 8):
                    self._reduce(work, lay[2], ns * 2, torch.int64, dist.ReduceOp.SUM)
                    self._reduce(work, lay[3], ns * 4, torch.int64, dist.ReduceOp.MIN)
                elif 3 <= phase <= 7:
                    self._reduce(work, lay[1], ns * 256, torch.int32, dist.ReduceOp.SUM)
            packed = torch.cat([lo, hi, done.to(torch.float64)], dim=1).cpu().numpy()
            for s, (pair, (k, gamma)) in enumerate(zip(grp, kg)):
                if not packed[:, 2 * S + s].all():
                    out[pair] = None
                    continue
                a = packed[:, s]
                b = packed[:, S + s] if k < M - 1 else a
                out[pair] = _lerp(a, b, gamma)
        return out

    def quantile_rows(self, V: torch.Tensor, q: float, m_global: int) -> np.ndarray:
        """``sieving.abstract.quantile_rows`` over the values of all ranks:
        ``V[P, m_local]`` = this rank's values of every problem."""
        import ctypes

        from . import _backend as be
        from .sieving.abstract import _lerp, _virtual_index
        V = V.contiguous()
        P, m_local = V.shape
        k, gamma = _virtual_index(m_global, q)
        L = be.lib()
        lay = (ctypes.c_int64 * 4)()
        be.check(L.fb_order_stats_dist8_layout(P, lay))
        work = torch.zeros((lay[3],), dtype=torch.uint8, device=V.device)
        lo, hi = be.empty((P,)), be.empty((P,))
        for phase in range(10):
            be.check(L.fb_order_stats_dist8(phase, V.data_ptr(), m_local, P, m_local, m_global,
                                            int(k), lo.data_ptr(), hi.data_ptr(), work.data_ptr(),
                                            be.stream_ptr()))
            if phase <= 7:
                self._reduce(work, lay[0], P * 256, torch.int32, dist.ReduceOp.SUM)
            elif phase == 8:
                self._reduce(work, lay[1], P * 2, torch.int64, dist.ReduceOp.SUM)
                self._reduce(work, lay[2], P, torch.int64, dist.ReduceOp.MIN)
        a = lo.cpu().numpy()
        b = hi.cpu().numpy() if k < m_global - 1 else a
        return _lerp(a, b, gamma)


def _draw_sample_positions(n_total: int, fit_sample_size) -> np.ndarray:
    """The draw of ``FruitSlice._select_fit_sample`` (fruits/fruit.py:430-438)."""
    if isinstance(fit_sample_size, int) and fit_sample_size == 1:
        return np.array([np.random.randint(0, n_total)], dtype=np.int64)
    s = max(int(fit_sample_size * n_total), 1)
    return np.random.choice(n_total, size=s, replace=False).astype(np.int64)


def _rows_ok(slc) -> bool:
    """Can this slice be fitted on a row-sharded sample?  Needs: preparateurs
    whose fit looks at one series at a time, one ISS, sieves whose fit is a
    quantile (SegmentSieve family, PPV) or nothing."""
    from .sieving.implicit import PPV
    from .sieving.segment import SegmentSieve
    if len(slc.get_iss()) != 1:
        return False
    if not all(p._row_independent_fit() and not p._needs_raw_cache()
               for p in slc.get_preparateurs()):
        return False
    for sv in slc.get_sieves():
        if sv.requires_fitting and not isinstance(sv, (SegmentSieve, PPV)):
            return False
    return True


def fit_sharded(fruit, X_local: torch.Tensor, n_total: Optional[int] = None, group=None,
                shard_nodes: bool = True, shard: str = "auto") -> None:
    """``Fruit.fit`` on a row-sharded batch; thresholds are bit-identical on
    all ranks and equal to a single-GPU fit of the whole batch.

    ``shard="rows"`` (the default wherever a slice allows it, ``_rows_ok``): the
    fit sample STAYS sharded -- every rank keeps the sampled rows it owns,
    materialises their iterated sums and the quantiles come from a radix select
    whose histograms are all-reduced (:class:`RowShard`).  No rank ever holds
    more than its share of the sample, slices whose sieves need no fitting touch
    no sample at all, and L1 / L2 weightings get the rows of the raw-input cache
    they need from their owners (reference quirk: fruits/cache.py:97-112).

    ``shard="nodes"``: the sample is gathered on every rank and the iterated
    sums are split over the ranks (``shard_nodes``), fitted sieves exchanged."""
    from . import _backend as be
    from .cache import SharedSeedCache
    if n_total is None:
        sizes = torch.tensor([X_local.shape[0]], dtype=torch.int64, device=X_local.device)
        dist.all_reduce(sizes, group=group)
        n_total = int(sizes.item())
    sync_numpy_rng(group)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_rows(n_total, world, rank)
    if hi - lo != X_local.shape[0]:
        raise ValueError(f"rank {rank} must hold the rows [{lo}, {hi}) of the batch "
                         f"(shard_rows), got {X_local.shape[0]} rows")
    for slc in fruit:
        rows_mode = shard in ("auto", "rows") and X_local.is_cuda and _rows_ok(slc)
        if shard == "rows" and not rows_mode:
            raise NotImplementedError("this slice cannot be fitted on a row-sharded sample")
        if rows_mode:
            _fit_slice_rows(slc, X_local, n_total, lo, hi, group)
            continue
        for iss in slc.get_iss():
            w = iss.weighting
            if w is not None and not getattr(w, "_on_prepared", True) and any(
                    s.requires_fitting for s in slc.get_sieves()):
                raise NotImplementedError(
                    "node-sharded fit of sieves on L1/L2-weighted sums needs the raw-input cache "
                    "of the whole batch (reference quirk: fruits/cache.py:97-112)")
        if any(p._needs_raw_cache() for p in slc.get_preparateurs()):
            raise NotImplementedError(
                "preparateurs that read the raw-input cache (WIN, SPE(step_transform=...)) "
                "cannot be fitted on a sharded batch")
        sample = gather_fit_sample(X_local, n_total, slc.fit_sample_size, group)

        def exchange(copies):
            parts = [None] * world
            dist.all_gather_object(parts, copies, group=group)
            return [row for part in parts for row in part]

        try:
            # the gathered rows ARE the sample: fit on all of them; the iterated
            # sums (each needs the whole sample for its quantiles) are split over
            # the ranks, the fitted thresholds exchanged
            slc._select_fit_sample = lambda X: X
            if world > 1 and shard_nodes:
                slc._fit_shard = (lambda n_emit, costs: shard_by_cost(n_emit, costs, world, rank),
                                  exchange)
            slc._fit_device(be.to_device(sample), SharedSeedCache(sample))
        finally:
            del slc._select_fit_sample
            slc.__dict__.pop("_fit_shard", None)
    fruit._fitted = True


class _ShardedCache:
    """``SharedSeedCache`` of a row-sharded batch, seen from the fit sample: the
    reference builds the cache on the WHOLE batch and reads row ``j`` of it for
    sample position ``j`` (fruits/cache.py:97-112), whoever owns that row."""

    def __init__(self, X_local, n_total, positions, group) -> None:
        from .cache import SharedSeedCache
        self._local = SharedSeedCache(X_local)
        self._n_total, self._positions, self._group = n_total, positions, group
        self._memo = {}

    def get_device(self, cache_id, key, X=None):
        k = (cache_id, key)
        if k not in self._memo:
            mine = self._local.get_device(cache_id, key)            # rows of this rank
            self._memo[k] = exchange_rows(mine, self._n_total, self._positions, self._group)
        return self._memo[k]

    def get(self, cache_id, key, X=None):
        return self.get_device(cache_id, key, X).cpu().numpy()


def _fit_slice_rows(slc, X_local, n_total, lo, hi, group) -> None:
    """One slice on a row-sharded sample (see ``fit_sharded``)."""
    idx = _draw_sample_positions(n_total, slc.fit_sample_size)      # same draw on every rank
    mine = np.nonzero((idx >= lo) & (idx < hi))[0]                  # sample positions held here
    needs_rows = (any(sv.requires_fitting for sv in slc.get_sieves())
                  or any(p.requires_fitting for p in slc.get_preparateurs()))
    cache_rows = mine
    if not needs_rows:
        # nothing to fit (e.g. experiments/fruit_twi.py): the draw above keeps the RNG in
        # step with the reference, the sample itself is never looked at
        sample, mine = X_local[:1], mine[:1]
        cache_rows = mine
    elif len(mine):
        rows = torch.as_tensor(idx[mine] - lo, device=X_local.device, dtype=torch.long)
        sample = X_local.index_select(0, rows)
    else:
        # none of the sampled rows lives here (small samples): this rank still takes part
        # in every collective, with a placeholder row that no selection looks at
        sample = torch.ones((1,) + tuple(X_local.shape[1:]), dtype=X_local.dtype,
                            device=X_local.device)
        cache_rows = np.zeros(1, dtype=np.int64)
    if sample.shape[0] == 0:
        sample = torch.ones((1,) + tuple(X_local.shape[1:]), dtype=X_local.dtype,
                            device=X_local.device)
        cache_rows = np.zeros(1, dtype=np.int64)
    rs = RowShard(mine, len(idx), group)
    cache = _ShardedCache(X_local, n_total, cache_rows, group)
    try:
        slc._select_fit_sample = lambda X: X
        slc._row_shard = rs
        slc._fit_device(sample, cache)
    finally:
        del slc._select_fit_sample
        slc.__dict__.pop("_row_shard", None)


class PeerGather:
    """Assembled feature matrix ``[world*S, n_feats]`` in NVLink peer memory.

    The buffer is a symmetric allocation (``torch.distributed._symmetric_memory``):
    every rank maps the buffers of all peers, so a rank *pushes* each finished
    row chunk straight into its rows of every peer's matrix with plain
    device-to-device copies.  Those run on the copy engines over NVLink 5 /
    NVSwitch: no SM is taken from the feature kernel, there is no staging
    buffer and no re-packing pass.  ``finish()`` is a device-side barrier over
    all ranks, after which every rank holds the complete matrix.

    Raises ``RuntimeError`` if symmetric memory is unavailable (the caller
    then falls back to the NCCL all-gather)."""

    def __init__(self, rows_per_rank: int, n_feats: int, group=None,
                 multicast: bool = True) -> None:
        import torch.distributed._symmetric_memory as symm
        from torch._C._autograd import DeviceType
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.S, self.F = rows_per_rank, n_feats
        dev = torch.device("cuda", torch.cuda.current_device())
        self.out = symm.empty((self.world * rows_per_rank, n_feats), dtype=torch.float64,
                              device=dev)
        self.handle = symm.rendezvous(self.out, self.group)
        shape = (self.world, rows_per_rank, n_feats)
        self.peers = [self.handle.get_buffer(r, shape, torch.float64)
                      for r in range(self.world)]
        self.local = self.peers[self.rank]
        self.copy_streams = [torch.cuda.Stream(device=dev) for _ in range(min(4, self.world - 1))]
        # NVSwitch multicast mapping of the same buffers (NVLS): a plain store to this
        # address is replicated by the switch into the matrix of EVERY rank, so the
        # feature kernel itself performs the all-gather (see ``multicast_rows``)
        self.mc_ptr = 0
        if multicast and self.world > 1:
            try:
                if self.handle.has_multicast_support(DeviceType.CUDA, dev.index):
                    self.mc_ptr = int(self.handle.multicast_ptr)
            except Exception:
                self.mc_ptr = 0

    @property
    def fused(self) -> bool:
        """True if the kernels can store straight into all ranks' matrices."""
        return self.mc_ptr != 0

    def multicast_rows(self, lo: int, hi: int):
        """``(address, row stride in elements)`` of this rank's rows ``[lo, hi)``
        in the multicast address space: every ``multimem.st`` to it lands in the
        same rows of the matrix of every rank (own included).  A raw address on
        purpose: multicast mappings may only be touched by ``multimem.*``
        instructions, so it is never wrapped into a tensor or handed to torch."""
        return self.mc_ptr + ((self.rank * self.S + lo) * self.F) * 8, self.F

    def rows(self, lo: int, hi: int) -> torch.Tensor:
        """This rank's rows ``[lo, hi)`` inside its own matrix (compute target)."""
        return self.local[self.rank, lo:hi]

    def push(self, lo: int, hi: int) -> None:
        """Send rows ``[lo, hi)`` (already computed on the current stream) to all peers."""
        cur = torch.cuda.current_stream()
        done = torch.cuda.Event()
        done.record(cur)
        src = self.local[self.rank, lo:hi]
        for i in range(1, self.world):
            peer = (self.rank + i) % self.world          # staggered: no two ranks hit one peer
            st = self.copy_streams[(i - 1) % len(self.copy_streams)]
            st.wait_event(done)
            with torch.cuda.stream(st):
                self.peers[peer][self.rank, lo:hi].copy_(src, non_blocking=True)

    def finish(self) -> torch.Tensor:
        cur = torch.cuda.current_stream()
        for st in self.copy_streams:
            cur.wait_stream(st)
        self.handle.barrier(channel=0)                   # all pushes of all ranks have landed
        return self.out


def _check_equal_shards(S: int, dev, group) -> None:
    """All ranks must hold the same number of rows: the assembled matrix is
    ``[world * S, F]`` with rank-major rows and every collective below uses
    equal sizes (uneven shards would hang or land rows at wrong offsets)."""
    t = torch.tensor([S, -S], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    if int(t[0]) != -int(t[1]):
        raise ValueError(f"transform_sharded needs equally sized shards (this rank holds {S} "
                         f"rows, the ranks hold between {-int(t[1])} and {int(t[0])}); pad the "
                         f"batch or use shard sizes that divide it")


def transform_sharded(fruit, X_local: torch.Tensor, n_feats: int, chunks: int = 8, group=None,
                      out=None) -> torch.Tensor:
    """All ranks hold ``S`` rows; returns the assembled ``[world*S, n_feats]``
    feature matrix (rank-major row order, identical on every rank).

    ``fruit`` is a fitted :class:`~fruits_b200.Fruit` (or, for tests, a callable
    ``compute(X_rows, out_rows)`` that writes the features of a row block).
    With ``out`` a :class:`PeerGather` whose multicast mapping is available the
    feature kernels store straight into the matrices of all ranks
    (``multimem.st``); else every finished row piece is pushed to the peers by
    the copy engines while the next piece is computed; without a
    ``PeerGather`` the pieces are all-gathered with NCCL on a side stream."""
    world = dist.get_world_size(group)
    S = X_local.shape[0]
    dev = X_local.device
    is_fruit = hasattr(fruit, "transform_device")
    compute = (lambda x, o: fruit.transform_device(x, out=o)) if is_fruit else fruit
    if world > 1:
        _check_equal_shards(S, dev, group)
    if isinstance(out, PeerGather):
        if out.fused and is_fruit:
            # the kernels store through the NVSwitch multicast mapping: compute and
            # all-gather are one launch, nothing is left to overlap
            if S:
                fruit.transform_device(X_local, out=out.rows(0, S),
                                       multicast=out.multicast_rows(0, S))
            return out.finish()
        chunks = max(1, min(chunks, S)) if S else 1
        for lo, hi in (shard_rows(S, chunks, c) for c in range(chunks)):
            compute(X_local[lo:hi], out.rows(lo, hi))
            out.push(lo, hi)
        return out.finish()
    if out is None:
        out = torch.empty((world * S, n_feats), dtype=torch.float64, device=dev)
    if world == 1:
        compute(X_local, out)
        return out
    chunks = max(1, min(chunks, S)) if S else 1
    bounds = [shard_rows(S, chunks, c) for c in range(chunks)]
    rows_max = max(hi - lo for lo, hi in bounds) if S else 0
    use_cuda = dev.type == "cuda"
    local = [torch.empty((rows_max, n_feats), dtype=torch.float64, device=dev) for _ in range(2)]
    stage = [torch.empty((world, rows_max, n_feats), dtype=torch.float64, device=dev)
             for _ in range(2)]
    out3 = out.view(world, S, n_feats)
    if use_cuda:
        cur = torch.cuda.current_stream(dev)
        comm = torch.cuda.Stream(device=dev)
        free_ev = [None, None]
    for c, (lo, hi) in enumerate(bounds):
        b, rows = c % 2, hi - lo
        if use_cuda and free_ev[b] is not None:
            cur.wait_event(free_ev[b])
        compute(X_local[lo:hi], local[b][:rows])
        if use_cuda:
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                dist.all_gather_into_tensor(stage[b].view(world * rows_max, n_feats), local[b],
                                            group=group)
                out3[:, lo:hi].copy_(stage[b][:, :rows])
                free_ev[b] = torch.cuda.Event()
                free_ev[b].record(comm)
        else:
            dist.all_gather_into_tensor(stage[b].view(world * rows_max, n_feats), local[b],
                                        group=group)
            out3[:, lo:hi].copy_(stage[b][:, :rows])
    if use_cuda:
        cur.wait_stream(comm)
    return out
