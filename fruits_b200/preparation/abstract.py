"""Reference: ``fruits/preparation/abstract.py:9-20``."""
from abc import ABC
from typing import Any

from ..seed import Seed


class Preparateur(Seed, ABC):
    """A preparateur maps ``X[n_series, n_dims, length]`` to another array of
    the same kind before the iterated sums are calculated."""

    def __eq__(self, other: Any) -> bool:
        return False

    __hash__ = object.__hash__

    def _fusable(self):
        """Per-dimension description ``[(source_dim, inc, std), ...]`` if the
        ISS kernel can apply this preparateur while loading X, else None."""
        return None

    def _row_independent_fit(self) -> bool:
        """True if ``fit`` looks at one series at a time (or at nothing), so a
        row-sharded fit sample needs no exchange (``parallel.fit_sharded``)."""
        return True

    def _row_independent_transform(self) -> bool:
        """True if ``transform`` treats every series on its own, so a host batch
        may be streamed through the GPU in row chunks (``Fruit._transform_host``);
        false for user code that sees the whole batch (``FUN``)."""
        return True

    def _needs_raw_cache(self) -> bool:
        """True if ``transform`` reads the cache of the raw batch (``WIN``,
        ``SPE(step_transform=...)``): a fit on a gathered sample cannot serve it."""
        return False
