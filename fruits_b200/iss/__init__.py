from . import semiring, weighting
from .cache import CachePlan
from .cos import CosWISS
from .iss import ISS, ISSMode
