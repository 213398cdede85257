from .abstract import FeatureSieve
from .segment import *
from .increment import *
from .implicit import *
from .wrapper import *
