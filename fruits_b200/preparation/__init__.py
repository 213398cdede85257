from .abstract import Preparateur
from .filter import *
from .transform import *
from .wrapper import *
