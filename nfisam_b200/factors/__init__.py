from .factors import *  # noqa: F401,F403
from .factors import FACTOR_CLASSES, oracle_descriptor  # noqa: F401
from .geometry import SE2Pose  # noqa: F401

