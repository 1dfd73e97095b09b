from .base import *  # noqa: F401,F403
from .quantizer import *  # noqa: F401,F403
