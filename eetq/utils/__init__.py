from eetq_b200.utils import *  # noqa: F401,F403
from eetq_b200.utils import base, quantizer  # noqa: F401
