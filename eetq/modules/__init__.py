from eetq_b200.modules import *  # noqa: F401,F403
from eetq_b200.modules import qlinear  # noqa: F401
