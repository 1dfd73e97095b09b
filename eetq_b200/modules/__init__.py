from .qlinear import *  # noqa: F401,F403
