"""Import-compatible alias of the reference package name ``eetq`` (/root/reference/python/eetq/__init__.py) for the w8a16
hot path: ``from eetq import W8A16Linear, EetqLinear, eet_quantize`` resolves to the B200 implementations in
:mod:`eetq_b200`.  ``AutoEETQForCausalLM`` (checkpoint packaging) is outside the hot path and is not provided."""
from eetq_b200 import *  # noqa: F401,F403
from eetq_b200 import __version__  # noqa: F401
