import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from eetq_b200 import _cabi
torch.cuda.set_device(0); torch.zeros(1, device="cuda")
a, b = ctypes.c_int(), ctypes.c_int()
_cabi.check(_cabi.lib().eetq_b200_decode_attention_occupancy(32, ctypes.byref(a), ctypes.byref(b)), "occ")
print("attention: CTAs/SM", a.value, "max active clusters (of 8)", b.value)
