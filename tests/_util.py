import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _from_bits(a: np.ndarray, dtype: str) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.view(torch.float16) if (dtype == "float16" and t.dtype == torch.int16) else t


def golden_cases():
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        z = np.load(path)
        dt = str(z["dtype"])
        case = {
            "name": os.path.basename(path)[:-4],
            "dtype": dt,
            "w": _from_bits(z["w"], dt),
            "q": torch.from_numpy(z["q"]),
            "w_ref": torch.from_numpy(z["w_ref"]),
            "scales": _from_bits(z["scales"], dt),
        }
        if "x" in z.files:
            case["x"] = _from_bits(z["x"], "float16")
            case["y"] = _from_bits(z["y"], "float16")
        out.append(case)
    assert out, "no golden fixtures found"
    return out


def golden_cases_int4():
    """tests/golden/int4/*.npz -- packed-int4 fixtures from the compiled reference (tests/golden/make_golden_int4.py)."""
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "int4", "*.npz"))):
        z = np.load(path)
        dt = str(z["dtype"])
        case = {
            "name": os.path.basename(path)[:-4],
            "dtype": dt,
            "w": _from_bits(z["w"], dt),
            "q4": torch.from_numpy(z["q4"]),
            "w4_ref": torch.from_numpy(z["w4_ref"]),
            "scales": _from_bits(z["scales"], dt),
        }
        if "x" in z.files:
            case["x"] = _from_bits(z["x"], "float16")
            case["y"] = _from_bits(z["y"], "float16")
        out.append(case)
    assert out, "no int4 golden fixtures found"
    return out
