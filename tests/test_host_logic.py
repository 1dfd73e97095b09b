"""CPU tests of the host-side mirror of the reference's Python surface (no kernels run)."""
import pytest
import torch
import torch.nn as nn

import eetq_b200
from eetq_b200 import EetqLinear, W8A16Linear, find_layers, set_op_by_name
from eetq_b200.utils.base import find_submodule, get_named_linears


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj = nn.Linear(64, 64, bias=False)
        self.mlp = nn.Sequential(nn.Linear(64, 128), nn.SiLU(), nn.Linear(128, 64))


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([Block(), Block()])
        self.lm_head = nn.Linear(64, 100, bias=False)


def test_surface_names_match_reference():
    # the names the reference exports for the hot path (python/eetq/__init__.py, csrc/eetpy.cpp:9-17)
    for name in ["W8A16Linear", "EetqLinear", "EetqLinearMMFunction", "quantize_and_preprocess_weights", "eet_quantize",
                 "find_layers", "set_op_by_name", "quant_weights", "preprocess_weights", "w8_a16_gemm", "w8_a16_gemm_"]:
        assert hasattr(eetq_b200, name), name
    import EETQ

    for name in ["quant_weights", "preprocess_weights", "w8_a16_gemm", "w8_a16_gemm_"]:
        assert hasattr(EETQ, name)


def test_find_layers_skips_lm_head():
    m = Tiny()
    found = find_layers(m)
    assert "lm_head" not in found
    assert sorted(found) == ["layers.0.mlp.0", "layers.0.mlp.2", "layers.0.q_proj", "layers.1.mlp.0", "layers.1.mlp.2",
                             "layers.1.q_proj"]
    assert sorted(get_named_linears(m)) == sorted(found)
    assert find_submodule(m, "layers") is m.layers


def test_set_op_by_name_handles_indices():
    m = Tiny()
    new = nn.Identity()
    set_op_by_name(m, "layers.1.mlp.0", new)
    assert m.layers[1].mlp[0] is new
    set_op_by_name(m, "lm_head", new)
    assert m.lm_head is new


def test_w8a16linear_buffers_and_state_dict_keys():
    q = W8A16Linear(128, 64, bias=True, dev="cpu")
    sd = q.state_dict()
    # qlinear.py:34-38 + the layout marker that tells b200 bytes from reference-layout bytes
    assert sorted(sd) == ["bias", "qweight", "weight_layout", "weight_scales"]
    assert int(sd["weight_layout"][0]) == eetq_b200.B200_LAYOUT
    assert sd["qweight"].shape == (128, 64) and sd["qweight"].dtype == torch.int8
    assert sd["weight_scales"].shape == (64,) and sd["weight_scales"].dtype == torch.float16
    q2 = W8A16Linear(128, 64, bias=False, dev="cpu")
    assert q2.bias is None and sorted(q2.state_dict()) == ["qweight", "weight_layout", "weight_scales"]


def test_eetqlinear_late_scale_registration():
    q = EetqLinear(128, 64, bias=False, device="cpu")
    assert sorted(q.state_dict()) == ["weight", "weight_layout"]         # qlinear.py:103 (+ marker)
    q.register_scale("cpu")
    assert sorted(q.state_dict()) == ["weight", "weight_layout", "weight_scales"]   # qlinear.py:113-116


@pytest.mark.parametrize("cls,key", [(W8A16Linear, "qweight"), (EetqLinear, "weight")])
def test_state_dict_layout_marker_decides_conversion(monkeypatch, cls, key):
    """A state dict WITHOUT the marker comes from a reference build (sm80 interleaved bytes): the weight goes through the
    layout converter on load.  One WITH the marker is loaded verbatim; an unknown marker value is an error.  The converter
    (a GPU kernel) is stubbed here: this checks the host logic."""
    import eetq_b200.modules.qlinear as ql

    calls = []

    def fake_convert(w):
        calls.append(w.clone())
        return (w.to(torch.int16) + 1).to(torch.int8)

    monkeypatch.setattr(ql, "convert_ref_checkpoint_weight", fake_convert)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    kw = dict(dev="cpu") if cls is W8A16Linear else dict(device="cpu")
    src = cls(64, 64, bias=False, **kw)
    if cls is EetqLinear:
        src.register_scale("cpu")
    getattr(src, key).copy_(torch.randint(-100, 100, (64, 64), dtype=torch.int8))
    sd = src.state_dict()

    same = cls(64, 64, bias=False, **kw)
    if cls is EetqLinear:
        same.register_scale("cpu")
    same.load_state_dict(sd)                                   # b200 checkpoint: verbatim
    assert not calls and torch.equal(getattr(same, key), getattr(src, key))

    ref_sd = {k: v for k, v in sd.items() if k != "weight_layout"}      # what a reference build would have written
    conv = cls(64, 64, bias=False, **kw)
    if cls is EetqLinear:
        conv.register_scale("cpu")
    conv.load_state_dict(ref_sd)                               # strict load succeeds: the marker is synthesised
    assert len(calls) == 1 and torch.equal(getattr(conv, key), getattr(src, key) + 1)
    assert "weight_layout" not in ref_sd                       # the caller's dict is left untouched

    bad = dict(sd)
    bad["weight_layout"] = torch.tensor([7], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="weight layout does not match"):
        cls(64, 64, bias=False, **kw).load_state_dict(bad, strict=False)


def test_eet_quantize_init_only_builds_skeleton():
    m = Tiny().half()
    eetq_b200.eet_quantize(m, init_only=True)
    assert isinstance(m.layers[0].q_proj, W8A16Linear) and isinstance(m.layers[1].mlp[2], W8A16Linear)
    assert isinstance(m.lm_head, nn.Linear)                              # excluded (quantizer.py:40)
    assert m.layers[0].mlp[0].bias is not None and m.layers[0].q_proj.bias is None


def test_argument_validation_before_any_launch():
    x = torch.zeros(1, 64, dtype=torch.float32)
    w = torch.zeros(64, 64, dtype=torch.int8)
    s = torch.zeros(64, dtype=torch.float16)
    with pytest.raises(RuntimeError):
        eetq_b200.w8_a16_gemm(x, w, s)
    with pytest.raises(RuntimeError, match="int4 or int8"):
        eetq_b200.quant_weights(torch.zeros(64, 64, dtype=torch.float16), torch.int32, False)
    if not torch.cuda.is_available():  # int4 is implemented; like every op it needs the device (no CPU path)
        with pytest.raises(RuntimeError, match="CUDA"):
            eetq_b200.preprocess_weights(w, is_int4=True)
        with pytest.raises(RuntimeError, match="CUDA"):
            eetq_b200.quant_weights(torch.zeros(64, 64, dtype=torch.float16), torch.quint4x2, False)
    with pytest.raises(RuntimeError, match="2-D int8"):
        eetq_b200.w4_a16_gemm(x.half(), w.float(), s)


def test_eet_quantize_int8_ingest_uses_scb_scales(monkeypatch):
    """bitsandbytes ingest (quantizer.py:46-48 in the reference): int8 weight + SCB -> scales = SCB / 127, weights only
    re-laid-out (no re-quantisation).  The native calls are stubbed: this checks the host logic on CPU."""
    import eetq_b200.modules.qlinear as ql

    calls = {}

    def fake_preprocess(w):
        calls["pre"] = w.clone()
        return w

    monkeypatch.setattr(ql, "preprocess_weights", fake_preprocess)

    class FakeBnbLinear(nn.Linear):
        pass

    lin = FakeBnbLinear(64, 128, bias=False)
    lin.weight = nn.Parameter(torch.randint(-127, 128, (128, 64), dtype=torch.int8), requires_grad=False)
    lin.register_buffer("SCB", torch.rand(128) + 0.5)

    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.proj = lin

    m = Holder()
    eetq_b200.eet_quantize(m, include=(FakeBnbLinear,))
    q = m.proj
    assert isinstance(q, W8A16Linear)
    assert torch.equal(calls["pre"], lin.weight.t().contiguous())          # [K, N] handed to the layout pass
    assert torch.allclose(q.weight_scales.float(), (lin.SCB / 127.0).half().float())
    assert q.qweight.dtype == torch.int8 and q.qweight.shape == (64, 128)


def test_quantize_and_preprocess_rejects_unknown_dtype():
    from eetq_b200.modules.qlinear import quantize_and_preprocess_weights

    with pytest.raises(ValueError, match="Unsupported data type"):
        quantize_and_preprocess_weights(torch.zeros(64, 64, dtype=torch.int32))
    with pytest.raises(AssertionError):
        quantize_and_preprocess_weights(torch.zeros(64, 64, dtype=torch.int8), None)


def test_reference_package_names_resolve():
    """`import eetq` / `from eetq.modules import W8A16Linear` / `from EETQ import w8_a16_gemm` -- the import lines the
    reference's own code and examples use (python/eetq/modules/qlinear.py:11, examples/layers/test_qlinear.py)."""
    import eetq
    from eetq.modules import W8A16Linear as A
    from eetq.utils import eet_quantize as q
    from EETQ import preprocess_weights, quant_weights, w8_a16_gemm  # noqa: F401

    assert A is eetq_b200.W8A16Linear and q is eetq_b200.eet_quantize and eetq.EetqLinear is eetq_b200.EetqLinear


def test_w4a16_module_surface_and_layout_marker():
    """W4A16Linear (extension): packed [in, out/2] buffer, its own layout marker, eet_quantize(bits=4) swaps it in; a state dict of
    the other bit width is refused instead of being mis-read."""
    from eetq_b200 import W4A16Linear

    m = nn.Sequential(nn.Linear(128, 64), nn.Linear(64, 128, bias=False))
    eetq_b200.eet_quantize(m, init_only=True, bits=4)
    assert isinstance(m[0], W4A16Linear) and isinstance(m[1], W4A16Linear)
    sd = m.state_dict()
    assert sd["0.qweight"].shape == (128, 32) and sd["0.qweight"].dtype == torch.int8
    assert int(sd["0.weight_layout"][0]) == eetq_b200.B200_LAYOUT_INT4 and "1.bias" not in sd
    q4 = W4A16Linear(128, 64, bias=True, dev="cpu")
    q4.load_state_dict({k[2:]: v for k, v in sd.items() if k.startswith("0.")})          # same layout: loads as is
    q8 = W8A16Linear(128, 64, bias=True, dev="cpu")
    bad = dict(q8.state_dict())
    bad["qweight"] = torch.zeros(128, 32, dtype=torch.int8)
    with pytest.raises(RuntimeError, match="weight layout does not match"):
        q4.load_state_dict(bad)                                                           # int8 marker into an int4 module
    with pytest.raises(ValueError, match="bits must be 8 or 4"):
        eetq_b200.eet_quantize(nn.Sequential(nn.Linear(64, 64)), init_only=True, bits=3)
    with pytest.raises(ValueError, match="multiples of 64"):
        W4A16Linear(100, 64, dev="cpu")
