"""GPU parity tests (call through the C ABI): quantiser and layout kernels, bit-exact vs the oracle, the committed
golden vectors (reference C++ output) and -- when oracle/_ref is present -- the live reference library."""
import pytest
import torch

import eetq_b200
from _util import golden_cases

pytestmark = pytest.mark.gpu
CASES = golden_cases()


def _bits(t):
    return t.view(torch.int16) if t.dtype in (torch.float16, torch.bfloat16) else t


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_quant_weights_matches_reference_golden(cuda, oracle, case):
    unp, pro, sc = eetq_b200.quant_weights(case["w"].to(cuda), torch.int8, True)
    assert torch.equal(unp.cpu(), case["q"])
    assert torch.equal(_bits(sc.cpu()), _bits(case["scales"]))
    assert torch.equal(pro.cpu(), oracle.b200_layout(case["q"]))
    # reference-layout converter reproduces the reference's processed bytes
    assert torch.equal(eetq_b200.to_ref_checkpoint_weight(pro).cpu(), case["w_ref"])
    assert torch.equal(eetq_b200.convert_ref_checkpoint_weight(case["w_ref"].to(cuda)).cpu(), pro.cpu())


def test_quant_weights_cpu_input_returns_cpu(cuda, oracle):
    """Reference contract: CPU in, CPU out (fpA_intB_gemm_wrapper.cu:33, :54-73); 2-tuple / 3-tuple order (:101-106)."""
    w = oracle.synth_weight(128, 64, seed=2)
    out2 = eetq_b200.quant_weights(w, torch.int8, False)
    out3 = eetq_b200.quant_weights(w, torch.int8, True)
    assert len(out2) == 2 and len(out3) == 3
    assert all(not t.is_cuda for t in out2 + out3)
    q, s, _ = oracle.quantize(w)
    assert torch.equal(out3[0], q) and torch.equal(out3[2], s) and torch.equal(out2[0], out3[1])
    assert out2[0].shape == w.shape and out2[0].dtype == torch.int8 and out2[1].dtype == w.dtype


@pytest.mark.parametrize("shape", [(4096, 4096), (4096, 11008), (11008, 4096)])
def test_quant_weights_full_size_bit_exact(cuda, oracle, shape):
    """BASELINE.json sizes; oracle (and the live reference when built) must agree bit for bit."""
    w = oracle.synth_weight(*shape, seed=1000)
    unp, pro, sc = eetq_b200.quant_weights(w.to(cuda), torch.int8, True)
    q, s, _ = oracle.quantize(w)
    assert torch.equal(unp.cpu(), q) and torch.equal(sc.cpu(), s)
    assert torch.equal(pro.cpu(), oracle.b200_layout(q))
    if oracle.ref_lib() is not None and shape == (4096, 4096):
        r_unp, r_pro, r_sc = oracle.ref_quantize(w)
        assert torch.equal(unp.cpu(), r_unp) and torch.equal(sc.cpu(), r_sc)
        assert torch.equal(eetq_b200.to_ref_checkpoint_weight(pro).cpu(), r_pro)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_quant_weights_other_dtypes(cuda, oracle, dtype):
    w32 = oracle.synth_weight(256, 192, seed=4, dtype=torch.float32)
    w = w32.to(dtype)
    unp, pro, sc = eetq_b200.quant_weights(w.to(cuda), torch.int8, True)
    q, s, s32 = oracle.quantize(w.float())        # arithmetic is fp32 whatever the storage dtype
    assert torch.equal(unp.cpu(), q)
    assert torch.equal(sc.cpu().float(), s32.to(dtype).float())
    assert sc.dtype == dtype


def test_quant_weights_3d_experts(cuda, oracle):
    w = torch.stack([oracle.synth_weight(128, 64, seed=s) for s in (1, 2, 3)])
    pro, sc = eetq_b200.quant_weights(w.to(cuda), torch.int8, False)
    assert pro.shape == w.shape and sc.shape == (3, 64)
    for e in range(3):
        q, s, _ = oracle.quantize(w[e])
        assert torch.equal(pro[e].cpu(), oracle.b200_layout(q)) and torch.equal(sc[e].cpu(), s)


def test_quant_edge_values(cuda, oracle):
    w = torch.zeros(64, 64, dtype=torch.float16)
    w[:, 0] = 1.0                      # all equal to amax -> 127 (128 clamps)
    w[:, 1] = -1.0                     # -> -128
    w[0, 2] = 6e-8                     # fp16 subnormal amax
    w[:, 3] = float("nan"); w[0, 3] = 1.0   # NaNs ignored by amax, quantise to 127 (NaN compare order)
    unp, pro, sc = eetq_b200.quant_weights(w.to(cuda), torch.int8, True)
    q, s, _ = oracle.quantize(w)
    assert torch.equal(unp.cpu(), q)
    assert torch.equal(sc.cpu().view(torch.int16), s.view(torch.int16))
    assert (unp[:, 0] == 127).all() and (unp[:, 1] == -128).all() and (unp[:, 4] == 127).all() and sc[4] == 0


def test_pack_unpack_roundtrip(cuda, oracle):
    q = torch.randint(-128, 128, (192, 320), dtype=torch.int8)
    p = eetq_b200.preprocess_weights(q.to(cuda))
    assert torch.equal(p.cpu(), oracle.b200_layout(q))
    assert torch.equal(eetq_b200.unpack_weights(p).cpu(), q)
    # CPU in -> CPU out like the reference (fpA_intB_gemm_wrapper.cu:113)
    p2 = eetq_b200.preprocess_weights(q)
    assert not p2.is_cuda and torch.equal(p2, p.cpu())


def test_ref_layout_converters_full_size(cuda, oracle):
    q = torch.randint(-128, 128, (4096, 4096), dtype=torch.int8)
    w_ref = oracle.ref_layout(q)
    p = eetq_b200.convert_ref_checkpoint_weight(w_ref.to(cuda))
    assert torch.equal(p.cpu(), oracle.b200_layout(q))
    assert torch.equal(eetq_b200.to_ref_checkpoint_weight(p).cpu(), w_ref)


def test_shape_constraints_raise(cuda):
    with pytest.raises(RuntimeError, match="multiples of 64"):
        eetq_b200.quant_weights(torch.zeros(64, 96, dtype=torch.float16, device=cuda), torch.int8, False)
    with pytest.raises(RuntimeError):
        eetq_b200.preprocess_weights(torch.zeros(100, 64, dtype=torch.int8, device=cuda))
