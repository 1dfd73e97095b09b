"""GPU parity tests of the packed-int4 path, through the C ABI: quantiser and layout kernels
bit-exact against the committed golden vectors (reference C++ output), the oracle and -- when oracle/_ref is present -- the
live reference; the w4a16 forward within the w8a16 tolerance (1e-3 norm-relative for fp16 against the oracle on IDENTICAL
quantised weights, 8e-3 for the bf16 extension) and at least as close to the exact sum as the reference's own Int4b decode
kernel rebuilt for sm_100a."""
import ctypes
import os

import pytest
import torch

import eetq_b200
from _util import golden_cases_int4

pytestmark = pytest.mark.gpu
CASES = golden_cases_int4()
TOL = {torch.float16: 1e-3, torch.bfloat16: 8e-3}
LLAMA7B = [(4096, 4096), (4096, 11008), (11008, 4096)]


def _bits(t):
    return t.view(torch.int16) if t.dtype in (torch.float16, torch.bfloat16) else t.view(torch.int32)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_quant_weights_int4_matches_reference_golden(cuda, oracle, case):
    unp, pro, sc = eetq_b200.quant_weights(case["w"].to(cuda), torch.quint4x2, True)
    K, N = case["w"].shape
    assert unp.shape == (K, N // 2) and pro.shape == (K, N // 2) and unp.dtype == torch.int8      # wrapper.cu:54-63
    assert torch.equal(unp.cpu(), case["q4"])
    assert torch.equal(_bits(sc.cpu()), _bits(case["scales"]))
    q = oracle.unpack_int4(case["q4"])
    assert torch.equal(pro.cpu(), oracle.b200_layout4(q))
    # layout converters reproduce the reference's processed bytes both ways
    assert torch.equal(eetq_b200.to_ref_checkpoint_weight4(pro).cpu(), case["w4_ref"])
    assert torch.equal(eetq_b200.convert_ref_checkpoint_weight4(case["w4_ref"].to(cuda)).cpu(), pro.cpu())
    # preprocess_weights(is_int4=True) and its inverse
    assert torch.equal(eetq_b200.preprocess_weights(case["q4"].to(cuda), True).cpu(), pro.cpu())
    assert torch.equal(eetq_b200.unpack_weights4(pro).cpu(), case["q4"])
    if "x" in case:
        y = eetq_b200.w4_a16_gemm(case["x"].to(cuda), pro, sc)
        assert oracle.norm_rel_err(y.cpu(), case["y"]) <= 1e-3


def test_quant_weights_int4_cpu_contract(cuda, oracle):
    """CPU in, CPU out; 2-tuple / 3-tuple order (fpA_intB_gemm_wrapper.cu:33, :101-106); 3-D = one matrix per expert."""
    w = oracle.synth_weight(128, 64, seed=2)
    out2 = eetq_b200.quant_weights(w, torch.quint4x2, False)
    out3 = eetq_b200.quant_weights(w, torch.quint4x2, True)
    assert len(out2) == 2 and len(out3) == 3 and all(not t.is_cuda for t in out2 + out3)
    packed, s, _, q = oracle.quantize4(w)
    assert torch.equal(out3[0], packed) and torch.equal(out3[2], s) and torch.equal(out2[0], out3[1])
    w3 = torch.stack([oracle.synth_weight(128, 64, seed=i) for i in range(3)])
    u3, p3, s3 = eetq_b200.quant_weights(w3.to(cuda), torch.quint4x2, True)
    assert u3.shape == (3, 128, 32) and s3.shape == (3, 64)
    for e in range(3):
        pk, se, _, qe = oracle.quantize4(w3[e])
        assert torch.equal(u3[e].cpu(), pk) and torch.equal(s3[e].cpu(), se) and torch.equal(p3[e].cpu(), oracle.b200_layout4(qe))


@pytest.mark.parametrize("shape", LLAMA7B)
def test_quant_weights_int4_full_size_bit_exact(cuda, oracle, shape):
    w = oracle.synth_weight(*shape, seed=1000)
    unp, pro, sc = eetq_b200.quant_weights(w.to(cuda), torch.quint4x2, True)
    packed, s, _, q = oracle.quantize4(w)
    assert torch.equal(unp.cpu(), packed) and torch.equal(sc.cpu(), s)
    assert torch.equal(pro.cpu(), oracle.b200_layout4(q))
    assert torch.equal(eetq_b200.to_ref_checkpoint_weight4(pro).cpu(), oracle.ref_layout4(q))
    if oracle.ref_lib() is not None and hasattr(oracle.ref_lib(), "ref_quant4_fp16") and shape == (4096, 4096):
        r_unp, r_pro, r_sc = oracle.ref_quantize4(w)
        assert torch.equal(unp.cpu(), r_unp) and torch.equal(sc.cpu(), r_sc)
        assert torch.equal(eetq_b200.to_ref_checkpoint_weight4(pro).cpu(), r_pro)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_quant_weights_int4_other_dtypes(cuda, oracle, dtype):
    w = oracle.synth_weight(256, 192, seed=4, dtype=torch.float32).to(dtype)
    w[:, 9] = 0
    unp, pro, sc = eetq_b200.quant_weights(w.to(cuda), torch.quint4x2, True)
    packed, _, s32, q = oracle.quantize4(w.float())        # arithmetic is fp32 whatever the storage dtype
    assert torch.equal(unp.cpu(), packed) and torch.equal(pro.cpu(), oracle.b200_layout4(q))
    assert torch.equal(_bits(sc.cpu()), _bits(s32.to(dtype)))


def make4(oracle, cuda, K, N, seed=1000, dtype=torch.float16):
    _, s, _, q = oracle.quantize4(oracle.synth_weight(K, N, seed))
    return q, s.to(dtype), oracle.b200_layout4(q).to(cuda), s.to(dtype).to(cuda)


def ref_out(oracle, x, q, s, bias=None):
    if x.dtype == torch.float16:
        return oracle.gemm(x, q, s, bias)
    y = (x.float() @ q.float()) * s.float()
    if bias is not None:
        y = y + bias.float()
    return y.to(x.dtype)


# K = 4096 / 8192 / 11008 / 16384 walk the 1..4 register-resident k-iterations, 32768 the L1 re-read path
@pytest.mark.parametrize("K,N", LLAMA7B + [(64, 64), (8192, 1024), (16384, 512), (32768, 256), (1024, 1728)])
@pytest.mark.parametrize("M", [1, 2, 3, 4])
def test_w4a16_gemv_matches_oracle(cuda, oracle, K, N, M):
    q, s, wq, sd = make4(oracle, cuda, K, N)
    x = oracle.synth_act(M, K)
    y = eetq_b200.w4_a16_gemm(x.to(cuda), wq, sd)
    assert y.shape == (M, N) and y.dtype == torch.float16
    assert oracle.norm_rel_err(y.cpu(), ref_out(oracle, x, q, s)) <= TOL[torch.float16]


@pytest.mark.parametrize("M", [1, 3])
def test_w4a16_gemv_bf16_and_bias(cuda, oracle, M):
    K, N = 4096, 1024
    q, s, wq, sd = make4(oracle, cuda, K, N, dtype=torch.bfloat16)
    x = oracle.synth_act(M, K, dtype=torch.bfloat16)
    bias = (torch.randn(N) * 0.1).to(torch.bfloat16)
    y = eetq_b200.w4_a16_gemm(x.to(cuda), wq, sd, bias.to(cuda))
    assert oracle.norm_rel_err(y.cpu(), ref_out(oracle, x, q, s, bias)) <= TOL[torch.bfloat16]
    q, s, wq, sd = make4(oracle, cuda, K, N)
    x = oracle.synth_act(M, K)
    bias = (torch.randn(N) * 0.1).half()
    y = eetq_b200.w4_a16_gemm(x.to(cuda), wq, sd, bias.to(cuda))
    assert oracle.norm_rel_err(y.cpu(), ref_out(oracle, x, q, s, bias)) <= TOL[torch.float16]


@pytest.mark.parametrize("K,N", [(4096, 4096), (11008, 4096), (1024, 1728)])
@pytest.mark.parametrize("M", [5, 8, 16, 64, 300])
def test_w4a16_gemm_batched_matches_oracle(cuda, oracle, K, N, M):
    """M <= 8: mma.sync streaming kernel; above: nibbles widened to the b200 int8 layout in the workspace, then the tcgen05 kernel
    (same arithmetic as w8a16)."""
    q, s, wq, sd = make4(oracle, cuda, K, N)
    x = oracle.synth_act(M, K)
    y = eetq_b200.w4_a16_gemm(x.to(cuda).view(1, M, K), wq, sd)
    assert y.shape == (1, M, N)
    assert oracle.norm_rel_err(y.cpu().view(M, N), ref_out(oracle, x, q, s)) <= TOL[torch.float16]
    # identity known answer: x = I  ->  y = fp16(fp16(q) * s) exactly
    if K == 1024:
        eye = torch.eye(K, dtype=torch.float16, device=cuda)
        assert torch.equal(eetq_b200.w4_a16_gemm(eye, wq, sd).cpu(), oracle.dequantize(q, s))


def test_w4a16_needs_workspace_for_large_m(cuda, oracle):
    from eetq_b200 import _cabi
    q, s, wq, sd = make4(oracle, cuda, 256, 128)
    x = oracle.synth_act(16, 256).to(cuda)
    y = torch.empty(16, 128, dtype=torch.float16, device=cuda)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _cabi.lib().eetq_b200_w4a16_gemm(vp(x), 256, vp(wq), vp(sd), None, vp(y), 128, 16, 128, 256, _cabi.F16, None, 0, 0, None)
    assert rc == -4 and b"workspace" in _cabi.lib().eetq_b200_last_error()
    assert _cabi.lib().eetq_b200_w4a16_workspace_bytes(8, 128, 256) == 0      # up to 8 rows stream through the mma.sync kernel


@pytest.mark.parametrize("K,N", LLAMA7B)
@pytest.mark.parametrize("M", [1, 4])
def test_w4a16_against_live_reference_gemv(cuda, oracle, K, N, M):
    """Same-box parity against the REFERENCE Int4b decode kernel (weightOnlyBatchedGemv rebuilt for sm_100a from unmodified
    sources, kernel.h:68-116) fed with reference-layout int4 weights produced by OUR converter from OUR quantiser output."""
    path = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "libref_gemv.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_gemv.so not built")
    lib = ctypes.CDLL(path)
    if not hasattr(lib, "ref_w4a16_gemv"):
        pytest.skip("libref_gemv.so predates the int4 shim")
    lib.ref_w4a16_gemv.restype = ctypes.c_int
    q, s, wq, sd = make4(oracle, cuda, K, N)
    w_ref = eetq_b200.to_ref_checkpoint_weight4(wq)
    assert torch.equal(w_ref.cpu(), oracle.ref_layout4(q))
    x = oracle.synth_act(M, K).to(cuda)
    y_ref = torch.empty(M, N, dtype=torch.float16, device=cuda)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.ref_w4a16_gemv(vp(x), vp(w_ref), vp(sd), vp(y_ref), M, N, K, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0
    y = eetq_b200.w4_a16_gemm(x, wq, sd)
    y_or = oracle.gemm(x.cpu(), q, s)
    e_ours, e_ref = oracle.norm_rel_err(y.cpu(), y_or), oracle.norm_rel_err(y_ref.cpu(), y_or)
    assert e_ours <= 1e-3
    assert oracle.norm_rel_err(y.cpu(), y_ref.cpu()) <= 5e-3, (e_ours, e_ref)   # reference's fp16 accumulation dominates
    assert e_ours <= e_ref + 1e-4


@pytest.mark.parametrize("K,N", LLAMA7B + [(128, 64), (8192, 1024), (1024, 1728)])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8])
def test_w4a16_mma_stream_kernel_matches_oracle(cuda, oracle, K, N, M):
    """The mma.sync streaming kernel on int4 nibbles (gemv_mma.cu), forced through the flag for every row count it takes."""
    from eetq_b200 import _cabi
    q, s, wq, sd = make4(oracle, cuda, K, N)
    x = oracle.synth_act(M, K)
    y = eetq_b200.w4_a16_gemm(x.to(cuda), wq, sd, flags=_cabi.FLAG_FORCE_MMA)
    assert oracle.norm_rel_err(y.cpu(), ref_out(oracle, x, q, s)) <= TOL[torch.float16]
    if (K, N) == (1024, 1728):
        xb = oracle.synth_act(M, K, dtype=torch.bfloat16)
        bias = (torch.randn(N) * 0.1).to(torch.bfloat16)
        yb = eetq_b200.w4_a16_gemm(xb.to(cuda), wq, sd.to(torch.bfloat16), bias.to(cuda), flags=_cabi.FLAG_FORCE_MMA)
        assert oracle.norm_rel_err(yb.cpu(), ref_out(oracle, xb, q, s.to(torch.bfloat16), bias)) <= TOL[torch.bfloat16]


@pytest.mark.parametrize("M", [2, 4, 6, 8])
def test_w4a16_k_not_multiple_of_128(cuda, oracle, M):
    """K % 128 != 0: the mma.sync kernel does not take int4 there -> SIMT up to 4 rows, widen + tcgen05 above."""
    K, N = 192, 128
    q, s, wq, sd = make4(oracle, cuda, K, N)
    x = oracle.synth_act(M, K)
    y = eetq_b200.w4_a16_gemm(x.to(cuda), wq, sd)
    assert oracle.norm_rel_err(y.cpu(), ref_out(oracle, x, q, s)) <= TOL[torch.float16]


def test_w4a16_module_end_to_end(cuda, oracle):
    """W4A16Linear.from_torch = quant_weights(w.T, quint4x2) + fused-bias w4_a16_gemm; reference-layout int4 bytes load through the
    state-dict hook and export writes them back."""
    import torch.nn as nn
    from eetq_b200.modules.qlinear import export_reference_state_dict
    lin = nn.Linear(1024, 512, bias=True).half().to(cuda)
    ql = eetq_b200.W4A16Linear.from_torch(lin)
    packed, s, _, q = oracle.quantize4(lin.weight.detach().t().contiguous().cpu())
    assert torch.equal(ql.qweight.cpu(), oracle.b200_layout4(q)) and torch.equal(ql.weight_scales.cpu(), s)
    for M in (1, 3, 40):
        x = oracle.synth_act(M, 1024).to(cuda)
        y_ref = oracle.gemm(x.cpu(), q, s, lin.bias.detach().cpu())
        assert oracle.norm_rel_err(ql(x).cpu(), y_ref) <= 1e-3
    # a state dict WITHOUT the marker holds the reference's processed int4 bytes
    ref_sd = {"qweight": oracle.ref_layout4(q), "weight_scales": s, "bias": lin.bias.detach().cpu()}
    q2 = eetq_b200.W4A16Linear(1024, 512, bias=True, dev=cuda)
    q2.load_state_dict(ref_sd)
    assert torch.equal(q2.qweight.cpu(), ql.qweight.cpu())
    out = export_reference_state_dict(q2)
    assert torch.equal(out["qweight"].cpu(), ref_sd["qweight"]) and "weight_layout" not in out
    # eet_quantize(bits=4) on a small model
    m = nn.Sequential(nn.Linear(256, 128), nn.Linear(128, 64)).half().to(cuda)
    x = oracle.synth_act(2, 256).to(cuda)
    y_fp = m(x)
    eetq_b200.eet_quantize(m, bits=4)
    assert isinstance(m[0], eetq_b200.W4A16Linear)
    assert (m(x) - y_fp).abs().max() <= 0.25 * y_fp.abs().max()     # int4 quantisation noise, two layers deep
