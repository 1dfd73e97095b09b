"""CPU tests: pin the oracle (oracle/w8a16_oracle.py and the scalar C port) against
  (1) the committed golden vectors, which were produced by the UNMODIFIED reference C++
      (cutlass_preprocessors.cc compiled into oracle/_ref/libref_oracle.so, tests/golden/make_golden.py), and
  (2) that same library live, when it is present (build container and GPU box; skipped otherwise).
"""
import pytest
import torch

from _util import golden_cases

CASES = golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_quantizer_matches_golden(oracle, case):
    q, scales, s32 = oracle.quantize(case["w"])
    assert torch.equal(q, case["q"])
    assert torch.equal(scales.view(torch.int16) if scales.dtype == torch.float16 else scales,
                       case["scales"].view(torch.int16) if case["scales"].dtype == torch.float16 else case["scales"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_ref_layout_matches_golden(oracle, case):
    assert torch.equal(oracle.ref_layout(case["q"]), case["w_ref"])
    assert torch.equal(oracle.ref_layout_inv(case["w_ref"]), case["q"])


@pytest.mark.parametrize("case", [c for c in CASES if "x" in c], ids=[c["name"] for c in CASES if "x" in c])
def test_gemm_matches_golden(oracle, case):
    y = oracle.gemm(case["x"], case["q"], case["scales"])
    assert torch.equal(y.view(torch.int16), case["y"].view(torch.int16))


def test_edge_case_semantics(oracle):
    """Reference behaviours worth naming: zero column -> (127, scale 0); +-amax -> 127 / -128; ties away from 0."""
    case = next(c for c in CASES if c["name"].endswith("edge"))
    q, s = case["q"], case["scales"]
    assert (q[:, 5] == 127).all() and s[5] == 0
    assert q[0, 6] == 127 and q[3, 6] == -128
    assert q[0, 7] == 127 and q[1, 7] == 2 and q[2, 7] == -3


def test_layout_spot_values(oracle):
    """q=-128 -> byte 0, q=0 -> byte 128, q=127 -> byte 255."""
    for v, b in ((-128, 0), (0, 128), (127, 255)):
        q = torch.full((64, 64), v, dtype=torch.int8)
        assert (oracle.ref_layout(q).view(torch.uint8) == b).all()


def test_b200_layout_roundtrip(oracle):
    q = torch.randint(-128, 128, (192, 128), dtype=torch.int8)
    w = oracle.b200_layout(q)
    assert w.shape == q.shape
    assert torch.equal(w.view(torch.uint8).view(128, 192).to(torch.int16) - 128, q.t().to(torch.int16))
    assert torch.equal(oracle.b200_layout_inv(w), q)


@pytest.mark.parametrize("shape", [(128, 64), (256, 192), (4096, 256)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_oracle_vs_live_reference(oracle, shape, dtype):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not built on this box")
    w = oracle.synth_weight(*shape, seed=11, dtype=dtype)
    unp, pro, sc = oracle.ref_quantize(w)
    q, s, _ = oracle.quantize(w)
    assert torch.equal(q, unp) and torch.equal(s, sc)
    assert torch.equal(oracle.ref_layout(q), pro)
    assert torch.equal(oracle.ref_preprocess(unp), pro)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_c_port_matches_torch_port(oracle, dtype):
    if oracle.port_lib() is None:
        pytest.skip("oracle/libw8a16_oracle.so not built")
    w = oracle.synth_weight(192, 128, seed=5, dtype=dtype)
    w[:, 3] = 0
    q, s, s32 = oracle.quantize(w)
    q2, s2, s322 = oracle.port_quantize(w)
    assert torch.equal(q, q2) and torch.equal(s, s2) and torch.equal(s32, s322)
    assert torch.equal(oracle.port_ref_layout(q), oracle.ref_layout(q))
    if dtype == torch.float16:
        x = oracle.synth_act(2, 192)
        y1 = oracle.port_gemm_f16(x, q, s)
        y2 = oracle.gemm(x, q, s)
        assert oracle.norm_rel_err(y1, y2) <= 1e-3


def test_nan_entries_are_ignored_by_absmax(oracle):
    """std::max(acc, NaN) keeps acc in the reference (cutlass_preprocessors.cc:626); a NaN weight quantises to 127."""
    w = oracle.synth_weight(64, 64, seed=6)
    w[3, 2] = float("nan")
    q, s, _ = oracle.quantize(w)
    assert not torch.isnan(s).any() and q[3, 2] == 127
    if oracle.ref_lib() is not None:
        unp, _, sc = oracle.ref_quantize(w)
        assert torch.equal(q, unp) and torch.equal(s.view(torch.int16), sc.view(torch.int16))
    if oracle.port_lib() is not None:
        q2, s2, _ = oracle.port_quantize(w)
        assert torch.equal(q, q2) and torch.equal(s.view(torch.int16), s2.view(torch.int16))


def test_oracle_3d_experts(oracle):
    w = torch.stack([oracle.synth_weight(64, 64, seed=s) for s in (1, 2)])
    q, s, _ = oracle.quantize(w)
    for e in range(2):
        qe, se, _ = oracle.quantize(w[e])
        assert torch.equal(q[e], qe) and torch.equal(s[e], se)


def test_identity_gemm_known_answer(oracle):
    """w8_a16_gemm(I_K, Wq, s)[k, n] == fp16(fp16(q[k,n]) * s[n]) exactly."""
    w = oracle.synth_weight(64, 64, seed=9)
    q, s, _ = oracle.quantize(w)
    y = oracle.gemm(torch.eye(64, dtype=torch.float16), q, s)
    assert torch.equal(y, oracle.dequantize(q, s))


def test_layout_roundtrips_property(oracle):
    """Size-independent properties: both layouts are bijections on the K*N bytes; the reference layout only moves and
    biases bytes (multiset of values preserved); column blocks of the b200 layout are contiguous byte ranges."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(kt=st.integers(1, 4), nt=st.integers(1, 4), seed=st.integers(0, 2**16))
    def prop(kt, nt, seed):
        K, N = 64 * kt, 64 * nt
        g = torch.Generator().manual_seed(seed)
        q = torch.randint(-128, 128, (K, N), generator=g, dtype=torch.int8)
        r = oracle.ref_layout(q)
        assert torch.equal(oracle.ref_layout_inv(r), q)
        assert torch.equal(torch.sort(r.view(torch.uint8).flatten().to(torch.int16) - 128).values,
                           torch.sort(q.flatten().to(torch.int16)).values)
        b = oracle.b200_layout(q)
        assert torch.equal(oracle.b200_layout_inv(b), q)
        n0 = 64 * (seed % nt)
        shard = oracle.b200_layout(q[:, n0:n0 + 64].contiguous())
        assert torch.equal(b.flatten()[n0 * K:(n0 + 64) * K], shard.flatten())

    prop()


def test_quantizer_properties(oracle):
    """Per-column independence and scale equivariance of the reference quantiser: scaling a column by a power of two
    scales its scale and leaves q unchanged; permuting rows permutes q."""
    w = oracle.synth_weight(128, 64, seed=21, dtype=torch.float32)
    q, s, _ = oracle.quantize(w)
    w2 = w.clone()
    w2[:, 7] *= 4.0
    q2, s2, _ = oracle.quantize(w2)
    assert torch.equal(q2, q) and torch.equal(s2[7], s[7] * 4) and torch.equal(s2[:7], s[:7])
    perm = torch.randperm(128, generator=torch.Generator().manual_seed(0))
    qp, sp, _ = oracle.quantize(w[perm])
    assert torch.equal(qp, q[perm]) and torch.equal(sp, s)
    assert int(q.abs().max()) >= 127          # every column uses the full int8 range by construction
