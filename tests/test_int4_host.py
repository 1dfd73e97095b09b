"""CPU tests of the packed-int4 widening of the hot path: the oracle against the committed golden
vectors (reference C++ output), against the live compiled reference when oracle/_ref is present, the scalar C restatement
against the vectorised one, and the product's OWN int4 index arithmetic (eetq_b200/csrc/int4_layout.cuh -- the functions
the CUDA kernels call) compiled for the host and compared with the oracle."""
import ctypes
import os
import subprocess

import pytest
import torch

from _util import golden_cases_int4

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = golden_cases_int4()
IDS = [c["name"] for c in CASES]


def _bits(t):
    return t.view(torch.int16) if t.dtype == torch.float16 else t.view(torch.int32)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_matches_reference_golden(oracle, case):
    packed, s, s32, q = oracle.quantize4(case["w"])
    assert torch.equal(packed, case["q4"])
    assert torch.equal(_bits(s), _bits(case["scales"]))
    assert torch.equal(oracle.unpack_int4(case["q4"]), q) and torch.equal(oracle.pack_int4(q), case["q4"])
    assert torch.equal(oracle.ref_layout4(q), case["w4_ref"])
    assert torch.equal(oracle.ref_layout4_inv(case["w4_ref"]), q)
    assert torch.equal(oracle.b200_layout4_inv(oracle.b200_layout4(q)), q)
    if "x" in case:
        assert torch.equal(oracle.gemm(case["x"], q, s), case["y"])


def test_edge_semantics_in_golden(oracle):
    """What the reference does at the corners (cutlass_preprocessors.cc:655-660), read back from its own output."""
    case = next(c for c in CASES if c["name"].endswith("edge"))
    q = oracle.unpack_int4(case["q4"])
    assert (q[:, 5] == -8).all() and case["scales"][5] == 0       # all-zero column: NaN -> INT_MIN -> -8
    assert (q[:, 6] == 7).sum() == 127 and q[3, 6] == -8           # +amax -> 8 clamps to 7, -amax -> -8
    assert q[1, 7] == 2 and q[2, 7] == -3                          # ties round half away from zero
    assert q[9, 11] == -8                                          # NaN weight


@pytest.mark.parametrize("shape", [(64, 64), (320, 128), (1024, 256)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_oracle_matches_live_reference(oracle, shape, dtype):
    lib = oracle.ref_lib()
    if lib is None or not hasattr(lib, "ref_quant4_fp16"):
        pytest.skip("oracle/_ref/libref_oracle.so (with int4 shims) not built")
    w = oracle.synth_weight(*shape, seed=sum(shape), dtype=dtype)
    w[:, 1] = 0
    unp, pro, sc = oracle.ref_quantize4(w)
    packed, s, _, q = oracle.quantize4(w)
    assert torch.equal(unp, packed) and torch.equal(_bits(sc), _bits(s))
    assert torch.equal(oracle.ref_layout4(q), pro)
    assert torch.equal(oracle.ref_preprocess4(packed), pro)


def test_c_port_matches_python_oracle(oracle):
    lib = oracle.port_lib()
    if lib is None or not hasattr(lib, "oracle_quantize4_f16"):
        pytest.skip("oracle/libw8a16_oracle.so not built")
    P, sz = oracle._ptr, ctypes.c_size_t
    lib.oracle_ref_layout4.restype = ctypes.c_int
    K, N = 192, 128
    w = oracle.synth_weight(K, N, seed=5)
    w[:, 3] = 0
    w[5, 7] = float("nan")
    packed, s, s32, q = oracle.quantize4(w)
    cp, cs, cs32 = torch.empty(K, N // 2, dtype=torch.uint8), torch.empty(N, dtype=torch.float16), torch.empty(N)
    lib.oracle_quantize4_f16(P(w), sz(K), sz(N), P(cp), P(cs), P(cs32))
    assert torch.equal(cp.view(torch.int8), packed) and torch.equal(_bits(cs), _bits(s)) and torch.equal(cs32, s32)
    out = torch.empty(K, N // 2, dtype=torch.uint8)
    assert lib.oracle_ref_layout4(P(cp), sz(K), sz(N), P(out)) == 0
    assert torch.equal(out.view(torch.int8), oracle.ref_layout4(q))
    lib.oracle_b200_layout4(P(cp), sz(K), sz(N), P(out))
    assert torch.equal(out.view(torch.int8), oracle.b200_layout4(q))
    keep = [n for n in range(N) if n != 3]      # column 3 has scale 0 and q = -8: 0 * anything, skip only for NaN hygiene
    x = oracle.synth_act(2, K)
    y = torch.empty(2, N, dtype=torch.float16)
    lib.oracle_gemm4_f16(P(x), P(cp), P(cs), P(y), sz(2), sz(N), sz(K))
    assert oracle.norm_rel_err(y[:, keep], oracle.gemm(x, q, s)[:, keep]) <= 1e-3


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    """eetq_b200/csrc/int4_layout.cuh compiled for the HOST through tests/host/int4_host.cc"""
    so = str(tmp_path_factory.mktemp("int4host") / "int4_host.so")
    src = os.path.join(ROOT, "tests", "host", "int4_host.cc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, src], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("shape", [(64, 64), (128, 192), (512, 320)])
def test_kernel_index_arithmetic_on_host(oracle, host_lib, shape):
    """pack4 / unpack4 / from_ref4 / to_ref4 / widen4to8 / word packing: the kernels' per-word functions, looped on the CPU."""
    K, N = shape
    P, i64 = oracle._ptr, ctypes.c_int64
    _, _, _, q = oracle.quantize4(oracle.synth_weight(K, N, seed=K + N))
    packed, b200, ref = oracle.pack_int4(q), oracle.b200_layout4(q), oracle.ref_layout4(q)

    def run(mode, src):
        dst = torch.empty(K, N // 2, dtype=torch.int8)
        assert host_lib.host_nibble_layout(mode, P(src.contiguous()), i64(K), i64(N), P(dst)) == 0
        return dst

    assert torch.equal(run(0, packed), b200)
    assert torch.equal(run(1, b200), packed)
    assert torch.equal(run(2, ref), b200)
    assert torch.equal(run(3, b200), ref)
    wide = torch.empty(K, N, dtype=torch.int8)
    host_lib.host_widen4to8(P(b200), i64(K * N // 8), P(wide))
    assert torch.equal(wide, oracle.b200_layout(q))
    u = ((q.t().contiguous().to(torch.int16) + 8) & 15).to(torch.uint8)
    words = torch.empty(K, N // 2, dtype=torch.int8)
    host_lib.host_pack_words(P(u), i64(K * N // 8), P(words))
    assert torch.equal(words, b200)


def test_int4_entry_points_validate_without_gpu():
    from eetq_b200 import _cabi

    L = _cabi.lib()
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.cast(buf, ctypes.c_void_p)
    p2 = ctypes.c_void_p(p.value + 32768)
    assert L.eetq_b200_quantize4(None, 0, 64, 64, None, None, None, None, None) == -1
    assert L.eetq_b200_pack4(p, 64, 96, p2, None) == -1 and b"multiples of 64" in L.eetq_b200_last_error()
    assert L.eetq_b200_pack4(p, 64, 64, p, None) == -1 and b"in-place" in L.eetq_b200_last_error()
    assert L.eetq_b200_w4a16_gemm(p, 64, p, p, None, p, 64, 0, 64, 64, 0, None, 0, 0, None) == 0      # empty batch
    assert L.eetq_b200_w4a16_gemm(p, 64, p, p, None, p, 64, 1, 64, 64, 7, None, 0, 0, None) == -1     # bad dtype
    assert L.eetq_b200_w4a16_workspace_bytes(4, 4096, 4096) == 0      # SIMT rows never need one
    need = L.eetq_b200_w4a16_workspace_bytes(64, 4096, 4096)
    assert need >= 4096 * 4096 + L.eetq_b200_workspace_bytes(64, 4096, 4096)
