"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/eetq_b200.h declares,
and rejects bad arguments with an error code + message (no compute is launched without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from eetq_b200 import _cabi

    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _cabi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "eetq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eetq_b200_\w+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/eetq_b200.h but not exported"


def test_binding_table_covers_header():
    from eetq_b200 import _cabi

    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_version_and_error_string(lib):
    assert lib.eetq_b200_version() == 100
    assert isinstance(lib.eetq_b200_last_error(), bytes)


def test_null_pointer_is_einval(lib):
    rc = lib.eetq_b200_w8a16_gemm(None, None, None, None, None, 1, 64, 64, 0, None, 0, None)
    assert rc == -1
    assert b"null" in lib.eetq_b200_last_error()


def test_shape_constraints_are_einval(lib):
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.cast(buf, ctypes.c_void_p)
    # K not a multiple of 64 (reference: cutlass_preprocessors.cc:230 / fpA_intB_gemm_template.h:139-142)
    rc = lib.eetq_b200_w8a16_gemm(p, p, p, None, p, 1, 64, 100, 0, None, 0, None)
    assert rc == -1 and b"multiples of 64" in lib.eetq_b200_last_error()
    rc = lib.eetq_b200_pack(p, 64, 96, p, None)
    assert rc == -1
    # bad dtype
    rc = lib.eetq_b200_w8a16_gemm(p, p, p, None, p, 1, 64, 64, 7, None, 0, None)
    assert rc == -1 and b"dtype" in lib.eetq_b200_last_error()


def test_empty_batch_is_ok_without_gpu(lib):
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.eetq_b200_w8a16_gemm(p, p, p, None, p, 0, 64, 64, 0, None, 0, None) == 0


def test_workspace_bytes(lib):
    assert lib.eetq_b200_workspace_bytes(1, 4096, 4096) >= 0
    small = lib.eetq_b200_workspace_bytes(16, 4096, 4096)
    big = lib.eetq_b200_workspace_bytes(1024, 4096, 4096)
    # stream-K: 4 KiB of flags + one fp32 partial tile [tokens per tile x 128] per persistent CTA (<= one CTA per SM)
    assert 4096 < small <= 4096 + 1024 * 16 * 128 * 4
    assert small < big <= 4096 + 1024 * 256 * 128 * 4
    assert lib.eetq_b200_workspace_bytes(0, 4096, 4096) == 0


def test_python_surface_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import eetq_b200

    x = torch.zeros(1, 64, dtype=torch.float16)
    w = torch.zeros(64, 64, dtype=torch.int8)
    s = torch.zeros(64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="CUDA"):
        eetq_b200.w8_a16_gemm(x, w, s)
    with pytest.raises(RuntimeError, match="CUDA|sm_100"):
        eetq_b200.quant_weights(torch.zeros(64, 64, dtype=torch.float16), torch.int8, False)


def test_header_is_plain_c_and_a_c_host_links(lib, tmp_path):
    """include/eetq_b200.h compiles as C99 (the boundary is a C ABI, not a C++ one) and a C program linked against the
    library can call it (validation paths only: no device needed)."""
    import subprocess

    from eetq_b200 import _cabi

    src = os.path.join(ROOT, "tests", "host", "abi_check.c")
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-o", exe, src, "-L" + libdir, "-leetq_b200",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_decode_entry_points_refuse_to_launch_without_sm100(lib):
    """every forward entry point checks the device before launching (error code + message, never a launch failure)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.eetq_b200_rmsnorm(p, 64, p, p, 64, 1, 64, ctypes.c_float(1e-5), 0, None) in (-2, -3)
    assert lib.eetq_b200_layernorm_forward(p, p, p, 1, 64, ctypes.c_float(1e-5), None) in (-2, -3)
    assert lib.eetq_b200_silu_mul(p, 128, p, 64, 1, 64, 0, None) in (-2, -3)
    assert lib.eetq_b200_decode_embed(p, p, p, 64, None, 0, None) in (-2, -3)
    assert b"device" in lib.eetq_b200_last_error()
