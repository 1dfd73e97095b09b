"""CPU tests of the decode harness' host-side pieces (no kernels): HF-convention RoPE, the random-init skeleton with
HF sub-module names, and the properties the sharded decoder relies on."""
import math

import pytest
import torch

from eetq_b200.decode import LlamaShape, LlamaSkeleton, apply_rope, rope_tables

TINY = LlamaShape(hidden=256, inter=512, layers=2, heads=2, vocab=128, name="tiny-cpu")


def test_rope_tables_match_hf_formula():
    cos, sin = rope_tables(TINY, 16, "cpu", torch.float32)
    D = TINY.head_dim
    assert cos.shape == (16, D // 2)
    for pos in (0, 3, 15):
        for i in (0, 5, D // 2 - 1):
            ang = pos / (TINY.theta ** (2 * i / D))
            assert math.isclose(cos[pos, i].item(), math.cos(ang), abs_tol=1e-5)
            assert math.isclose(sin[pos, i].item(), math.sin(ang), abs_tol=1e-5)


def test_apply_rope_is_rotate_half_convention():
    T, H, D = 4, 2, TINY.head_dim
    x = torch.randn(T, H, D)
    cos, sin = rope_tables(TINY, T, "cpu", torch.float32)
    y = apply_rope(x, cos, sin)
    half = D // 2
    # pairs (i, i + D/2) rotate by the position-dependent angle; norms are preserved
    exp_lo = x[..., :half] * cos[:, None, :] - x[..., half:] * sin[:, None, :]
    exp_hi = x[..., half:] * cos[:, None, :] + x[..., :half] * sin[:, None, :]
    assert torch.allclose(y[..., :half], exp_lo, atol=1e-6) and torch.allclose(y[..., half:], exp_hi, atol=1e-6)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-5)
    assert torch.equal(apply_rope(x[:1], cos[:1], sin[:1]), x[:1])      # position 0 is the identity


def test_skeleton_has_hf_names_and_eet_quantize_targets():
    from eetq_b200 import find_layers

    m = LlamaSkeleton(TINY, device="cpu", dtype=torch.float32, seed=1)
    names = sorted(find_layers(m))
    assert "lm_head" not in names
    assert "model.layers.0.self_attn.q_proj" in names and "model.layers.1.mlp.down_proj" in names
    assert len(names) == 7 * TINY.layers                                    # q k v o gate up down per layer
    logits = m(torch.tensor([1, 2, 3, 4]))
    assert logits.shape == (4, TINY.vocab) and torch.isfinite(logits).all()
    # causal: the logits of a prefix do not depend on later tokens
    l2 = m(torch.tensor([1, 2, 9, 9]))
    assert torch.allclose(logits[:2], l2[:2], atol=1e-5)


def test_same_seed_gives_identical_models_on_every_rank():
    a = LlamaSkeleton(TINY, device="cpu", dtype=torch.float32, seed=7)
    b = LlamaSkeleton(TINY, device="cpu", dtype=torch.float32, seed=7)
    for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert na == nb and torch.equal(pa, pb)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_llama_shapes_shard_in_multiples_of_64(world):
    """Column shards of every fused Llama-2-7B / 13B linear must stay multiples of 64 rows (kernel constraint)."""
    from eetq_b200.decode import LLAMA2_7B, LLAMA2_13B

    for s in (LLAMA2_7B, LLAMA2_13B):
        for n in (3 * s.hidden, s.hidden, 2 * s.inter, s.hidden):
            assert n % (64 * world) == 0, (s.name, n, world)
