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


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_plan_llama_shapes(world):
    """Every rank's block of every (fused) Llama-2-7B / 13B linear stays a multiple of 64 rows (kernel constraint), the
    blocks tile the full matrices, and the exchange indices of a decode step are all distinct."""
    from eetq_b200.decode import LLAMA2_7B, LLAMA2_13B, shard_plan

    for s in (LLAMA2_7B, LLAMA2_13B):
        heads, hidden, inter, vocab = [], [], [], []
        for r in range(world):
            p = shard_plan(s, r, world)
            heads.append(p["heads"]); hidden.append(p["hidden"]); inter.append(p["inter"]); vocab.append(p["vocab"])
            hl = p["heads"][1] - p["heads"][0]
            for n in (3 * hl * s.head_dim, p["hidden"][1] - p["hidden"][0], 2 * (p["inter"][1] - p["inter"][0])):
                assert n % 64 == 0, (s.name, world, n)
        for spans, total in ((heads, s.heads), (hidden, s.hidden), (inter, s.inter), (vocab, s.vocab)):
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        p = shard_plan(s, 0, world)
        idx = [0] + [f(l) for l in range(s.layers) for f in (p["attn"], p["x2"], p["act"], p["x_out"])] + [p["cand"]]
        assert len(set(idx)) == len(idx) and max(idx) < p["per_step"]
        assert all(p["x_in"](l + 1) == p["x_out"](l) for l in range(s.layers - 1)) and p["x_in"](0) == 0


def test_shard_plan_rejects_impossible_splits():
    from eetq_b200.decode import shard_plan

    with pytest.raises(ValueError):
        shard_plan(LlamaShape(hidden=4096, inter=11008, layers=2, heads=32), 0, 3)
    with pytest.raises(ValueError):
        shard_plan(TINY, 0, 4)      # 2 heads cannot be split 4-way
