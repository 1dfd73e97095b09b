"""GPU tests of the decode-side kernels (fused GEMV prologues/epilogues, RoPE/KV append + attention, lm_head + arg-max,
the reference-named glue ops) and of the decode harness against a plain PyTorch forward of the same quantised model."""
import ctypes
import math

import pytest
import torch

import eetq_b200
from eetq_b200 import _cabi
from eetq_b200.decode import LlamaShape, LlamaSkeleton, W8A16LlamaDecoder, apply_rope, rope_tables

pytestmark = pytest.mark.gpu

SMALL = LlamaShape(hidden=512, inter=1408, layers=2, heads=4, vocab=1024, name="tiny")


def vp(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemv_fused(x, ldx, wq, sd, y, N, K, *, M=1, norm_w=None, eps=1e-5, xmode=0, epi=0, residual=None, pdl=0):
    o = _cabi.GemvOpts()
    o.norm_weight = 0 if norm_w is None else norm_w.data_ptr()
    o.eps, o.xmode, o.epi = eps, xmode, epi
    o.residual = 0 if residual is None else residual.data_ptr()
    o.ldr = N
    rc = _cabi.lib().eetq_b200_w8a16_gemv_fused(vp(x), ldx, vp(wq), vp(sd), None, vp(y), y.shape[-1], M, N, K, _cabi.F16, ctypes.byref(o),
                                                pdl, stream())
    _cabi.check(rc, "gemv_fused")
    torch.cuda.synchronize()


@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 12288), (11008, 4096), (512, 1408)])
def test_fused_rmsnorm_prologue_and_residual(cuda, oracle, K, N):
    w = oracle.synth_weight(K, N, 5)
    q, s, _ = oracle.quantize(w)
    wq, sd = oracle.b200_layout(q).to(cuda), s.to(cuda)
    x = (torch.randn(K) * 1.5).half().to(cuda)
    nw = (1 + 0.1 * torch.randn(K)).half().to(cuda)
    res = torch.randn(N).half().to(cuda)
    y = torch.empty(N, dtype=torch.float16, device=cuda)
    gemv_fused(x, K, wq, sd, y, N, K, norm_w=nw, xmode=1, residual=res)
    xf = x.float()
    xn = (nw * (xf * torch.rsqrt(xf.pow(2).mean() + 1e-5)).half()).cpu()          # HF LlamaRMSNorm arithmetic
    ref = oracle.gemm(xn[None], q, s)[0] + res.cpu()
    assert oracle.norm_rel_err(y.cpu(), ref) <= 1.5e-3


def test_fused_silu_mul_prologue(cuda, oracle):
    K, N = 11008, 4096
    w = oracle.synth_weight(K, N, 6)
    q, s, _ = oracle.quantize(w)
    wq, sd = oracle.b200_layout(q).to(cuda), s.to(cuda)
    gu = torch.randn(2 * K).half().to(cuda)
    y = torch.empty(N, dtype=torch.float16, device=cuda)
    gemv_fused(gu, 2 * K, wq, sd, y, N, K, xmode=2, pdl=1)
    act = (torch.nn.functional.silu(gu[:K]) * gu[K:]).cpu()
    ref = oracle.gemm(act[None], q, s)[0]
    assert oracle.norm_rel_err(y.cpu(), ref) <= 1.5e-3


@pytest.mark.parametrize("K,I", [(4096, 11008), (512, 1408), (5120, 1728)])
def test_silu_up_pair_epilogue(cuda, oracle, K, I):
    """Interleaved (gate_i, up_i) rows -> the kernel emits fp16(silu(fp16 gate)) * fp16 up directly (HF MLP arithmetic)."""
    wg, wu = oracle.synth_weight(K, I, 21), oracle.synth_weight(K, I, 22)
    qg, sg, _ = oracle.quantize(wg)
    qu, su, _ = oracle.quantize(wu)
    rows = torch.stack([oracle.b200_layout(qg).view(torch.uint8).view(I, K), oracle.b200_layout(qu).view(torch.uint8).view(I, K)], 1)
    wq = rows.reshape(2 * I, K).contiguous().view(torch.int8).view(K, 2 * I).to(cuda)
    sd = torch.stack([sg, su], 1).reshape(-1).contiguous().to(cuda)
    x = oracle.synth_act(1, K, seed=9)
    y = torch.empty(I, dtype=torch.float16, device=cuda)
    gemv_fused(x.to(cuda), K, wq, sd, y, 2 * I, K, epi=1)
    g = oracle.gemm(x, qg, sg)[0]
    u = oracle.gemm(x, qu, su)[0]
    ref = torch.nn.functional.silu(g.float()).half() * u
    # the gate feeds an exponential: allow one fp16 ulp of the gate to move the product
    assert oracle.norm_rel_err(y.cpu(), ref) <= 2e-3


def torch_attention_reference(q, kc, vc, cos, sin, pos):
    """fp32 reference: rotate q/k at `pos` in fp16 like the kernel, append, softmax(q k^T / sqrt(d)) v in fp32."""
    heads, D = q.shape
    qr = apply_rope(q[None], cos[pos:pos + 1], sin[pos:pos + 1])[0]
    k = kc[:, :pos + 1].float()
    v = vc[:, :pos + 1].float()
    sc = torch.einsum("hd,htd->ht", qr.float(), k) / math.sqrt(D)
    p = torch.softmax(sc, dim=-1)
    return torch.einsum("ht,htd->hd", p, v)


@pytest.mark.parametrize("max_ctx,pos", [(160, 0), (160, 63), (160, 64), (1280, 1024), (1280, 1153), (4224, 4160), (4224, 700), (1100, 1099)])
def test_decode_attention_matches_fp32_reference(cuda, max_ctx, pos):
    """The fused RoPE + append + attention kernel where the bench runs it (ctx ~1024..1153) and far beyond (4160: every
    CTA walks several chunks), against an fp32 PyTorch reference; tolerance 2e-3 norm-relative."""
    heads, D = 8, 128
    H = heads * D
    shape = LlamaShape(hidden=H, heads=heads)
    g = torch.Generator(device=cuda).manual_seed(pos + 1)
    kc = (torch.randn(heads, max_ctx, D, generator=g, device=cuda) * 0.8).half()
    vc = torch.randn(heads, max_ctx, D, generator=g, device=cuda).half()
    qkv = torch.randn(3 * H, generator=g, device=cuda).half()
    cos, sin = rope_tables(shape, max_ctx, cuda, torch.float16)
    pos_t = torch.tensor([pos], dtype=torch.int32, device=cuda)
    out = torch.zeros(H, dtype=torch.float16, device=cuda)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    knew = apply_rope(qkv[H:2 * H].view(1, heads, D), cos[pos:pos + 1], sin[pos:pos + 1])[0]
    kc_ref[:, pos] = knew
    vc_ref[:, pos] = qkv[2 * H:].view(heads, D)
    rc = _cabi.lib().eetq_b200_decode_attention(vp(qkv), vp(cos), vp(sin), vp(pos_t), vp(kc), vp(vc), vp(out), H, D, max_ctx, None, None, 0, 0, 0, stream())
    _cabi.check(rc, "decode_attention")
    torch.cuda.synchronize()
    ref = torch_attention_reference(qkv[:H].view(heads, D), kc_ref, vc_ref, cos, sin, pos).reshape(-1)
    err = (out.float() - ref).abs().max() / ref.abs().max()
    assert err <= 2e-3, float(err)
    # the new row was appended (and nothing else was touched)
    assert torch.equal(kc[:, pos], kc_ref[:, pos]) and torch.equal(vc[:, pos], vc_ref[:, pos])
    assert torch.equal(kc[:, :pos], kc_ref[:, :pos]) and torch.equal(kc[:, pos + 1:], kc_ref[:, pos + 1:])


@pytest.mark.parametrize("V,H", [(32000, 4096), (1024, 512), (4000, 5120)])
def test_lm_head_argmax(cuda, V, H):
    L = _cabi.lib()
    g = torch.Generator(device=cuda).manual_seed(V)
    w = (torch.randn(V, H, generator=g, device=cuda) * 0.02).half()
    x = torch.randn(H, generator=g, device=cuda).half()
    nw = (1 + 0.1 * torch.randn(H, generator=g, device=cuda)).half()
    logits = torch.zeros(V, dtype=torch.float16, device=cuda)
    scratch = torch.zeros(int(L.eetq_b200_lm_head_scratch_bytes()), dtype=torch.uint8, device=cuda)
    token = torch.zeros(1, dtype=torch.int64, device=cuda)
    pos = torch.tensor([5], dtype=torch.int32, device=cuda)
    step = torch.tensor([7], dtype=torch.int32, device=cuda)
    for _ in range(2):  # second call: the scratch ticket was left clean
        rc = L.eetq_b200_lm_head_argmax(vp(x), None, vp(nw), 1e-5, vp(w), V, H, 0, vp(logits), vp(scratch), vp(token), vp(pos), vp(step),
                                        None, 0, 0, stream())
        _cabi.check(rc, "lm_head_argmax")
    torch.cuda.synchronize()
    xf = x.float()
    xn = nw * (xf * torch.rsqrt(xf.pow(2).mean() + 1e-5)).half()
    ref = (xn.float() @ w.float().t())
    assert (logits.float() - ref).abs().max() <= 2e-3 * ref.abs().max() + 1e-3
    assert int(token.item()) == int(torch.argmax(logits).item())          # first maximum of the fp16 logits
    assert int(pos.item()) == 7 and int(step.item()) == 9


def test_reference_named_glue_ops(cuda, oracle):
    """rotary_embedding_neox / layernorm_forward (csrc/eetpy.cpp:18-19) against their CPU restatements."""
    T, heads, D = 5, 4, 128
    g = torch.Generator().manual_seed(3)
    q = torch.randn(T, heads, D, generator=g).half()
    k = torch.randn(T, heads, D, generator=g).half()
    pos = torch.tensor([0, 3, 17, 100, 511])
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.arange(512).float()[:, None] * inv[None]
    cache = torch.cat([fr.cos(), fr.sin()], -1).half()
    qr, kr = oracle.rotary_embedding_neox(pos, q, k, D, cache)
    qd, kd = q.to(cuda), k.to(cuda)
    eetq_b200.rotary_embedding_neox(pos.to(cuda), qd, kd, D, cache.to(cuda))
    torch.cuda.synchronize()
    assert torch.equal(qd.cpu(), qr) and torch.equal(kd.cpu(), kr)

    x = (torch.randn(2, 3, 1024, generator=g) * 2).half()
    gamma = (1 + 0.2 * torch.randn(1024, generator=g)).half()
    out = torch.empty_like(x).to(cuda)
    eetq_b200.layernorm_forward(x.to(cuda), gamma.to(cuda), out, 1e-6)
    torch.cuda.synchronize()
    ref = oracle.layernorm_forward(x, gamma, 1e-6)
    # rsqrtf on the device vs 1/sqrt on the host: at most one fp16 ulp apart
    assert (out.cpu().float() - ref.float()).abs().max() <= 2e-3 * ref.float().abs().max()


def test_decode_matches_torch_forward(cuda):
    torch.manual_seed(0)
    model = LlamaSkeleton(SMALL, device=cuda, seed=3, std=0.05)
    eetq_b200.eet_quantize(model)
    T = 40
    tokens = torch.randint(0, SMALL.vocab, (T,), device=cuda)
    dec = W8A16LlamaDecoder.from_model(model, max_ctx=128)
    first = dec.prefill(tokens[:T - 8])
    ref_logits = model(tokens[:T - 8])
    assert int(first.item()) == int(ref_logits[-1].argmax().item())
    # teacher-forced decode of the remaining tokens: compare logits with the full-sequence forward at each position
    full = model(tokens)
    dec.capture()
    for i in range(T - 8, T):
        dec.token.fill_(int(tokens[i].item()))
        dec.step()
        torch.cuda.synchronize()
        got = dec.logits[0].float()
        exp = full[i].float()
        assert (got - exp).abs().max() <= 3e-2 * exp.abs().max() + 1e-3, i
        assert int(dec.pos.item()) == i + 1
        assert int(dec.token.item()) == int(torch.argmax(dec.logits[0]).item())


def test_pdl_and_plain_launch_agree(cuda):
    model = LlamaSkeleton(SMALL, device=cuda, seed=4, std=0.05)
    eetq_b200.eet_quantize(model)
    prompt = torch.randint(0, SMALL.vocab, (16,), device=cuda)
    outs = [W8A16LlamaDecoder.from_model(model, max_ctx=64, pdl=pdl).generate(prompt, 12) for pdl in (True, False)]
    assert outs[0] == outs[1]


def test_tagged_word_exchange_on_one_gpu_matches_plain_buffers(cuda):
    """The multi-GPU activation exchange ("LL" words: {2 x fp16, tag} pushed by the producing kernel's epilogue, polled by the
    consumer's prologue) also runs with a single rank.  Same kernels, same arithmetic: the tokens must equal the plain-buffer path."""
    shape = LlamaShape(hidden=1024, inter=2816, layers=3, heads=8, vocab=2048, name="ll-1gpu")
    model = LlamaSkeleton(shape, device=cuda, seed=21, std=0.05)
    eetq_b200.eet_quantize(model)
    prompt = torch.randint(0, shape.vocab, (70,), device=cuda)
    plain = W8A16LlamaDecoder.from_model(model, max_ctx=128, exchange="none")
    tagged = W8A16LlamaDecoder.from_model(model, max_ctx=128, exchange="ll")
    assert plain.exchange == "none" and tagged.exchange == "ll"
    assert plain.generate(prompt, 24) == tagged.generate(prompt, 24)


def test_decode_on_7b_shaped_layers_long_context(cuda):
    """Two Llama-2-7B-shaped layers decoded at a context where every attention CTA walks several chunks; tokens must be
    reproducible and equal between a graph replay and a fresh decoder."""
    shape = LlamaShape(hidden=4096, inter=11008, layers=2, heads=32, vocab=512, name="7b-2layer")
    model = LlamaSkeleton(shape, device=cuda, seed=9)
    eetq_b200.eet_quantize(model)
    prompt = torch.randint(0, shape.vocab, (600,), device=cuda)
    a = W8A16LlamaDecoder.from_model(model, max_ctx=700).generate(prompt, 20)
    b = W8A16LlamaDecoder.from_model(model, max_ctx=1400, pdl=False).generate(prompt, 20)
    assert a == b
    ref_first = int(model(prompt)[-1].argmax().item())
    assert a[0] == ref_first


def test_step_host_roundtrip_and_cache_guard(cuda):
    model = LlamaSkeleton(SMALL, device=cuda, seed=5, std=0.05)
    eetq_b200.eet_quantize(model)
    dec = W8A16LlamaDecoder.from_model(model, max_ctx=14)
    prompt = torch.randint(0, SMALL.vocab, (8,), device=cuda)
    t0 = dec.prefill(prompt)
    a = torch.zeros(1, dtype=torch.int64).pin_memory()
    b = torch.zeros(1, dtype=torch.int64).pin_memory()
    a[0] = int(t0.item())
    seq = []
    for _ in range(5):
        dec.step_host(a, b)
        seq.append(int(b[0]))
        a.copy_(b)
    dec2 = W8A16LlamaDecoder.from_model(model, max_ctx=64)
    assert dec2.generate(prompt, 6)[1:] == seq
    # embed + per layer (4 fused GEMVs + 1 fused attention) + final norm/lm_head/arg-max
    assert dec.launches_per_step == 1 + SMALL.layers * 5 + 1
    dec.step()                      # position 13 -> 14 == max_ctx
    with pytest.raises(RuntimeError, match="KV cache is full"):
        dec.step()
