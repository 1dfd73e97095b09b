"""Llama-family decode harness on top of the w8a16 kernels -- what bench.py measures (BASELINE.json: "decode
tokens/sec Llama-2-7B w8a16 @1/2/4/8 B200").

This is the caller either side of the hot path (SURVEY.md section 8 "next"), kept deliberately small:

* ``LlamaSkeleton``      -- an ``nn.Module`` tree with the Hugging Face Llama sub-module names (random-init, there is no
                            network for checkpoints) so that ``eet_quantize`` walks it exactly like it walks a HF model
                            (/root/reference/python/eetq/utils/quantizer.py:40-61).  Its ``forward`` is a plain PyTorch
                            implementation used as the numerical reference in tests.
* ``W8A16LlamaDecoder``  -- takes the quantised model, fuses q|k|v and gate|up row-wise (trivial in the b200 layout:
                            rows are output features; gate and up rows are INTERLEAVED so the GEMV epilogue can emit
                            silu(gate) * up directly), and runs single-token decode as ONE CUDA graph of native kernels:
                            per layer 4 streaming GEMVs (RMSNorm / residual / SiLU*up fused into them) and ONE fused
                            RoPE + KV-append + attention kernel, then ONE final-norm + lm_head + arg-max kernel, all chained
                            with programmatic dependent launch.

Multi-GPU (``world_size > 1``): every linear is column-sharded (rank r owns a contiguous block of
output features = a contiguous byte range of the b200 layout), attention is sharded by head, the lm_head by vocabulary.  The
vectors every rank needs (attention output, residual stream, MLP activation, arg-max candidates) are exchanged through "LL"
buffers in symmetric memory: the producing kernel's epilogue stores {2 x fp16, tag} words straight into every rank's copy over
NVLink and the consuming kernel's prologue polls the words it needs (``exchange="ll"``, default).  ``exchange="nccl"`` keeps plain
buffers and one ``ncclAllGather`` per sharded linear -- the baseline the fused exchange is measured against.

The reference's own end-to-end path is HF ``generate`` over ``W8A16Linear`` modules
(/root/reference/examples/models/llama_transformers_example.py:22-90); its attention side
(/root/reference/python/eetq/modules/llama_modules.py) is outside the w8a16 hot path.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .modules.qlinear import W8A16Linear
from .ops import w8_a16_gemm_bias, w8_a16_gemm_residual

__all__ = ["LlamaShape", "LlamaSkeleton", "W8A16LlamaDecoder", "LLAMA2_7B", "LLAMA2_13B", "shard_plan"]


@dataclass
class LlamaShape:
    hidden: int = 4096
    inter: int = 11008
    layers: int = 32
    heads: int = 32
    vocab: int = 32000
    eps: float = 1e-5
    theta: float = 10000.0
    name: str = "llama-2-7b"

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


LLAMA2_7B = LlamaShape()
LLAMA2_13B = LlamaShape(hidden=5120, inter=13824, layers=40, heads=40, name="llama-2-13b")


# ---------------------------------------------------------------------------------------------------------------------
# skeleton with HF sub-module names
# ---------------------------------------------------------------------------------------------------------------------
class _RMSNorm(nn.Module):
    def __init__(self, n, eps, device, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n, device=device, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        return self.weight * xf.to(x.dtype)


class _Attn(nn.Module):
    def __init__(self, s: LlamaShape, device, dtype):
        super().__init__()
        H = s.hidden
        self.q_proj = nn.Linear(H, H, bias=False, device=device, dtype=dtype)
        self.k_proj = nn.Linear(H, H, bias=False, device=device, dtype=dtype)
        self.v_proj = nn.Linear(H, H, bias=False, device=device, dtype=dtype)
        self.o_proj = nn.Linear(H, H, bias=False, device=device, dtype=dtype)


class _MLP(nn.Module):
    def __init__(self, s: LlamaShape, device, dtype):
        super().__init__()
        self.gate_proj = nn.Linear(s.hidden, s.inter, bias=False, device=device, dtype=dtype)
        self.up_proj = nn.Linear(s.hidden, s.inter, bias=False, device=device, dtype=dtype)
        self.down_proj = nn.Linear(s.inter, s.hidden, bias=False, device=device, dtype=dtype)


class _Layer(nn.Module):
    def __init__(self, s: LlamaShape, device, dtype):
        super().__init__()
        self.self_attn = _Attn(s, device, dtype)
        self.mlp = _MLP(s, device, dtype)
        self.input_layernorm = _RMSNorm(s.hidden, s.eps, device, dtype)
        self.post_attention_layernorm = _RMSNorm(s.hidden, s.eps, device, dtype)


class _Model(nn.Module):
    def __init__(self, s: LlamaShape, device, dtype):
        super().__init__()
        self.embed_tokens = nn.Embedding(s.vocab, s.hidden, device=device, dtype=dtype)
        self.layers = nn.ModuleList([_Layer(s, device, dtype) for _ in range(s.layers)])
        self.norm = _RMSNorm(s.hidden, s.eps, device, dtype)


def rope_tables(s: LlamaShape, max_pos: int, device, dtype):
    """cos/sin [max_pos, D/2] exactly as HF LlamaRotaryEmbedding builds them (fp32 math, cast to the model dtype)."""
    D = s.head_dim
    inv_freq = 1.0 / (s.theta ** (torch.arange(0, D, 2, dtype=torch.float32, device=device) / D))
    freqs = torch.arange(max_pos, dtype=torch.float32, device=device)[:, None] * inv_freq[None, :]
    return freqs.cos().to(dtype).contiguous(), freqs.sin().to(dtype).contiguous()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [T, heads, D]; cos/sin [T, D/2]; HF rotate_half convention, evaluated in the model dtype."""
    half = x.shape[-1] // 2
    c = torch.cat([cos, cos], -1)[:, None, :]
    s = torch.cat([sin, sin], -1)[:, None, :]
    rot = torch.cat([-x[..., half:], x[..., :half]], -1)
    return x * c + rot * s


class LlamaSkeleton(nn.Module):
    """Random-init Llama with HF sub-module names (``model.layers.N.self_attn.q_proj`` ... ``lm_head``)."""

    def __init__(self, shape: LlamaShape, device="cuda", dtype=torch.float16, seed: int = 1000, std: float = 0.02):
        super().__init__()
        self.shape = shape
        self.model = _Model(shape, device, dtype)
        self.lm_head = nn.Linear(shape.hidden, shape.vocab, bias=False, device=device, dtype=dtype)
        g = torch.Generator(device=device).manual_seed(seed)
        with torch.no_grad():
            for p in self.parameters():
                if p.dim() >= 2:  # linears + embedding: N(0, 0.02^2) like the Llama init
                    p.copy_((torch.randn(p.shape, generator=g, device=device, dtype=torch.float32) * std).to(dtype))

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """Plain PyTorch causal forward for a 1-D token tensor; returns logits [T, vocab].  Works before and after
        eet_quantize (the linears are called as modules)."""
        s = self.shape
        T = tokens.shape[0]
        x = self.model.embed_tokens(tokens)
        cos, sin = rope_tables(s, T, x.device, x.dtype)
        for layer in self.model.layers:
            h = layer.input_layernorm(x)
            a = layer.self_attn
            q = a.q_proj(h).view(T, s.heads, s.head_dim)
            k = a.k_proj(h).view(T, s.heads, s.head_dim)
            v = a.v_proj(h).view(T, s.heads, s.head_dim)
            q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
            o = F.scaled_dot_product_attention(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), is_causal=True)
            x = x + a.o_proj(o.transpose(0, 1).reshape(T, s.hidden))
            h = layer.post_attention_layernorm(x)
            m = layer.mlp
            x = x + m.down_proj(F.silu(m.gate_proj(h)) * m.up_proj(h))
        return self.lm_head(self.model.norm(x))


# ---------------------------------------------------------------------------------------------------------------------
# sharding plan (pure arithmetic: also exercised by the CPU tests)
# ---------------------------------------------------------------------------------------------------------------------
def shard_plan(shape: LlamaShape, rank: int, world: int) -> dict:
    """Which output features of every (fused) linear rank `rank` of `world` owns, and the exchange indices of a decode step.

    q|k|v: the rows of this rank's heads (q, then k, then v);  o / down: rows [r H/P, (r+1) H/P);  gate|up: the interleaved
    rows (g_i, u_i) for i in [r I/P, (r+1) I/P);  lm_head: vocabulary rows [r V/P, (r+1) V/P)."""
    H, I, V, P = shape.hidden, shape.inter, shape.vocab, world
    if shape.heads % P or H % (64 * P) or (2 * I) % (64 * P) or V % P or shape.head_dim != 128:
        raise ValueError(f"{shape.name} cannot be sharded {P}-way (heads {shape.heads}, hidden {H}, inter {I}, vocab {V})")
    hl = shape.heads // P
    return dict(
        heads=(rank * hl, (rank + 1) * hl), hidden=(rank * H // P, (rank + 1) * H // P), inter=(rank * I // P, (rank + 1) * I // P),
        vocab=(rank * V // P, (rank + 1) * V // P),
        # exchange indices inside one decode step: 0 = embedding, per layer 1..4 = attention out, x + o(...), act, x + down(...);
        # the last one = arg-max candidates
        per_step=4 * shape.layers + 2, x_in=lambda l: 4 * l, attn=lambda l: 4 * l + 1, x2=lambda l: 4 * l + 2, act=lambda l: 4 * l + 3,
        x_out=lambda l: 4 * l + 4, cand=4 * shape.layers + 1)


# ---------------------------------------------------------------------------------------------------------------------
# decoder
# ---------------------------------------------------------------------------------------------------------------------
def _vp(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _rows(lin: W8A16Linear) -> torch.Tensor:
    """b200 bytes of a quantised linear as a [N, K] uint8 matrix (row n = output feature n)."""
    K, N = lin.qweight.shape
    return lin.qweight.view(torch.uint8).view(N, K)


class _Shard:
    """A contiguous block of rows of a (possibly fused) quantised linear: int8 [K, N_local] (nominal) + scales."""

    def __init__(self, rows: torch.Tensor, scales: torch.Tensor):
        self.N, self.K = rows.shape
        self.w = rows.contiguous().view(torch.int8).view(self.K, self.N)  # nominal [K, N_local] like the reference
        self.scales = scales.contiguous()


class W8A16LlamaDecoder:
    def __init__(self, model: nn.Module, shape: LlamaShape, max_ctx: int = 1280, pdl: bool = True, rank: int = 0, world_size: int = 1,
                 group=None, exchange: Optional[str] = None):
        self.shape, self.max_ctx, self.pdl = shape, max_ctx, bool(pdl)
        self.rank, self.world, self.group = rank, world_size, group
        if world_size > 1:
            self.exchange = exchange or os.environ.get("EETQ_B200_EXCHANGE", "ll")
        else:
            # one GPU: plain vectors + griddepcontrol.wait ("none").  The tagged-word buffers of the multi-GPU path also work here
            # ("ll", no kernel waits for its predecessor to drain) but measured slower: 509 vs 633 tok/s -- 296 CTAs polling for the
            # activation words slow the producer's weight stream down more than the skipped drain saves.
            self.exchange = exchange or os.environ.get("EETQ_B200_LOCAL_EXCHANGE", "none")
            if self.exchange == "nccl":
                self.exchange = "none"
        assert self.exchange in ("none", "ll", "nccl")
        plan = self.plan = shard_plan(shape, rank, world_size)
        m = model.model
        dev = self.device = m.embed_tokens.weight.device
        dt = torch.float16
        H, I, L, D = shape.hidden, shape.inter, shape.layers, shape.head_dim
        h0, h1 = plan["heads"]
        n0, n1 = plan["hidden"]
        i0, i1 = plan["inter"]
        v0, v1 = plan["vocab"]
        self.Hl, self.Il, self.Vl = (h1 - h0) * D, i1 - i0, v1 - v0
        self.embed = m.embed_tokens.weight.detach()
        self.lm_head_w = model.lm_head.weight.detach()[v0:v1].contiguous()  # fp16 [V/P, H], not quantised (quantizer.py:40)
        self.norm_w = m.norm.weight.detach()
        self.layers = []
        for layer in m.layers:
            a, p = layer.self_attn, layer.mlp
            for lin in (a.q_proj, a.k_proj, a.v_proj, a.o_proj, p.gate_proj, p.up_proj, p.down_proj):
                assert isinstance(lin, W8A16Linear), "run eet_quantize(model) first"
            hs = slice(h0 * D, h1 * D)
            qkv_rows = torch.cat([_rows(a.q_proj)[hs], _rows(a.k_proj)[hs], _rows(a.v_proj)[hs]], 0)
            qkv_scales = torch.cat([a.q_proj.weight_scales[hs], a.k_proj.weight_scales[hs], a.v_proj.weight_scales[hs]], 0)
            gu_rows = torch.stack([_rows(p.gate_proj)[i0:i1], _rows(p.up_proj)[i0:i1]], 1).reshape(2 * (i1 - i0), H)
            gu_scales = torch.stack([p.gate_proj.weight_scales[i0:i1], p.up_proj.weight_scales[i0:i1]], 1).reshape(-1)
            self.layers.append(dict(
                qkv=_Shard(qkv_rows, qkv_scales), o=_Shard(_rows(a.o_proj)[n0:n1], a.o_proj.weight_scales[n0:n1]),
                gu=_Shard(gu_rows, gu_scales), down=_Shard(_rows(p.down_proj)[n0:n1], p.down_proj.weight_scales[n0:n1]),
                ln1=layer.input_layernorm.weight.detach(), ln2=layer.post_attention_layernorm.weight.detach()))
        self.cos, self.sin = rope_tables(shape, max_ctx, dev, dt)
        # KV cache of this rank's heads, head-major: [layer][head][max_ctx][head_dim]
        self.kcache = torch.zeros(L, h1 - h0, max_ctx, D, dtype=dt, device=dev)
        self.vcache = torch.zeros(L, h1 - h0, max_ctx, D, dtype=dt, device=dev)
        # decode-step state (device resident; the graph reads/writes these)
        self.token = torch.zeros(1, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_ctr = torch.zeros(1, dtype=torch.int32, device=dev)  # LL tag base, advanced by the lm_head kernel
        self.qkv = torch.zeros(3 * self.Hl, dtype=dt, device=dev)
        self.logits = torch.zeros(1, shape.vocab, dtype=dt, device=dev)
        self._L = _cabi.lib()
        self.lm_scratch = torch.zeros(int(self._L.eetq_b200_lm_head_scratch_bytes()), dtype=torch.uint8, device=dev)
        self._ll = None
        if self.exchange == "ll":
            try:
                self._setup_ll(H, I, dev)
            except Exception as e:  # symmetric memory unavailable -> NCCL all-gather (still a GPU path, never a CPU one)
                if rank == 0:
                    print(f"[eetq_b200] LL exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather", flush=True)
                self.exchange = "nccl"
        if self._ll is None:
            self.x = torch.zeros(H, dtype=dt, device=dev)
            self.x2 = torch.zeros(H, dtype=dt, device=dev)
            self.attn = torch.zeros(H, dtype=dt, device=dev)
            self.act = torch.zeros(I, dtype=dt, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------- LL exchange buffers
    def _setup_ll(self, H, I, dev):
        """LL buffers (8-byte words {2 x fp16, tag}) in ONE symmetric-memory arena mapped into every rank."""
        if self.world > 1:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

        def rnd(n):
            return (n + 255) // 256 * 256

        sizes = [("x", H * 4), ("x2", H * 4), ("attn", H * 4), ("act", I * 4), ("cand", 2 * 8 * 8)]
        offs, total = {}, 0
        for name, nb in sizes:
            offs[name] = total
            total += rnd(nb)
        if self.world == 1:
            arena = torch.zeros(total, dtype=torch.uint8, device=dev)  # tag 0 is never a valid tag
            self._ll = dict(arena=arena, hdl=None, bases=[arena.data_ptr()], offs=offs)
            return
        arena = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        arena.zero_()  # tag 0 is never a valid tag
        group = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(arena, group)
        bases = [int(b) for b in hdl.buffer_ptrs]
        assert len(bases) == self.world
        self._ll = dict(arena=arena, hdl=hdl, bases=bases, offs=offs)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # every rank's buffers are zero before anybody pushes

    def _ll_local(self, name: str) -> int:
        return self._ll["arena"].data_ptr() + self._ll["offs"][name]

    def _tag(self, index: int):
        return _cabi.LL(self.step_ctr.data_ptr(), self.plan["per_step"], index)

    def _push(self, name: str, elem_off: int, index: int):
        ll = self._ll
        peers = (ctypes.c_uint64 * 8)(*[b + ll["offs"][name] for b in ll["bases"]] + [0] * (8 - self.world))
        p = _cabi.LLPush(self.world, ctypes.cast(peers, ctypes.POINTER(ctypes.c_uint64)), self._ll_local(name), elem_off,
                         self.step_ctr.data_ptr(), self.plan["per_step"], index)
        p._keep = peers  # keep the host array alive as long as the struct
        return p

    # ------------------------------------------------------------------------------------------------- construction
    @classmethod
    def from_model(cls, model: nn.Module, **kw) -> "W8A16LlamaDecoder":
        return cls(model, model.shape, **kw)

    # ------------------------------------------------------------------------------------------------- helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _all_gather(self, full: torch.Tensor, part: torch.Tensor):
        import torch.distributed as dist

        dist.all_gather_into_tensor(full, part, group=self.group)

    def _gemv(self, x, ldx, lin: _Shard, y, *, norm_w=None, xmode=0, epi=0, residual=None, x_tag=None, res_tag=None, res_off=0, push=None,
              nxt: Optional["_Shard"] = None):
        """One fused decode GEMV over this rank's rows.  x / residual: tensors (plain) or raw LL addresses (int) with their tags."""
        o = _cabi.GemvOpts()
        o.norm_weight = 0 if norm_w is None else norm_w.data_ptr()
        o.eps, o.xmode, o.epi = float(self.shape.eps), xmode, epi
        o.residual = 0 if residual is None else (residual if isinstance(residual, int) else residual.data_ptr())
        o.ldr = lin.N
        o.x_ll = ctypes.pointer(x_tag) if x_tag is not None else None
        o.residual_ll = ctypes.pointer(res_tag) if res_tag is not None else None
        o.residual_off = res_off
        o.push = ctypes.pointer(push) if push is not None else None
        if nxt is not None:  # L2 staging hint: the GEMV that runs next (eetq_b200_gemv_opts.next_w)
            o.next_w, o.next_n, o.next_k = nxt.w.data_ptr(), nxt.N, nxt.K
        xp = ctypes.c_void_p(x if isinstance(x, int) else x.data_ptr())
        rc = self._L.eetq_b200_w8a16_gemv_fused(xp, ldx, _vp(lin.w), _vp(lin.scales), None, _vp(y), lin.N, 1, lin.N, lin.K, _cabi.F16,
                                                ctypes.byref(o), 1 if self.pdl else 0, self._stream())
        _cabi.check(rc, "eetq_b200_w8a16_gemv_fused")

    # ------------------------------------------------------------------------------------------------- one decode step
    def _enqueue_step(self):
        """Enqueue one token's worth of kernels on the current stream (captured once into a CUDA graph)."""
        s, L, pdl, plan = self.shape, self._L, 1 if self.pdl else 0, self.plan
        H, I, D = s.hidden, s.inter, s.head_dim
        st = self._stream
        n0 = plan["hidden"][0]
        i0 = plan["inter"][0]
        ll = self._ll is not None
        nccl = self.exchange == "nccl"

        if ll:
            xt = self._tag(plan["x_in"](0))
            _cabi.check(L.eetq_b200_decode_embed(_vp(self.embed), _vp(self.token), ctypes.c_void_p(self._ll_local("x")), H, ctypes.byref(xt), pdl,
                                                 st()), "decode_embed")
        else:
            _cabi.check(L.eetq_b200_decode_embed(_vp(self.embed), _vp(self.token), _vp(self.x), H, None, pdl, st()), "decode_embed")

        for li, w in enumerate(self.layers):
            if ll:
                x_in, x_tag = self._ll_local("x"), self._tag(plan["x_in"](li))
                # q|k|v of this rank's heads (plain, local)
                self._gemv(x_in, H, w["qkv"], self.qkv, norm_w=w["ln1"], xmode=1, x_tag=x_tag)
                push = self._push("attn", plan["heads"][0] * D, plan["attn"](li))
                _cabi.check(L.eetq_b200_decode_attention(_vp(self.qkv), _vp(self.cos), _vp(self.sin), _vp(self.pos), _vp(self.kcache[li]),
                                                         _vp(self.vcache[li]), None, self.Hl, D, self.max_ctx, ctypes.byref(push),
                                                         _vp(w["o"].w), w["o"].N, w["o"].K, pdl, st()),
                            "decode_attention")
                # x2 = x + o_proj(attn)
                nxt_qkv = self.layers[li + 1]["qkv"] if li + 1 < len(self.layers) else None
                self._gemv(self._ll_local("attn"), H, w["o"], None, x_tag=self._tag(plan["attn"](li)), residual=x_in, res_tag=x_tag, res_off=n0,
                           push=self._push("x2", n0, plan["x2"](li)), nxt=w["gu"])
                # act = silu(gate(norm(x2))) * up(norm(x2))
                x2_tag = self._tag(plan["x2"](li))
                self._gemv(self._ll_local("x2"), H, w["gu"], None, norm_w=w["ln2"], xmode=1, epi=1, x_tag=x2_tag,
                           push=self._push("act", i0, plan["act"](li)), nxt=w["down"])
                # x = x2 + down(act)
                self._gemv(self._ll_local("act"), I, w["down"], None, x_tag=self._tag(plan["act"](li)), residual=self._ll_local("x2"),
                           res_tag=x2_tag, res_off=n0, push=self._push("x", n0, plan["x_out"](li)), nxt=nxt_qkv)
            else:
                nxt_qkv = self.layers[li + 1]["qkv"] if li + 1 < len(self.layers) else None
                self._gemv(self.x, H, w["qkv"], self.qkv, norm_w=w["ln1"], xmode=1)
                attn_out = self.attn[n0:n0 + self.Hl] if nccl else self.attn
                _cabi.check(L.eetq_b200_decode_attention(_vp(self.qkv), _vp(self.cos), _vp(self.sin), _vp(self.pos), _vp(self.kcache[li]),
                                                         _vp(self.vcache[li]), _vp(attn_out), self.Hl, D, self.max_ctx, None,
                                                         _vp(w["o"].w), w["o"].N, w["o"].K, pdl, st()),
                            "decode_attention")
                if nccl:
                    self._all_gather(self.attn, attn_out)
                sl = slice(n0, n0 + w["o"].N)
                self._gemv(self.attn, H, w["o"], self.x2[sl], residual=self.x[sl], nxt=w["gu"])
                if nccl:
                    self._all_gather(self.x2, self.x2[sl])
                al = slice(i0, i0 + self.Il)
                self._gemv(self.x2, H, w["gu"], self.act[al], norm_w=w["ln2"], xmode=1, epi=1, nxt=w["down"])
                if nccl:
                    self._all_gather(self.act, self.act[al])
                self._gemv(self.act, I, w["down"], self.x[sl], residual=self.x2[sl], nxt=nxt_qkv)
                if nccl:
                    self._all_gather(self.x, self.x[sl])

        v0 = plan["vocab"][0]
        if ll:
            xt = self._tag(plan["x_out"](len(self.layers) - 1))
            cand = self._push("cand", 0, plan["cand"]) if self.world > 1 else None
            rc = L.eetq_b200_lm_head_argmax(ctypes.c_void_p(self._ll_local("x")), ctypes.byref(xt), _vp(self.norm_w), float(s.eps),
                                            _vp(self.lm_head_w), self.Vl, H, v0, _vp(self.logits), _vp(self.lm_scratch), _vp(self.token),
                                            _vp(self.pos), _vp(self.step_ctr), ctypes.byref(cand) if cand is not None else None,
                                            self.rank, pdl, st())
            _cabi.check(rc, "lm_head_argmax")
        else:
            rc = L.eetq_b200_lm_head_argmax(_vp(self.x), None, _vp(self.norm_w), float(s.eps), _vp(self.lm_head_w), self.Vl, H, v0,
                                            _vp(self.logits), _vp(self.lm_scratch), _vp(self.token), _vp(self.pos), _vp(self.step_ctr), None,
                                            self.rank, pdl, st())
            _cabi.check(rc, "lm_head_argmax")
            if nccl:
                # baseline path: gather the logits slices and pick the winner with framework ops
                self._all_gather(self.logits.view(-1), self.logits.view(-1)[v0:v0 + self.Vl])
                self.token.copy_(torch.argmax(self.logits, dim=-1))

    def capture(self):
        """Warm up (lazy module loads) and capture the decode step into a CUDA graph."""
        saved = (self.pos.clone(), self.token.clone())
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._check_room()
                self._enqueue_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.pos.copy_(saved[0])
        self.token.copy_(saved[1])
        before = _cabi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue_step()
        self.launches_per_step = _cabi.launch_count() - before
        torch.cuda.synchronize(self.device)
        self.pos.copy_(saved[0])
        self.token.copy_(saved[1])
        self._host_pos = int(self.pos.item())

    def _check_room(self):
        """The attention kernel appends at row *pos: refuse to step past the end of the KV cache (host-tracked position)."""
        hp = getattr(self, "_host_pos", None)
        if hp is None:
            hp = self._host_pos = int(self.pos.item())
        if hp + 1 > self.max_ctx:
            raise RuntimeError(f"KV cache is full: position {hp} + 1 > max_ctx {self.max_ctx}")

    def step(self):
        """Advance one token entirely on the device (token / position / step counter are updated by the graph)."""
        if self.graph is None:
            self.capture()
        self._check_room()
        self.graph.replay()
        self._host_pos += 1

    def step_host(self, token_host: torch.Tensor, out_host: torch.Tensor):
        """Public end-to-end call: token id in PINNED host memory -> next token id in pinned host memory
        (H2D copy, one decode step, D2H copy, synchronise)."""
        self.token.copy_(token_host, non_blocking=True)
        self.step()
        out_host.copy_(self.token, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host

    # ------------------------------------------------------------------------------------------------- prefill
    @torch.no_grad()
    def prefill(self, tokens: torch.Tensor) -> torch.Tensor:
        """Process a prompt [T] with the batched (tcgen05) kernels, fill the KV cache, return the first new token.
        Native kernels: w8a16 GEMMs (residual adds fused), RMSNorm, RoPE + KV-cache write, SiLU*up, final norm + lm_head +
        arg-max; causal attention over the prompt uses the framework's SDPA (attention is outside the w8a16 hot path)."""
        s, L, plan = self.shape, self._L, self.plan
        T, H, D = tokens.shape[0], s.hidden, s.head_dim
        assert 0 < T < self.max_ctx
        dev, dt = self.device, torch.float16
        st = self._stream()
        hl = plan["heads"][1] - plan["heads"][0]
        n0, n1 = plan["hidden"]
        x = self.embed[tokens].contiguous()

        def rms(v, w):
            out = torch.empty_like(v)
            _cabi.check(L.eetq_b200_rmsnorm(_vp(v), v.stride(0), _vp(w), _vp(out), out.stride(0), v.shape[0], H, float(s.eps), 0, st), "rmsnorm")
            return out

        def gather_cols(part):
            """[T, n_local] on every rank -> [T, n_local * P] (rank-major feature order)."""
            if self.world == 1:
                return part
            import torch.distributed as dist

            buf = torch.empty(self.world, *part.shape, dtype=part.dtype, device=dev)
            dist.all_gather_into_tensor(buf, part.contiguous(), group=self.group)
            return buf.permute(1, 0, 2).reshape(part.shape[0], -1).contiguous()

        for li, w in enumerate(self.layers):
            qkv = w8_a16_gemm_bias(rms(x, w["ln1"]), w["qkv"].w, w["qkv"].scales, None)          # [T, 3 Hl]
            _cabi.check(L.eetq_b200_prefill_rope_kv(_vp(qkv), qkv.stride(0), _vp(self.cos), _vp(self.sin), _vp(self.kcache[li]),
                                                    _vp(self.vcache[li]), T, hl, D, self.max_ctx, 0, st), "prefill_rope_kv")
            q = qkv[:, :self.Hl].view(T, hl, D).transpose(0, 1)
            k = self.kcache[li, :, :T]
            v = self.vcache[li, :, :T]
            o = F.scaled_dot_product_attention(q, k, v, is_causal=True)                             # [hl, T, D]
            attn = gather_cols(o.transpose(0, 1).reshape(T, self.Hl))
            x2 = gather_cols(w8_a16_gemm_residual(attn, w["o"].w, w["o"].scales, x[:, n0:n1]))
            gu = w8_a16_gemm_bias(rms(x2, w["ln2"]), w["gu"].w, w["gu"].scales, None)             # [T, 2 Il] interleaved (g, u)
            act = torch.empty(T, self.Il, dtype=dt, device=dev)
            _cabi.check(L.eetq_b200_silu_mul(_vp(gu), gu.stride(0), _vp(act), act.stride(0), T, self.Il, 1, st), "silu_mul")
            act = gather_cols(act)
            x = gather_cols(w8_a16_gemm_residual(act, w["down"].w, w["down"].scales, x2[:, n0:n1]))
        # first new token: final norm + lm_head + arg-max on the last row (the kernel advances pos and the step counter)
        last = x[-1].contiguous()
        self.pos.fill_(T - 1)
        cand = self._push("cand", 0, plan["cand"]) if (self._ll is not None and self.world > 1) else None
        rc = L.eetq_b200_lm_head_argmax(_vp(last), None, _vp(self.norm_w), float(s.eps), _vp(self.lm_head_w), self.Vl, H, plan["vocab"][0],
                                        _vp(self.logits), _vp(self.lm_scratch), _vp(self.token), _vp(self.pos), _vp(self.step_ctr),
                                        ctypes.byref(cand) if cand is not None else None, self.rank, 0, st)
        _cabi.check(rc, "lm_head_argmax")
        if self.exchange == "nccl":
            v0 = plan["vocab"][0]
            self._all_gather(self.logits.view(-1), self.logits.view(-1)[v0:v0 + self.Vl])
            self.token.copy_(torch.argmax(self.logits, dim=-1))
        self._host_pos = T
        return self.token.clone()

    def set_context(self, ctx_len: int, token: int = 1):
        """Pretend `ctx_len` tokens are already cached (cache content stays as is) -- for benchmarks/tests."""
        assert 0 <= ctx_len < self.max_ctx
        self.pos.fill_(ctx_len)
        self.token.fill_(token)
        self._host_pos = ctx_len

    @torch.no_grad()
    def generate(self, prompt: torch.Tensor, max_new_tokens: int) -> List[int]:
        out = [int(self.prefill(prompt).item())]
        for _ in range(max_new_tokens - 1):
            self.step()
            out.append(int(self.token.item()))
        return out

    # ------------------------------------------------------------------------------------------------- accounting
    def weight_bytes_per_token(self) -> int:
        """int8 weight bytes THIS rank streams per decoded token (quantised linears only)."""
        return sum(l.N * l.K for w in self.layers for l in (w["qkv"], w["o"], w["gu"], w["down"]))

    def lm_head_bytes_per_token(self) -> int:
        return self.lm_head_w.numel() * 2
