"""Llama-family decode harness on top of the w8a16 kernels -- what bench.py measures (BASELINE.json: "decode
tokens/sec Llama-2-7B w8a16 @1/2/4/8 B200").

This is the caller either side of the hot path (SURVEY.md section 8 "next"), kept deliberately small:

* ``LlamaSkeleton``      -- an ``nn.Module`` tree with the Hugging Face Llama sub-module names (random-init, there is no
                            network for checkpoints) so that ``eet_quantize`` walks it exactly like it walks a HF model
                            (/root/reference/python/eetq/utils/quantizer.py:40-61).  Its ``forward`` is a plain PyTorch
                            implementation used as the numerical reference in tests.
* ``W8A16LlamaDecoder``  -- takes the quantised model, fuses q|k|v and gate|up row-wise (trivial in the b200 layout:
                            rows are output features), and runs single-token decode as ONE CUDA graph of native kernels:
                            per layer 4 streaming GEMVs (RMSNorm / SiLU*up / residual fused into them) and ONE fused
                            RoPE + KV-append + split-KV attention kernel, chained with programmatic dependent launch so that the weight
                            stream of kernel i+1 starts while kernel i drains.
                            With ``world_size > 1`` every linear is column-sharded (rank r owns rows
                            [r*N/P, (r+1)*N/P) of each fused weight -- a contiguous byte range) and the activations are
                            all-gathered after each linear.

The reference's own end-to-end path is HF ``generate`` over ``W8A16Linear`` modules
(/root/reference/examples/models/llama_transformers_example.py:22-90); its attention side
(/root/reference/python/eetq/modules/llama_modules.py) is outside the w8a16 hot path.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .modules.qlinear import W8A16Linear
from .ops import w8_a16_gemm_bias

__all__ = ["LlamaShape", "LlamaSkeleton", "W8A16LlamaDecoder", "LLAMA2_7B", "LLAMA2_13B"]


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
# decoder
# ---------------------------------------------------------------------------------------------------------------------
def _vp(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _rows(lin: W8A16Linear) -> torch.Tensor:
    """b200 bytes of a quantised linear as a [N, K] uint8 matrix (row n = output feature n)."""
    K, N = lin.qweight.shape
    return lin.qweight.view(torch.uint8).view(N, K)


class _ShardedLinear:
    """Rows [r*N/P, (r+1)*N/P) of a (possibly fused) quantised linear on this rank."""

    def __init__(self, lins: List[W8A16Linear], rank: int, world: int):
        rows = torch.cat([_rows(l) for l in lins], 0)
        scales = torch.cat([l.weight_scales for l in lins], 0)
        N, K = rows.shape
        assert N % (64 * world) == 0, f"N={N} cannot be column-sharded {world}-way in multiples of 64"
        self.N, self.K = N, K
        self.n_local = N // world
        self.n_begin = rank * self.n_local
        sl = slice(self.n_begin, self.n_begin + self.n_local)
        self.w = rows[sl].contiguous().view(torch.int8).view(K, self.n_local)  # nominal [K, N_local] like the reference
        self.scales = scales[sl].contiguous()


class W8A16LlamaDecoder:
    def __init__(self, model: nn.Module, shape: LlamaShape, max_ctx: int = 1280, pdl: bool = True, rank: int = 0, world_size: int = 1,
                 group=None, allgather: Optional[str] = None, chain: Optional[bool] = None):
        self.shape, self.max_ctx, self.pdl = shape, max_ctx, bool(pdl)
        self.rank, self.world, self.group = rank, world_size, group
        # "p2p": all-gather fused into the GEMV epilogue over NVLink peer memory; "nccl": one ncclAllGather per linear
        # (default: NCCL measured 388 tok/s vs 344 for the first p2p protocol at N=2)
        self.allgather = (allgather or os.environ.get("EETQ_B200_ALLGATHER", "nccl")) if world_size > 1 else "none"
        m = model.model
        dev = m.embed_tokens.weight.device
        self.device = dev
        dt = torch.float16
        self.embed = m.embed_tokens.weight.detach()
        self.lm_head_w = model.lm_head.weight.detach()  # fp16 [V, H], not quantised (quantizer.py:40 excludes lm_head)
        self.norm_w = m.norm.weight.detach()
        self.layers = []
        for layer in m.layers:
            a, p = layer.self_attn, layer.mlp
            for lin in (a.q_proj, a.k_proj, a.v_proj, a.o_proj, p.gate_proj, p.up_proj, p.down_proj):
                assert isinstance(lin, W8A16Linear), "run eet_quantize(model) first"
            self.layers.append(dict(
                qkv=_ShardedLinear([a.q_proj, a.k_proj, a.v_proj], rank, world_size),
                o=_ShardedLinear([a.o_proj], rank, world_size),
                gu=_ShardedLinear([p.gate_proj, p.up_proj], rank, world_size),
                down=_ShardedLinear([p.down_proj], rank, world_size),
                ln1=layer.input_layernorm.weight.detach(), ln2=layer.post_attention_layernorm.weight.detach()))
        H, I, L = shape.hidden, shape.inter, shape.layers
        self.cos, self.sin = rope_tables(shape, max_ctx, dev, dt)
        # KV cache, head-major: [layer][head][max_ctx][head_dim] (each attention CTA streams one contiguous block)
        self.kcache = torch.zeros(L, shape.heads, max_ctx, shape.head_dim, dtype=dt, device=dev)
        self.vcache = torch.zeros(L, shape.heads, max_ctx, shape.head_dim, dtype=dt, device=dev)
        # decode-step buffers (device resident; the graph reads/writes these)
        self.token = torch.zeros(1, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)   # strictly increasing step counter (p2p flags)
        self._p2p = None
        if self.allgather == "p2p":
            try:
                self._setup_p2p(H, I, L, dt, dev)
            except Exception as e:  # symmetric memory unavailable -> NCCL all-gather (still a GPU path, never a CPU one)
                if rank == 0:
                    print(f"[eetq_b200] p2p all-gather unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)
                self.allgather, self._p2p = "nccl", None
        if self._p2p is None:
            self.x = torch.zeros(H, dtype=dt, device=dev)
            self.x2 = torch.zeros(H, dtype=dt, device=dev)
            self.qkv = torch.zeros(3 * H, dtype=dt, device=dev)
            self.gu = torch.zeros(2 * I, dtype=dt, device=dev)
        self.attn = torch.zeros(H, dtype=dt, device=dev)
        # single GPU: o_proj -> gate|up -> down -> next q|k|v run as ONE chained launch per layer (grid barriers inside)
        if chain is None:
            chain = os.environ.get("EETQ_B200_CHAIN", "0") == "1"   # measured slower than PDL-chained launches
        self.chain = bool(chain) and world_size == 1
        self.chain_counters = torch.zeros(L, 4, dtype=torch.int32, device=dev)
        # the q|k|v GEMV of a layer prefetches that layer's KV cache rows into L2 for the attention kernel that follows
        # (measured slower, 551 vs 563 tok/s: off by default)
        self.kv_prefetch = os.environ.get("EETQ_B200_KV_PREFETCH", "0") == "1"
        self.xn = torch.zeros(1, H, dtype=dt, device=dev)
        self.logits = torch.zeros(1, shape.vocab, dtype=dt, device=dev)
        self._L = _cabi.lib()
        splits = int(self._L.eetq_b200_decode_attention_splits(max_ctx))
        self.partial = torch.zeros(shape.heads * splits * (shape.head_dim + 2), dtype=torch.float32, device=dev)
        self.tickets = torch.zeros(shape.heads, dtype=torch.int32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------- p2p all-gather
    def _setup_p2p(self, H, I, L, dt, dev):
        """Activation buffers + flags in ONE symmetric-memory arena mapped into every rank (NVLink peer pointers)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        def rnd(n):
            return (n + 255) // 256 * 256

        nslot = 4 * L
        sizes = [("x", H * 2), ("x2", H * 2), ("qkv", 3 * H * 2), ("gu", 2 * I * 2), ("flags", nslot * 8 * 4)]
        offs, total = {}, 0
        for name, nb in sizes:
            offs[name] = total
            total += rnd(nb)
        arena = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        arena.zero_()
        group = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(arena, group)
        bases = [int(b) for b in hdl.buffer_ptrs]
        assert len(bases) == self.world
        view = lambda name, nb, dtype: arena[offs[name]:offs[name] + nb].view(dtype)
        self.x, self.x2 = view("x", H * 2, dt), view("x2", H * 2, dt)
        self.qkv, self.gu = view("qkv", 3 * H * 2, dt), view("gu", 2 * I * 2, dt)
        flags = view("flags", nslot * 8 * 4, torch.int32)
        self._p2p = dict(arena=arena, hdl=hdl, bases=bases, offs=offs, flags=flags, nslot=nslot,
                         ticket=torch.zeros(1, dtype=torch.int32, device=dev), slot=0)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)   # every rank's flags are zero before anybody signals

    def _p2p_args(self, y_full: torch.Tensor, lin, slot: int):
        """ctypes arrays of peer pointers for output buffer `y_full` (a view into the arena) and flag slot `slot`."""
        p = self._p2p
        arena_base = p["arena"].data_ptr()
        y_off = y_full.data_ptr() - arena_base + lin.n_begin * 2            # this rank's first row inside the buffer
        f_off = p["offs"]["flags"] + (slot * 8) * 4
        peer_y = (ctypes.c_uint64 * 8)(*[b + y_off for b in p["bases"]] + [0] * (8 - self.world))
        peer_f = (ctypes.c_uint64 * 8)(*[b + f_off + self.rank * 4 for b in p["bases"]] + [0] * (8 - self.world))
        local_flags = ctypes.c_void_p(arena_base + f_off)
        return peer_y, peer_f, local_flags

    # ------------------------------------------------------------------------------------------------- construction
    @classmethod
    def from_model(cls, model: nn.Module, **kw) -> "W8A16LlamaDecoder":
        return cls(model, model.shape, **kw)

    # ------------------------------------------------------------------------------------------------- helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _flags_ptr(self, slot):
        """Address of this rank's flags[slot][0..world) inside the symmetric arena (None -> NULL)."""
        if slot is None or self._p2p is None:
            return ctypes.c_void_p(0)
        p = self._p2p
        return ctypes.c_void_p(p["arena"].data_ptr() + p["offs"]["flags"] + (slot * 8) * 4)

    def _gemv(self, x, ldx, lin: _ShardedLinear, y_full, *, norm_w=None, xmode=0, residual_full=None, wait_slot=None, kv_layer=None):
        """y_full[n_begin : n_begin + n_local] = fused GEMV over this rank's rows; then all-gather if sharded.
        p2p mode: returns the flag slot this call publishes; `wait_slot` is the slot of the call that produced `x`."""
        off = lin.n_begin
        y = y_full[off:off + lin.n_local]
        res = None if residual_full is None else residual_full[off:off + lin.n_local]
        if self._p2p is not None:
            p = self._p2p
            slot = p["slot"] % p["nslot"]
            p["slot"] += 1
            peer_y, peer_f, local_flags = self._p2p_args(y_full, lin, slot)
            rc = self._L.eetq_b200_w8a16_gemv_fused_p2p(_vp(x), ldx, _vp(lin.w), _vp(lin.scales), _vp(norm_w), float(self.shape.eps), xmode,
                                                        _vp(res), lin.N, 1, lin.n_local, lin.K, _cabi.F16, self.world, peer_y, peer_f,
                                                        local_flags, self._flags_ptr(wait_slot), _vp(p["ticket"]), _vp(self.epoch), lin.N,
                                                        1 if self.pdl else 0, self._stream())
            _cabi.check(rc, "eetq_b200_w8a16_gemv_fused_p2p")
            return slot
        if kv_layer is not None and self.kv_prefetch and self.world == 1:
            rc = self._L.eetq_b200_w8a16_gemv_fused_kvprefetch(_vp(x), ldx, _vp(lin.w), _vp(lin.scales), _vp(norm_w), float(self.shape.eps),
                                                               xmode, _vp(res), lin.N, _vp(y), lin.N, 1, lin.n_local, lin.K, _cabi.F16,
                                                               _vp(self.kcache[kv_layer]), _vp(self.vcache[kv_layer]), _vp(self.pos),
                                                               self.shape.heads, self.max_ctx, 1 if self.pdl else 0, self._stream())
            _cabi.check(rc, "eetq_b200_w8a16_gemv_fused_kvprefetch")
            return None
        rc = self._L.eetq_b200_w8a16_gemv_fused(_vp(x), ldx, _vp(lin.w), _vp(lin.scales), None, _vp(norm_w), float(self.shape.eps),
                                                xmode, _vp(res), lin.N, _vp(y), lin.N, 1, lin.n_local, lin.K, _cabi.F16,
                                                1 if self.pdl else 0, self._stream())
        _cabi.check(rc, "eetq_b200_w8a16_gemv_fused")
        if self.world > 1:
            import torch.distributed as dist

            dist.all_gather_into_tensor(y_full[:lin.N], y, group=self.group)

    # ------------------------------------------------------------------------------------------------- one decode step
    def _enqueue_step(self):
        """Enqueue one token's worth of kernels on the current stream (captured once into a CUDA graph)."""
        s, L, pdl = self.shape, self._L, 1 if self.pdl else 0
        H, I, D = s.hidden, s.inter, s.head_dim
        st = self._stream
        if self._p2p is not None:
            self._p2p["slot"] = 0
            self.epoch.add_(1)   # one epoch per decode step; flags[slot] only ever increase
        _cabi.check(L.eetq_b200_decode_embed(_vp(self.embed), _vp(self.token), _vp(self.x), H, pdl, st()), "decode_embed")
        def attention(li, wait_slot=None):
            _cabi.check(L.eetq_b200_decode_attention_p2p(_vp(self.qkv), _vp(self.cos), _vp(self.sin), _vp(self.pos), _vp(self.kcache[li]),
                                                         _vp(self.vcache[li]), _vp(self.partial), _vp(self.tickets), _vp(self.attn), H, D,
                                                         self.max_ctx, self._flags_ptr(wait_slot), self.world, _vp(self.epoch), pdl, st()),
                        "decode_attention")

        if self.chain:
            nl = len(self.layers)
            self.epoch.add_(1)
            self._gemv(self.x, H, self.layers[0]["qkv"], self.qkv, norm_w=self.layers[0]["ln1"], xmode=1)
            for li, w in enumerate(self.layers):
                attention(li)
                # x2 = x + o_proj(attn); gu = gate|up(norm(x2)); x = x2 + down(silu(gate)*up); qkv = q|k|v(norm(x)) of the next layer
                phases = [
                    (self.attn, H, w["o"], self.x2, None, self.x, 0),
                    (self.x2, H, w["gu"], self.gu, w["ln2"], None, 1),
                    (self.gu, 2 * I, w["down"], self.x, None, self.x2, 2),
                ]
                if li + 1 < nl:
                    nxt = self.layers[li + 1]
                    phases.append((self.x, H, nxt["qkv"], self.qkv, nxt["ln1"], None, 1))
                arr = (_cabi.GemvPhase * len(phases))()
                for i, (xin, ldx, lin, y, nw, res, xmode) in enumerate(phases):
                    arr[i] = _cabi.GemvPhase(xin.data_ptr(), ldx, lin.w.data_ptr(), lin.scales.data_ptr(), y.data_ptr(), lin.N, lin.K,
                                             0 if nw is None else nw.data_ptr(), 0 if res is None else res.data_ptr(), float(s.eps), xmode)
                _cabi.check(L.eetq_b200_w8a16_gemv_chain(ctypes.byref(arr), len(phases), _vp(self.chain_counters[li]), _vp(self.epoch),
                                                          pdl, st()), "eetq_b200_w8a16_gemv_chain")
        else:
            # p2p mode: each call returns the flag slot it publishes; the kernel that consumes its output waits on that slot
            s_x = None   # slot of the call that produced self.x (None: produced locally by the embedding gather)
            for li, w in enumerate(self.layers):
                s_qkv = self._gemv(self.x, H, w["qkv"], self.qkv, norm_w=w["ln1"], xmode=1, wait_slot=s_x, kv_layer=li)
                attention(li, wait_slot=s_qkv)
                # x2 = x + o_proj(attn); x = x2 + down(silu(gate) * up)   (ping-pong so no kernel reads what it writes)
                s_o = self._gemv(self.attn, H, w["o"], self.x2, residual_full=self.x)       # attn is local, x already waited for
                s_gu = self._gemv(self.x2, H, w["gu"], self.gu, norm_w=w["ln2"], xmode=1, wait_slot=s_o)
                s_x = self._gemv(self.gu, 2 * I, w["down"], self.x, xmode=2, residual_full=self.x2, wait_slot=s_gu)
        last_slot = s_x if (self._p2p is not None and not self.chain) else None
        _cabi.check(L.eetq_b200_decode_rmsnorm_p2p(_vp(self.x), _vp(self.norm_w), _vp(self.xn), 1, H, float(s.eps), self._flags_ptr(last_slot),
                                                   self.world, _vp(self.epoch), pdl, st()), "decode_rmsnorm")
        torch.matmul(self.xn, self.lm_head_w.t(), out=self.logits)       # fp16 lm_head (library GEMV; not quantised)
        self.token.copy_(torch.argmax(self.logits, dim=-1))               # greedy
        self.pos.add_(1)

    def capture(self):
        """Warm up (lazy module loads, cuBLAS handles) and capture the decode step into a CUDA graph."""
        saved_pos, saved_tok = self.pos.clone(), self.token.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._enqueue_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.pos.copy_(saved_pos)
        self.token.copy_(saved_tok)
        before = _cabi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue_step()
        self.launches_per_step = _cabi.launch_count() - before
        torch.cuda.synchronize(self.device)
        self.pos.copy_(saved_pos)
        self.token.copy_(saved_tok)

    def step(self):
        """Advance one token entirely on the device (token/pos buffers are updated by the graph)."""
        if self.graph is None:
            self.capture()
        self.graph.replay()

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
        """Process a prompt [T] with the batched (tcgen05) kernels, fill the KV cache, return the first new token."""
        s = self.shape
        T, H = tokens.shape[0], s.hidden
        assert T < self.max_ctx
        x = self.embed[tokens]
        cos, sin = self.cos[:T], self.sin[:T]

        def lin(inp, l: _ShardedLinear):
            y = w8_a16_gemm_bias(inp, l.w, l.scales, None)
            if self.world > 1:
                import torch.distributed as dist

                parts = [torch.empty_like(y) for _ in range(self.world)]
                dist.all_gather(parts, y, group=self.group)
                y = torch.cat(parts, -1)
            return y

        def rms(v, w):
            vf = v.float()
            return w * (vf * torch.rsqrt(vf.pow(2).mean(-1, keepdim=True) + s.eps)).to(v.dtype)

        for li, w in enumerate(self.layers):
            qkv = lin(rms(x, w["ln1"]), w["qkv"])
            q = apply_rope(qkv[:, :H].reshape(T, s.heads, s.head_dim), cos, sin)
            k = apply_rope(qkv[:, H:2 * H].reshape(T, s.heads, s.head_dim), cos, sin)
            v = qkv[:, 2 * H:].reshape(T, s.heads, s.head_dim)
            self.kcache[li, :, :T] = k.transpose(0, 1)
            self.vcache[li, :, :T] = v.transpose(0, 1)
            o = F.scaled_dot_product_attention(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), is_causal=True)
            x = x + lin(o.transpose(0, 1).reshape(T, H), w["o"])
            gu = lin(rms(x, w["ln2"]), w["gu"])
            x = x + lin(F.silu(gu[:, :s.inter]) * gu[:, s.inter:], w["down"])
        logits = torch.matmul(rms(x[-1:], self.norm_w), self.lm_head_w.t())
        self.token.copy_(torch.argmax(logits, dim=-1))
        self.pos.fill_(T)
        return self.token.clone()

    def set_context(self, ctx_len: int, token: int = 1):
        """Pretend `ctx_len` tokens are already cached (cache content stays as is) -- for benchmarks/tests."""
        self.pos.fill_(ctx_len)
        self.token.fill_(token)

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
        return sum(l.n_local * l.K for w in self.layers for l in (w["qkv"], w["o"], w["gu"], w["down"]))
