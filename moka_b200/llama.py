"""Frozen Llama-2 host wrapper that threads the MokA masks to the 7 projections of every layer.

This is the CALLER of the hot path, not part of it: the frozen math is
stock PyTorch (cuBLAS GEMMs, F.scaled_dot_product_attention -> flash kernels, RMSNorm, RoPE,
SiLU).  What it mirrors from the reference is the mask plumbing:

* AVT ``AudioVisualText/models/modeling_llama.py``: ``LlamaModel.forward`` unpacks
  ``inputs_embeds = [emb, text, video, audio, question]`` (:647-654) and every decoder layer
  passes the same list to ``q/k/v/o_proj`` (:326-328, :384) and ``gate/up/down_proj`` (:222-224);
  per-layer gradient checkpointing (:719-733).
* VT ``VisualText/modified_models/modeling_llama.py``: ``my_text_mask / my_image_mask /
  question_mask`` ride in kwargs (:495) to the same 7 projections (:152-159, :251-253, :279).

* Generation: ``KVCache`` + ``masks_for_step`` mirror the hand-off of
  ``VisualText/modified_models/modeling_llama.py:314-329`` - the masks reach the projections only while
  the cache is empty (training, or the prefill call of a generation); every later call runs the no-mask
  branch (``layer.py:672-678`` plain text-adapter LoRA; AVT ``lora.py:373-381`` for one-token steps).

Module names match HF Llama (``model.layers.N.self_attn.q_proj`` ...) so that
``get_peft_model`` target matching and checkpoint keys are the reference's.
Weights are random-initialised from the config (no checkpoints on the box).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from .layer import MokaLinearAVT, MokaLinearVT
from .metadata import TokenMeta, build_token_meta_avt, build_token_meta_vt


@dataclass
class LlamaConfig:
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    vocab_size: int = 32000
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    initializer_range: float = 0.02
    max_position_embeddings: int = 4096

    @staticmethod
    def llama2_7b() -> 'LlamaConfig':
        return LlamaConfig()

    @staticmethod
    def llama2_13b() -> 'LlamaConfig':
        return LlamaConfig(hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                           num_attention_heads=40)

    @staticmethod
    def tiny(layers: int = 2, hidden: int = 256, inter: int = 512, heads: int = 4, vocab: int = 512) -> 'LlamaConfig':
        return LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                           num_attention_heads=heads, vocab_size=vocab)


class KVCache:
    """Keys and values of the frozen attention, per layer, as ``[B, S_total, H, D]`` (the role HF's
    ``DynamicCache`` plays in the reference's generate loop)."""

    def __init__(self, num_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * num_layers
        self.v: List[Optional[torch.Tensor]] = [None] * num_layers

    def get_seq_length(self) -> int:
        return 0 if self.k[0] is None else int(self.k[0].shape[1])

    def update(self, layer_idx: int, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.k[layer_idx] is None:
            self.k[layer_idx], self.v[layer_idx] = k, v
        else:
            self.k[layer_idx] = torch.cat((self.k[layer_idx], k), dim=1)
            self.v[layer_idx] = torch.cat((self.v[layer_idx], v), dim=1)
        return self.k[layer_idx], self.v[layer_idx]


def masks_for_step(past_key_values: Optional[KVCache], mask_args: Optional[tuple]) -> Optional[tuple]:
    """The mask hand-off of ``VisualText/modified_models/modeling_llama.py:314-329``: no cache (training) or an
    empty cache (first call of a generation) -> the masks go to the 7 projections; a cache that already holds
    tokens (decode steps) -> ``None``, i.e. the no-mask branch of every MokaLinear."""
    if past_key_values is None or past_key_values.get_seq_length() == 0:
        return mask_args
    return None


def _proj(mod: nn.Module, x: torch.Tensor, mask_args: Optional[tuple]) -> torch.Tensor:
    """Call a projection the way the reference decoder layers do: MokA layers get the masks."""
    if mask_args is not None and isinstance(mod, (MokaLinearAVT, MokaLinearVT)):
        return mod(x, *mask_args)
    if isinstance(mod, MokaLinearVT):
        return mod(x, None, None, None)
    return mod(x)


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.variance_epsilon = eps

    def forward(self, x):
        return F.rms_norm(x, (x.shape[-1],), self.weight, self.variance_epsilon)


def _rope_tables(S: int, head_dim: int, theta: float, device, dtype):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float32) / head_dim))
    t = torch.arange(S, device=device, dtype=torch.float32)
    fr = torch.outer(t, inv)
    emb = torch.cat((fr, fr), dim=-1)
    return emb.cos().to(dtype)[None, :, None, :], emb.sin().to(dtype)[None, :, None, :]


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


class _RopeInplace(torch.autograd.Function):
    """Rotary embedding in place on a contiguous bf16 [B, S, H, D] tensor (libmoka.so helper
    kernel); the backward is the inverse rotation of the incoming gradient."""

    @staticmethod
    def forward(ctx, x, theta):
        from . import _cabi
        B, S, H, D = x.shape
        ctx.dims, ctx.theta = (B, S, H, D), theta
        _cabi.check(_cabi.load().moka_rope(x.data_ptr(), x.data_ptr(), B * S, S, H, D, theta, 0,
                                           torch.cuda.current_stream(x.device).cuda_stream), 'moka_rope')
        ctx.mark_dirty(x)
        return x

    @staticmethod
    def backward(ctx, g):
        from . import _cabi
        B, S, H, D = ctx.dims
        g = g.contiguous()
        out = torch.empty_like(g)                  # the incoming gradient is not ours to overwrite
        _cabi.check(_cabi.load().moka_rope(g.data_ptr(), out.data_ptr(), B * S, S, H, D, ctx.theta, 1,
                                           torch.cuda.current_stream(g.device).cuda_stream), 'moka_rope')
        return out, None


def apply_rope(x: torch.Tensor, rope, theta: float) -> torch.Tensor:
    """x: [B, S, H, D].  CUDA bf16 takes the fused in-place kernel; anything else the eager form."""
    if x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[-1] % 16 == 0:
        if x.requires_grad and x.is_leaf:
            x = x.clone()
        return _RopeInplace.apply(x, theta)
    cos, sin = rope
    return x * cos + _rotate_half(x) * sin


class LlamaAttention(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.num_heads = cfg.num_attention_heads
        self.head_dim = cfg.hidden_size // cfg.num_attention_heads
        self.rope_theta = cfg.rope_theta
        d = cfg.hidden_size
        self.q_proj = nn.Linear(d, d, bias=False)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=False)
        self.o_proj = nn.Linear(d, d, bias=False)

    def forward(self, h, rope, mask_args, cache: Optional[KVCache] = None, layer_idx: int = 0, pos0: int = 0):
        B, S, _ = h.shape
        H, D = self.num_heads, self.head_dim
        # keep the [B, S, H, D] memory layout of the projections; SDPA gets transposed views
        q = _proj(self.q_proj, h, mask_args).view(B, S, H, D)
        k = _proj(self.k_proj, h, mask_args).view(B, S, H, D)
        v = _proj(self.v_proj, h, mask_args).view(B, S, H, D)
        if pos0 == 0:
            q = apply_rope(q, rope, self.rope_theta)
            k = apply_rope(k, rope, self.rope_theta)
        else:
            # decode steps: `rope` holds the tables of the absolute positions pos0 .. pos0 + S - 1
            cos, sin = rope
            q = q * cos + _rotate_half(q) * sin
            k = k * cos + _rotate_half(k) * sin
        if cache is not None:
            k, v = cache.update(layer_idx, k, v)
        if pos0 == 0:
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
        else:
            # new token i (absolute position pos0 + i) sees the cached tokens and the new ones up to itself
            Sk = k.shape[1]
            allowed = torch.ones(S, Sk, dtype=torch.bool, device=h.device).tril(diagonal=Sk - S)
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                               attn_mask=allowed)
        o = o.transpose(1, 2).reshape(B, S, H * D)
        return _proj(self.o_proj, o, mask_args)


class LlamaMLP(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.gate_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.up_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.down_proj = nn.Linear(cfg.intermediate_size, cfg.hidden_size, bias=False)

    def forward(self, x, mask_args):
        return _proj(self.down_proj, F.silu(_proj(self.gate_proj, x, mask_args)) * _proj(self.up_proj, x, mask_args),
                     mask_args)


class LlamaDecoderLayer(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.self_attn = LlamaAttention(cfg)
        self.mlp = LlamaMLP(cfg)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, h, cos, sin, *mask_args, cache: Optional[KVCache] = None, layer_idx: int = 0, pos0: int = 0):
        ma = mask_args if mask_args else None
        h = h + self.self_attn(self.input_layernorm(h), (cos, sin), ma, cache, layer_idx, pos0)
        return h + self.mlp(self.post_attention_layernorm(h), ma)


class LlamaModel(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.config = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([LlamaDecoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.gradient_checkpointing = False
        # The reference's VT hand-off runs inside EACH decoder layer and asks ``past_key_value.get_seq_length()``,
        # which reads layer 0's cache (VisualText/modified_models/modeling_llama.py:314-329): during a generation
        # prefill layer 0 has already stored its keys when layer 1 asks, so layers 1..L-1 take the no-mask
        # (plain text-adapter LoRA) branch on the prompt.  Default here: every layer sees the masks on the prompt,
        # i.e. the prefill computes what training computed (a deliberate fix of that quirk).  Set this flag to get
        # the reference's behaviour bit for bit (masks to layer 0 only whenever a cache is passed).
        self.reference_compat_prefill = False

    def forward(self, input_ids=None, inputs_embeds=None, mask_args: Optional[tuple] = None,
                past_key_values: Optional[KVCache] = None):
        if isinstance(inputs_embeds, (list, tuple)):
            # AVT packing: [embeddings, text, video, audio, question] (modeling_llama.py:647-654)
            packed = list(inputs_embeds)
            inputs_embeds, mask_args = packed[0], (packed[1:5],)
        h = self.embed_tokens(input_ids) if inputs_embeds is None else inputs_embeds
        B, S, _ = h.shape
        hd = self.config.hidden_size // self.config.num_attention_heads
        pos0 = past_key_values.get_seq_length() if past_key_values is not None else 0
        mask_args = masks_for_step(past_key_values, mask_args)
        cos, sin = _rope_tables(pos0 + S, hd, self.config.rope_theta, h.device, h.dtype)
        cos, sin = cos[:, pos0:], sin[:, pos0:]
        ma = tuple(mask_args) if mask_args is not None else ()
        for i, layer in enumerate(self.layers):
            if past_key_values is not None:
                ma_i = ma if (i == 0 or not self.reference_compat_prefill) else ()
                h = layer(h, cos, sin, *ma_i, cache=past_key_values, layer_idx=i, pos0=pos0)
            elif self.gradient_checkpointing and self.training:
                h = checkpoint(layer, h, cos, sin, *ma, use_reentrant=False)
            else:
                h = layer(h, cos, sin, *ma)
        return self.norm(h)


class LlamaForCausalLM(nn.Module):
    """``forward(input_ids | inputs_embeds, labels, modality_masks=[t,v,a,q] | my_text_mask=...,
    my_image_mask=..., question_mask=... | token_meta=TokenMeta)`` -> (loss, logits)."""

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.config = cfg
        self.model = LlamaModel(cfg)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        std = self.config.initializer_range
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)

    def gradient_checkpointing_enable(self, enabled: bool = True):
        self.model.gradient_checkpointing = enabled

    def forward(self, input_ids=None, inputs_embeds=None, labels=None, modality_masks: Optional[Sequence] = None,
                my_text_mask=None, my_image_mask=None, question_mask=None,
                token_meta: Optional[TokenMeta] = None, return_logits: bool = False,
                past_key_values: Optional[KVCache] = None):
        if past_key_values is not None and past_key_values.get_seq_length() > 0:
            mask_args: Optional[tuple] = None          # decode step: the masks are not even looked at (:326-329)
        elif token_meta is not None:
            mask_args = (token_meta,)
        elif modality_masks is not None:
            # build the token metadata ONCE per step and hand it to all 7 x L projections
            mask_args = (build_token_meta_avt(list(modality_masks)),)
        elif my_text_mask is not None:
            mask_args = (build_token_meta_vt(my_text_mask, my_image_mask, question_mask),)
        else:
            mask_args = None
        h = self.model(input_ids=input_ids, inputs_embeds=inputs_embeds, mask_args=mask_args,
                       past_key_values=past_key_values)
        loss = None
        logits = None
        if labels is not None:
            loss = chunked_causal_lm_loss(h, self.lm_head.weight, labels)
        if return_logits or labels is None:
            logits = self.lm_head(h)
        return loss, logits

    @torch.no_grad()
    def generate(self, input_ids=None, inputs_embeds=None, max_new_tokens: int = 16, eos_token_id: Optional[int] = None,
                 **mask_kwargs) -> torch.Tensor:
        """Greedy decoding with a KV cache.  The prefill call carries the masks (``modality_masks=`` /
        ``my_text_mask=...`` / ``token_meta=``), every later step is a one-token call whose projections take the
        no-mask branch - the call pattern of the reference's ``model.generate`` (inference.py of both variants).
        AVT layers need ``loramethod`` containing 'test' for the one-token branch (lora.py:373)."""
        cache = KVCache(len(self.model.layers))
        _, logits = self.forward(input_ids=input_ids, inputs_embeds=inputs_embeds, past_key_values=cache, **mask_kwargs)
        nxt = logits[:, -1].argmax(-1)
        out = [nxt]
        done = torch.zeros_like(nxt, dtype=torch.bool)
        for _ in range(max_new_tokens - 1):
            if eos_token_id is not None:
                done |= nxt == eos_token_id
                if bool(done.all()):
                    break
            _, logits = self.forward(input_ids=nxt[:, None], past_key_values=cache, **mask_kwargs)
            nxt = logits[:, -1].argmax(-1)
            out.append(nxt)
        return torch.stack(out, dim=1)


def chunked_causal_lm_loss(h: torch.Tensor, lm_weight: torch.Tensor, labels: torch.Tensor,
                           chunk: int = 2048) -> torch.Tensor:
    """Shifted next-token cross entropy, computed in token chunks so the fp32 logits never exist
    for the whole batch at once (HF computes the same loss un-chunked)."""
    B, S, d = h.shape
    hs = h[:, :-1, :].reshape(-1, d)
    tg = labels[:, 1:].reshape(-1)
    total = hs.new_zeros((), dtype=torch.float32)
    count = (tg != -100).sum().clamp_min(1)
    for i in range(0, hs.shape[0], chunk):
        logits = F.linear(hs[i:i + chunk], lm_weight).float()
        total = total + F.cross_entropy(logits, tg[i:i + chunk], ignore_index=-100, reduction='sum')
    return total / count
