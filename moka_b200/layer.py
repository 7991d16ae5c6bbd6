"""MokaLinear: the host-side mirror of the two reference MokA layers, driving libmoka.so.

* ``MokaLinearAVT`` mirrors ``AudioVisualText/peft_hyper/tuners/lora.py:277-532``
  (``class Linear(nn.Linear, LoraLayer)``): same constructor arguments, same parameter names
  (``weight``, ``lora_A{i}.weight``, ``lora_B0.weight``), ``forward(x, modality_mask)``.
* ``MokaLinearVT`` mirrors ``VisualText/modified_peft/tuners/lora/layer.py:389-681``
  (``class Linear(nn.Module, LoraLayer)``): ``base_layer``, ``lora_A`` / ``lora_B`` ModuleDicts,
  ``update_layer``, ``set_adapter``, ``forward(x, my_text_mask, my_image_mask, question_mask)``.

Both route through one autograd Function whose forward/backward are the C-ABI calls
``moka_forward`` / ``moka_backward``.  There is no PyTorch or CPU implementation of the adapter
math in this package: CPU tensors or a missing library raise.
"""
from __future__ import annotations

import math
import os
from typing import Any, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .config import parse_avt_rank
from .metadata import (META_CACHE, TokenMeta, all_text_meta, build_token_meta_avt, build_token_meta_vt)

_DTYPES = {torch.bfloat16: _cabi.MOKA_BF16, torch.float32: _cabi.MOKA_F32}

# one scratch buffer per device, grown on demand (calls are stream-ordered on the current stream)
_WORKSPACE = {}

# optional per-call CUDA-event timing of the C-ABI calls (bench.py turns it on)
TIMING = {'enabled': False, 'events': [], 'marks': False}      # marks: tools/instep_timeline.py (globaltimer markers)

# Frozen-GEMM epilogue fusion (MOKA_FLAG_DELTA_OUT): the adapter kernels write their term
# alone (no read of W0 x / W0^T dy) and the cuBLAS GEMM accumulates onto it with beta = 1.  Opt-in
# (MOKA_DELTA=1 in the environment or FUSION['enabled'] = True) wherever the tcgen05 chain serves the shape:
# measured on B200 it moves 32 % of the adapter's bytes into the GEMM epilogue but gains only ~0.5 % of the 7B
# step (profiles/r01_delta_out.md), so the default stays the read-modify-write order (GEMM first, adapter
# added in place), whose numbers are the ones SURVEY 8(d)'s byte model describes.
FUSION = {'enabled': os.environ.get('MOKA_DELTA', '0') == '1', 'unsupported': set()}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WORKSPACE[key] = ws
    return ws


class FusedScalars:
    """pre / post / w of SURVEY.md A.1-A.2.  AVT scales between A and B (lora.py:477), VT after B
    (layer.py:662, 668)."""

    def __init__(self, pre: float, post: Sequence[float], w: float, p: float):
        self.pre, self.post, self.w, self.p = float(pre), tuple(float(v) for v in post), float(w), float(p)


def _make_desc(x2d_shape, B, S, d_out, r, M, dtype, Lk, sc: FusedScalars, flags: int) -> _cabi.MokaDesc:
    d = _cabi.MokaDesc()
    d.B, d.S, d.d_in, d.d_out, d.r, d.M = B, S, x2d_shape[-1], d_out, r, M
    d.dtype, d.Lk = _DTYPES[dtype], Lk
    d.pre = sc.pre
    post = list(sc.post) + [0.0] * (3 - len(sc.post))
    d.post[0], d.post[1], d.post[2] = post[0], post[1], post[2]
    d.attn_w, d.dropout_p, d.flags = sc.w, sc.p, flags
    return d


class _MokaFn(torch.autograd.Function):
    """y = W0 x (+bias) + adapter(x).  The frozen GEMM is cuBLAS (F.linear), everything else is
    libmoka.so.  Adapter-weight gradients are accumulated in fp32; if a parameter carries a
    ``main_grad`` fp32 tensor (the flat data-parallel bucket, moka_b200/dp.py) the kernels write
    straight into it and autograd receives ``None`` for that parameter."""

    @staticmethod
    def forward(ctx, x, weight, bias, Bw, meta: TokenMeta, sc: FusedScalars, flags: int, n_a: int, *A):
        lib = _cabi.load()
        if not x.is_cuda:
            raise _cabi.MokaError('MokaLinear runs on CUDA only (sm_100a); there is no CPU fallback')
        if x.dtype not in _DTYPES:
            raise _cabi.MokaError(f'unsupported dtype {x.dtype}: bf16 or fp32')
        for t in (weight, Bw, *A):
            if t.dtype != x.dtype:
                raise _cabi.MokaError(f'adapter/base weights must have the dtype of x ({x.dtype}), got {t.dtype}')
        shp = x.shape
        d_in, d_out, r = shp[-1], weight.shape[0], A[0].shape[0]
        B, S = meta.B, meta.S
        if x.numel() != B * S * d_in:
            raise ValueError(f'x has shape {tuple(shp)} but the masks describe {B}x{S} tokens')
        x2 = x.reshape(B * S, d_in)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        desc = _make_desc(x2.shape, B, S, d_out, r, len(A), x.dtype, meta.Lk, sc, flags)
        shape_key = (B, S, d_in, d_out, r, len(A), x.dtype, meta.Lk, flags, sc.p > 0.0)
        delta = (FUSION['enabled'] and x.dtype == torch.bfloat16 and shape_key not in FUSION['unsupported'])
        if delta and not lib.moka_supports_delta(desc, 1):
            FUSION['unsupported'].add(shape_key)
            delta = False
        if delta:
            # adapter term first (write-only), frozen GEMM accumulated onto it below (beta = 1)
            y = torch.empty(*shp[:-1], d_out, dtype=x.dtype, device=x.device)
            desc.flags = flags | _cabi.FLAG_DELTA_OUT
        else:
            # frozen base, cuBLAS; the adapter is added in place.  Computed in the caller's shape so the
            # returned tensor is not a view made inside this Function (callers may modify it in place).
            y = F.linear(x2.view(shp), weight, bias)
        h = torch.empty(2 * B * S * r + B * S, dtype=torch.float32, device=x.device)   # h, g, lse
        seed = offset = 0
        if sc.p > 0.0:
            # CPU generator draws are restored by torch.utils.checkpoint, so a recomputed forward
            # reproduces its dropout mask; no device sync involved.
            seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
            offset = int(torch.randint(0, 2 ** 62, (1,)).item())
        Ac = [a if a.is_contiguous() else a.contiguous() for a in A]
        Bc = Bw if Bw.is_contiguous() else Bw.contiguous()
        ws = _workspace(x.device, int(lib.moka_workspace_bytes(desc)))
        stream = torch.cuda.current_stream(x.device).cuda_stream
        ev = None
        if TIMING['enabled']:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        with torch.cuda.device(x.device):
            code = lib.moka_forward(desc, x2.data_ptr(), y.data_ptr(), y.data_ptr(),
                                    _cabi.ptr_array([a.data_ptr() for a in Ac]), Bc.data_ptr(),
                                    meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                    h.data_ptr(), seed, offset, ws.data_ptr(), ws.numel(), stream)
        if delta and code == _cabi.MOKA_EUNSUPPORTED:
            # a shape the write-only kernels cannot tile after all: remember it, redo in the usual order
            FUSION['unsupported'].add(shape_key)
            delta = False
            desc.flags = flags
            y = F.linear(x2.view(shp), weight, bias)
            with torch.cuda.device(x.device):
                code = lib.moka_forward(desc, x2.data_ptr(), y.data_ptr(), y.data_ptr(),
                                        _cabi.ptr_array([a.data_ptr() for a in Ac]), Bc.data_ptr(),
                                        meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                        h.data_ptr(), seed, offset, ws.data_ptr(), ws.numel(), stream)
        _cabi.check(code, 'moka_forward')
        if ev is not None:
            ev[1].record()
            TIMING['events'].append(('fwd', B * S, d_in, d_out, x.element_size(), ev, delta))
        if delta:
            y.view(B * S, d_out).addmm_(x2, weight.t())          # y = adapter term + W0 x, one rounding
            if bias is not None:
                y.add_(bias)
        ctx.save_for_backward(x2, h, weight, Bc, *Ac)
        ctx.meta, ctx.sc, ctx.flags, ctx.seed, ctx.offset = meta, sc, flags, seed, offset
        ctx.delta = delta
        ctx.x_shape = shp
        ctx.param_refs = (Bw, A)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.load()
        x2, h, weight, Bc, *Ac = ctx.saved_tensors
        meta, sc = ctx.meta, ctx.sc
        B, S = meta.B, meta.S
        d_in, d_out, r = x2.shape[1], weight.shape[0], Ac[0].shape[0]
        dy2 = dy.reshape(B * S, d_out)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        need_dx = ctx.needs_input_grad[0]
        delta = ctx.delta and need_dx
        if delta:
            dx = torch.empty(B * S, d_in, dtype=dy2.dtype, device=dy2.device)   # adapter term first, GEMM onto it
        else:
            dx = torch.matmul(dy2, weight) if need_dx else None    # frozen dgrad, cuBLAS
        Bw_param, A_params = ctx.param_refs
        dev = x2.device

        def grad_buf(param, shape):
            mg = getattr(param, 'main_grad', None)
            if mg is not None:
                return mg, True
            return torch.zeros(shape, dtype=torch.float32, device=dev), False

        dB, dB_main = grad_buf(Bw_param, (d_out, r))
        dAs = [grad_buf(p, (r, d_in)) for p in A_params]
        desc = _make_desc(x2.shape, B, S, d_out, r, len(Ac), x2.dtype, meta.Lk, sc,
                          ctx.flags | (_cabi.FLAG_DELTA_OUT if delta else 0))
        nbytes = int(lib.moka_workspace_bytes(desc))
        ws = _workspace(dev, nbytes)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ev = None
        if TIMING['enabled']:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if TIMING['marks']:
            lib.moka_debug_mark(0, stream)
        with torch.cuda.device(dev):
            code = lib.moka_backward(desc, dy2.data_ptr(), x2.data_ptr(), h.data_ptr(),
                                     _cabi.ptr_array([a.data_ptr() for a in Ac]), Bc.data_ptr(),
                                     meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                     0 if dx is None else dx.data_ptr(),
                                     _cabi.ptr_array([g.data_ptr() for g, _ in dAs]), dB.data_ptr(),
                                     ctx.seed, ctx.offset, ws.data_ptr(), ws.numel(), stream)
        _cabi.check(code, 'moka_backward')
        if TIMING['marks']:
            lib.moka_debug_mark(1, stream)
        if ev is not None:
            ev[1].record()
            TIMING['events'].append(('bwd', B * S, d_in, d_out, x2.element_size(), ev, delta))
        if delta:
            dx.addmm_(dy2, weight)                                 # dx = adapter term + W0^T dy, one rounding
        # tell the data-parallel engine that these parameters' gradients are complete
        for prm in (Bw_param, *A_params):
            cb = getattr(prm, '_moka_dp_ready', None)
            if cb is not None:
                cb()
        gB = None if dB_main else dB.to(Bw_param.dtype)
        gA = [None if main else g.to(p.dtype) for (g, main), p in zip(dAs, A_params)]
        gx = None if dx is None else dx.view(ctx.x_shape)
        # bias='all' / 'lora_only' (lora.py:234-245) makes the frozen layer's bias trainable: its gradient is the
        # column sum of dy.  A base weight that still requires grad (a layer used outside get_peft_model, which
        # freezes it: lora.py:333, layer.py:124-132) gets what autograd would give it: dW0 = dy^T x (cuBLAS).
        gW = torch.matmul(dy2.t(), x2) if ctx.needs_input_grad[1] else None
        gbias = dy2.float().sum(0).to(dy2.dtype) if ctx.needs_input_grad[2] else None
        return (gx, gW, gbias, gB, None, None, None, None, *gA)


def moka_linear_function(x, weight, bias, A: Sequence[torch.Tensor], Bw, meta: TokenMeta,
                         sc: FusedScalars, flags: int = 0):
    return _MokaFn.apply(x, weight, bias, Bw, meta, sc, flags, len(A), *A)


# =========================================================================== AVT facade
class MokaLinearAVT(nn.Linear):
    """Drop-in for ``peft_hyper.tuners.lora.Linear`` (lora.py:277-532).

    ``r`` follows the reference's digit encoding (``r=444, lora_nums=3`` -> rank 4 x 3) unless
    ``rank`` is given explicitly, which is how ranks >= 10 are expressed."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1,
                 lora_nums: int = 2, blc_alpha: float = 0.0, blc_weight: float = 0.0,
                 lora_dropout: float = 0.0, reserved_modality='text', loramethod='uni',
                 fan_in_fan_out: bool = False, merge_weights: bool = True,
                 rank: Optional[int] = None, kernel_flags: int = 0, **kwargs):
        super().__init__(in_features, out_features, **kwargs)
        if fan_in_fan_out:
            raise NotImplementedError('fan_in_fan_out (Conv1D weights) is not used by the Llama targets')
        ranks = [rank] * lora_nums if rank is not None else parse_avt_rank(r, lora_nums)
        if len(set(ranks)) != 1:
            raise ValueError(f'per-modality ranks must be equal, got {ranks} (lora.py:525 sums them)')
        if not (2 <= lora_nums <= 3):
            raise ValueError('lora_nums must be 2 or 3')
        self.r = list(ranks)
        self.d_k = ranks[0]
        self.lora_alpha = lora_alpha
        self.lora_num = lora_nums
        self.blc_alpha = blc_alpha
        self.blc_weight = blc_weight
        self.reserved_modality = reserved_modality
        self.loramethod = loramethod
        self.fan_in_fan_out = fan_in_fan_out
        self.lora_dropout_p = float(lora_dropout)
        self.merged = False
        self.merge_weights = merge_weights
        self.disable_adapters = False
        self.kernel_flags = kernel_flags
        if ranks[0] > 0:
            for i in range(lora_nums):
                setattr(self, f'lora_A{i}', nn.Linear(in_features, ranks[i], bias=False))
            self.lora_B0 = nn.Linear(ranks[0], out_features, bias=False)
            self.scaling = [self.lora_alpha / ranks[0]]
            self.weight.requires_grad = False
        self.reset_parameters()

    def reset_parameters(self):
        nn.Linear.reset_parameters(self)
        if hasattr(self, 'lora_A0'):
            for i in range(self.lora_num):
                nn.init.kaiming_uniform_(getattr(self, f'lora_A{i}').weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B0.weight)

    def _scalars(self) -> FusedScalars:
        p = self.lora_dropout_p if self.training else 0.0
        return FusedScalars(self.scaling[0], (1.0, 1.0, 1.0), self.blc_weight, p)

    def forward(self, x: torch.Tensor, modality_mask: Any = None):
        if self.disable_adapters or self.r[0] == 0:
            return F.linear(x, self.weight, self.bias)
        A = [getattr(self, f'lora_A{i}').weight for i in range(self.lora_num)]
        if 'test' in self.loramethod and x.size(1) == 1:
            # decode step: text adapter only (lora.py:373-381)
            meta = all_text_meta(x.shape[0], 1, x.device)
            return moka_linear_function(x, self.weight, self.bias, A[:1], self.lora_B0.weight, meta,
                                        self._scalars(), self.kernel_flags)
        if isinstance(modality_mask, TokenMeta):
            meta = modality_mask
        else:
            if modality_mask is None:
                raise TypeError('modality_mask=[text, video, audio, question] is required (lora.py:462-465)')
            masks = list(modality_mask)
            meta = META_CACHE.get(_cabi.VARIANT_AVT, masks, lambda: build_token_meta_avt(masks))
        return moka_linear_function(x, self.weight, self.bias, A, self.lora_B0.weight, meta,
                                    self._scalars(), self.kernel_flags)

    def extra_repr(self) -> str:
        return (f'{super().extra_repr()}, moka=avt, r={self.r}, alpha={self.lora_alpha}, '
                f'blc_weight={self.blc_weight}, dropout={self.lora_dropout_p}')


# =========================================================================== VT facade
class MokaLinearVT(nn.Module):
    """Drop-in for ``modified_peft.tuners.lora.layer.Linear`` (layer.py:389-681) restricted to
    what MokA uses: a wrapped ``base_layer``, named adapters in ``lora_A`` / ``lora_B``
    ModuleDicts, per-adapter ``scaling`` / ``r`` / ``lora_alpha`` dicts, ``attn_weight``."""

    adapter_layer_names = ('lora_A', 'lora_B', 'lora_embedding_A', 'lora_embedding_B')
    other_param_names = ('r', 'lora_alpha', 'scaling', 'lora_dropout')

    def __init__(self, base_layer: nn.Module, adapter_name: str, r: int = 0, lora_alpha: int = 1,
                 lora_dropout: float = 0.0, fan_in_fan_out: bool = False,
                 is_target_conv_1d_layer: bool = False, init_lora_weights=True, use_rslora: bool = False,
                 use_dora: bool = False, kernel_flags: int = 0, **kwargs) -> None:
        super().__init__()
        if not isinstance(base_layer, nn.Linear):
            raise TypeError('MokaLinearVT wraps nn.Linear base layers only')
        if use_dora:
            raise NotImplementedError('DoRA is not part of the MokA path')
        if fan_in_fan_out:
            raise NotImplementedError('fan_in_fan_out is not used by the Llama targets')
        self.base_layer = base_layer
        self.in_features, self.out_features = base_layer.in_features, base_layer.out_features
        self.r, self.lora_alpha, self.scaling = {}, {}, {}
        self.lora_dropout = nn.ModuleDict({})
        self.lora_A = nn.ModuleDict({})
        self.lora_B = nn.ModuleDict({})
        self.lora_embedding_A = nn.ParameterDict({})
        self.lora_embedding_B = nn.ParameterDict({})
        self._disable_adapters = False
        self.merged_adapters: List[str] = []
        self.fan_in_fan_out = fan_in_fan_out
        self.attn_weight = kwargs.get('attn_weight', 1)        # layer.py:408
        self.kernel_flags = kernel_flags
        self._active_adapter = adapter_name
        self.update_layer(adapter_name, r, lora_alpha=lora_alpha, lora_dropout=lora_dropout,
                          init_lora_weights=init_lora_weights, use_rslora=use_rslora, use_dora=use_dora)
        self.is_target_conv_1d_layer = is_target_conv_1d_layer

    # ---- the slice of BaseTunerLayer the injection code touches (tuners_utils.py:438-582)
    def get_base_layer(self) -> nn.Module:
        return self.base_layer

    @property
    def weight(self) -> torch.Tensor:
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    @property
    def merged(self) -> bool:
        return bool(self.merged_adapters)

    @property
    def disable_adapters(self) -> bool:
        return self._disable_adapters

    @property
    def active_adapter(self):
        return self._active_adapter

    @property
    def active_adapters(self):
        return [self._active_adapter] if isinstance(self._active_adapter, str) else list(self._active_adapter)

    def enable_adapters(self, enabled: bool) -> None:
        self._disable_adapters = not enabled
        for name in self.lora_A:
            self.lora_A[name].requires_grad_(enabled)
            self.lora_B[name].requires_grad_(enabled)

    def set_adapter(self, adapter_names) -> None:
        if isinstance(adapter_names, str):
            adapter_names = [adapter_names]
        for name in self.lora_A:
            on = name in adapter_names
            self.lora_A[name].requires_grad_(on)
            self.lora_B[name].requires_grad_(on)
        self._active_adapter = list(adapter_names)

    def update_layer(self, adapter_name, r, lora_alpha, lora_dropout, init_lora_weights=True,
                     use_rslora=False, use_dora: bool = False):
        """layer.py:93-140 (kaiming A / zero B, scaling alpha/r, adapters in the base dtype)."""
        if r <= 0:
            raise ValueError(f'`r` should be a positive integer value but the value passed is {r}')
        for other in self.r.values():
            if other != r:
                raise ValueError('MokA shares one projection width between the adapters (layer.py:600): '
                                 f'rank {r} != {other}')
        self.r[adapter_name] = r
        self.lora_alpha[adapter_name] = lora_alpha
        self.lora_dropout.update(nn.ModuleDict(
            {adapter_name: nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else nn.Identity()}))
        w = self.base_layer.weight
        self.lora_A[adapter_name] = nn.Linear(self.in_features, r, bias=False, device=w.device, dtype=w.dtype)
        self.lora_B[adapter_name] = nn.Linear(r, self.out_features, bias=False, device=w.device, dtype=w.dtype)
        self.scaling[adapter_name] = lora_alpha / math.sqrt(r) if use_rslora else lora_alpha / r
        if init_lora_weights:
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.set_adapter(self.active_adapters)

    def merge(self, *a, **k):
        raise NotImplementedError('the MokA update is not a weight delta (cross-modal attention): cannot merge')

    def unmerge(self, *a, **k):
        raise NotImplementedError('nothing to unmerge')

    def _dropout_p(self, name: str) -> float:
        mod = self.lora_dropout[name]
        return float(mod.p) if (isinstance(mod, nn.Dropout) and self.training) else 0.0

    def forward(self, x: torch.Tensor, my_text_mask: Any = None, my_image_mask: Optional[torch.Tensor] = None,
                question_mask: Optional[torch.Tensor] = None, *args: Any, **kwargs: Any) -> torch.Tensor:
        if self._disable_adapters or 'text' not in self.lora_A:
            return self.base_layer(x)
        s_text = self.scaling['text']
        bw = self.base_layer
        if my_text_mask is None:
            # decode steps / no masks: plain LoRA with the text adapter (layer.py:672-678)
            x3 = x if x.dim() == 3 else x.unsqueeze(0)
            meta = all_text_meta(x3.shape[0], x3.shape[1], x.device)
            sc = FusedScalars(1.0, (s_text, s_text, s_text), 0.0, self._dropout_p('text'))
            y = moka_linear_function(x3, bw.weight, bw.bias, [self.lora_A['text'].weight],
                                     self.lora_B['text'].weight, meta, sc, self.kernel_flags)
            return y if x.dim() == 3 else y.squeeze(0)
        if 'image' not in self.lora_A:
            raise KeyError("MokA needs both the 'text' and the 'image' adapter (train.py:559-561)")
        if isinstance(my_text_mask, TokenMeta):
            meta = my_text_mask
        else:
            masks = [my_text_mask, my_image_mask, question_mask]
            meta = META_CACHE.get(_cabi.VARIANT_VT, masks,
                                  lambda: build_token_meta_vt(my_text_mask, my_image_mask, question_mask))
        p_drop = self._dropout_p('text')
        if self._dropout_p('image') != p_drop:
            # layer.py:607 / :617 apply lora_dropout['text'] to text rows and lora_dropout['image'] to image rows;
            # the kernels take one p (both reference recipes configure the two adapters identically)
            raise NotImplementedError(f"lora_dropout differs between the adapters (text {p_drop}, image "
                                      f"{self._dropout_p('image')}): the fused kernels take one dropout rate")
        sc = FusedScalars(1.0, (s_text, self.scaling['image'], 0.0), self.attn_weight, p_drop)
        A = [self.lora_A['text'].weight, self.lora_A['image'].weight]
        return moka_linear_function(x, bw.weight, bw.bias, A, self.lora_B['text'].weight, meta, sc,
                                    self.kernel_flags)

    def __repr__(self) -> str:
        return 'moka.' + super().__repr__()
