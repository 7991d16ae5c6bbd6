"""Token metadata: the reference's mask tensors -> (tokmeta, kidx, klen), built once per
training step on the device by ``moka_build_token_meta`` and shared by every MokaLinear call
of that step (the reference re-derives indices inside every layer call:
AudioVisualText/peft_hyper/tuners/lora.py:489, VisualText/modified_peft/tuners/lora/layer.py:603-637).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi


@dataclass
class TokenMeta:
    """Device-resident token metadata of one batch.

    tokmeta uint8 [B,S]: low 3 bits modality code (0 none, 1 text, 2 video/image, 3 audio),
    bit 3 = this token's projection is a usable key row.  kidx int32 [B,Lk]: in-sample
    positions of the key rows.  klen int32 [B]."""
    tokmeta: torch.Tensor
    kidx: torch.Tensor
    klen: torch.Tensor
    variant: int
    Lk: int
    max_klen: Optional[int] = None      # known only after a host check
    status: Optional[torch.Tensor] = None   # int32 [4] on the device: {MOKA_META_* bits, longest key list, 0, 0}

    def raise_if_error(self) -> None:
        """The deferred error check of a metadata object built with ``check=False``: ONE device -> host read, to be
        made where a synchronisation happens anyway (e.g. next to the loss read at the end of the step).  Raises what
        the reference raises: IndexError for an AVT sample without question tokens (lora.py:490)."""
        if self.status is None:
            return
        _raise_for_status(self.status.tolist(), self.Lk)

    @property
    def B(self) -> int:
        return self.tokmeta.shape[0]

    @property
    def S(self) -> int:
        return self.tokmeta.shape[1]


def _flat_mask(m: torch.Tensor, B: int, S: int) -> torch.Tensor:
    if m.numel() != B * S:
        raise ValueError(f'mask has {m.numel()} elements, expected {B}x{S}')
    m = m.reshape(B, S)
    if m.dtype == torch.bool:
        m = m.view(torch.uint8) if m.is_contiguous() else m.contiguous().view(torch.uint8)
    elif m.dtype not in (torch.uint8, torch.int8, torch.int32, torch.int64):
        m = m.to(torch.int32)
    return m.contiguous()


def _raise_for_status(st, cap: int) -> None:
    if st[0] & _cabi.META_BAD_SEGMENT:
        raise ValueError('segment table: unknown kind, negative length or more tokens than the sequence length')
    if st[0] & _cabi.META_OVERLAP:
        raise ValueError('modality masks overlap: a token may belong to at most one modality')
    if st[0] & _cabi.META_EMPTY_QUESTION:
        raise IndexError('index 0 is out of bounds for dimension 0 with size 0 '
                         '(a sample has no question token; reference: peft_hyper/tuners/lora.py:490)')
    if st[0] & _cabi.META_KEY_OVERFLOW:
        raise ValueError(f'a sample has {st[1]} key rows but Lk={cap}')


def build_token_meta(variant: int, text: torch.Tensor, vis: torch.Tensor, aud: Optional[torch.Tensor],
                     question: torch.Tensor, *, check: bool = True, Lk: Optional[int] = None) -> TokenMeta:
    """Run the device-side builder.  ``check=True`` costs ONE host sync per step and gives the
    reference's error behaviour: AVT raises IndexError on a sample without question tokens
    (lora.py:490); overlapping modality masks raise ValueError.  It also trims the key-list
    stride to the longest key list."""
    if not text.is_cuda:
        raise _cabi.MokaError('token metadata is built on the GPU: masks must be CUDA tensors (no CPU fallback)')
    lib = _cabi.load()
    B, S = text.shape[0], text.shape[1]
    masks = [_flat_mask(text, B, S), _flat_mask(vis, B, S), None if aud is None else _flat_mask(aud, B, S),
             _flat_mask(question, B, S)]
    eb = masks[0].element_size()
    for m in masks:
        if m is not None and m.element_size() != eb:
            raise ValueError('all masks must share one dtype')
    dev = text.device
    cap = S if Lk is None else int(Lk)
    tokmeta = torch.empty(B, S, dtype=torch.uint8, device=dev)
    kidx = torch.empty(B, cap, dtype=torch.int32, device=dev)
    klen = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    ptrs = _cabi.ptr_array([0 if m is None else m.data_ptr() for m in masks])
    with torch.cuda.device(dev):
        code = lib.moka_build_token_meta(variant, B, S, cap, ptrs, eb, tokmeta.data_ptr(), kidx.data_ptr(),
                                         klen.data_ptr(), status.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(code, 'moka_build_token_meta')
    meta = TokenMeta(tokmeta, kidx, klen, variant, cap, status=status)
    if check:
        st = status.tolist()                       # the one sync of the step
        _raise_for_status(st, cap)
        mk = max(int(st[1]), 1)
        meta.max_klen = mk
        if mk < cap:
            meta.kidx = kidx[:, :mk].contiguous()
            meta.Lk = mk
    return meta


def build_token_meta_avt(modality_mask: Sequence[torch.Tensor], *, check: bool = True) -> TokenMeta:
    """AVT: ``modality_mask = [text, video, audio, question]`` each [B,S,1] int32
    (AudioVisualText/models/unified_arch.py:336, consumed at lora.py:462-465)."""
    text, video, audio, question = modality_mask[0], modality_mask[1], modality_mask[2], modality_mask[3]
    return build_token_meta(_cabi.VARIANT_AVT, text, video, audio, question, check=check)


def build_token_meta_vt(my_text_mask: torch.Tensor, my_image_mask: torch.Tensor,
                        question_mask: torch.Tensor, *, check: bool = True) -> TokenMeta:
    """VT: three [B,S] bool masks (VisualText/train/train.py:206-231, consumed at layer.py:548)."""
    return build_token_meta(_cabi.VARIANT_VT, my_text_mask, my_image_mask, None, question_mask, check=check)


def _finish(meta: TokenMeta, check: bool) -> TokenMeta:
    if check:
        st = meta.status.tolist()
        _raise_for_status(st, meta.Lk)
        mk = max(int(st[1]), 1)
        meta.max_klen = mk
        if mk < meta.Lk:
            meta.kidx = meta.kidx[:, :mk].contiguous()
            meta.Lk = mk
    return meta


def build_token_meta_vt_from_ids(input_ids: torch.Tensor, labels: torch.Tensor, image_token_id: int, *,
                                 lengths: Optional[torch.Tensor] = None, ignore_index: int = -100,
                                 max_keys: int = 128, check: bool = False) -> TokenMeta:
    """SURVEY 8(f) N1, VT: the token metadata straight from the collated ``input_ids`` / ``labels`` [B, S] (image
    positions still hold ``image_token_id``) and the unpadded ``lengths`` [B] -- the rule of
    VisualText/train/train.py:206-231 evaluated on the device, no mask tensors, and with ``check=False`` (default) no
    host synchronisation: the status word stays on the device (``meta.raise_if_error()`` reads it when convenient).
    ``max_keys`` bounds the question length (the kernels pick their attention form from it)."""
    if not input_ids.is_cuda:
        raise _cabi.MokaError('token metadata is built on the GPU: input_ids must be a CUDA tensor (no CPU fallback)')
    lib = _cabi.load()
    B, S = input_ids.shape
    dev = input_ids.device
    ids = input_ids.to(torch.int64).contiguous()
    lab = labels.to(torch.int64).contiguous()
    ln = None if lengths is None else lengths.to(device=dev, dtype=torch.int32).contiguous()
    cap = int(max_keys)
    tokmeta = torch.empty(B, S, dtype=torch.uint8, device=dev)
    kidx = torch.empty(B, cap, dtype=torch.int32, device=dev)
    klen = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        code = lib.moka_build_token_meta_from_ids(B, S, cap, ids.data_ptr(), lab.data_ptr(), int(image_token_id),
                                                  int(ignore_index), 0 if ln is None else ln.data_ptr(),
                                                  tokmeta.data_ptr(), kidx.data_ptr(), klen.data_ptr(),
                                                  status.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(code, 'moka_build_token_meta_from_ids')
    return _finish(TokenMeta(tokmeta, kidx, klen, _cabi.VARIANT_VT, cap, status=status), check)


def build_token_meta_avt_from_segments(segments: torch.Tensor, S: int, *, max_keys: int = 128,
                                       check: bool = False) -> TokenMeta:
    """SURVEY 8(f) N1, AVT: ``segments`` int32 [B, max_segments, 2] = (kind, length) per segment in sequence order
    (``_cabi.SEG_TEXT / SEG_VIDEO / SEG_AUDIO / SEG_QUESTION``, ``SEG_END``-terminated) -- the table
    ``prepare_multimodal_inputs`` walks (AudioVisualText/models/unified_arch.py:139-240) -- left-padded to ``S`` like
    the reference pads (:304-312).  No mask tensors; no host synchronisation unless ``check=True``."""
    if not segments.is_cuda:
        raise _cabi.MokaError('token metadata is built on the GPU: segments must be a CUDA tensor (no CPU fallback)')
    lib = _cabi.load()
    seg = segments.to(torch.int32).contiguous()
    B, NS = seg.shape[0], seg.shape[1]
    dev = seg.device
    cap = int(max_keys)
    tokmeta = torch.empty(B, S, dtype=torch.uint8, device=dev)
    kidx = torch.empty(B, cap, dtype=torch.int32, device=dev)
    klen = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        code = lib.moka_build_token_meta_from_segments(B, int(S), cap, NS, seg.data_ptr(), tokmeta.data_ptr(),
                                                       kidx.data_ptr(), klen.data_ptr(), status.data_ptr(),
                                                       torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(code, 'moka_build_token_meta_from_segments')
    return _finish(TokenMeta(tokmeta, kidx, klen, _cabi.VARIANT_AVT, cap, status=status), check)


def all_text_meta(B: int, S: int, device) -> TokenMeta:
    """Every token handled by the text adapter, no keys: the plain-LoRA decode branches
    (lora.py:373-381, layer.py:672-678)."""
    tokmeta = torch.full((B, S), _cabi.MOD_TEXT, dtype=torch.uint8, device=device)
    kidx = torch.full((B, 1), -1, dtype=torch.int32, device=device)
    klen = torch.zeros(B, dtype=torch.int32, device=device)
    return TokenMeta(tokmeta, kidx, klen, _cabi.VARIANT_VT, 1, 0)


class _MetaCache:
    """The reference hands the SAME mask objects to all 7 projections of all layers
    (AudioVisualText/models/modeling_llama.py:222-224,326-328,384).  Key the metadata on the
    identity + version of those tensors so it is built once per step, not 224 times."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self.items: List[Tuple[tuple, TokenMeta, tuple]] = []

    @staticmethod
    def _key(variant: int, masks: Sequence[Optional[torch.Tensor]]) -> tuple:
        return (variant,) + tuple((0, 0, 0) if m is None else (id(m), m.data_ptr(), m._version) for m in masks)

    def get(self, variant: int, masks: Sequence[Optional[torch.Tensor]], builder) -> TokenMeta:
        key = self._key(variant, masks)
        for k, meta, _keep in self.items:
            if k == key:
                return meta
        meta = builder()
        # keep the mask tensors alive while cached so ids cannot be recycled
        self.items.append((key, meta, tuple(masks)))
        if len(self.items) > self.capacity:
            self.items.pop(0)
        return meta

    def clear(self) -> None:
        self.items.clear()


META_CACHE = _MetaCache()
