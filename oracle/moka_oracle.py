"""CPU oracle for the MokA adapter layer (MokaLinear) hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``moka_b200``) never routes through this module
and fails loudly when ``libmoka.so`` is missing.

Parity status: PINNED.  Every function below is checked against the reference's own
modules, imported unmodified from ``/root/reference`` in the build container by
``oracle/gen_golden.py``; the resulting input/output vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py`` (the reference itself
ships no tests or golden vectors -- SURVEY.md section 4).

It restates, in plain torch CPU ops, the algorithm of the two reference layers:

* AVT (3 modalities): ``AudioVisualText/peft_hyper/tuners/lora.py:367-369, 460-532``
  (train branch; the 'test' prefill branch :385-457 is the same math; decode :373-381)
* VT  (2 modalities): ``VisualText/modified_peft/tuners/lora/layer.py:580-678``

and then the *kernel form* of the same computation -- per-token modality codes, a key
index list per sample and an explicit hand-derived backward --
which is what the CUDA kernels implement.  ``*_reference_form`` functions follow the
reference op by op (so they can also serve as the timed CPU "port" baseline);
``fused_forward`` / ``fused_backward`` are the fp64 truth the GPU parity tests compare to.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

# ----------------------------------------------------------------------------------------
# token metadata (the kernel-side encoding of the reference's mask tensors)
# ----------------------------------------------------------------------------------------
MOD_NONE = 0          # token in no modality mask (pad): y = y0
MOD_TEXT = 1          # adapter 0  (AVT lora_A0 / VT lora_A['text'])
MOD_VIS = 2           # adapter 1  (AVT lora_A1 video / VT lora_A['image'])  -- query modality
MOD_AUD = 3           # adapter 2  (AVT lora_A2 audio)                       -- query modality
MOD_MASK = 0x7
KEY_VALID = 0x8       # this token's (pre-attention) projection is a usable key/value row

VARIANT_AVT = 0
VARIANT_VT = 1


@dataclass
class TokenMeta:
    """Per-step token metadata shared by every MokaLinear call of the step.

    tokmeta : uint8 [B, S]   low 3 bits modality code, bit 3 KEY_VALID
    kidx    : int32 [B, Lk]  in-sample positions of the key rows (padded with -1)
    klen    : int32 [B]      number of key rows of each sample
    """
    tokmeta: torch.Tensor
    kidx: torch.Tensor
    klen: torch.Tensor

    @property
    def B(self) -> int:
        return self.tokmeta.shape[0]

    @property
    def S(self) -> int:
        return self.tokmeta.shape[1]


def _check_disjoint(masks: Sequence[torch.Tensor]) -> None:
    tot = sum(m.to(torch.int32) for m in masks)
    if int(tot.max()) > 1:
        raise ValueError("modality masks overlap: a token may belong to at most one modality")


def build_token_meta_avt(modality_mask: Sequence[torch.Tensor]) -> TokenMeta:
    """AVT masks -> metadata.  ``modality_mask = [text, video, audio, question]``, each
    ``[B,S,1]`` int32 in {0,1} (AudioVisualText/models/unified_arch.py:159-204, :336).

    Key rows of sample b are the CONTIGUOUS span first..last question position
    (lora.py:489-491); rows inside the span with question==0 (or not text) are zero rows
    that still take part in the softmax.  An empty question mask raises IndexError exactly
    like ``indices[0]`` does in the reference (lora.py:490).
    """
    text, video, audio, question = [m.reshape(m.shape[0], m.shape[1]).to(torch.int32) for m in modality_mask]
    _check_disjoint([text, video, audio])
    B, S = text.shape
    code = text * MOD_TEXT + video * MOD_VIS + audio * MOD_AUD
    keyvalid = (question == 1) & (text == 1)
    tokmeta = (code + keyvalid.to(torch.int32) * KEY_VALID).to(torch.uint8)
    spans = []
    for b in range(B):
        idx = torch.where(question[b] == 1)[0]
        if idx.numel() == 0:
            raise IndexError(f"sample {b}: empty question mask (index 0 is out of bounds)")
        spans.append((int(idx[0]), int(idx[-1]) + 1))
    Lk = max(hi - lo for lo, hi in spans)
    kidx = torch.full((B, Lk), -1, dtype=torch.int32)
    klen = torch.zeros(B, dtype=torch.int32)
    for b, (lo, hi) in enumerate(spans):
        kidx[b, : hi - lo] = torch.arange(lo, hi, dtype=torch.int32)
        klen[b] = hi - lo
    return TokenMeta(tokmeta, kidx, klen)


def build_token_meta_vt(text_mask: torch.Tensor, image_mask: torch.Tensor,
                        question_mask: torch.Tensor) -> TokenMeta:
    """VT masks -> metadata.  Masks are ``[B,S]`` bool (VisualText/train/train.py:206-231).

    Key rows of sample b are the EXACT gather of its question positions (layer.py:635,
    :643-644).  A sample with no question token or no image token skips its attention
    (layer.py:631-637): here an empty key list / no query token has the same effect.
    """
    text = text_mask.reshape(text_mask.shape[0], -1).to(torch.int32)
    image = image_mask.reshape(image_mask.shape[0], -1).to(torch.int32)
    question = question_mask.reshape(question_mask.shape[0], -1).to(torch.int32)
    _check_disjoint([text, image])
    B, S = text.shape
    code = text * MOD_TEXT + image * MOD_VIS
    keyvalid = (question == 1) & (code != MOD_NONE)
    tokmeta = (code + keyvalid.to(torch.int32) * KEY_VALID).to(torch.uint8)
    lists = [torch.where(question[b] == 1)[0].to(torch.int32) for b in range(B)]
    Lk = max(1, max(int(l.numel()) for l in lists))
    kidx = torch.full((B, Lk), -1, dtype=torch.int32)
    klen = torch.zeros(B, dtype=torch.int32)
    for b, l in enumerate(lists):
        kidx[b, : l.numel()] = l
        klen[b] = l.numel()
    return TokenMeta(tokmeta, kidx, klen)


# ----------------------------------------------------------------------------------------
# mask producers: what the reference's data pipeline does BEFORE the layer
# ----------------------------------------------------------------------------------------
def vt_masks_from_ids(input_ids: torch.Tensor, labels: torch.Tensor, image_token_id: int,
                      lengths: Optional[Sequence[int]] = None, ignore_index: int = -100):
    """VisualText/train/train.py:206-231 per sample, then the collator's right-padding with False
    (:259-297).  ``input_ids`` / ``labels`` are the PADDED [B,S] tensors, ``lengths`` the unpadded lengths.
    Returns (my_text_mask, my_image_mask, question_mask), each [B,S] bool."""
    B, S = input_ids.shape
    text = torch.zeros(B, S, dtype=torch.bool)
    image = torch.zeros(B, S, dtype=torch.bool)
    question = torch.zeros(B, S, dtype=torch.bool)
    for b in range(B):
        n = S if lengths is None else int(lengths[b])
        ids = input_ids[b:b + 1, :n]
        tgt = labels[b:b + 1, :n]
        my_image_mask = ids == image_token_id                                  # :210
        my_text_mask = ids != image_token_id                                   # :211
        image_positions = torch.where(my_image_mask)[1]                        # :220
        if len(image_positions) > 0:
            last = image_positions[-1]
            position_mask = (torch.arange(n) > last).unsqueeze(0)              # :224-225
        else:
            position_mask = torch.zeros_like(my_image_mask, dtype=torch.bool)  # :227
        q = ~my_image_mask & (tgt == ignore_index) & position_mask             # :231
        text[b, :n], image[b, :n], question[b, :n] = my_text_mask[0], my_image_mask[0], q[0]
    return text, image, question


SEG_END, SEG_TEXT, SEG_VIDEO, SEG_AUDIO, SEG_QUESTION = 0, 1, 2, 3, 4


def avt_masks_from_segments(segments: Sequence[Sequence[Tuple[int, int]]], S: int):
    """AudioVisualText/models/unified_arch.py:139-240 (per-segment mask pieces, concatenated :262-275) and the
    LEFT padding of :304-312.  ``segments[b]`` = [(kind, length), ...].  Returns [text, video, audio, question],
    each [B,S,1] int32."""
    B = len(segments)
    out = [torch.zeros(B, S, 1, dtype=torch.int32) for _ in range(4)]
    for b, segs in enumerate(segments):
        pieces = [[], [], [], []]
        for kind, n in segs:
            if kind == SEG_END:
                break
            one, zero = torch.ones(n, 1, dtype=torch.int32), torch.zeros(n, 1, dtype=torch.int32)
            pieces[0].append(one if kind in (SEG_TEXT, SEG_QUESTION) else zero)   # :159, :171, :189 ...
            pieces[1].append(one if kind == SEG_VIDEO else zero)
            pieces[2].append(one if kind == SEG_AUDIO else zero)
            pieces[3].append(one if kind == SEG_QUESTION else zero)               # :163 vs :175
        cat = [torch.cat(p, 0) if p else torch.zeros(0, 1, dtype=torch.int32) for p in pieces]
        n = cat[0].shape[0]
        for k in range(4):
            out[k][b, S - n:] = cat[k]                                            # left padding (:304-312)
    return out


# ----------------------------------------------------------------------------------------
# reference-form restatements (op by op; dtype follows the inputs like the reference does)
# ----------------------------------------------------------------------------------------
def _drop(x: torch.Tensor, keep: Optional[torch.Tensor], p: float) -> torch.Tensor:
    """nn.Dropout with an explicit keep mask (1 = kept), scaled by 1/(1-p)."""
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype) * (1.0 / (1.0 - p))


def avt_reference_form(x: torch.Tensor, weight: torch.Tensor, A: Sequence[torch.Tensor],
                       B0: torch.Tensor, modality_mask: Sequence[torch.Tensor], *,
                       lora_alpha: float, blc_weight: float, bias: Optional[torch.Tensor] = None,
                       keep: Optional[torch.Tensor] = None, p: float = 0.0) -> torch.Tensor:
    """AVT train branch, lora.py:367-369 and :460-532 (the 'test' prefill branch :385-457
    is the same computation).  ``A[i]`` is ``lora_A{i}.weight [r, d_in]``, ``B0`` is
    ``lora_B0.weight [d_out, r]``.  ``keep`` is one dropout keep mask ``[B,S,d_in]`` shared
    by the three modality calls (each token belongs to one modality, so only that call's
    mask matters for it)."""
    r = A[0].shape[0]
    s = lora_alpha / r                                   # lora.py:329  scaling[0]
    d_k = r                                              # lora.py:314
    result = torch.nn.functional.linear(x, weight, bias)  # lora.py:369
    text, video, audio, question = modality_mask
    xs = [x * text.to(x.dtype), x * video.to(x.dtype), x * audio.to(x.dtype)]          # :468
    out_a = [torch.nn.functional.linear(_drop(xs[i], keep, p), A[i]) * s for i in range(len(A))]  # :477
    q_tok = out_a[0] * question.to(x.dtype)              # :482
    new = []
    for j, m in ((1, video), (2, audio)):
        if j >= len(A):
            break
        tok = out_a[j]
        upd = torch.zeros_like(tok)                      # :483, :504
        for b in range(x.shape[0]):                      # :485-499, :506-521
            idx = torch.where(question[b, :, 0] == 1)[0]
            lo, hi = int(idx[0]), int(idx[-1]) + 1       # IndexError when empty, like :490
            kv = q_tok[b, lo:hi, :]
            score = torch.matmul(tok[b], kv.transpose(-2, -1)) / math.sqrt(d_k)
            score = torch.softmax(score, dim=-1)
            att = torch.matmul(score, kv)
            upd[b] = tok[b] + (m[b].to(x.dtype) * att) * blc_weight
        new.append(upd)
    g = out_a[0]
    for n in new:                                        # :524-525
        g = g + n
    return torch.nn.functional.linear(g, B0) + result    # :528-530


def vt_reference_form(x: torch.Tensor, weight: torch.Tensor, A_text: torch.Tensor,
                      A_image: torch.Tensor, B_text: torch.Tensor,
                      text_mask: torch.Tensor, image_mask: torch.Tensor,
                      question_mask: torch.Tensor, *, lora_alpha: float, attn_weight: float,
                      bias: Optional[torch.Tensor] = None, keep: Optional[torch.Tensor] = None,
                      p: float = 0.0) -> torch.Tensor:
    """VT MokA branch, layer.py:580-671.  ``lora_B['image']`` exists in the reference but
    is never read (layer.py:657), so it is not an argument here."""
    r = A_text.shape[0]
    s_text = s_image = lora_alpha / r                    # layer.py:114
    Bsz, S, d_in = x.shape
    result = torch.nn.functional.linear(x, weight, bias)  # :580
    xf = x.reshape(-1, d_in)
    kf = None if keep is None else keep.reshape(-1, d_in)
    tf = text_mask.reshape(-1)
    imf = image_mask.reshape(-1)
    rf = result.reshape(-1, result.shape[-1]).clone()
    h = torch.zeros(xf.shape[0], r, dtype=x.dtype)       # :600
    if bool((tf == 1).any()):                            # :603-611
        it = (tf == 1).nonzero(as_tuple=True)[0]
        h[it] = torch.nn.functional.linear(_drop(xf[it], None if kf is None else kf[it], p), A_text)
    if bool((imf == 1).any()):                           # :613-621
        ii = (imf == 1).nonzero(as_tuple=True)[0]
        h[ii] = torch.nn.functional.linear(_drop(xf[ii], None if kf is None else kf[ii], p), A_image)
    h = h.reshape(Bsz, S, r)
    if bool((imf == 1).any()):                           # :627-653
        h = h.clone()
        for b in range(Bsz):
            qi = torch.where(image_mask[b] == 1)[0]
            if len(qi) == 0:
                continue
            ki = torch.where(question_mask[b] == 1)[0]
            if len(ki) == 0:
                continue
            q = h[b, qi, :]
            kv = h[b, ki, :]
            score = torch.softmax(torch.matmul(q, kv.transpose(-2, -1)) / math.sqrt(r), dim=-1)
            h[b, qi, :] = q + attn_weight * torch.matmul(score, kv)
    hf = h.reshape(-1, r)
    if bool((tf == 1).any()):                            # :660-663
        it = (tf == 1).nonzero(as_tuple=True)[0]
        rf[it] = rf[it] + torch.nn.functional.linear(hf[it], B_text) * s_text
    if bool((imf == 1).any()):                           # :666-669
        ii = (imf == 1).nonzero(as_tuple=True)[0]
        rf[ii] = rf[ii] + torch.nn.functional.linear(hf[ii], B_text) * s_image
    return rf.reshape(Bsz, S, -1)


def plain_lora_reference_form(x: torch.Tensor, weight: torch.Tensor, A_text: torch.Tensor,
                              B_text: torch.Tensor, *, lora_alpha: float, variant: int,
                              bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decode-step branches: AVT lora.py:373-381 (scale between A and B), VT
    layer.py:672-678 (scale after B).  Text adapter only, no masks."""
    s = lora_alpha / A_text.shape[0]
    base = torch.nn.functional.linear(x, weight, bias)
    if variant == VARIANT_AVT:
        return torch.nn.functional.linear(torch.nn.functional.linear(x, A_text) * s, B_text) + base
    return base + torch.nn.functional.linear(torch.nn.functional.linear(x, A_text), B_text) * s


# ----------------------------------------------------------------------------------------
# kernel form: what the CUDA kernels compute (SURVEY.md sections A.1-A.3)
# ----------------------------------------------------------------------------------------
@dataclass
class FusedScalars:
    """pre  : scale applied to A_m x      (AVT alpha/r, VT 1)          lora.py:477
    post : per-adapter scale after B      (AVT 1,       VT alpha/r)  layer.py:662,668
    w    : attention weight               (AVT blc_weight lora.py:499, VT attn_weight layer.py:653)
    """
    pre: float
    post: Tuple[float, float, float]
    w: float

    @staticmethod
    def for_variant(variant: int, *, lora_alpha: float, r: int, w: float) -> "FusedScalars":
        s = lora_alpha / r
        if variant == VARIANT_AVT:
            return FusedScalars(pre=s, post=(1.0, 1.0, 1.0), w=w)
        return FusedScalars(pre=1.0, post=(s, s, s), w=w)


def _keys_of(meta: TokenMeta, b: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n = int(meta.klen[b])
    pos = meta.kidx[b, :n].long()
    valid = (meta.tokmeta[b, pos].to(torch.int32) & KEY_VALID) != 0
    return pos, valid


def fused_forward(meta: TokenMeta, x: torch.Tensor, y0: torch.Tensor, A: Sequence[torch.Tensor],
                  Bw: torch.Tensor, sc: FusedScalars, keep: Optional[torch.Tensor] = None,
                  p: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """y = y0 + post_m * B (h + w * attn(h, K_b)) in the dtype of ``x`` (use fp64 for truth).

    Returns (y, h) where ``h [B,S,r]`` is the pre-attention projection (the stash the
    backward re-uses).  Equivalent to avt_reference_form / vt_reference_form under the
    metadata built by build_token_meta_*; verified in tests/test_oracle_golden.py (both forms against
    the reference-generated fixtures).  Device-agnostic: tensors and ``meta`` may live on a GPU (the
    bench-geometry parity tests run this fp64 truth there)."""
    Bsz, S, d_in = x.shape
    r = A[0].shape[0]
    code = (meta.tokmeta.to(torch.int32) & MOD_MASK)
    u = _drop(x, keep, p)
    h = torch.zeros(Bsz, S, r, dtype=x.dtype, device=x.device)
    for m in range(len(A)):
        sel = code == (m + 1)
        if bool(sel.any()):
            h[sel] = torch.nn.functional.linear(u[sel], A[m]) * sc.pre
    g = h.clone()
    for b in range(Bsz):
        pos, valid = _keys_of(meta, b)
        if pos.numel() == 0:
            continue
        K = h[b, pos, :] * valid.to(x.dtype)[:, None]
        qsel = code[b] >= MOD_VIS
        if not bool(qsel.any()):
            continue
        q = h[b, qsel, :]
        P = torch.softmax(q @ K.t() / math.sqrt(r), dim=-1)
        g[b, qsel, :] = q + sc.w * (P @ K)
    post = torch.zeros(Bsz, S, dtype=x.dtype, device=x.device)
    for m in range(len(A)):
        post[code == (m + 1)] = sc.post[m]
    y = y0 + (g @ Bw.t()) * post[..., None]
    return y, h


def fused_backward(meta: TokenMeta, x: torch.Tensor, dy: torch.Tensor, h: torch.Tensor,
                   A: Sequence[torch.Tensor], Bw: torch.Tensor, sc: FusedScalars,
                   keep: Optional[torch.Tensor] = None, p: float = 0.0):
    """Hand-derived backward of ``fused_forward`` w.r.t. x (adapter part only: the frozen
    base contributes W0^T dy separately), A_m and B.  SURVEY.md section A.3.

    Returns (dx_adapter [B,S,d_in], [dA_m [r,d_in]], dB [d_out,r], dh [B,S,r])."""
    Bsz, S, d_in = x.shape
    r = A[0].shape[0]
    rs = 1.0 / math.sqrt(r)
    code = (meta.tokmeta.to(torch.int32) & MOD_MASK)
    u = _drop(x, keep, p)
    post = torch.zeros(Bsz, S, dtype=x.dtype, device=x.device)
    for m in range(len(A)):
        post[code == (m + 1)] = sc.post[m]
    # recompute g (cheap) exactly like the forward
    g = h.clone()
    dg = (dy @ Bw) * post[..., None]                       # [B,S,r]
    dh = dg.clone()
    for b in range(Bsz):
        pos, valid = _keys_of(meta, b)
        if pos.numel() == 0:
            continue
        vm = valid.to(x.dtype)[:, None]
        K = h[b, pos, :] * vm
        qsel = code[b] >= MOD_VIS
        if not bool(qsel.any()):
            continue
        q = h[b, qsel, :]
        P = torch.softmax(q @ K.t() * rs, dim=-1)
        g[b, qsel, :] = q + sc.w * (P @ K)
        da = sc.w * dg[b, qsel, :]                          # d(attn out)
        dP = da @ K.t()
        dS = P * (dP - (P * dP).sum(-1, keepdim=True))
        dh[b, qsel, :] = dg[b, qsel, :] + (dS @ K) * rs
        dK = (P.t() @ da + (dS.t() @ q) * rs) * vm          # zero-filled / invalid rows dropped
        # key tokens with no modality have a constant zero projection: no gradient path
        live = (code[b, pos] != MOD_NONE).to(x.dtype)[:, None]
        dh[b].index_add_(0, pos, dK * live)
    dB = torch.einsum('bso,bsr->or', dy * post[..., None], g)
    dA, dxa = [], torch.zeros_like(x)
    for m in range(len(A)):
        sel = code == (m + 1)
        dhm = dh[sel] * sc.pre
        dA.append(dhm.t() @ u[sel])
        dxa[sel] = dhm @ A[m]
    if keep is not None and p != 0.0:
        dxa = dxa * keep.to(x.dtype) * (1.0 / (1.0 - p))
    return dxa, dA, dB, dh


# ----------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests, smoke and bench
# ----------------------------------------------------------------------------------------
def layout_masks(variant: int, B: int, S: int, layout: str):
    """Span layouts of SURVEY.md 8(d).  Returns the reference-shaped mask tensors:
    AVT -> [text, video, audio, question] each [B,S,1] int32; VT -> (text, image, question)
    each [B,S] bool."""
    text = torch.zeros(B, S, dtype=torch.bool)
    vis = torch.zeros(B, S, dtype=torch.bool)
    aud = torch.zeros(B, S, dtype=torch.bool)
    qst = torch.zeros(B, S, dtype=torch.bool)

    def sc(v: int) -> int:          # scale the S=2048 layouts to other S
        return (v * S) // 2048

    if layout == 'c1':              # VT, S=128
        assert S >= 128
        text[:, 0:10] = True; vis[:, 10:42] = True; text[:, 42:128] = True
        qst[:, 50:70] = True
        if B > 1:
            qst[1] = False; qst[1, 60:90] = True
    elif layout == 'c2':            # VT S=2048: text, 576 image, 64 question, answer
        text[:, 0:sc(35)] = True; vis[:, sc(35):sc(611)] = True; text[:, sc(611):] = True
        qst[:, sc(611):sc(675)] = True
    elif layout == 'c3':            # AVT S=2048: left pad, text, video, text, audio, text, question, answer
        text[:, sc(16):sc(32)] = True; vis[:, sc(32):sc(1056)] = True; text[:, sc(1056):sc(1064)] = True
        aud[:, sc(1064):sc(1576)] = True; text[:, sc(1576):] = True
        qst[:, sc(1584):sc(1680)] = True
    elif layout == 'c4':            # VT S=4096
        f = S / 4096.0
        a, b_, c = int(35 * f), int(2339 * f), int(2467 * f)
        text[:, 0:a] = True; vis[:, a:b_] = True; text[:, b_:] = True; qst[:, b_:c] = True
    else:
        raise ValueError(layout)
    if variant == VARIANT_AVT:
        return [m.to(torch.int32).unsqueeze(-1) for m in (text, vis, aud, qst)]
    assert not bool(aud.any())
    return text, vis, qst


def kaiming_uniform_a5(rows: int, cols: int, gen: torch.Generator) -> torch.Tensor:
    """nn.init.kaiming_uniform_(a=sqrt(5)) for a [rows, cols] Linear weight: U(-1/sqrt(cols), 1/sqrt(cols))."""
    bound = 1.0 / math.sqrt(cols)
    return (torch.rand(rows, cols, generator=gen, dtype=torch.float64) * 2 - 1) * bound


def synth_layer_inputs(variant: int, B: int, S: int, d_in: int, d_out: int, r: int, M: int,
                       layout: str, seed: int = 1234, store_dtype: torch.dtype = torch.bfloat16):
    """Seeded synthetic (x, W0, A_m, B, dy) with values representable in ``store_dtype``,
    returned as fp64 tensors (the truth is computed on exactly those values)."""
    gen = torch.Generator().manual_seed(seed)

    def q(t):
        return t.to(store_dtype).to(torch.float64)

    x = q(torch.randn(B, S, d_in, generator=gen, dtype=torch.float64))
    W0 = q(torch.randn(d_out, d_in, generator=gen, dtype=torch.float64) * 0.02)
    A = [q(kaiming_uniform_a5(r, d_in, gen)) for _ in range(M)]
    Bw = q(torch.randn(d_out, r, generator=gen, dtype=torch.float64) * 0.02)
    dy = q(torch.randn(B, S, d_out, generator=gen, dtype=torch.float64))
    masks = layout_masks(variant, B, S, layout)
    return x, W0, A, Bw, dy, masks


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """rel(a,b) = ||a-b||_2 / ||b||_2 (SURVEY.md 8d parity metric)."""
    a = a.to(torch.float64); b = b.to(torch.float64)
    den = float(b.norm())
    return float((a - b).norm()) / den if den > 0 else float((a - b).norm())
