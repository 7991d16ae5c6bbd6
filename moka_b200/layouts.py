"""Synthetic token-span layouts of SURVEY.md 8(d) (C1-C4), in the shapes the reference's mask producers emit
(AVT ``AudioVisualText/models/unified_arch.py:159-163``: four ``[B,S,1]`` int32 tensors, left-padded;
VT ``VisualText/train/train.py:210-231``: three ``[B,S]`` bool tensors).  Used by bench.py and tools/ to build
workloads; the oracle keeps its own copy for the tests (``tests/test_host_logic.py`` checks they agree)."""
from __future__ import annotations

import torch

VARIANT_AVT, VARIANT_VT = 0, 1


def layout_masks(variant: int, B: int, S: int, layout: str):
    """AVT -> [text, video, audio, question] each [B,S,1] int32; VT -> (text, image, question) each [B,S] bool."""
    text, vis, aud, qst = (torch.zeros(B, S, dtype=torch.bool) for _ in range(4))
    at = lambda v: (v * S) // 2048                      # the S = 2048 layouts, scaled to other lengths
    if layout == 'c1':                                  # VT, S = 128; sample 1 has its own question span
        if S < 128:
            raise ValueError('layout c1 needs S >= 128')
        text[:, :10] = True
        vis[:, 10:42] = True
        text[:, 42:128] = True
        qst[:, 50:70] = True
        if B > 1:
            qst[1] = False
            qst[1, 60:90] = True
    elif layout == 'c2':                                # VT: text, 576 image tokens, 64 question tokens, answer
        text[:, :at(35)] = True
        vis[:, at(35):at(611)] = True
        text[:, at(611):] = True
        qst[:, at(611):at(675)] = True
    elif layout == 'c3':                                # AVT: pad, text, video, text, audio, text, question, answer
        text[:, at(16):at(32)] = True
        vis[:, at(32):at(1056)] = True
        text[:, at(1056):at(1064)] = True
        aud[:, at(1064):at(1576)] = True
        text[:, at(1576):] = True
        qst[:, at(1584):at(1680)] = True
    elif layout == 'c4':                                # VT, S = 4096: four images
        f = S / 4096.0
        a, b, c = int(35 * f), int(2339 * f), int(2467 * f)
        text[:, :a] = True
        vis[:, a:b] = True
        text[:, b:] = True
        qst[:, b:c] = True
    else:
        raise ValueError(f'unknown layout {layout!r}')
    if variant == VARIANT_AVT:
        return [m.to(torch.int32).unsqueeze(-1) for m in (text, vis, aud, qst)]
    if bool(aud.any()):
        raise ValueError('the VT variant has no audio span')
    return text, vis, qst


# ---- the same layouts as the data pipeline holds them BEFORE any mask tensor exists
SEG_END, SEG_TEXT, SEG_VIDEO, SEG_AUDIO, SEG_QUESTION = 0, 1, 2, 3, 4


def layout_segments(B: int, S: int, layout: str = 'c3', max_segments: int = 16) -> torch.Tensor:
    """AVT: the per-sample segment table ``prepare_multimodal_inputs`` walks
    (AudioVisualText/models/unified_arch.py:139-240) as int32 [B, max_segments, 2] = (kind, length); the left
    padding of :304-312 is implied (S - sum of lengths).  Equivalent to ``layout_masks(VARIANT_AVT, ...)``."""
    if layout != 'c3':
        raise ValueError('segment tables are defined for the AVT layout c3')
    at = lambda v: (v * S) // 2048
    cuts = [(SEG_TEXT, at(16), at(32)), (SEG_VIDEO, at(32), at(1056)), (SEG_TEXT, at(1056), at(1064)),
            (SEG_AUDIO, at(1064), at(1576)), (SEG_TEXT, at(1576), at(1584)), (SEG_QUESTION, at(1584), at(1680)),
            (SEG_TEXT, at(1680), S)]
    seg = torch.zeros(B, max_segments, 2, dtype=torch.int32)
    for i, (kind, lo, hi) in enumerate(cuts):
        seg[:, i, 0] = kind
        seg[:, i, 1] = hi - lo
    return seg


def layout_ids_labels(B: int, S: int, layout: str, image_token_id: int, vocab: int = 32000, seed: int = 0):
    """VT: collated ``input_ids`` / ``labels`` [B,S] int64 whose masks (VisualText/train/train.py:206-231) are the
    layout's: image positions carry ``image_token_id``; labels are -100 up to the last image token and on the
    question span, and equal the ids elsewhere.  Equivalent to ``layout_masks(VARIANT_VT, ...)``."""
    text, vis, qst = layout_masks(VARIANT_VT, B, S, layout)
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (B, S), generator=g)
    ids[ids == image_token_id] = image_token_id - 1
    ids[vis] = image_token_id
    labels = ids.clone()
    for b in range(B):
        im = torch.where(vis[b])[0]
        if len(im):
            labels[b, :int(im[-1]) + 1] = -100
        labels[b, qst[b]] = -100
    return ids, labels
