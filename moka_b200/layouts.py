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
