"""Generate tests/golden/*.pt from the UNMODIFIED reference layers (build container only).

TEST INFRASTRUCTURE.  Run from the repo root, in the container that has /root/reference:

    python oracle/gen_golden.py

It imports
  * AudioVisualText/peft_hyper/tuners/lora.py:277  ``Linear``  (AVT, 3 modalities)
  * VisualText/modified_peft/tuners/lora/layer.py:389 ``Linear`` (VT, 2 modalities)
with only ``accelerate`` stubbed (it is imported by the package __init__ chains, never used
by the layers), runs them in fp64 / fp32 / bf16 on seeded inputs, and stores inputs,
outputs and autograd gradients.  /root/reference does not exist on the GPU box, so the
vectors are committed; nothing in tests/, smoke() or bench.py reads the reference tree.
"""
from __future__ import annotations

import json
import math
import os
import sys
import tempfile
from unittest import mock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import moka_oracle as mo  # noqa: E402

REF = '/root/reference'
OUT = os.environ.get('MOKA_GOLDEN_OUT', os.path.join(ROOT, 'tests', 'golden'))


from oracle import ref_loader  # noqa: E402


def _stub_accelerate():
    ref_loader.stub_accelerate()


def import_avt_linear():
    return ref_loader.import_avt_linear(prefer_tree=True)      # the reference tree itself, not the vendored copy


def import_vt_linear():
    return ref_loader.import_vt_linear(prefer_tree=True)


def make_avt(Linear, d_in, d_out, r, M, alpha, w, W0, A, Bw, dtype):
    return ref_loader.make_avt(Linear, d_in, d_out, r, M, alpha, w, W0, A, Bw, dtype)


def make_vt(Linear, d_in, d_out, r, alpha, w, W0, A, Bw, dtype):
    return ref_loader.make_vt(Linear, d_in, d_out, r, alpha, w, W0, A, Bw, dtype)


def run_avt(l, x, masks, dy, M):
    x = x.clone().requires_grad_(True)
    y = l(x, masks)
    y.backward(dy)
    return (y.detach(), x.grad.detach(),
            [getattr(l, f'lora_A{i}').weight.grad.detach().clone() for i in range(M)],
            l.lora_B0.weight.grad.detach().clone())


def run_vt(l, x, masks, dy):
    x = x.clone().requires_grad_(True)
    y = l(x, masks[0], masks[1], masks[2])
    y.backward(dy)
    gimg = l.lora_B['image'].weight.grad
    assert gimg is None, 'lora_B.image unexpectedly received a gradient'
    return (y.detach(), x.grad.detach(),
            [l.lora_A['text'].weight.grad.detach().clone(), l.lora_A['image'].weight.grad.detach().clone()],
            l.lora_B['text'].weight.grad.detach().clone())


def edit_masks(variant, masks, edit):
    """Edge cases of SURVEY.md A.4."""
    if variant == mo.VARIANT_AVT:
        text, vis, aud, qst = [m.clone() for m in masks]
        if edit == 'noncontig_question':
            qst[:, qst.shape[1] * 1600 // 2048: qst.shape[1] * 1640 // 2048] = 0
        elif edit == 'no_video_sample':
            text[0] = text[0] + vis[0]; vis[0] = 0
        elif edit == 'pad_holes':
            for m in (text, vis, aud):
                m[:, 5::17] = 0
            qst = qst * text
            qst[:, qst.shape[1] * 1584 // 2048] = 1      # keep a question token that is pad (not text)
        return [text, vis, aud, qst]
    text, vis, qst = [m.clone() for m in masks]
    if edit == 'noncontig_question':
        qst[:, 55:60] = False
    elif edit == 'no_question_sample':
        qst[0] = False
    elif edit == 'no_image_sample':
        text[0] = text[0] | vis[0]; vis[0] = False
    elif edit == 'pad_tail':
        text[:, -9:] = False; vis[:, -9:] = False; qst[:, -9:] = False
        qst[1, 100] = True; text[1, 100] = False          # a question token in no modality
    return text, vis, qst


CASES = [
    # name, variant, B, S, d_in, d_out, r, M, layout, alpha, w, edit
    ('vt_c1_small', mo.VARIANT_VT, 2, 128, 128, 96, 8, 2, 'c1', 16, 0.05, None),
    ('vt_r4', mo.VARIANT_VT, 2, 128, 64, 160, 4, 2, 'c1', 16, 0.5, None),
    ('vt_r16_noncontig', mo.VARIANT_VT, 2, 128, 96, 64, 16, 2, 'c1', 16, 0.05, 'noncontig_question'),
    ('vt_no_question_sample', mo.VARIANT_VT, 2, 128, 64, 64, 8, 2, 'c1', 16, 0.05, 'no_question_sample'),
    ('vt_no_image_sample', mo.VARIANT_VT, 2, 128, 64, 64, 8, 2, 'c1', 16, 0.05, 'no_image_sample'),
    ('vt_pad_tail', mo.VARIANT_VT, 2, 128, 64, 64, 8, 2, 'c1', 16, 0.05, 'pad_tail'),
    ('avt_r4', mo.VARIANT_AVT, 2, 128, 128, 96, 4, 3, 'c3', 16, 1.0, None),
    ('avt_r16', mo.VARIANT_AVT, 2, 128, 64, 160, 16, 3, 'c3', 16, 1.0, None),
    ('avt_noncontig', mo.VARIANT_AVT, 2, 128, 64, 64, 4, 3, 'c3', 16, 1.0, 'noncontig_question'),
    ('avt_no_video_sample', mo.VARIANT_AVT, 2, 128, 64, 64, 8, 3, 'c3', 16, 0.7, 'no_video_sample'),
    ('avt_pad_holes', mo.VARIANT_AVT, 2, 128, 64, 64, 4, 3, 'c3', 16, 1.0, 'pad_holes'),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    _stub_accelerate()
    AVT = import_avt_linear()
    VT = import_vt_linear()
    torch.manual_seed(0)
    index = {}
    for (name, variant, B, S, d_in, d_out, r, M, layout, alpha, w, edit) in CASES:
        x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(variant, B, S, d_in, d_out, r, M, layout, seed=1234)
        if edit:
            masks = edit_masks(variant, masks, edit)
        rec = dict(name=name, variant=variant, B=B, S=S, d_in=d_in, d_out=d_out, r=r, M=M,
                   alpha=alpha, w=w, layout=layout, edit=edit,
                   x=x.to(torch.bfloat16), W0=W0.to(torch.bfloat16), A=[a.to(torch.bfloat16) for a in A],
                   Bw=Bw.to(torch.bfloat16), dy=dy.to(torch.bfloat16),
                   masks=[m.clone() for m in masks])
        for dt, tag in ((torch.float64, 'f64'), (torch.bfloat16, 'bf16')):
            if variant == mo.VARIANT_AVT:
                l = make_avt(AVT, d_in, d_out, r, M, alpha, w, W0, A, Bw, dt)
                y, dx, dA, dB = run_avt(l, x.to(dt), masks, dy.to(dt), M)
            else:
                l = make_vt(VT, d_in, d_out, r, alpha, w, W0, A, Bw, dt)
                y, dx, dA, dB = run_vt(l, x.to(dt), masks, dy.to(dt))
            rec[f'y_{tag}'] = y
            rec[f'dx_{tag}'] = dx
            rec[f'dA_{tag}'] = dA
            rec[f'dB_{tag}'] = dB
        torch.save(rec, os.path.join(OUT, f'{name}.pt'))
        index[name] = dict(variant=variant, B=B, S=S, d_in=d_in, d_out=d_out, r=r, M=M, edit=edit)
        print('wrote', name)

    # error conventions
    x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(mo.VARIANT_AVT, 2, 256, 64, 64, 4, 3, 'c3')
    masks[3][0] = 0
    l = make_avt(AVT, 64, 64, 4, 3, 16, 1.0, W0, A, Bw, torch.float64)
    try:
        l(x, masks)
        index['avt_empty_question_error'] = 'none'
    except Exception as e:  # noqa: BLE001
        index['avt_empty_question_error'] = type(e).__name__
    print('AVT empty question ->', index['avt_empty_question_error'])

    # decode branches (a5, a9): plain LoRA with the text adapter
    x, W0, A, Bw, dy, _ = mo.synth_layer_inputs(mo.VARIANT_VT, 3, 128, 64, 96, 8, 2, 'c1')
    x1 = x[:, :1, :].contiguous()
    lvt = make_vt(VT, 64, 96, 8, 16, 0.05, W0, A, Bw, torch.float64)
    y_vt = lvt(x1, None, None, None).detach()
    A3 = A + [A[0]]
    lavt = make_avt(AVT, 64, 96, 8, 3, 16, 1.0, W0, A3, Bw, torch.float64)
    lavt.loramethod = 'test'
    y_avt = lavt(x1, None).detach()
    torch.save(dict(x=x1.to(torch.bfloat16), W0=W0.to(torch.bfloat16), A_text=A[0].to(torch.bfloat16),
                    Bw=Bw.to(torch.bfloat16), alpha=16, y_vt=y_vt, y_avt=y_avt),
               os.path.join(OUT, 'decode_plain_lora.pt'))
    print('wrote decode_plain_lora')

    # the reference's own bf16 noise floor at C1 full size (numbers only; inputs are seeded)
    B, S, d, r = 2, 128, 4096, 8
    x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(mo.VARIANT_VT, B, S, d, d, r, 2, 'c1', seed=1234)
    l64 = make_vt(VT, d, d, r, 16, 0.05, W0, A, Bw, torch.float64)
    t = run_vt(l64, x, masks, dy)
    l16 = make_vt(VT, d, d, r, 16, 0.05, W0, A, Bw, torch.bfloat16)
    o = run_vt(l16, x.to(torch.bfloat16), masks, dy.to(torch.bfloat16))
    floor = dict(y=mo.rel_l2(o[0], t[0]), dx=mo.rel_l2(o[1], t[1]), dA_text=mo.rel_l2(o[2][0], t[2][0]),
                 dA_image=mo.rel_l2(o[2][1], t[2][1]), dB=mo.rel_l2(o[3], t[3]),
                 y_checksum=float(t[0].sum()), dx_checksum=float(t[1].sum()),
                 dB_checksum=float(t[3].sum()))
    # adapter-only figures (base output removed) are what the fused kernel is responsible for
    y0 = torch.nn.functional.linear(x, W0)
    dx0 = dy @ W0
    floor['delta_y'] = mo.rel_l2(o[0].double() - y0, t[0] - y0)
    floor['delta_dx'] = mo.rel_l2(o[1].double() - dx0, t[1] - dx0)
    index['c1_reference_bf16_floor'] = floor
    print('C1 bf16 floor', floor)

    x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(mo.VARIANT_AVT, 2, 256, d, d, 4, 3, 'c3', seed=1234)
    l64 = make_avt(AVT, d, d, 4, 3, 16, 1.0, W0, A, Bw, torch.float64)
    t = run_avt(l64, x, masks, dy, 3)
    l16 = make_avt(AVT, d, d, 4, 3, 16, 1.0, W0, A, Bw, torch.bfloat16)
    o = run_avt(l16, x.to(torch.bfloat16), masks, dy.to(torch.bfloat16), 3)
    index['avt_d4096_reference_bf16_floor'] = dict(
        y=mo.rel_l2(o[0], t[0]), dx=mo.rel_l2(o[1], t[1]), dB=mo.rel_l2(o[3], t[3]),
        dA=[mo.rel_l2(o[2][i], t[2][i]) for i in range(3)],
        y_checksum=float(t[0].sum()), dx_checksum=float(t[1].sum()), dB_checksum=float(t[3].sum()))
    print('AVT d4096 bf16 floor', index['avt_d4096_reference_bf16_floor'])

    with open(os.path.join(OUT, 'index.json'), 'w') as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
