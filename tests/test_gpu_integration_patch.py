"""INTEGRATION.md sections 3 and 4, executed: the two ``Linear.forward`` patches are read out of INTEGRATION.md
verbatim, installed on the vendored reference classes (oracle/_ref = byte-identical copies of
AudioVisualText/peft_hyper and VisualText/modified_peft) and compared, on the GPU, with the UNPATCHED reference
modules on the same inputs: outputs, input gradients and every adapter gradient (lora.py:367, layer.py:548)."""
import os
import re
import textwrap

import pytest
import torch
import torch.nn as nn

from oracle import moka_oracle as mo
from oracle import ref_loader as rl

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rl.available(), reason='oracle/_ref not vendored and /root/reference absent')]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _patch_from_integration_md(section: str):
    """The first ```python block after the heading that starts with `section`, compiled to a function."""
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    start = text.index(section)
    m = re.search(r"```python\n(.*?)```", text[start:], re.S)
    src = textwrap.dedent(m.group(1))
    ns = {'nn': nn, 'torch': torch}
    exec(compile(src, 'INTEGRATION.md:' + section, 'exec'), ns)
    return ns['forward']


def _grads(layer, names):
    out = {}
    for n, p in layer.named_parameters():
        if any(k in n for k in names) and p.grad is not None:
            out[n] = p.grad.detach().float().clone()
            p.grad = None
    return out


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.bfloat16, 8e-3)], ids=['fp32', 'bf16'])
def test_avt_patch_of_integration_md_section_3(dtype, tol):
    from moka_b200.metadata import META_CACHE
    Linear = rl.import_avt_linear()
    B, S, d_in, d_out, r, M = 2, 256, 256, 384, 4, 3
    x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(mo.VARIANT_AVT, B, S, d_in, d_out, r, M, 'c3')
    layer = rl.make_avt(Linear, d_in, d_out, r, M, 16, 1.0, W0, A, Bw, dtype, device='cuda')
    dm = [m.cuda() for m in masks]
    xd = x.to(dtype).cuda()
    dyd = dy.to(dtype).cuda()

    def run():
        xi = xd.clone().requires_grad_(True)
        y = layer(xi, dm)
        y.backward(dyd)
        return y.detach().float(), xi.grad.float(), _grads(layer, ['lora_'])

    y0, dx0, g0 = run()                                            # the reference, eager on the GPU
    fwd = _patch_from_integration_md('## 3. Patch to')
    orig = Linear.forward
    Linear.forward = fwd
    try:
        META_CACHE.clear()
        y1, dx1, g1 = run()
    finally:
        Linear.forward = orig
    assert mo.rel_l2(y1, y0) <= tol and mo.rel_l2(dx1, dx0) <= tol
    assert set(g0) == set(g1) and len(g0) == M + 1
    for k in g0:
        assert mo.rel_l2(g1[k], g0[k]) <= max(tol, 5e-5) * (4 if dtype == torch.bfloat16 else 1), k


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.bfloat16, 8e-3)], ids=['fp32', 'bf16'])
def test_vt_patch_of_integration_md_section_4(dtype, tol):
    from moka_b200.metadata import META_CACHE
    Linear = rl.import_vt_linear()
    B, S, d_in, d_out, r = 2, 128, 256, 384, 8
    x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(mo.VARIANT_VT, B, S, d_in, d_out, r, 2, 'c1')
    layer = rl.make_vt(Linear, d_in, d_out, r, 16, 0.05, W0, A, Bw, dtype, device='cuda')
    dm = [m.cuda() for m in masks]
    xd = x.to(dtype).cuda()
    dyd = dy.to(dtype).cuda()

    def run(with_masks=True):
        xi = xd.clone().requires_grad_(True)
        y = layer(xi, *dm) if with_masks else layer(xi[:, :1].contiguous(), None, None, None)
        y.backward(dyd if with_masks else dyd[:, :1].contiguous())
        return y.detach().float(), xi.grad.float(), _grads(layer, ['lora_'])

    ref = run(), run(False)
    fwd = _patch_from_integration_md('## 4. Patch to')
    orig = Linear.forward
    Linear.forward = fwd
    try:
        META_CACHE.clear()
        got = run(), run(False)
    finally:
        Linear.forward = orig
    for (y0, dx0, g0), (y1, dx1, g1) in zip(ref, got):
        assert mo.rel_l2(y1, y0) <= tol and mo.rel_l2(dx1, dx0) <= tol
        assert set(g0) == set(g1)                                  # lora_B.image gets no gradient in either (layer.py:657)
        assert not any('lora_B.image' in k for k in g1)
        for k in g0:
            assert mo.rel_l2(g1[k], g0[k]) <= max(tol, 5e-5) * (4 if dtype == torch.bfloat16 else 1), k
