"""Import and instantiate the UNMODIFIED reference MokA layers (test / baseline infrastructure).

Two places can provide them: the reference tree itself (``/root/reference``, build container only) and its
byte-identical vendored copy ``oracle/_ref`` (made by ``oracle/vendor_ref.py``; travels to the GPU box).  Only
``accelerate`` is stubbed: the package ``__init__`` chains import it, the layers never use it.

  AVT  ``peft_hyper.tuners.lora.Linear``        AudioVisualText/peft_hyper/tuners/lora.py:277-532
  VT   ``peft.tuners.lora.layer.Linear``        VisualText/modified_peft/tuners/lora/layer.py:389-685
"""
from __future__ import annotations

import os
import sys
import tempfile
from unittest import mock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED = os.path.join(HERE, '_ref')
REFERENCE = os.environ.get('MOKA_REFERENCE_ROOT', '/root/reference')
_cache = {}


def available() -> str:
    """'vendored', 'tree' or '' (neither present)."""
    if os.path.isfile(os.path.join(VENDORED, 'avt', 'peft_hyper', 'tuners', 'lora.py')) and \
            os.path.isfile(os.path.join(VENDORED, 'vt', 'peft', 'tuners', 'lora', 'layer.py')):
        return 'vendored'
    if os.path.isdir(os.path.join(REFERENCE, 'AudioVisualText', 'peft_hyper')):
        return 'tree'
    return ''


def stub_accelerate() -> None:
    if _cache.get('stubbed'):
        return
    import transformers  # noqa: F401  (must be imported before the stub)
    import transformers.pytorch_utils  # noqa: F401
    from transformers import PreTrainedModel  # noqa: F401
    from transformers.modeling_outputs import SequenceClassifierOutput  # noqa: F401
    if 'accelerate' not in sys.modules or isinstance(sys.modules['accelerate'], mock.MagicMock):
        acc = mock.MagicMock()
        acc.__version__ = '0.30.0'                       # VT/modified_peft/utils/other.py:50 parses it
        sys.modules['accelerate'] = acc
        for n in ('accelerate.hooks', 'accelerate.utils', 'accelerate.utils.imports'):
            sys.modules[n] = mock.MagicMock()
    _cache['stubbed'] = True


def import_avt_linear(prefer_tree: bool = False):
    """``Linear`` of AudioVisualText/peft_hyper/tuners/lora.py:277."""
    if 'avt' in _cache:
        return _cache['avt']
    stub_accelerate()
    kind = available()
    if kind == 'tree' or (prefer_tree and os.path.isdir(REFERENCE)):
        sys.path.insert(0, os.path.join(REFERENCE, 'AudioVisualText'))
    elif kind == 'vendored':
        sys.path.insert(0, os.path.join(VENDORED, 'avt'))
    else:
        raise ImportError('reference not available: neither oracle/_ref nor /root/reference exists')
    prev = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from peft_hyper.tuners.lora import Linear
    finally:
        sys.dont_write_bytecode = prev
    _cache['avt'] = Linear
    return Linear


def import_vt_linear(prefer_tree: bool = False):
    """``Linear`` of VisualText/modified_peft/tuners/lora/layer.py:389 (the package imports itself as ``peft``)."""
    if 'vt' in _cache:
        return _cache['vt']
    stub_accelerate()
    kind = available()
    if kind == 'tree' or (prefer_tree and os.path.isdir(REFERENCE)):
        tmp = tempfile.mkdtemp(prefix='oracle_vt_')
        os.symlink(os.path.join(REFERENCE, 'VisualText', 'modified_peft'), os.path.join(tmp, 'peft'))
        sys.path.insert(0, tmp)
    elif kind == 'vendored':
        sys.path.insert(0, os.path.join(VENDORED, 'vt'))
    else:
        raise ImportError('reference not available: neither oracle/_ref nor /root/reference exists')
    prev = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from peft.tuners.lora.layer import Linear
    finally:
        sys.dont_write_bytecode = prev
    _cache['vt'] = Linear
    return Linear


def make_avt(Linear, d_in, d_out, r, M, alpha, w, W0=None, A=None, Bw=None, dtype=torch.float32, dropout=0.0,
             device=None):
    """Reference AVT layer with rank ``r`` per modality.  Ranks >= 10 are not expressible through the reference's
    digit encoding (lora.py:256-259): the sub-modules are replaced after construction, forward only reads
    ``lora_A{i}``, ``lora_B0``, ``scaling``, ``d_k`` (verified by the golden vectors)."""
    digits = int(str(min(r, 9)) * M)
    l = Linear(d_in, d_out, r=digits, lora_alpha=alpha, lora_nums=M, blc_alpha=1, blc_weight=w,
               lora_dropout=dropout, reserved_modality=None, loramethod='train',
               merge_weights=False, bias=False)
    for i in range(M):
        setattr(l, f'lora_A{i}', torch.nn.Linear(d_in, r, bias=False))
    l.lora_B0 = torch.nn.Linear(r, d_out, bias=False)
    l.r = [r] * M
    l.d_k = r
    l.scaling = [alpha / r]
    l = l.to(dtype)
    if device is not None:
        l = l.to(device)
    with torch.no_grad():
        if W0 is not None:
            l.weight.copy_(W0.to(dtype))
        for i in range(M):
            if A is not None:
                getattr(l, f'lora_A{i}').weight.copy_(A[i].to(dtype))
        if Bw is not None:
            l.lora_B0.weight.copy_(Bw.to(dtype))
    l.weight.requires_grad_(False)
    return l


def make_vt(Linear, d_in, d_out, r, alpha, w, W0=None, A=None, Bw=None, dtype=torch.float32, dropout=0.0,
            device=None):
    base = torch.nn.Linear(d_in, d_out, bias=False).to(dtype)
    l = Linear(base, 'image', r=r, lora_alpha=alpha, lora_dropout=dropout, attn_weight=w)
    l.update_layer('text', r, lora_alpha=alpha, lora_dropout=dropout, init_lora_weights=True,
                   use_rslora=False, use_dora=False)
    l.set_adapter(['image', 'text'])
    l = l.to(dtype)                                  # update_layer hard-codes bf16 for A/B (layer.py:109-110)
    if device is not None:
        l = l.to(device)
    with torch.no_grad():
        if W0 is not None:
            l.base_layer.weight.copy_(W0.to(dtype))
        if A is not None:
            l.lora_A['text'].weight.copy_(A[0].to(dtype))
            l.lora_A['image'].weight.copy_(A[1].to(dtype))
        if Bw is not None:
            l.lora_B['text'].weight.copy_(Bw.to(dtype))
    l.base_layer.weight.requires_grad_(False)
    return l
