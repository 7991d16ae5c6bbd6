"""the checkpoint wire format, pinned to the reference.  tests/golden/ckpt_keys.json was produced by
oracle/gen_ckpt_keys.py from the UNMODIFIED reference PEFT forks wrapped around oracle/tiny_tree.TinyCausalLM; the same
tree wrapped by moka_b200.inject must expose the same state-dict keys / shapes / trainable flags and write / read the
same files (adapter_model.bin, non_lora_trainables.bin, model.safetensors)."""
import json
import os

import pytest
import torch

from moka_b200.config import MokaConfig
from moka_b200 import inject
from oracle.tiny_tree import TARGETS, TinyCausalLM

KEYS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'ckpt_keys.json')))


def _avt():
    m = inject.get_peft_model(TinyCausalLM(), MokaConfig(variant='avt', r=4, num_modalities=3, lora_alpha=16,
                                                         blc_weight=1.0, lora_dropout=0.05, target_modules=TARGETS))
    inject.apply_save_modules_trainable(m, 'vl_projector,al_projector,lora')       # finetune.py:154-160
    return m


def _vt():
    c = MokaConfig(variant='vt', r=4, lora_alpha=16, attn_weight=0.05, lora_dropout=0.05, target_modules=TARGETS)
    m = inject.PeftMixedModel(TinyCausalLM(), c, 'image')
    m.add_adapter('text', c)
    m.set_adapter(['image', 'text'])
    return m


def _describe(model):
    train = {n for n, p in model.named_parameters() if p.requires_grad}
    return [[k, list(v.shape), k in train] for k, v in model.state_dict().items()]


@pytest.mark.parametrize('variant', ['avt', 'vt'])
def test_state_dict_keys_shapes_and_trainable_flags_match_the_reference(variant):
    got = _describe(_avt() if variant == 'avt' else _vt())
    want = KEYS[variant]['state_dict']
    assert [g[0] for g in got] == [w[0] for w in want]            # same names, same order
    assert got == want


def test_avt_checkpoint_files_have_the_reference_key_sets(tmp_path):
    m = _avt()
    inject.save_reference_checkpoint(m, str(tmp_path))
    a = torch.load(tmp_path / 'adapter_model.bin')
    n = torch.load(tmp_path / 'non_lora_trainables.bin')
    assert sorted(a) == KEYS['avt']['adapter_model_bin']
    assert sorted(n) == KEYS['avt']['non_lora_trainables_bin']
    # round trip through the reference's loader order (inference_cut.py:174-184), into a fresh model
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad:
                p.add_(1.0)
    inject.save_reference_checkpoint(m, str(tmp_path))
    fresh = _avt()
    rep = inject.load_reference_checkpoint(fresh, str(tmp_path))
    assert rep == {'non_lora_trainables.bin': [], 'adapter_model.bin': []}
    for (k, v), (k2, v2) in zip(m.state_dict().items(), fresh.state_dict().items()):
        if 'lora_' in k or 'projector' in k:
            assert torch.equal(v, v2), k


def test_vt_model_safetensors_strict_round_trip(tmp_path):
    m = _vt()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.5)
    inject.save_safetensors_checkpoint(m, str(tmp_path))
    from safetensors.torch import load_file
    assert sorted(load_file(str(tmp_path / 'model.safetensors'))) == KEYS['vt']['model_safetensors']
    fresh = _vt()
    res = inject.load_safetensors_checkpoint(fresh, str(tmp_path), strict=True)    # mmbench.py:576-577
    assert not res.missing_keys and not res.unexpected_keys
    for (k, v), (_, v2) in zip(m.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(v, v2), k
