"""CPU tests of the host side: config / rank encoding, PEFT-style injection and parameter naming,
checkpoint key layouts, the C-ABI library surface (symbols only -- no compute without a GPU),
and the loud failure when CUDA is not available."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

from moka_b200 import _cabi
from moka_b200.config import LLAMA_TARGETS, MokaConfig, parse_avt_rank
from moka_b200.inject import (PeftMixedModel, apply_save_modules_trainable, get_mm_adapter_state, get_peft_model,
                              get_peft_model_state_dict, set_peft_model_state_dict)
from moka_b200.layer import MokaLinearAVT, MokaLinearVT
from moka_b200.llama import LlamaConfig, LlamaForCausalLM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    """libmoka.so loads and exports exactly the functions include/moka.h declares."""
    lib = _cabi.load()
    header = open(os.path.join(ROOT, 'include', 'moka.h')).read()
    declared = set(re.findall(r'\b(moka_[a-z_0-9]+)\s*\(', header))
    declared.discard('moka_stream_t')
    assert declared == set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.moka_version() >= 100
    assert lib.moka_strerror(0) == b'ok'
    assert ctypes.sizeof(_cabi.MokaDesc) == 15 * 4             # POD layout of MokaDesc (15 x 32-bit fields)


def test_workspace_and_path_queries_need_no_gpu():
    from moka_b200.layer import FusedScalars, _make_desc
    lib = _cabi.load()
    sc = FusedScalars(1.0, (2.0, 2.0, 0.0), 0.05, 0.0)
    d = _make_desc((8192, 4096), 4, 2048, 4096, 16, 2, torch.bfloat16, 64, sc, 0)
    assert lib.moka_workspace_bytes(d) >= 8192 * 16 * 4 + 4 * 64 * 16 * 4
    d.r = 0
    assert lib.moka_workspace_bytes(d) == 0
    d = _make_desc((8192, 4096), 4, 2048, 4096, 16, 2, torch.float32, 64, sc, 0)
    assert lib.moka_uses_tensor_cores(d) == 0                   # fp32 storage keeps exact fp32 FMA


def test_avt_rank_digit_encoding():
    assert parse_avt_rank(444, 3) == [4, 4, 4]
    assert parse_avt_rank(16, 2) == [1, 6]                      # the reference's quirk
    with pytest.raises(IndexError):
        parse_avt_rank(16, 3)
    cfg = MokaConfig.from_avt_lora_config(r=444, lora_nums=3, lora_alpha=16, blc_weight=1)
    assert (cfg.variant, cfg.r, cfg.num_modalities) == ('avt', 4, 3)
    with pytest.raises(ValueError):
        MokaConfig.from_avt_lora_config(r=448, lora_nums=3)
    cfg = MokaConfig(variant='avt', r=16)                       # ranks >= 10 need the explicit form
    assert cfg.num_modalities == 3


def _tiny():
    torch.manual_seed(0)
    return LlamaForCausalLM(LlamaConfig.tiny(layers=2, hidden=64, inter=128, heads=4, vocab=97))


def test_inject_avt_names_sharing_and_trainable_flags():
    model = _tiny()
    w_before = model.model.layers[0].self_attn.q_proj.weight
    pm = get_peft_model(model, MokaConfig(variant='avt', r=4, lora_alpha=16, blc_weight=1.0))
    layers = pm.moka_layers()
    assert len(layers) == 2 * 7 and all(isinstance(l, MokaLinearAVT) for l in layers)
    q = pm.base_model.model.model.layers[0].self_attn.q_proj
    assert q.weight is w_before                                  # frozen weight shared, not copied (lora.py:178)
    names = dict(pm.named_parameters())
    key = 'base_model.model.model.layers.0.self_attn.q_proj.'
    for suffix in ('weight', 'lora_A0.weight', 'lora_A1.weight', 'lora_A2.weight', 'lora_B0.weight'):
        assert key + suffix in names
    assert names[key + 'lora_A1.weight'].shape == (4, 64) and names[key + 'lora_B0.weight'].shape == (64, 4)
    for n, p in names.items():
        assert p.requires_grad == ('lora_' in n), n
    assert float(names[key + 'lora_B0.weight'].abs().max()) == 0.0        # B starts at zero (lora.py:348)
    assert float(names[key + 'lora_A0.weight'].abs().max()) <= 1.0 / 8 + 1e-6   # kaiming a=sqrt(5): U(+-1/sqrt(64))
    tr, al = pm.get_nb_trainable_parameters()
    assert tr == 2 * (4 * (3 * 4 * 64 + 64 * 4) + 2 * (3 * 4 * 64 + 128 * 4) + (3 * 4 * 128 + 64 * 4))


def test_inject_vt_two_adapters_like_train_py():
    model = _tiny()
    cfg = MokaConfig(variant='vt', r=4, lora_alpha=16, lora_dropout=0.05, attn_weight=0.05)
    pm = PeftMixedModel(model, cfg, 'image')
    pm.add_adapter('text', cfg)
    pm.set_adapter(['image', 'text'])
    q = pm.base_model.model.model.layers[1].mlp.down_proj
    assert isinstance(q, MokaLinearVT) and set(q.lora_A.keys()) == {'image', 'text'}
    assert q.scaling == {'image': 4.0, 'text': 4.0} and q.attn_weight == 0.05
    names = dict(pm.named_parameters())
    key = 'base_model.model.model.layers.1.mlp.down_proj.'
    for suffix in ('base_layer.weight', 'lora_A.image.weight', 'lora_A.text.weight', 'lora_B.image.weight',
                   'lora_B.text.weight'):
        assert key + suffix in names
    assert not names[key + 'base_layer.weight'].requires_grad
    assert names[key + 'lora_B.image.weight'].requires_grad       # trainable but never used
    with pytest.raises(ValueError):
        q.update_layer('extra', 8, lora_alpha=16, lora_dropout=0.0)   # ranks must match


def test_missing_targets_raise_like_reference():
    with pytest.raises(ValueError, match='not found in the base model'):
        get_peft_model(_tiny(), MokaConfig(variant='avt', r=4, target_modules=['no_such_proj']))


def test_checkpoint_key_layouts_round_trip(tmp_path):
    model = _tiny()
    pm = get_peft_model(model, MokaConfig(variant='avt', r=4))
    with torch.no_grad():
        for n, p in pm.named_parameters():
            if 'lora_' in n:
                p.normal_()
    sd = get_peft_model_state_dict(pm)
    assert sd and all('lora_' in k for k in sd)
    pm.save_pretrained(str(tmp_path))
    loaded = torch.load(tmp_path / 'adapter_model.bin')
    assert set(loaded) == set(sd)
    pm2 = get_peft_model(_tiny(), MokaConfig(variant='avt', r=4))
    res = set_peft_model_state_dict(pm2, loaded)
    assert not res.unexpected_keys
    for k, v in get_peft_model_state_dict(pm2).items():
        assert torch.equal(v, sd[k])
    # finetune_weights.bin convention (trainer.py:183-218): names containing any save_modules key
    apply_save_modules_trainable(pm, 'vl_projector,al_projector,lora')
    st = get_mm_adapter_state(pm.named_parameters(), 'vl_projector,al_projector,lora'.split(','))
    assert set(st) == set(sd)


def test_cpu_tensors_fail_loudly_no_fallback():
    layer = MokaLinearAVT(64, 64, rank=4, lora_alpha=16, lora_nums=3, blc_weight=1.0, loramethod='train', bias=False)
    x = torch.randn(1, 8, 64)
    masks = [torch.ones(1, 8, 1, dtype=torch.int32)] + [torch.zeros(1, 8, 1, dtype=torch.int32)] * 2 + \
            [torch.ones(1, 8, 1, dtype=torch.int32)]
    with pytest.raises(_cabi.MokaError):
        layer(x, masks)
    base = nn.Linear(64, 64, bias=False)
    vt = MokaLinearVT(base, 'image', r=4, lora_alpha=16, attn_weight=0.05)
    vt.update_layer('text', 4, lora_alpha=16, lora_dropout=0.0)
    with pytest.raises(_cabi.MokaError):
        vt(x, torch.ones(1, 8, dtype=torch.bool), torch.zeros(1, 8, dtype=torch.bool), torch.ones(1, 8, dtype=torch.bool))
    with pytest.raises(_cabi.MokaError):
        vt(x, None, None, None)


def test_disabled_adapters_are_the_frozen_base():
    base = nn.Linear(32, 48, bias=False)
    vt = MokaLinearVT(base, 'image', r=4, lora_alpha=16, attn_weight=0.05)
    vt.enable_adapters(False)
    x = torch.randn(2, 5, 32)
    assert torch.equal(vt(x, None, None, None), base(x))
    avt = MokaLinearAVT(32, 48, rank=4, lora_nums=3, bias=False)
    avt.disable_adapters = True
    assert torch.equal(avt(x, None), nn.functional.linear(x, avt.weight))


def test_meta_cache_builds_once_per_mask_identity():
    from moka_b200.metadata import _MetaCache
    cache = _MetaCache(capacity=2)
    calls = []
    m = [torch.zeros(2, 4, dtype=torch.bool) for _ in range(3)]
    build = lambda: calls.append(1) or 'meta'
    for _ in range(5):
        assert cache.get(1, m, build) == 'meta'
    assert len(calls) == 1
    m[0].fill_(True)                                             # in-place edit bumps the version
    cache.get(1, m, build)
    assert len(calls) == 2
    cache.get(1, [t.clone() for t in m], build)                   # different tensors, same values
    assert len(calls) == 3


def test_kv_cache_mask_handoff_and_incremental_decode_cpu():
    """SURVEY 8f N4, host side (no MokaLinear on the CPU: the frozen caller alone).  masks_for_step mirrors
    VisualText/modified_models/modeling_llama.py:314-329; prefill + one-token (and multi-token) steps over the KV
    cache reproduce the full-sequence logits."""
    from moka_b200.llama import KVCache, LlamaConfig, LlamaForCausalLM, masks_for_step
    masks = ('t', 'i', 'q')
    assert masks_for_step(None, masks) is masks                        # training: no cache object at all
    c = KVCache(2)
    assert c.get_seq_length() == 0 and masks_for_step(c, masks) is masks   # first call of a generation
    torch.manual_seed(0)
    m = LlamaForCausalLM(LlamaConfig.tiny()).eval()
    ids = torch.randint(0, 500, (2, 12))
    with torch.no_grad():
        _, full = m(input_ids=ids)
        _, a = m(input_ids=ids[:, :8], past_key_values=c)
        assert c.get_seq_length() == 8 and masks_for_step(c, masks) is None   # decode steps: no masks
        steps = [a[:, -1]]
        for t in range(8, 12):
            _, b = m(input_ids=ids[:, t:t + 1], past_key_values=c)
            steps.append(b[:, -1])
        assert torch.allclose(torch.stack(steps, 1), full[:, 7:], atol=1e-5)
        c2 = KVCache(2)
        m(input_ids=ids[:, :8], past_key_values=c2)
        _, chunk = m(input_ids=ids[:, 8:], past_key_values=c2)
        assert torch.allclose(chunk, full[:, 8:], atol=1e-5)
        toks = m.generate(input_ids=ids[:, :8], max_new_tokens=4)
        assert toks.shape == (2, 4) and torch.equal(toks[:, 0], full[:, 7].argmax(-1))


def test_product_span_layouts_equal_the_oracles():
    """bench.py and tools/ build their workloads from moka_b200.layouts (the GPU arm never imports oracle/);
    the oracle keeps its own generator for the tests.  They must describe the same spans."""
    from moka_b200 import layouts
    from oracle import moka_oracle as mo
    assert (layouts.VARIANT_AVT, layouts.VARIANT_VT) == (mo.VARIANT_AVT, mo.VARIANT_VT)
    for variant, layout, S in ((mo.VARIANT_VT, 'c1', 128), (mo.VARIANT_VT, 'c2', 2048), (mo.VARIANT_VT, 'c2', 256),
                               (mo.VARIANT_AVT, 'c3', 2048), (mo.VARIANT_AVT, 'c3', 1024), (mo.VARIANT_VT, 'c4', 4096)):
        a = layouts.layout_masks(variant, 3, S, layout)
        b = mo.layout_masks(variant, 3, S, layout)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)


def test_flag_constants_match_the_header_and_fusion_is_opt_in():
    """_cabi's flag values are the header's; the GEMM-epilogue fusion (MOKA_FLAG_DELTA_OUT) is off unless asked
    for, so every default number is the read-modify-write order that SURVEY 8(d)'s byte model describes."""
    hdr = open(os.path.join(ROOT, 'include', 'moka.h')).read()
    vals = {k: int(v) for k, v in re.findall(r'(MOKA_FLAG_[A-Z_]+)\s*=\s*(\d+)\s*,', hdr)}
    assert vals['MOKA_FLAG_FORCE_SIMT'] == _cabi.FLAG_FORCE_SIMT
    assert vals['MOKA_FLAG_FORCE_TC'] == _cabi.FLAG_FORCE_TC
    assert vals['MOKA_FLAG_DELTA_OUT'] == _cabi.FLAG_DELTA_OUT
    assert int(re.search(r'MOKA_EUNSUPPORTED\s*=\s*(-?\d+)', hdr).group(1)) == _cabi.MOKA_EUNSUPPORTED
    from moka_b200 import layer
    if os.environ.get('MOKA_DELTA', '0') != '1':
        assert layer.FUSION['enabled'] is False


def test_bench_workloads_name_the_baseline_configs():
    """bench.py: the default workload is BASELINE.json's metric shape (7B, r = 16, three modalities), the other two
    named workloads are configs[1] and configs[3], and the metric string names the model the workload really runs."""
    import bench
    wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert (wl['model'], wl['r'], wl['M'], wl['S']) == ('7b', 16, 3, 2048)
    assert bench.metric_name(wl) == 'tokens/sec Llama-2-7B MokA fine-tune step'
    c2, c4 = bench.WORKLOADS['llama2-7b-vt-m2-r8-s2048'], bench.WORKLOADS['llama2-13b-vt-m2-r64-s4096']
    assert (c2['variant'], c2['r'], c2['S']) == ('vt', 8, 2048)
    assert (c4['variant'], c4['r'], c4['S']) == ('vt', 64, 4096)
    assert bench.metric_name(c4) == 'tokens/sec Llama-2-13B MokA fine-tune step'
