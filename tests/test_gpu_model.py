"""GPU tests of the layers above the C ABI: the two MokaLinear facades under autograd against the
reference-generated golden vectors, the fused AdamW kernel, the data-parallel engine at world size
1, and a tiny frozen-Llama fine-tune step end to end."""
import glob
import os

import pytest
import torch

from oracle import moka_oracle as mo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*vt_*.pt')))
IDS = [os.path.basename(p)[:-3] for p in GOLD]


def _facade(g, dtype):
    from moka_b200.layer import MokaLinearAVT, MokaLinearVT
    f = lambda t: t.to(torch.float64)
    W0, A, Bw = f(g['W0']), [f(a) for a in g['A']], f(g['Bw'])
    if g['variant'] == mo.VARIANT_VT:
        base = torch.nn.Linear(g['d_in'], g['d_out'], bias=False)
        layer = MokaLinearVT(base, 'image', r=g['r'], lora_alpha=g['alpha'], lora_dropout=0.0, attn_weight=g['w'])
        layer.update_layer('text', g['r'], lora_alpha=g['alpha'], lora_dropout=0.0)
        layer.set_adapter(['image', 'text'])
        layer = layer.to(DEV, dtype)
        with torch.no_grad():
            layer.base_layer.weight.copy_(W0)
            layer.lora_A['text'].weight.copy_(A[0]); layer.lora_A['image'].weight.copy_(A[1])
            layer.lora_B['text'].weight.copy_(Bw)
        args = [m.to(DEV) for m in g['masks']]
        params = [layer.lora_A['text'].weight, layer.lora_A['image'].weight, layer.lora_B['text'].weight]
        unused = layer.lora_B['image'].weight
    else:
        layer = MokaLinearAVT(g['d_in'], g['d_out'], rank=g['r'], lora_alpha=g['alpha'], lora_nums=g['M'],
                              blc_weight=g['w'], loramethod='train', bias=False).to(DEV, dtype)
        with torch.no_grad():
            layer.weight.copy_(W0)
            for i in range(g['M']):
                getattr(layer, f'lora_A{i}').weight.copy_(A[i])
            layer.lora_B0.weight.copy_(Bw)
        args = [[m.to(DEV) for m in g['masks']]]
        params = [getattr(layer, f'lora_A{i}').weight for i in range(g['M'])] + [layer.lora_B0.weight]
        unused = None
    return layer, args, params, unused


@pytest.mark.parametrize('path', GOLD, ids=IDS)
def test_facades_match_reference_autograd_fp32(path):
    """MokaLinearAVT / MokaLinearVT (frozen GEMM by cuBLAS + libmoka.so) vs the reference layer's
    own forward and autograd gradients, fp32, rel <= 1e-5."""
    g = torch.load(path, weights_only=False)
    layer, args, params, unused = _facade(g, torch.float32)
    x = g['x'].to(torch.float32).to(DEV).requires_grad_(True)
    y = layer(x, *args)
    y.backward(g['dy'].to(torch.float32).to(DEV))
    assert mo.rel_l2(y.detach().cpu(), g['y_f64']) <= 1e-5
    assert mo.rel_l2(x.grad.cpu(), g['dx_f64']) <= 1e-5
    for m in range(g['M']):
        ref = g['dA_f64'][m]
        got = params[m].grad
        if float(ref.norm()) == 0:
            assert got is None or float(got.norm()) == 0
        else:
            assert mo.rel_l2(got.cpu(), ref) <= 1e-5
    assert mo.rel_l2(params[-1].grad.cpu(), g['dB_f64']) <= 1e-5
    if unused is not None:
        assert unused.grad is None                       # lora_B.image never receives a gradient


@pytest.mark.parametrize('variant', ['avt', 'vt'])
def test_gemm_epilogue_fusion_under_autograd(variant):
    """layer.FUSION (MOKA_FLAG_DELTA_OUT + beta = 1 GEMMs) through the facades and autograd, bf16:
    same bars against the oracle as the default order, and the two orders agree with each other."""
    from moka_b200 import layer as L
    v = mo.VARIANT_AVT if variant == 'avt' else mo.VARIANT_VT
    M, layout, w = (3, 'c3', 1.0) if variant == 'avt' else (2, 'c2', 0.05)
    B, S, d_in, d_out, r = 2, 2048, 512, 768, 16
    x, W0, A, Bw, dy, masks = mo.synth_layer_inputs(v, B, S, d_in, d_out, r, M, layout, seed=3)
    g = dict(variant=v, d_in=d_in, d_out=d_out, r=r, alpha=16, w=w, M=M, W0=W0, A=A, Bw=Bw, masks=masks)
    meta = mo.build_token_meta_avt(masks) if variant == 'avt' else mo.build_token_meta_vt(*masks)
    sc = mo.FusedScalars.for_variant(v, lora_alpha=16, r=r, w=w)
    bf = lambda t: t.to(torch.bfloat16).double()
    xb, dyb = bf(x), bf(dy)
    W0b, Ab, Bwb = bf(W0), [bf(a) for a in A], bf(Bw)
    ty, th = mo.fused_forward(meta, xb, xb @ W0b.t(), Ab, Bwb, sc)
    tdxa, tdA, tdB, _ = mo.fused_backward(meta, xb, dyb, th, Ab, Bwb, sc)
    tdx = tdxa + dyb @ W0b
    outs = {}
    was = L.FUSION['enabled']
    try:
        for mode in (False, True):
            L.FUSION['enabled'] = mode
            layer, args, params, _ = _facade(g, torch.bfloat16)
            xd = x.to(torch.bfloat16).to(DEV).requires_grad_(True)
            y = layer(xd, *args)
            y.backward(dy.to(torch.bfloat16).to(DEV))
            torch.cuda.synchronize()
            outs[mode] = (y.detach().cpu(), xd.grad.cpu(), [p.grad.float().cpu() for p in params])
            # the frozen product is rounded once more in the default order; the bar on the totals covers both
            assert mo.rel_l2(outs[mode][0], bf(ty)) <= 4e-3
            assert mo.rel_l2(outs[mode][1], bf(tdx)) <= 4e-3
            for m in range(M):
                assert mo.rel_l2(outs[mode][2][m], tdA[m]) <= 1e-2            # bf16 parameter gradients
            assert mo.rel_l2(outs[mode][2][-1], tdB) <= 1e-2
    finally:
        L.FUSION['enabled'] = was
    assert not L.FUSION['unsupported']                                          # the fused order really ran
    # one rounding of W0 x + term (fused) against two (default): the fused result is at least as close
    assert mo.rel_l2(outs[True][0], ty) <= mo.rel_l2(outs[False][0], ty) * 1.02 + 1e-6
    assert mo.rel_l2(outs[True][0], outs[False][0]) <= 5e-3
    assert mo.rel_l2(outs[True][1], outs[False][1]) <= 5e-3


def test_decode_branches_match_reference(golden_dir):
    from moka_b200.layer import MokaLinearAVT, MokaLinearVT
    g = torch.load(os.path.join(golden_dir, 'decode_plain_lora.pt'), weights_only=False)
    f = lambda t: t.to(torch.float32).to(DEV)
    x = f(g['x'])
    base = torch.nn.Linear(64, 96, bias=False)
    vt = MokaLinearVT(base, 'image', r=8, lora_alpha=g['alpha'], attn_weight=0.05)
    vt.update_layer('text', 8, lora_alpha=g['alpha'], lora_dropout=0.0)
    vt = vt.to(DEV)
    with torch.no_grad():
        vt.base_layer.weight.copy_(f(g['W0'])); vt.lora_A['text'].weight.copy_(f(g['A_text']))
        vt.lora_B['text'].weight.copy_(f(g['Bw']))
    assert mo.rel_l2(vt(x, None, None, None).cpu(), g['y_vt']) <= 1e-5
    avt = MokaLinearAVT(64, 96, rank=8, lora_alpha=g['alpha'], lora_nums=3, blc_weight=1.0, loramethod='test',
                        bias=False).to(DEV)
    with torch.no_grad():
        avt.weight.copy_(f(g['W0'])); avt.lora_A0.weight.copy_(f(g['A_text'])); avt.lora_B0.weight.copy_(f(g['Bw']))
    assert mo.rel_l2(avt(x, None).cpu(), g['y_avt']) <= 1e-5


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
def test_fused_adamw_matches_torch(dtype):
    from moka_b200.dp import _cuda_adamw
    n = 100003
    gen = torch.Generator(device=DEV).manual_seed(1)
    w0 = torch.randn(n, device=DEV, generator=gen)
    ref_p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.AdamW([ref_p], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    master, m, v = w0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    out = torch.empty(n, device=DEV, dtype=dtype)
    for step in range(1, 6):
        g = torch.randn(n, device=DEV, generator=gen)
        ref_p.grad = g.clone()
        opt.step()
        _cuda_adamw(n, master, m, v, g, out, 3e-3, 0.9, 0.999, 1e-8, 0.05, step, 1.0)
    torch.cuda.synchronize()
    assert mo.rel_l2(master.cpu(), ref_p.detach().cpu()) <= 2e-6
    assert torch.equal(out, master.to(dtype))


def _tiny_model(variant, dtype=torch.bfloat16, r=8, dropout=0.0, ckpt=False):
    from moka_b200.config import MokaConfig
    from moka_b200.inject import PeftMixedModel, get_peft_model
    from moka_b200.llama import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    model = LlamaForCausalLM(LlamaConfig.tiny(layers=2, hidden=256, inter=512, heads=4, vocab=512)).to(DEV, dtype)
    if variant == 'avt':
        pm = get_peft_model(model, MokaConfig(variant='avt', r=r, lora_alpha=16, blc_weight=1.0, lora_dropout=dropout))
    else:
        cfg = MokaConfig(variant='vt', r=r, lora_alpha=16, attn_weight=0.05, lora_dropout=dropout)
        pm = PeftMixedModel(model, cfg, 'image')
        pm.add_adapter('text', cfg)
        pm.set_adapter(['image', 'text'])
    model.gradient_checkpointing_enable(ckpt)
    pm.train()
    return pm


def _batch(variant, B=2, S=256):
    gen = torch.Generator().manual_seed(5)
    ids = torch.randint(3, 500, (B, S), generator=gen)
    if variant == 'avt':
        masks = [m.to(DEV) for m in mo.layout_masks(mo.VARIANT_AVT, B, S, 'c3')]
        kw = dict(modality_masks=masks)
    else:
        t, i, q = [m.to(DEV) for m in mo.layout_masks(mo.VARIANT_VT, B, S, 'c2')]
        kw = dict(my_text_mask=t, my_image_mask=i, question_mask=q)
    labels = ids.clone()
    labels[:, : S // 2] = -100
    return ids.to(DEV), labels.to(DEV), kw


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-4), (torch.bfloat16, 4e-2)], ids=['fp32', 'bf16'])
@pytest.mark.parametrize('variant', ['avt', 'vt'])
def test_generation_with_kv_cache_matches_full_forward(variant, dtype, tol):
    """the prefill call takes the MokA branch with the masks, the decode steps take the no-mask
    branch behind the KV cache (VT modeling_llama.py:314-329 -> layer.py:672-678; AVT lora.py:373-381).  Because
    generated tokens are text tokens (text adapter, no attention term), the step-by-step logits must equal those
    of ONE full forward over prompt + generated tokens whose masks mark the new tokens as text."""
    from moka_b200.llama import KVCache
    pm = _tiny_model(variant, dtype=dtype)
    pm.eval()
    with torch.no_grad():
        for n, p in pm.named_parameters():
            if 'lora_B' in n:
                p.normal_(std=0.02)
    for mod in pm.modules():
        if hasattr(mod, 'loramethod'):
            mod.loramethod = 'test'                       # inference.py builds the AVT model with the test method
    B, S, K = 2, 256, 4
    ids, _, kw = _batch(variant, B, S)
    gen = torch.Generator().manual_seed(9)
    new = torch.randint(3, 500, (B, K), generator=gen).to(DEV)
    all_ids = torch.cat((ids, new), dim=1)

    def extend(m, value):
        pad = torch.full((B, K) + tuple(m.shape[2:]), value, dtype=m.dtype, device=m.device)
        return torch.cat((m, pad), dim=1)
    if variant == 'avt':
        t, v, a, q = kw['modality_masks']
        kw_full = dict(modality_masks=[extend(t, 1), extend(v, 0), extend(a, 0), extend(q, 0)])
    else:
        kw_full = dict(my_text_mask=extend(kw['my_text_mask'], True), my_image_mask=extend(kw['my_image_mask'], False),
                       question_mask=extend(kw['question_mask'], False))
    with torch.no_grad():
        _, full = pm(input_ids=all_ids, return_logits=True, **kw_full)
        cache = KVCache(2)
        _, pre = pm(input_ids=ids, past_key_values=cache, **kw)          # masks -> MokA branch
        steps = [pre[:, -1]]
        for j in range(K):
            _, lg = pm(input_ids=new[:, j:j + 1], past_key_values=cache, **kw)   # masks passed, not used (:326-329)
            steps.append(lg[:, -1])
        inc = torch.stack(steps, dim=1)
    assert cache.get_seq_length() == S + K
    ref = full[:, S - 1:].float()
    assert mo.rel_l2(inc.float().cpu(), ref.cpu()) <= tol
    # and the prefill really used the masks: without them the prompt's logits differ
    with torch.no_grad():
        if variant == 'vt':
            _, nomask = pm(input_ids=ids, return_logits=True)
            assert mo.rel_l2(nomask[:, -1].float().cpu(), pre[:, -1].float().cpu()) > 10 * tol or dtype == torch.bfloat16
        toks = pm.generate(input_ids=ids, max_new_tokens=3, **kw)
    assert toks.shape == (B, 3) and torch.equal(toks[:, 0], pre[:, -1].argmax(-1))


@pytest.mark.parametrize('variant', ['avt', 'vt'])
def test_tiny_llama_finetune_step_trains(variant):
    """Frozen tiny Llama + injected MokaLinear + flat-bucket AdamW (world size 1): the loss goes
    down, only lora_ parameters change, gradients land in the flat fp32 bucket."""
    from moka_b200 import _cabi
    from moka_b200.dp import AdamWConfig, AdapterDataParallel
    pm = _tiny_model(variant)
    with torch.no_grad():       # B starts at zero (reference init): give it signal so every path is live
        for n, p in pm.named_parameters():
            if 'lora_B' in n:
                p.normal_(std=0.02)
    frozen_before = {n: p.detach().clone() for n, p in pm.named_parameters() if not p.requires_grad}
    eng = AdapterDataParallel(pm, AdamWConfig(lr=2e-3), zero_stage=2, num_buckets=2)
    ids, labels, kw = _batch(variant)
    n0 = _cabi.launch_count()
    losses = []
    for _ in range(6):
        loss, _ = pm(input_ids=ids, labels=labels, **kw)
        loss.backward()
        assert float(eng.flat_grad.abs().max()) > 0
        eng.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] - 1e-3, losses
    assert _cabi.launch_count() - n0 >= 6 * (14 * 4 + 1)
    for n, p in pm.named_parameters():
        if not p.requires_grad:
            assert torch.equal(p, frozen_before[n]), n


def test_checkpointed_forward_with_dropout_is_consistent():
    """Per-layer gradient checkpointing (AVT recipe) re-runs the forward in backward; the
    dropout mask must be the one of the first run, so gradients equal the non-checkpointed ones."""
    grads = []
    for ckpt in (False, True):
        pm = _tiny_model('avt', dropout=0.1, ckpt=ckpt)
        with torch.no_grad():
            for n, p in pm.named_parameters():
                if 'lora_B' in n:
                    p.normal_(std=0.02, generator=torch.Generator(device=DEV).manual_seed(len(n)))
        ids, labels, kw = _batch('avt')
        torch.manual_seed(123)
        loss, _ = pm(input_ids=ids, labels=labels, **kw)
        loss.backward()
        grads.append({n: p.grad.detach().float().clone() for n, p in pm.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 0
    for n in grads[0]:
        assert mo.rel_l2(grads[1][n].cpu(), grads[0][n].cpu()) <= 2e-2, n


def test_fused_rope_matches_eager():
    from moka_b200.llama import _rope_tables, _rotate_half, apply_rope
    B, S, H, D = 2, 300, 4, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(B, S, H, D, device=DEV, generator=g).to(torch.bfloat16)
    cos, sin = _rope_tables(S, D, 10000.0, DEV, torch.float32)
    ref = x.float() * cos + _rotate_half(x.float()) * sin
    xr = x.clone().requires_grad_(True)
    y = apply_rope(xr * 1.0, (cos, sin), 10000.0)
    assert mo.rel_l2(y.detach().cpu(), ref.cpu()) <= 4e-3          # bf16 output rounding
    dy = torch.randn(B, S, H, D, device=DEV, generator=g).to(torch.bfloat16)
    y.backward(dy)
    gref = dy.float() * cos - _rotate_half(dy.float() * sin)       # transpose of the rotation
    assert mo.rel_l2(xr.grad.cpu(), gref.cpu()) <= 4e-3
