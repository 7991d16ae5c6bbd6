"""the mask producers.  CPU part: the oracle's restatement of the reference's data-pipeline
rules (VisualText/train/train.py:206-231 + collator; AudioVisualText/models/unified_arch.py:139-240, 304-312)
reproduces the synthetic layouts and the masks stored in the reference-generated golden fixtures.  GPU part: the
device builders (no mask tensors, no host sync) are bit-exact against the oracle's metadata."""
import glob
import os

import pytest
import torch

from moka_b200 import layouts as lay
from oracle import moka_oracle as mo

IMG = 31999
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('layout,B,S', [('c1', 2, 128), ('c2', 3, 2048), ('c2', 2, 512), ('c4', 2, 4096)])
def test_vt_rule_reproduces_the_layout_masks(layout, B, S):
    ids, labels = lay.layout_ids_labels(B, S, layout, IMG)
    got = mo.vt_masks_from_ids(ids, labels, IMG)
    want = lay.layout_masks(lay.VARIANT_VT, B, S, layout)
    for g, w in zip(got, want):
        assert torch.equal(g, w)


def test_vt_rule_padding_and_sample_without_image():
    B, S = 3, 256
    ids, labels = lay.layout_ids_labels(B, S, 'c2', IMG)
    lengths = [256, 200, 256]
    ids[1, 200:] = 2; labels[1, 200:] = -100                  # collator padding (train.py:266-268)
    ids[2][ids[2] == IMG] = 7                                 # no image token: position mask all False (:227)
    t, i, q = mo.vt_masks_from_ids(ids, labels, IMG, lengths)
    assert not bool(t[1, 200:].any() | i[1, 200:].any() | q[1, 200:].any())
    assert bool(t[1, :200].any()) and int(q[1].sum()) > 0
    assert int(i[2].sum()) == 0 and int(q[2].sum()) == 0 and int(t[2].sum()) == S


@pytest.mark.parametrize('S', [128, 2048])
def test_avt_segment_rule_reproduces_the_layout_masks(S):
    seg = lay.layout_segments(2, S)
    got = mo.avt_masks_from_segments([[(int(k), int(n)) for k, n in s] for s in seg], S)
    want = lay.layout_masks(lay.VARIANT_AVT, 2, S, 'c3')
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.parametrize('name', ['vt_c1_small', 'vt_r16_noncontig', 'vt_no_question_sample', 'vt_no_image_sample'])
def test_vt_rule_reproduces_golden_fixture_masks(name):
    """The masks the reference layers were run on (tests/golden, generated from /root/reference) are what the rule
    yields for ids / labels that encode them."""
    g = torch.load(os.path.join(GOLD, name + '.pt'), weights_only=False)
    text, image, question = g['masks']
    B, S = text.shape
    ids = torch.full((B, S), 11, dtype=torch.int64)
    ids[image] = IMG
    labels = ids.clone()
    for b in range(B):
        im = torch.where(image[b])[0]
        if len(im):
            labels[b, :int(im[-1]) + 1] = -100
        labels[b, question[b]] = -100
    t, i, q = mo.vt_masks_from_ids(ids, labels, IMG)
    assert torch.equal(t, text) and torch.equal(i, image)
    if name == 'vt_no_image_sample':
        # train.py:227: without an image token the sample gets NO question token, whatever the labels say
        assert int(q[0].sum()) == 0 and torch.equal(q[1], question[1])
    else:
        assert torch.equal(q, question)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('layout,B,S', [('c1', 2, 128), ('c2', 4, 2048), ('c4', 2, 4096)])
def test_gpu_meta_from_ids_bit_exact(layout, B, S):
    from moka_b200.metadata import build_token_meta_vt_from_ids
    ids, labels = lay.layout_ids_labels(B, S, layout, IMG)
    lengths = torch.full((B,), S, dtype=torch.int32)
    lengths[0] = S - 13
    ids[0, S - 13:] = 2; labels[0, S - 13:] = -100
    masks = mo.vt_masks_from_ids(ids, labels, IMG, lengths.tolist())
    om = mo.build_token_meta_vt(*masks)
    gm = build_token_meta_vt_from_ids(ids.cuda(), labels.cuda(), IMG, lengths=lengths.cuda(), max_keys=256)
    assert gm.status is not None and gm.max_klen is None          # check=False: nothing was read back
    assert torch.equal(gm.tokmeta.cpu(), om.tokmeta)
    assert torch.equal(gm.klen.cpu(), om.klen)
    L = int(om.klen.max())
    assert torch.equal(gm.kidx.cpu()[:, :L], om.kidx[:, :L])
    assert bool((gm.kidx.cpu()[:, L:] == -1).all())
    gm.raise_if_error()
    small = build_token_meta_vt_from_ids(ids.cuda(), labels.cuda(), IMG, max_keys=8)
    with pytest.raises(ValueError):
        small.raise_if_error()                                    # key overflow is reported, late but loudly


@pytest.mark.gpu
@pytest.mark.parametrize('S', [128, 2048])
def test_gpu_meta_from_segments_bit_exact(S):
    from moka_b200.metadata import build_token_meta_avt_from_segments
    B = 3
    seg = lay.layout_segments(B, S)
    seg[1, 6, 1] -= 5                                             # a shorter sample: more left padding
    seg[2, 0, 1] += 3; seg[2, 6, 1] -= 3
    masks = mo.avt_masks_from_segments([[(int(k), int(n)) for k, n in s] for s in seg], S)
    om = mo.build_token_meta_avt(masks)
    gm = build_token_meta_avt_from_segments(seg.cuda(), S, max_keys=128, check=True)
    assert torch.equal(gm.tokmeta.cpu(), om.tokmeta)
    assert torch.equal(gm.klen.cpu(), om.klen)
    L = int(om.klen.max())
    assert gm.Lk == L and torch.equal(gm.kidx.cpu()[:, :L], om.kidx[:, :L])
    # a sample without a question segment: the reference raises IndexError (lora.py:490) - here at the deferred check
    seg[0, 5, 0] = lay.SEG_TEXT
    bad = build_token_meta_avt_from_segments(seg.cuda(), S, max_keys=128)
    with pytest.raises(IndexError):
        bad.raise_if_error()
    seg[0, 5, 0] = 9
    with pytest.raises(ValueError):
        build_token_meta_avt_from_segments(seg.cuda(), S, max_keys=128, check=True)
