"""The vendored copy of the reference layers (oracle/_ref, made by oracle/vendor_ref.py) is what times as the CPU
baseline on the GPU box: it must reproduce the committed golden vectors (which came from /root/reference itself),
and its MANIFEST must describe every file.  Skipped when oracle/_ref has not been built."""
import glob
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, 'oracle', '_ref')
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REFDIR, 'MANIFEST.json')),
                                reason='oracle/_ref not vendored (run python oracle/vendor_ref.py in the build container)')

_SCRIPT = r'''
import os, sys, json, torch
sys.path.insert(0, os.environ['MOKA_ROOT'])
from oracle import ref_loader as rl, moka_oracle as mo
assert rl.available() == 'vendored', rl.available()
out = {}
for name in ('vt_c1_small', 'vt_pad_tail', 'avt_r16', 'avt_noncontig'):
    g = torch.load(os.path.join(os.environ['MOKA_ROOT'], 'tests', 'golden', name + '.pt'), weights_only=False)
    f = lambda t: t.to(torch.float64)
    x = f(g['x']).requires_grad_(True)
    if g['variant'] == mo.VARIANT_AVT:
        l = rl.make_avt(rl.import_avt_linear(), g['d_in'], g['d_out'], g['r'], g['M'], g['alpha'], g['w'],
                        f(g['W0']), [f(a) for a in g['A']], f(g['Bw']), torch.float64)
        y = l(x, g['masks'])
    else:
        l = rl.make_vt(rl.import_vt_linear(), g['d_in'], g['d_out'], g['r'], g['alpha'], g['w'],
                       f(g['W0']), [f(a) for a in g['A']], f(g['Bw']), torch.float64)
        y = l(x, *g['masks'])
    y.backward(f(g['dy']))
    out[name] = [bool(torch.equal(y.detach(), g['y_f64'])), bool(torch.equal(x.grad, g['dx_f64']))]
import peft_hyper, peft
out['paths'] = [peft_hyper.__file__, peft.__file__]
print(json.dumps(out))
'''


def test_vendored_reference_reproduces_golden_vectors_bitwise():
    env = dict(os.environ, MOKA_ROOT=ROOT, MOKA_REFERENCE_ROOT='/nonexistent', PYTHONDONTWRITEBYTECODE='1')
    r = subprocess.run([sys.executable, '-c', _SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    for p in out.pop('paths'):
        assert os.path.realpath(p).startswith(os.path.realpath(REFDIR)), p
    for name, (ey, edx) in out.items():
        assert ey and edx, name


def test_manifest_covers_every_vendored_file():
    man = json.load(open(os.path.join(REFDIR, 'MANIFEST.json')))['files']
    files = [os.path.relpath(p, REFDIR) for p in glob.glob(os.path.join(REFDIR, '**', '*.py'), recursive=True)]
    assert sorted(files) == sorted(man)
    assert 'avt/peft_hyper/tuners/lora.py' in man and 'vt/peft/tuners/lora/layer.py' in man
