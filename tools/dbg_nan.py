"""Locate non-finite outputs of one C-ABI case (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from moka_b200 import _cabi
if os.environ.get('MOKA_LIB'):                       # A/B against another build of the library
    _cabi.LIB_PATH = os.environ['MOKA_LIB']
from oracle import moka_oracle as mo
from tests import test_gpu_parity as tp

variant, M, layout, r, d_in, d_out, S = mo.VARIANT_AVT, 3, 'c3', 128, 512, 256, 1024
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
o = tp._oracle_case(variant, 2, S, d_in, d_out, r, M, layout, 16, 1.0, torch.bfloat16, flags=flags)
for k in ('y', 'dx', 'dB', 'h', 'g'):
    t = o[k].float()
    bad = ~torch.isfinite(t)
    print(k, tuple(t.shape), 'nonfinite', int(bad.sum()))
    if int(bad.sum()):
        idx = bad.nonzero()
        print('  first', idx[:5].tolist(), 'last', idx[-3:].tolist())
        if t.dim() == 3:
            rows = bad.any(-1).nonzero()
            cols = bad.any(0).any(0).nonzero().flatten()
            print('  rows with nan', len(rows), rows[:8].tolist(), ' cols', len(cols), cols[:8].tolist(), cols[-4:].tolist())
for m in range(M):
    t = o['dA'][m]
    print('dA', m, int((~torch.isfinite(t)).sum()), float(mo.rel_l2(t, o['tdA'][m])) if float(o['tdA'][m].norm()) else 0)
bf = lambda t: t.to(torch.bfloat16).double()
t = o['dx'].double(); ref = bf(o['tdx'])
err = (t - ref).abs()
err[~torch.isfinite(err)] = 1e9
rows = (err.amax(-1) > 0.1).nonzero()
print('rows with big dx error', len(rows), rows[:10].tolist())
print('meta codes of those rows', [(int(o['ometa'].tokmeta.view(2, S)[b, s]) & 7) for b, s in rows[:10].tolist()])
bad = (err > 0.05)
idx = bad.nonzero()
print('bad elements', len(idx))
for b, s_, c in idx[:12].tolist():
    print('  row', b, s_, 'col', c, 'got', float(t[b, s_, c]), 'ref', float(ref[b, s_, c]), 'dx0', float(o['dx0'][b, s_, c]))
import collections
print('by (tile, slice):', sorted(collections.Counter(((b * S + s_) // 128, c // 128) for b, s_, c in idx.tolist()).items())[:20])
