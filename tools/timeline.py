"""Wall-clock timeline (globaltimer) of one moka_forward + moka_backward chain: per kernel the first CTA
entry and the last CTA exit, relative to the first kernel's entry.  MOKA_DEBUG_TIMING=1 is set here.
usage: python tools/timeline.py [avt|vt] [r] [B]      (MOKA_TL_D / MOKA_TL_S / MOKA_TL_P: width, sequence length, lora_dropout)"""
import ctypes, os, sys
os.environ['MOKA_DEBUG_TIMING'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from moka_b200 import _cabi
from moka_b200.layer import FusedScalars, _make_desc
from moka_b200.metadata import build_token_meta_avt, build_token_meta_vt
from moka_b200.layouts import VARIANT_AVT, VARIANT_VT, layout_masks

variant = VARIANT_VT if (len(sys.argv) > 1 and sys.argv[1] == 'vt') else VARIANT_AVT
r = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lib = _cabi.load()
dev = 'cuda:0'
S, d = int(os.environ.get('MOKA_TL_S', 2048)), int(os.environ.get('MOKA_TL_D', 4096))
P_DROP = float(os.environ.get('MOKA_TL_P', 0.0))
T = B * S
M = 3 if variant == VARIANT_AVT else 2
masks = [m.to(dev) for m in layout_masks(variant, B, S, 'c3' if variant == VARIANT_AVT else 'c2')]
meta = build_token_meta_avt(masks) if variant == VARIANT_AVT else build_token_meta_vt(*masks)
s = 16.0 / r
sc = FusedScalars(s, (1, 1, 1), 1.0, P_DROP) if variant == VARIANT_AVT else FusedScalars(1.0, (s, s, s), 0.05, P_DROP)
bufs = [[torch.randn(T, d, device=dev).bfloat16() for _ in range(4)] for _ in range(3)]
A = [(torch.randn(r, d, device=dev) * d ** -0.5).bfloat16() for _ in range(M)]
Bw = (torch.randn(d, r, device=dev) * 0.02).bfloat16()
dA = [torch.zeros(r, d, device=dev) for _ in range(M)]
dB = torch.zeros(d, r, device=dev)
h = torch.empty(2 * T * r + T, device=dev)
st = torch.cuda.current_stream().cuda_stream
Ap = _cabi.ptr_array([a.data_ptr() for a in A]); dAp = _cabi.ptr_array([a.data_ptr() for a in dA])
desc = _make_desc((T, d), B, S, d, r, M, torch.bfloat16, meta.Lk, sc, 0)
ws = torch.empty(int(lib.moka_workspace_bytes(desc)), dtype=torch.uint8, device=dev)


def run(i):
    x, y, dy, dx = bufs[i % 3]
    _cabi.check(lib.moka_forward(desc, x.data_ptr(), y.data_ptr(), y.data_ptr(), Ap, Bw.data_ptr(),
                                 meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                 h.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), st), 'fwd')
    _cabi.check(lib.moka_backward(desc, dy.data_ptr(), x.data_ptr(), h.data_ptr(), Ap, Bw.data_ptr(),
                                  meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                  dx.data_ptr(), dAp, dB.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), st), 'bwd')


for i in range(4):
    run(i)
torch.cuda.synchronize()
run(4)
torch.cuda.synchronize()
n = 148 * 4 * 128
buf = (ctypes.c_ulonglong * n)()
_cabi.check(lib.moka_debug_read(buf, n), 'debug_read')
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 128).astype(np.int64)
slots = [('down', 100, 101), ('attn_fwd', 102, 103), ('up', 104, 105), ('prep', 106, 107), ('dgrad', 108, 109),
         ('attn_pack', 110, 111), ('wgrad', 112, 113)]
t_ref = None
rows = []
for name, se, sx in slots:
    ent = a[:, se][a[:, se] > 0]
    ext = a[:, sx][a[:, sx] > 0]
    if len(ent) == 0:
        continue
    # keep only the stamps of the last run (within 1 ms of the newest)
    ent = ent[ent > ent.max() - 1_000_000]
    ext = ext[ext > ext.max() - 1_000_000]
    if t_ref is None:
        t_ref = ent.min()
    rows.append((name, len(ent), (ent.min() - t_ref) / 1e3, (np.median(ent) - t_ref) / 1e3, (ent.max() - t_ref) / 1e3,
                 (ext.min() - t_ref) / 1e3, (np.median(ext) - t_ref) / 1e3, (ext.max() - t_ref) / 1e3))
print(f'{"kernel":10s} {"CTAs":>5s} | entry first/median/last (us) | exit first/median/last (us)')
for rrow in rows:
    print('%-10s %5d | %8.2f %8.2f %8.2f | %8.2f %8.2f %8.2f' % rrow)

# attn_pack phases (per CTA, relative to t_ref)
names = [(110, 'entry'), (114, 'staged (pre-wait)'), (115, 'wait passed'), (116, 'phase A done'), (117, 'phase B done / compute done'),
         (118, 'packed/parked'), (119, 'arrived'), (111, 'exit')]
ent = a[:, 110]
sel = ent > ent.max() - 1_000_000
q = sel & (a[:, 117] > a[:, 110])          # CTAs that ran the attention (query phase)
print('attn_pack CTAs: %d recorded, %d with queries' % (sel.sum(), q.sum()))
for slot, nm in names:
    for lab, m in (('all', sel), ('query', q)):
        v = a[m, slot]
        v = v[v > 0]
        if len(v) == 0:
            continue
        v = (v - t_ref) / 1e3
        print('  %-20s %-6s first %8.2f median %8.2f last %8.2f' % (nm, lab, v.min(), np.median(v), v.max()))

fin = sel & (a[:, 120] > a[:, 110])
print('finishers:', fin.sum())
for row in a[fin]:
    print('  arrived %8.2f | loads done %8.2f | listed %8.2f | rows in %8.2f | folded %8.2f | exit %8.2f' %
          tuple((row[k] - t_ref) / 1e3 for k in (119, 120, 121, 122, 123, 111)))
