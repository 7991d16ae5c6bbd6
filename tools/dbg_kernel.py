"""In-kernel stamp breakdown of ONE kernel of the tensor-core chain (MOKA_DEBUG_TIMING=1).
usage: python tools/dbg_kernel.py {down|up|dgrad|wgrad} [avt|vt] [r]
Runs a full fwd+bwd to populate every buffer, then launches only the selected kernel (the
MOKA_FLAG_DBG_SKIP bits veto the rest) between two device synchronisations and prints the stamps:
clock64 cycles relative to the CTA's first stamp, globaltimer ns for CTA entry/exit."""
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

kern = sys.argv[1] if len(sys.argv) > 1 else 'up'
variant = VARIANT_VT if (len(sys.argv) > 2 and sys.argv[2] == 'vt') else VARIANT_AVT
r = int(sys.argv[3]) if len(sys.argv) > 3 else 16
BARG = int(sys.argv[4]) if len(sys.argv) > 4 else 4
KID = {'down': 0, 'up': 1, 'prep': 2, 'dgrad': 3, 'attn_pack': 4, 'wgrad': 5, 'attn_fwd': 6}[kern]
lib = _cabi.load()
dev = 'cuda:0'
B, S, d = BARG, 2048, 4096
T = B * S
M = 3 if variant == VARIANT_AVT else 2
masks = [m.to(dev) for m in layout_masks(variant, B, S, 'c3' if variant == VARIANT_AVT else 'c2')]
meta = build_token_meta_avt(masks) if variant == VARIANT_AVT else build_token_meta_vt(*masks)
s = 16.0 / r
PDROP = float(os.environ.get('MOKA_DBG_P', '0'))
sc = FusedScalars(s, (1, 1, 1), 1.0, PDROP) if variant == VARIANT_AVT else FusedScalars(1.0, (s, s, s), 0.05, PDROP)
x = torch.randn(T, d, device=dev).bfloat16(); y = torch.randn(T, d, device=dev).bfloat16()
dy = torch.randn(T, d, device=dev).bfloat16(); dx = torch.randn(T, d, device=dev).bfloat16()
A = [(torch.randn(r, d, device=dev) * d ** -0.5).bfloat16() for _ in range(M)]
Bw = (torch.randn(d, r, device=dev) * 0.02).bfloat16()
dA = [torch.zeros(r, d, device=dev) for _ in range(M)]
dB = torch.zeros(d, r, device=dev)
h = torch.empty(2 * T * r + T, device=dev)
st = torch.cuda.current_stream().cuda_stream
Ap = _cabi.ptr_array([a.data_ptr() for a in A]); dAp = _cabi.ptr_array([a.data_ptr() for a in dA])


def run(flags):
    desc = _make_desc((T, d), B, S, d, r, M, torch.bfloat16, meta.Lk, sc, flags)
    ws = run.ws
    if ws is None:
        ws = run.ws = torch.empty(int(lib.moka_workspace_bytes(desc)), dtype=torch.uint8, device=dev)
    if flags == 0 or KID < 2 or KID == 6:
        _cabi.check(lib.moka_forward(desc, x.data_ptr(), y.data_ptr(), y.data_ptr(), Ap, Bw.data_ptr(),
                                     meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                     h.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), st), 'fwd')
    if flags == 0 or 2 <= KID < 6:
        _cabi.check(lib.moka_backward(desc, dy.data_ptr(), x.data_ptr(), h.data_ptr(), Ap, Bw.data_ptr(),
                                      meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                      dx.data_ptr(), dAp, dB.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), st), 'bwd')


run.ws = None
for _ in range(3):
    run(0)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
only = (0x7f & ~(1 << KID)) << 16
for _ in range(2):
    flush.zero_()                      # evict L2
    torch.cuda.synchronize()
    run(only)
    torch.cuda.synchronize()
n = 148 * 4 * 128
buf = (ctypes.c_ulonglong * n)()
_cabi.check(lib.moka_debug_read(buf, n), 'debug_read')
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 128).astype(np.int64)
ES = {'down': 100, 'attn_fwd': 102, 'up': 104, 'prep': 106, 'dgrad': 108, 'attn_pack': 110, 'wgrad': 112}[kern]
a = a[a[:, ES] > 0] if (a[:, ES] > 0).any() else a[a[:, 0] > 0]
print(kern, 'CTAs', len(a))
if (a[:, ES] > 0).any():
    t0 = a[:, ES].min()
    ent = a[:, ES] - t0
    ext = a[:, ES + 1] - t0
    print('globaltimer ns: entry min/med/max %d %d %d | exit min/med/max %d %d %d' %
          (ent.min(), np.median(ent), ent.max(), ext.min(), np.median(ext), ext.max()))
rel = a - a[:, :1]
names = {
    'down': {1: 'prologue done', 2: 'after pdl_wait', 3: 'first stage landed', 4: 'last mma issued', 5: 'tmem_full seen', 6: 'exit'},
    'up': {6: 'after pdl_wait (rel worker start)', 1: 'B/h staged', 2: 'attention done', 3: 'A operand staged', 4: 'loop end', 5: 'exit'},
    'wgrad': {1: 'entry (rel worker loop start)', 4: 'loop end', 5: 'a_full', 6: 'dA flushed', 7: 'exit'},
    'dgrad': {2: 'G^T staged (mma thread)', 3: 'first stage landed', 4: 'last mma issued', 5: 'dg_full seen', 6: 'exit'},
}.get(kern, {})
for slot, nm in names.items():
    v = rel[:, slot]
    print(f'{nm:36s} mean {v.mean():9.0f}  min {v.min():9.0f}  max {v.max():9.0f}   (cycles)')
if kern == 'up':
    for j in range(6):
        b = 8 + j * 4
        m = lambda i: rel[:, i].mean()
        print(f'chunk {j}: t_full {m(b):8.0f} y_full {m(b+1):8.0f} epilogue {m(b+2):8.0f} stored {m(b+3):8.0f} | mma b_full {m(32+j*3):8.0f} t_empty {m(33+j*3):8.0f} committed {m(34+j*3):8.0f}')
if kern == 'wgrad':
    for j in range(8):
        b = 8 + j * 4
        m = lambda i: rel[:, i].mean()
        print(f'tile {j}: t_full {m(b):8.0f} d_full {m(b+1):8.0f} epilogue {m(b+2):8.0f} done {m(b+3):8.0f} | mma op_full {m(40+j*3):8.0f} x_full {m(41+j*3):8.0f} t_empty {m(42+j*3):8.0f}')

if kern == 'dgrad':
    for j in range(8):
        m = lambda i: rel[:, i].mean()
        print(f'stage {j}: mma full {m(8+j*3):8.0f} d2_empty {m(9+j*3):8.0f} issued {m(10+j*3):8.0f} | workers d2_full {m(32+j*2):8.0f} epilogue done {m(33+j*2):8.0f}')

if kern == 'dgrad':
    m = lambda i: rel[:, i].mean()
    print(f'stage 2 detail: d2_empty {m(15):8.0f} fence {m(48):8.0f} mmas issued {m(49):8.0f} commit1 {m(50):8.0f} commit2 {m(16):8.0f}')

if kern == 'down':
    for j in range(8):
        m = lambda i: rel[:, i].mean()
        print(f'stage {j}: mma full {m(8+j*3):8.0f} ready {m(9+j*3):8.0f} committed {m(10+j*3):8.0f} | mask full {m(32+j*3):8.0f} masked {m(33+j*3):8.0f} arrived {m(34+j*3):8.0f}')
