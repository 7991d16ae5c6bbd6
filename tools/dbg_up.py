"""Run one k_up_tc launch with MOKA_DEBUG_TIMING=1 and print the per-phase cycle breakdown."""
import ctypes, os, sys
os.environ['MOKA_DEBUG_TIMING'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from moka_b200 import _cabi
from tools.microbench import run_point
from moka_b200.layouts import VARIANT_AVT, VARIANT_VT

variant = VARIANT_AVT if (len(sys.argv) < 2 or sys.argv[1] == 'avt') else VARIANT_VT
row = run_point(variant, 4, 2048, 4096, 4096, 16, torch.bfloat16, 0, 2, 1)
print(row)
lib = _cabi.load()
n = 148 * 2 * 128
buf = (ctypes.c_ulonglong * n)()
torch.cuda.synchronize()
_cabi.check(lib.moka_debug_read(buf, n), 'debug_read')
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 128).astype(np.int64)
a = a[a[:, 0] > 0][:128]
rel = a - a[:, :1]
def col(i): return rel[:, i].mean(), rel[:, i].max()
print('CTAs', len(a))
print('attn start  %8.0f  (max %8.0f)' % col(1))
print('attn end    %8.0f  (max %8.0f)' % col(2))
print('prologue end%8.0f  (max %8.0f)' % col(3))
for j in range(6):
    b = 8 + j * 4
    print(f'chunk {j}: t_full {col(b)[0]:8.0f} y_full {col(b+1)[0]:8.0f} epilogue {col(b+2)[0]:8.0f} stored {col(b+3)[0]:8.0f}')
for j in range(6):
    b = 32 + j * 3
    print(f'mma {j}: b_full {col(b)[0]:8.0f} t_empty {col(b+1)[0]:8.0f} committed {col(b+2)[0]:8.0f}')
print('loop end    %8.0f  (max %8.0f)' % col(4))
