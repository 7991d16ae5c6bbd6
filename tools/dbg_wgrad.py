"""Cycle-stamp breakdown of k_wgrad_tc (MOKA_DEBUG_TIMING=1).  Run the forward with the stamps of
k_up_tc disabled by reading only after a backward-only pass."""
import ctypes, os, sys
os.environ['MOKA_DEBUG_TIMING'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from moka_b200 import _cabi
from tools.microbench import run_point
from moka_b200.layouts import VARIANT_AVT

row = run_point(VARIANT_AVT, 4, 2048, 4096, 4096, 16, torch.bfloat16, 0, 2, 1)   # last timed pass is fwd+bwd: wgrad stamps win
print(row)
lib = _cabi.load()
n = 148 * 2 * 128
buf = (ctypes.c_ulonglong * n)()
torch.cuda.synchronize()
_cabi.check(lib.moka_debug_read(buf, n), 'debug_read')
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 128).astype(np.int64)
a = a[a[:, 0] > 0][:128]
rel = a - a[:, :1]
m = lambda i: rel[:, i].mean()
print('CTAs', len(a))
for j in range(8):
    b = 8 + j * 4
    print(f'tile {j}: t_full {m(b):8.0f} d_full {m(b+1):8.0f} epilogue {m(b+2):8.0f} done {m(b+3):8.0f}   | mma: op_full {m(40+j*3):8.0f} x_full {m(41+j*3):8.0f} t_empty {m(42+j*3):8.0f}')
print(f'loop end {m(4):8.0f}  a_full {m(5):8.0f}  dA flushed {m(6):8.0f}')
