"""Per-shape adapter timing at the 7B projection shapes (4096->4096, 4096->11008, 11008->4096), T tokens."""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tools.microbench as mb
from moka_b200.layouts import VARIANT_AVT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # 4 = MOKA_FLAG_DELTA_OUT
for d_in, d_out in ((4096, 4096), (4096, 11008), (11008, 4096)):
    row = mb.run_point(VARIANT_AVT, B, 2048, d_in, d_out, 16, torch.bfloat16, flags, 10, 3, p, nbuf=2, isolate=True)
    print(json.dumps(row), flush=True)
