"""Per-kernel times (isolated, via the timing-aid skip flags) at the high-rank shapes: C4 (VT, r = 64, d = 5120,
4 x 4096 tokens) and the r = 64 / 128 points of the C5 sweep.   usage: python tools/rank_breakdown.py [p]"""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tools.microbench as mb
from moka_b200.layouts import VARIANT_AVT, VARIANT_VT
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
for variant, B, S, d, r in ((VARIANT_VT, 4, 4096, 5120, 64), (VARIANT_VT, 4, 2048, 5120, 64), (VARIANT_AVT, 4, 2048, 5120, 64),
                           (VARIANT_AVT, 4, 2048, 5120, 128), (VARIANT_VT, 4, 2048, 5120, 128), (VARIANT_AVT, 4, 2048, 5120, 32)):
    row = mb.run_point(variant, B, S, d, d, r, torch.bfloat16, 0, 10, 3, p, nbuf=2, isolate=True)
    print(json.dumps({k: row[k] for k in ('variant', 'T', 'd_in', 'r', 'fwd_us', 'bwd_us', 'frac', 'isolated_us')}), flush=True)
