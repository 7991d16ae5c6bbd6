"""Is the C-ABI call host-bound?  Times the enqueue (host) cost of moka_forward / moka_backward
next to the GPU time of the same loop (microbench shapes)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tools.microbench as mb  # noqa: E402
from moka_b200 import _cabi  # noqa: E402
from moka_b200.layouts import VARIANT_AVT


def main():
    lib = _cabi.load()
    r = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    # re-use run_point's setup by monkeypatching: simplest is to copy the few lines we need
    import types
    res = {}
    orig_event = torch.cuda.Event
    # run the standard point for GPU time
    row = mb.run_point(VARIANT_AVT, 4, 2048, 4096, 4096, r, torch.bfloat16, 0, 50, 5)
    res['gpu'] = row
    # host enqueue cost: same calls, but time the host loop only (GPU runs behind)
    dev = 'cuda:0'
    from moka_b200.layer import FusedScalars, _make_desc
    from moka_b200.metadata import build_token_meta_avt
    from moka_b200.layouts import layout_masks
    B, S, d, M = 4, 2048, 4096, 3
    T = B * S
    masks = [m.to(dev) for m in layout_masks(VARIANT_AVT, B, S, 'c3')]
    meta = build_token_meta_avt(masks)
    sc = FusedScalars(16.0 / r, (1, 1, 1), 1.0, 0.0)
    x = torch.randn(T, d, device=dev).bfloat16(); y = torch.randn(T, d, device=dev).bfloat16()
    A = [torch.randn(r, d, device=dev).bfloat16() for _ in range(M)]
    Bw = torch.randn(d, r, device=dev).bfloat16()
    dA = [torch.zeros(r, d, device=dev) for _ in range(M)]
    dB = torch.zeros(d, r, device=dev)
    h = torch.empty(2 * T * r + T, device=dev)
    desc = _make_desc((T, d), B, S, d, r, M, torch.bfloat16, meta.Lk, sc, 0)
    ws = torch.empty(int(lib.moka_workspace_bytes(desc)), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    Ap = _cabi.ptr_array([a.data_ptr() for a in A]); dAp = _cabi.ptr_array([a.data_ptr() for a in dA])
    xp, yp, hp, bp = x.data_ptr(), y.data_ptr(), h.data_ptr(), Bw.data_ptr()
    tm, ki, kl = meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr()
    dbp, wsp, wsn = dB.data_ptr(), ws.data_ptr(), ws.numel()

    def fwd():
        lib.moka_forward(desc, xp, yp, yp, Ap, bp, tm, ki, kl, hp, 1, 0, wsp, wsn, st)

    def bwd():
        lib.moka_backward(desc, yp, xp, hp, Ap, bp, tm, ki, kl, xp, dAp, dbp, 1, 0, wsp, wsn, st)

    for name, fn in (('fwd', fwd), ('bwd', bwd)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n = 200
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res[name + '_host_enqueue_us'] = round((t1 - t0) / n * 1e6, 2)
        res[name + '_total_us'] = round((t2 - t0) / n * 1e6, 2)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
