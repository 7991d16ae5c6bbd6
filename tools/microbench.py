"""MokaLinear microbench (BASELINE config 5): adapter fwd+bwd through the C ABI, CUDA-event timed,
inputs larger than L2 rotated between iterations.  Prints one JSON line per point."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moka_b200 import _cabi  # noqa: E402
from moka_b200.layer import FusedScalars, _make_desc  # noqa: E402
from moka_b200.metadata import build_token_meta_avt, build_token_meta_vt  # noqa: E402
from moka_b200.layouts import VARIANT_AVT, VARIANT_VT, layout_masks

PEAK = 6550.4
try:
    PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass


def run_point(variant, B, S, d_in, d_out, r, dtype, flags, iters, warmup, p=0.0, nbuf=3, isolate=False):
    lib = _cabi.load()
    dev = 'cuda:0'
    M = 3 if variant == VARIANT_AVT else 2
    layout = 'c3' if variant == VARIANT_AVT else 'c2'
    masks = [m.to(dev) for m in layout_masks(variant, B, S, layout)]
    meta = build_token_meta_avt(masks) if variant == VARIANT_AVT else build_token_meta_vt(*masks)
    s = 16.0 / r
    sc = FusedScalars(s, (1, 1, 1), 1.0, p) if variant == VARIANT_AVT else FusedScalars(1.0, (s, s, s), 0.05, p)
    T = B * S
    e = 2 if dtype == torch.bfloat16 else 4
    bufs = []
    for i in range(nbuf):      # rotate buffers so that nothing stays in the 126 MB L2
        g = torch.Generator(device=dev).manual_seed(i)
        x = torch.randn(T, d_in, device=dev, generator=g).to(dtype)
        y = torch.randn(T, d_out, device=dev, generator=g).to(dtype)
        dy = torch.randn(T, d_out, device=dev, generator=g).to(dtype)
        dx = torch.randn(T, d_in, device=dev, generator=g).to(dtype)
        bufs.append((x, y, dy, dx))
    A = [(torch.rand(r, d_in, device=dev) * 2 - 1).mul_(d_in ** -0.5).to(dtype) for _ in range(M)]
    Bw = (torch.randn(d_out, r, device=dev) * 0.02).to(dtype)
    dA = [torch.zeros(r, d_in, device=dev) for _ in range(M)]
    dB = torch.zeros(d_out, r, device=dev)
    h = torch.empty(2 * T * r + T, device=dev)
    desc = _make_desc((T, d_in), B, S, d_out, r, M, dtype, meta.Lk, sc, flags)
    ws = torch.empty(int(lib.moka_workspace_bytes(desc)), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    Ap = _cabi.ptr_array([a.data_ptr() for a in A])
    dAp = _cabi.ptr_array([a.data_ptr() for a in dA])

    def fwd(i):
        x, y, dy, dx = bufs[i % nbuf]
        _cabi.check(lib.moka_forward(desc, x.data_ptr(), y.data_ptr(), y.data_ptr(), Ap, Bw.data_ptr(),
                                     meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                     h.data_ptr(), 1, i, ws.data_ptr(), ws.numel(), st), 'fwd')

    def bwd(i):
        x, y, dy, dx = bufs[i % nbuf]
        _cabi.check(lib.moka_backward(desc, dy.data_ptr(), x.data_ptr(), h.data_ptr(), Ap, Bw.data_ptr(),
                                      meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                      dx.data_ptr(), dAp, dB.data_ptr(), 1, i, ws.data_ptr(), ws.numel(), st), 'bwd')

    for i in range(warmup):
        fwd(i); bwd(i)
    torch.cuda.synchronize()
    res = {}
    for name, fn in (('fwd', fwd), ('bwd', bwd), ('fwd+bwd', lambda i: (fwd(i), bwd(i)))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / iters * 1e3       # us
    iso = {}
    if isolate:
        # per-kernel hot times: run the call with every other kernel of the tensor-core chain skipped
        # (MOKA_FLAG_DBG_SKIP bits; results are garbage, timings are what we want)
        names = ['down', 'up', 'prep', 'dgrad', 'attn_pack', 'wgrad', 'attn_fwd']
        for k, nm in enumerate(names):
            mask = (0x7f & ~(1 << k)) << 16
            desc_k = _make_desc((T, d_in), B, S, d_out, r, M, dtype, meta.Lk, sc, flags | mask)
            is_fwd = k < 2 or k == 6

            def call(i, desc_k=desc_k, is_fwd=is_fwd):
                x, y, dy, dx = bufs[i % nbuf]
                if is_fwd:
                    lib.moka_forward(desc_k, x.data_ptr(), y.data_ptr(), y.data_ptr(), Ap, Bw.data_ptr(),
                                     meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                     h.data_ptr(), 1, i, ws.data_ptr(), ws.numel(), st)
                else:
                    lib.moka_backward(desc_k, dy.data_ptr(), x.data_ptr(), h.data_ptr(), Ap, Bw.data_ptr(),
                                      meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                                      dx.data_ptr(), dAp, dB.data_ptr(), 1, i, ws.data_ptr(), ws.numel(), st)
            for i in range(3):
                call(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(iters):
                call(i)
            e1.record()
            torch.cuda.synchronize()
            iso[nm] = round(e0.elapsed_time(e1) / iters * 1e3, 2)
    delta = bool(flags & _cabi.FLAG_DELTA_OUT)      # write-only outputs: the residual read is the frozen GEMM's
    bytes_f = e * ((d_in + d_out) if delta else (d_in + 2 * d_out)) * T
    bytes_b = e * ((2 * d_in + d_out) if delta else (3 * d_in + d_out)) * T
    out = dict(variant='avt' if variant == VARIANT_AVT else 'vt', T=T, d_in=d_in, d_out=d_out, r=r,
               delta_out=delta, dtype=str(dtype).split('.')[-1], tc=bool(lib.moka_uses_tensor_cores(desc)), dropout=p,
               fwd_us=round(res['fwd'], 2), bwd_us=round(res['bwd'], 2), both_us=round(res['fwd+bwd'], 2),
               fwd_gbs=round(bytes_f / res['fwd'] / 1e3, 1), bwd_gbs=round(bytes_b / res['bwd'] / 1e3, 1),
               both_gbs=round((bytes_f + bytes_b) / res['fwd+bwd'] / 1e3, 1))
    out['frac'] = round(out['both_gbs'] / PEAK, 4)
    if iso:
        out['isolated_us'] = iso
    return out


def torch_split_gemm_us(variant, B, S, d_in, d_out, r, dtype, iters=10, warmup=3):
    """The comparator of BASELINE config 5: the same layer as plain PyTorch on the GPU - cuBLAS split GEMMs
    (one skinny GEMM per modality over all rows, the gather-free masked formulation), eager softmax attention
    batched over samples, autograd backward.  fwd+bwd microseconds (adapter only, like the fused numbers:
    the frozen W0 GEMM is excluded, y0 / dy are given)."""
    dev = 'cuda:0'
    M = 3 if variant == VARIANT_AVT else 2
    layout = 'c3' if variant == VARIANT_AVT else 'c2'
    masks = [m.to(dev) for m in layout_masks(variant, B, S, layout)]
    if variant == VARIANT_AVT:
        mod = [masks[0].view(B, S, 1).to(dtype), masks[1].view(B, S, 1).to(dtype), masks[2].view(B, S, 1).to(dtype)]
        q = masks[3].view(B, S).bool()
        pre, post, w = 16.0 / r, 1.0, 1.0
    else:
        mod = [masks[0].view(B, S, 1).to(dtype), masks[1].view(B, S, 1).to(dtype)]
        q = masks[2].view(B, S).bool()
        pre, post, w = 1.0, 16.0 / r, 0.05
    x = torch.randn(B, S, d_in, device=dev).to(dtype).requires_grad_(True)
    y0 = torch.randn(B, S, d_out, device=dev).to(dtype)
    dy = torch.randn(B, S, d_out, device=dev).to(dtype)
    A = [((torch.rand(r, d_in, device=dev) * 2 - 1) * d_in ** -0.5).to(dtype).requires_grad_(True) for _ in range(M)]
    Bw = (torch.randn(d_out, r, device=dev) * 0.02).to(dtype).requires_grad_(True)
    # every sample has the same (contiguous) question span in these layouts: keys = that slice of h_text
    idx = q[0].nonzero().flatten()
    lo, hi = int(idx[0]), int(idx[-1]) + 1

    def step():
        h = [pre * ((x * mod[m]) @ A[m].t()) for m in range(M)]
        K = h[0][:, lo:hi, :]                                   # [B, Lk, r]
        g = h[0]
        for m in range(1, M):
            p = torch.softmax(h[m] @ K.transpose(1, 2) / r ** 0.5, dim=-1)
            g = g + h[m] + mod[m] * (p @ K) * w
        y = y0 + post * (g @ Bw.t())
        y.backward(dy)
        x.grad = None
        for a in A:
            a.grad = None
        Bw.grad = None

    for _ in range(warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ranks', default='4,8,16')
    ap.add_argument('--dims', default='4096')
    ap.add_argument('--variants', default='vt,avt')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--B', type=int, default=4)
    ap.add_argument('--S', type=int, default=2048)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--flags', type=int, default=0)
    ap.add_argument('--dropout', type=float, default=0.0)
    ap.add_argument('--isolate', action='store_true')
    ap.add_argument('--baseline', action='store_true', help='also time the plain-PyTorch split-GEMM comparator')
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    dtype = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    rows = []
    for v in a.variants.split(','):
        for d in [int(t) for t in a.dims.split(',')]:
            for r in [int(t) for t in a.ranks.split(',')]:
                row = run_point(VARIANT_AVT if v == 'avt' else VARIANT_VT, a.B, a.S, d, d, r, dtype, a.flags,
                                a.iters, a.warmup, a.dropout, isolate=a.isolate)
                if a.baseline:
                    row['torch_split_gemm_us'] = round(torch_split_gemm_us(
                        VARIANT_AVT if v == 'avt' else VARIANT_VT, a.B, a.S, d, d, r, dtype), 1)
                    row['speedup_vs_torch'] = round(row['torch_split_gemm_us'] / row['both_us'], 2)
                print(json.dumps(row), flush=True)
                rows.append(row)
    if a.out:
        with open(a.out, 'w') as f:
            for row in rows:
                f.write(json.dumps(row) + '\n')


if __name__ == '__main__':
    main()
