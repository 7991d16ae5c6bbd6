"""Helpers for the -m gpu parity tests: drive libmoka.so through its C ABI on explicit tensors."""
import torch

from moka_b200 import _cabi
from moka_b200.layer import FusedScalars, _make_desc
from moka_b200.metadata import TokenMeta, build_token_meta_avt, build_token_meta_vt
from oracle import moka_oracle as mo

DEV = 'cuda:0'


def scalars_for(variant, alpha, r, w, p=0.0):
    s = alpha / r
    if variant == mo.VARIANT_AVT:
        return FusedScalars(s, (1.0, 1.0, 1.0), w, p), mo.FusedScalars(s, (1.0, 1.0, 1.0), w)
    return FusedScalars(1.0, (s, s, s), w, p), mo.FusedScalars(1.0, (s, s, s), w)


def gpu_meta(variant, masks, check=True):
    dm = [m.to(DEV) for m in masks]
    if variant == mo.VARIANT_AVT:
        return build_token_meta_avt(dm, check=check)
    return build_token_meta_vt(*dm, check=check)


def abi_forward(meta: TokenMeta, x, y0, A, Bw, sc: FusedScalars, flags=0, seed=0, offset=0, inplace=True):
    """x [B,S,d_in], y0 [B,S,d_out] on the GPU in the kernel dtype.  Returns (y, h_stash)."""
    lib = _cabi.load()
    B, S, d_in = x.shape
    d_out, r = Bw.shape
    x2 = x.reshape(B * S, d_in).contiguous()
    y0c = y0.reshape(B * S, d_out).contiguous()
    delta = bool(flags & _cabi.FLAG_DELTA_OUT)      # write-only output: y starts as NaN, y0 is not passed at all
    y = torch.full_like(y0c, float('nan')) if delta else (y0c.clone() if inplace else torch.empty_like(y0c))
    src = y if inplace else y0c
    h = torch.full((2 * B * S * r + B * S,), float('nan'), dtype=torch.float32, device=x.device)
    desc = _make_desc(x2.shape, B, S, d_out, r, len(A), x.dtype, meta.Lk, sc, flags)
    ws = torch.empty(int(lib.moka_workspace_bytes(desc)), dtype=torch.uint8, device=x.device)
    code = lib.moka_forward(desc, x2.data_ptr(), None if delta else src.data_ptr(), y.data_ptr(),
                            _cabi.ptr_array([a.data_ptr() for a in A]), Bw.data_ptr(),
                            meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                            h.data_ptr(), seed, offset, ws.data_ptr(), ws.numel(),
                            torch.cuda.current_stream().cuda_stream)
    _cabi.check(code, 'moka_forward')
    torch.cuda.synchronize()
    return y.view(B, S, d_out), h


def abi_backward(meta: TokenMeta, x, dy, h, A, Bw, sc: FusedScalars, dx0=None, flags=0, seed=0, offset=0):
    """Returns (dx (dx0 + adapter part, kernel dtype) or None, [dA fp32], dB fp32)."""
    lib = _cabi.load()
    B, S, d_in = x.shape
    d_out, r = Bw.shape
    x2 = x.reshape(B * S, d_in).contiguous()
    dy2 = dy.reshape(B * S, d_out).contiguous()
    dx = None if dx0 is None else dx0.reshape(B * S, d_in).clone().contiguous()
    if dx is not None and (flags & _cabi.FLAG_DELTA_OUT):
        dx.fill_(float('nan'))                       # write-only output: previous contents must not matter
    dA = [torch.zeros(r, d_in, dtype=torch.float32, device=x.device) for _ in A]
    dB = torch.zeros(d_out, r, dtype=torch.float32, device=x.device)
    desc = _make_desc(x2.shape, B, S, d_out, r, len(A), x.dtype, meta.Lk, sc, flags)
    nbytes = int(lib.moka_workspace_bytes(desc))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    code = lib.moka_backward(desc, dy2.data_ptr(), x2.data_ptr(), h.contiguous().data_ptr(),
                             _cabi.ptr_array([a.data_ptr() for a in A]), Bw.data_ptr(),
                             meta.tokmeta.data_ptr(), meta.kidx.data_ptr(), meta.klen.data_ptr(),
                             0 if dx is None else dx.data_ptr(),
                             _cabi.ptr_array([g.data_ptr() for g in dA]), dB.data_ptr(),
                             seed, offset, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _cabi.check(code, 'moka_backward')
    torch.cuda.synchronize()
    return (None if dx is None else dx.view(B, S, d_in)), dA, dB


def keep_mask(B, S, d_in, d_out, r, M, dtype, p, seed, offset):
    lib = _cabi.load()
    sc = FusedScalars(1.0, (1.0, 1.0, 1.0), 0.0, p)
    desc = _make_desc((B * S, d_in), B, S, d_out, r, M, dtype, 1, sc, 0)
    keep = torch.empty(B * S, d_in, dtype=torch.uint8, device=DEV)
    _cabi.check(lib.moka_dropout_keep_mask(desc, seed, offset, keep.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), 'moka_dropout_keep_mask')
    torch.cuda.synchronize()
    return keep.view(B, S, d_in)


def oracle_meta(variant, masks):
    return mo.build_token_meta_avt(masks) if variant == mo.VARIANT_AVT else mo.build_token_meta_vt(*masks)


def effective_p(p):
    """The drop probability the kernels realise (round(4096 p) / 4096): what the oracle's 1/(1-p) scale must use."""
    return float(_cabi.load().moka_effective_dropout_p(float(p))) if p > 0 else 0.0
