"""world_size-2 gloo tests (CPU) of the data-parallel engine: flat buckets, gradient averaging,
sharded AdamW + all-gather for ZeRO stage 0/1/2, against a single-process torch.optim.AdamW on the
averaged gradient.  The optimizer update itself is injected (the product uses the CUDA kernel
moka_adamw_step, which refuses CPU tensors)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from moka_b200.dp import AdamWConfig, AdapterDataParallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ref_adamw(n, master, m, v, grad, out, lr, b1, b2, eps, wd, step, gscale):
    """torch.optim.AdamW math on flat fp32 tensors (test-side stand-in for the CUDA kernel)."""
    g = grad[:n] * gscale
    master[:n].mul_(1 - lr * wd)
    m[:n].mul_(b1).add_(g, alpha=1 - b1)
    v[:n].mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v[:n].sqrt() / (bc2 ** 0.5)).add_(eps)
    master[:n].addcdiv_(m[:n], denom, value=-lr / bc1)
    out[:n].copy_(master[:n].to(out.dtype))


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.frozen = nn.Parameter(torch.randn(5, 5, generator=g), requires_grad=False)
        self.a = nn.ParameterDict({f'lora_A{i}': nn.Parameter(torch.randn(4, 33, generator=g)) for i in range(3)})
        self.b = nn.ParameterDict({'lora_B0': nn.Parameter(torch.randn(17, 4, generator=g))})


def _grads_for(rank, shapes, step):
    g = torch.Generator().manual_seed(100 + 7 * rank + step)
    return [torch.randn(s, generator=g) for s in shapes]


def _worker(rank, world, port, stage, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _Toy()
        eng = AdapterDataParallel(model, AdamWConfig(lr=1e-2, weight_decay=0.01), zero_stage=stage, num_buckets=2,
                                  update_fn=_ref_adamw, overlap=False)
        params = [p for p in model.parameters() if p.requires_grad]
        shapes = [p.shape for p in params]
        for step in range(3):
            for p, g in zip(params, _grads_for(rank, shapes, step)):
                p.main_grad.add_(g)                       # what moka_backward does
                p._moka_dp_ready()
            eng.step()
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            ret['flat'] = flat.clone()
            ret['same'] = all(torch.equal(gathered[0], t) for t in gathered)
            ret['views'] = all(p.data_ptr() >= eng.flat_param.data_ptr() for p in params)
            ret['zeroed'] = float(eng.flat_grad.abs().max()) == 0.0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('stage', [0, 1, 2])
def test_two_rank_step_matches_single_process_adamw(stage):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, stage, ret), nprocs=world, join=True)
    assert ret['same'], 'ranks diverged'
    assert ret['views'] and ret['zeroed']
    # single-process reference: torch.optim.AdamW on the rank-averaged gradients
    model = _Toy()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    shapes = [p.shape for p in params]
    for step in range(3):
        gs = [_grads_for(r, shapes, step) for r in range(world)]
        for i, p in enumerate(params):
            p.grad = sum(g[i] for g in gs) / world
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in params])
    assert torch.allclose(ret['flat'], ref, rtol=1e-5, atol=1e-6)


def test_single_process_engine_layout():
    model = _Toy()
    eng = AdapterDataParallel(model, AdamWConfig(), zero_stage=2, num_buckets=3, update_fn=_ref_adamw)
    assert eng.world == 1 and eng.trainable_numel() == 3 * 4 * 33 + 17 * 4
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    assert names == eng.names and all('lora_' in n for n in names)
    for p in (q for q in model.parameters() if q.requires_grad):
        assert p.main_grad.dtype == torch.float32 and p.main_grad.shape == p.shape
        assert p.grad is None
    assert not hasattr(model.frozen, 'main_grad')
    with pytest.raises(ValueError):
        AdapterDataParallel(model, zero_stage=3)


def _worker_accum(rank, world, port, stage, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _Toy()
        eng = AdapterDataParallel(model, AdamWConfig(lr=1e-2, weight_decay=0.01, max_grad_norm=1.0), zero_stage=stage,
                                  num_buckets=2, update_fn=_ref_adamw, overlap=False)
        params = [p for p in model.parameters() if p.requires_grad]
        shapes = [p.shape for p in params]
        accum = 2
        for step in range(3):
            for mb in range(accum):
                with eng.accumulate(mb, accum):
                    for p, g in zip(params, _grads_for(rank, shapes, 10 * step + mb)):
                        p.main_grad.add_(g)
                        p._moka_dp_ready()
            eng.step(grad_scale=1.0 / accum)
        # a second un-guarded backward after the collectives were armed must be refused, not silently dropped
        eng2 = AdapterDataParallel(_Toy(), AdamWConfig(), zero_stage=stage, num_buckets=2, update_fn=_ref_adamw,
                                   overlap=False)
        eng2.overlap = True                               # arm the launch-on-ready path (gloo, same stream)
        ps = [p for p in eng2.model.parameters() if p.requires_grad]
        for p in ps:
            p._moka_dp_ready()
        refused = False
        try:
            ps[0]._moka_dp_ready()
        except RuntimeError:
            refused = True
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        if rank == 0:
            ret['flat'] = flat.clone()
            ret['refused'] = refused
            ret['norm'] = float(eng.last_grad_norm)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('stage', [0, 1, 2])
def test_two_rank_gradient_accumulation_and_clipping(stage):
    """two micro-batches per step (VisualText/shell/train.sh:72) under no_sync(), mean over micro-batches
    and ranks, global-norm clipping at 1.0 -- against single-process torch.optim.AdamW + clip_grad_norm_."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_accum, args=(world, port, stage, ret), nprocs=world, join=True)
    assert ret['refused']
    model = _Toy()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    shapes = [p.shape for p in params]
    norm = None
    for step in range(3):
        for i, p in enumerate(params):
            p.grad = sum(_grads_for(r, shapes, 10 * step + mb)[i] for r in range(world) for mb in range(2)) / (world * 2)
        norm = torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in params])
    assert torch.allclose(ret['flat'], ref, rtol=1e-5, atol=1e-6)
    assert abs(ret['norm'] - float(norm)) <= 1e-4 * float(norm)
