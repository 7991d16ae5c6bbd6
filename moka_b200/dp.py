"""Data-parallel step for the adapter parameters: the ZeRO-1/2-equivalent the reference gets
from HF Trainer + DeepSpeed (VisualText/zero_stage{1,2}_config.json, AudioVisualText/deepspeed/
stage2-offload.json).  None of that stack is installed here, so this module
provides the same semantics natively:

* every trainable (``lora_``) parameter lives in ONE flat working buffer (model dtype) and has a
  view ``main_grad`` into ONE flat fp32 gradient buffer; ``moka_backward`` accumulates straight
  into those views (no per-parameter ``.grad`` tensors, no flatten/copy before the collective);
* the flat gradient is cut into buckets in backward order; a bucket's collective is launched on
  a side stream as soon as its last layer has run backward (overlap with the remaining backward);
* stage 2: reduce-scatter(mean) -> each rank updates its 1/N shard with fused AdamW
  (``moka_adamw_step``: fp32 master + moments, rounded working copy) -> all-gather;
  stage 1: all-reduce(mean) then the same sharded update + all-gather; stage 0: all-reduce and a
  replicated update (plain DDP).  All three produce identical parameters.
* world size 1 degenerates to the fused AdamW alone.

Collectives are torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU
tests).  The path has no other exchange step: samples are independent.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _cabi


@dataclass
class AdamWConfig:
    lr: float = 1e-4                 # VisualText/shell/train.sh:77
    betas: Tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 0.0        # VisualText/zero_stage2_config.json:23-34 ("auto" -> trainer 0.0)


def _cuda_adamw(n, master, m, v, grad, out, lr, b1, b2, eps, wd, step, gscale):
    lib = _cabi.load()
    if not master.is_cuda:
        raise _cabi.MokaError('moka_adamw_step runs on CUDA only (tests inject their own update_fn for gloo)')
    dt = _cabi.MOKA_BF16 if out.dtype == torch.bfloat16 else _cabi.MOKA_F32
    code = lib.moka_adamw_step(n, master.data_ptr(), m.data_ptr(), v.data_ptr(), grad.data_ptr(),
                               out.data_ptr(), dt, lr, b1, b2, eps, wd, step, gscale,
                               torch.cuda.current_stream(master.device).cuda_stream)
    _cabi.check(code, 'moka_adamw_step')


class AdapterDataParallel:
    """Flat-bucket data parallel engine over the trainable parameters of ``model``."""

    def __init__(self, model: nn.Module, optim: AdamWConfig = AdamWConfig(), zero_stage: int = 2,
                 num_buckets: int = 4, process_group=None, update_fn: Optional[Callable] = None,
                 overlap: bool = True):
        if zero_stage not in (0, 1, 2):
            raise ValueError('zero_stage must be 0, 1 or 2 (ZeRO-3 / offload are out of scope: the frozen base fits)')
        self.model, self.cfg, self.stage = model, optim, zero_stage
        self.pg = process_group
        self.world = dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(self.pg) if self.world > 1 else 0
        self.update_fn = update_fn or _cuda_adamw
        self.step_count = 0
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError('no trainable parameters')
        self.names = [n for n, _ in named]
        params = [p for _, p in named]
        dev, dtype = params[0].device, params[0].dtype
        for n, p in named:
            if p.dtype != dtype or p.device != dev:
                raise ValueError(f'{n}: all trainable parameters must share one dtype/device')
        self.device, self.dtype = dev, dtype
        self.overlap = overlap and dev.type == 'cuda' and self.world > 1

        # ---- bucket layout: parameters in module order, buckets are contiguous ranges; backward
        # visits them last-to-first, so bucket K-1 completes first.
        total = sum(p.numel() for p in params)
        num_buckets = max(1, min(num_buckets, len(params)))
        target = (total + num_buckets - 1) // num_buckets
        align = self.world * 64
        self.buckets: List[Dict] = []
        cur, cur_n = [], 0
        for p in params:
            cur.append(p); cur_n += p.numel()
            if cur_n >= target and len(self.buckets) < num_buckets - 1:
                self.buckets.append(dict(params=cur)); cur, cur_n = [], 0
        if cur:
            self.buckets.append(dict(params=cur))
        off = 0
        for b in self.buckets:
            n = sum(p.numel() for p in b['params'])
            padded = (n + align - 1) // align * align
            b.update(offset=off, numel=n, padded=padded, shard=padded // self.world, pending=0)
            off += padded
        self.total_padded = off
        self.flat_param = torch.zeros(off, dtype=dtype, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.param_bucket: Dict[int, int] = {}
        for bi, b in enumerate(self.buckets):
            o = b['offset']
            for p in b['params']:
                n = p.numel()
                self.flat_param[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + n].view(p.shape)
                p.main_grad = self.flat_grad[o:o + n].view(p.shape)
                p.grad = None
                self.param_bucket[id(p)] = bi
                p._moka_dp_ready = self._make_ready_cb(bi)
                o += n
        # ---- optimizer shards (stage 0: the whole vector on every rank)
        shard_total = sum(b['padded'] if self.stage == 0 or self.world == 1 else b['shard'] for b in self.buckets)
        self.shard_numel = shard_total
        self.master = torch.empty(shard_total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(shard_total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(shard_total, dtype=torch.float32, device=dev)
        self.shard_grad = torch.zeros(shard_total, dtype=torch.float32, device=dev)
        self.shard_out = torch.zeros(shard_total, dtype=dtype, device=dev)
        so = 0
        for b in self.buckets:
            n = b['padded'] if self.stage == 0 or self.world == 1 else b['shard']
            lo = b['offset'] + (0 if n == b['padded'] else self.rank * b['shard'])
            b.update(shard_off=so, shard_len=n, own_lo=lo)
            self.master[so:so + n].copy_(self.flat_param[lo:lo + n].float())
            so += n
        self.comm_stream = torch.cuda.Stream(device=dev) if self.overlap else None
        self._handles: List = []
        self._launched = [False] * len(self.buckets)
        self.reset_pending()

    # ------------------------------------------------------------------ bookkeeping
    def reset_pending(self):
        for b in self.buckets:
            b['pending'] = len(b['params'])
        self._launched = [False] * len(self.buckets)

    def _make_ready_cb(self, bi: int):
        def cb():
            b = self.buckets[bi]
            b['pending'] -= 1
            if b['pending'] == 0 and self.overlap:
                self._launch_bucket(bi)
        return cb

    def trainable_numel(self) -> int:
        return sum(b['numel'] for b in self.buckets)

    def zero_grad(self):
        self.flat_grad.zero_()
        self.reset_pending()

    # ------------------------------------------------------------------ collectives
    def _reduce_bucket(self, bi: int):
        b = self.buckets[bi]
        g = self.flat_grad[b['offset']:b['offset'] + b['padded']]
        out = self.shard_grad[b['shard_off']:b['shard_off'] + b['shard_len']]
        if self.world == 1:
            return
        backend = dist.get_backend(self.pg)
        if self.stage == 2 and backend != 'gloo':
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            # stage 0/1 (and gloo, which has no reduce-scatter): all-reduce then keep the owned slice
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
            g.mul_(1.0 / self.world)
            if self.stage != 0:
                out.copy_(g[self.rank * b['shard']:(self.rank + 1) * b['shard']])

    def _launch_bucket(self, bi: int):
        if self._launched[bi]:
            return
        self._launched[bi] = True
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self._reduce_bucket(bi)
        else:
            self._reduce_bucket(bi)

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        """Gradient sync (whatever was not already overlapped) + sharded AdamW + parameter all-gather."""
        for bi in reversed(range(len(self.buckets))):
            self._launch_bucket(bi)
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        self.step_count += 1
        c = self.cfg
        if self.world == 1 or self.stage == 0:
            grad = self.flat_grad                       # full vector, already averaged
            out = self.flat_param
        else:
            grad, out = self.shard_grad, self.shard_out
        self.update_fn(self.shard_numel, self.master, self.exp_avg, self.exp_avg_sq, grad, out,
                       c.lr, c.betas[0], c.betas[1], c.eps, c.weight_decay, self.step_count, grad_scale)
        if self.world > 1 and self.stage != 0:
            for b in self.buckets:
                full = self.flat_param[b['offset']:b['offset'] + b['padded']]
                mine = self.shard_out[b['shard_off']:b['shard_off'] + b['shard_len']]
                dist.all_gather_into_tensor(full, mine, group=self.pg)
        self.zero_grad()

    # ------------------------------------------------------------------ state (adapter-only checkpoints)
    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(master=self.master.clone(), exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(),
                    step=torch.tensor(self.step_count))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.master.copy_(sd['master']); self.exp_avg.copy_(sd['exp_avg']); self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self.step_count = int(sd['step'])
