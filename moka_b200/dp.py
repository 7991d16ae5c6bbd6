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
    # "gradient_clipping": "auto" in both DeepSpeed configs (zero_stage2_config.json:14) = HF Trainer's
    # max_grad_norm (1.0 by default); 0 disables.  The global norm is taken over the rank-averaged gradient.
    max_grad_norm: float = 0.0


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
        self._sync = True                # False inside no_sync(): micro-batches accumulate locally, nothing is reduced
        self._micro_batches = 0          # backward passes since the last step()
        self.last_grad_norm: Optional[torch.Tensor] = None
        self.reset_pending()

    # ------------------------------------------------------------------ bookkeeping
    def reset_pending(self):
        for b in self.buckets:
            b['pending'] = len(b['params'])
        self._launched = [False] * len(self.buckets)

    def _make_ready_cb(self, bi: int):
        def cb():
            if not self._sync:
                return                   # accumulating micro-batch: the gradient stays local until the last one
            if self._launched[bi]:
                # a second backward arrived after this bucket's collective was launched: its contribution would be
                # dropped (stage 1/2) or added un-reduced (stage 0).  Accumulate under no_sync() instead.
                raise RuntimeError('gradient accumulation without no_sync(): run every micro-batch but the last '
                                   'inside `with engine.no_sync():` (or use engine.accumulate(i, n))')
            b = self.buckets[bi]
            b['pending'] -= 1
            if b['pending'] == 0 and self.overlap:
                self._launch_bucket(bi)
        return cb

    def no_sync(self):
        """Context manager for gradient accumulation (``--gradient_accumulation_steps 2``,
        VisualText/shell/train.sh:72): backward passes inside it only accumulate into the flat fp32 gradient;
        the bucket collectives are armed by the first backward outside it and ``step()`` reduces whatever is
        left.  Pass ``grad_scale = 1 / accumulation_steps`` to ``step()`` for the mean over micro-batches."""
        eng = self

        class _NoSync:
            def __enter__(self_inner):
                self_inner.prev = eng._sync
                eng._sync = False

            def __exit__(self_inner, *exc):
                eng._sync = self_inner.prev
                return False
        return _NoSync()

    def accumulate(self, micro_step: int, accumulation_steps: int):
        """``with engine.accumulate(i, n):`` = no_sync() for every micro-batch but the last of the n."""
        import contextlib
        return self.no_sync() if (micro_step + 1) % accumulation_steps else contextlib.nullcontext()

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
            # SUM here; the 1 / world of the mean rides in the optimizer's grad_scale (no extra pass over the bucket)
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
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
    def _grads_are_sums(self) -> bool:
        """True when the reduced gradient still has to be divided by the world size (all-reduce SUM paths)."""
        if self.world == 1:
            return False
        return not (self.stage == 2 and dist.get_backend(self.pg) != 'gloo')

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        """Gradient sync (whatever was not already overlapped) + global-norm clipping + sharded AdamW +
        parameter all-gather (per bucket, overlapped with the update of the next bucket)."""
        if not self._sync:
            raise RuntimeError('step() inside no_sync(): leave the context before the optimizer step')
        for bi in reversed(range(len(self.buckets))):
            self._launch_bucket(bi)
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        self.step_count += 1
        c = self.cfg
        sharded = self.world > 1 and self.stage != 0
        grad, out = (self.shard_grad, self.shard_out) if sharded else (self.flat_grad, self.flat_param)
        scale = grad_scale * (1.0 / self.world if self._grads_are_sums() else 1.0)
        if c.max_grad_norm and c.max_grad_norm > 0:
            # global l2 norm of the (scaled, rank-averaged) gradient: shards partition the vector, so the squared
            # norms add across ranks.  Stays on the device: the clip coefficient is applied as a tensor multiply.
            sq = grad.float().pow(2).sum()
            if sharded:
                dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=self.pg)
            norm = sq.sqrt() * abs(scale)
            self.last_grad_norm = norm
            coef = torch.clamp(c.max_grad_norm / (norm + 1e-6), max=1.0)
            grad.mul_(coef)
        if not sharded:
            self.update_fn(self.shard_numel, self.master, self.exp_avg, self.exp_avg_sq, grad, out,
                           c.lr, c.betas[0], c.betas[1], c.eps, c.weight_decay, self.step_count, scale)
        else:
            # bucket by bucket: the all-gather of bucket i (side stream) overlaps the update of bucket i + 1
            for b in self.buckets:
                so, n = b['shard_off'], b['shard_len']
                self.update_fn(n, self.master[so:so + n], self.exp_avg[so:so + n], self.exp_avg_sq[so:so + n],
                               grad[so:so + n], out[so:so + n], c.lr, c.betas[0], c.betas[1], c.eps,
                               c.weight_decay, self.step_count, scale)
                full = self.flat_param[b['offset']:b['offset'] + b['padded']]
                mine = out[so:so + n]
                if self.comm_stream is not None:
                    self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(self.comm_stream):
                        dist.all_gather_into_tensor(full, mine, group=self.pg)
                else:
                    dist.all_gather_into_tensor(full, mine, group=self.pg)
            if self.comm_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        self.zero_grad()

    # ------------------------------------------------------------------ state (adapter-only checkpoints)
    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(master=self.master.clone(), exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(),
                    step=torch.tensor(self.step_count))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.master.copy_(sd['master']); self.exp_avg.copy_(sd['exp_avg']); self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self.step_count = int(sd['step'])
