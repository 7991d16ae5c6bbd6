#!/usr/bin/env python
"""bench.py -- Llama-2-7B MokA fine-tune step throughput on B200 (tokens/s, whole job).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (N=1 plain, N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own layers on the host cores (oracle/_ref)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what every key means.
A "step" = forward + backward + adapter-gradient sync + sharded AdamW of one micro-batch of
synthetic image/audio/text token sequences through a frozen, randomly initialised Llama-2-7B whose
224 projections are MokaLinear layers.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] shape (per GPU): the metric's "Llama-2-7B MokA r=16", 3 modalities
    # micro-batch 16 x 2048 tokens per GPU (fits without gradient checkpointing; measured round 1: the step rate
    # is flat from 4 to 16 sequences, the per-call launch ramps of the adapter kernels amortise better)
    'llama2-7b-avt-m3-r16-s2048': dict(variant='avt', model='7b', r=16, M=3, S=2048, B=16, layout='c3', w=1.0),
    # BASELINE.json configs[1]
    'llama2-7b-vt-m2-r8-s2048': dict(variant='vt', model='7b', r=8, M=2, S=2048, B=4, layout='c2', w=0.05),
    # BASELINE.json configs[3] (13B, r=64, seq 4096)
    'llama2-13b-vt-m2-r64-s4096': dict(variant='vt', model='13b', r=64, M=2, S=4096, B=4, layout='c4', w=0.05),
    # smoke-sized
    'tiny-avt': dict(variant='avt', model='tiny', r=16, M=3, S=512, B=2, layout='c3', w=1.0),
}
DEFAULT_WORKLOAD = 'llama2-7b-avt-m3-r16-s2048'


def metric_name(wl) -> str:
    """BASELINE.json's metric; names the model the workload really runs (13B for configs[3])."""
    model = {'7b': 'Llama-2-7B', '13b': 'Llama-2-13B'}.get(wl['model'], wl['model'])
    return f'tokens/sec {model} MokA fine-tune step'


def peak_hbm():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7),
                              ('sw_power_cap', 8)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


def _pct(xs):
    xs = sorted(xs)
    if not xs:
        return [0.0, 0.0, 0.0]
    return [xs[int(0.1 * (len(xs) - 1))], xs[len(xs) // 2], xs[int(0.9 * (len(xs) - 1))]]


def ncu_traffic(wl, d):
    """roofline.traffic: measured DRAM bytes of one fwd+bwd call pair at the dominant (hidden -> hidden) shape of this
    workload, looked up in profiles/r02_ncu_traffic.json by (T, d_in, d_out, r, M, p); None if never captured."""
    key = f"T={wl['B'] * wl['S']},d_in={d},d_out={d},r={wl['r']},M={wl['M']},p={wl['dropout']}"
    try:
        rec = json.load(open(os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')))[key]
        alg = 2 * wl['B'] * wl['S'] * 7 * d
        return dict(traffic=rec['traffic_bytes'], traffic_shape=f'one fwd+bwd call pair, {key} (ncu: {rec["source"]})',
                    traffic_algorithmic=alg)
    except Exception:
        return dict(traffic=None, traffic_shape=f'no ncu capture for {key}')


def kernel_path_name(wl, d, flags):
    """Which kernels serve the step's square projections, as the library itself reports it."""
    import torch
    from moka_b200 import _cabi
    from moka_b200.layer import FusedScalars, _make_desc
    sc = FusedScalars(1.0, (1.0, 1.0, 1.0), 1.0, wl['dropout'])
    desc = _make_desc((wl['B'] * wl['S'], d), wl['B'], wl['S'], d, wl['r'], wl['M'], torch.bfloat16, 128, sc, flags)
    return 'tcgen05+TMA' if _cabi.load().moka_uses_tensor_cores(desc) else 'warp-FMA'


# --------------------------------------------------------------------------- CPU baseline (the reference itself)
_MODEL_DIMS = {'7b': (4096, 11008, 32), '13b': (5120, 13824, 40), 'tiny': (256, 512, 2)}


def _reference_layers(wl, shapes, dt, device='cpu'):
    """One reference MokA layer per projection shape.  kind 'reference': the UNMODIFIED reference modules from
    oracle/_ref (vendored copy of AudioVisualText/peft_hyper and VisualText/modified_peft, see
    oracle/vendor_ref.py) -- lora.py:277-532 / layer.py:389-685, lora_dropout as configured, train mode.
    kind 'port': the oracle's op-by-op restatement (only when oracle/_ref is absent)."""
    import torch
    from oracle import moka_oracle as mo
    from oracle import ref_loader as rl
    r, M = wl['r'], wl['M']
    avt = wl['variant'] == 'avt'
    if rl.available():
        Lin = rl.import_avt_linear() if avt else rl.import_vt_linear()
        layers = []
        for d_in, d_out in shapes:
            if avt:
                l = rl.make_avt(Lin, d_in, d_out, r, M, 16, wl['w'], dtype=dt, dropout=wl['dropout'], device=device)
                torch.nn.init.normal_(l.lora_B0.weight, std=0.02)
            else:
                l = rl.make_vt(Lin, d_in, d_out, r, 16, wl['w'], dtype=dt, dropout=wl['dropout'], device=device)
                torch.nn.init.normal_(l.lora_B['text'].weight, std=0.02)
            l.train()
            layers.append(l)
        return 'reference', layers
    gen = torch.Generator().manual_seed(0)
    layers = []
    for d_in, d_out in shapes:
        W0 = (torch.randn(d_out, d_in, generator=gen) * 0.02).to(dt).to(device)
        A = [(torch.rand(r, d_in, generator=gen) - 0.5).to(dt).to(device).requires_grad_(True) for _ in range(M)]
        Bw = (torch.randn(d_out, r, generator=gen) * 0.02).to(dt).to(device).requires_grad_(True)
        if avt:
            layers.append(lambda x, masks, W0=W0, A=A, Bw=Bw: mo.avt_reference_form(
                x, W0, A, Bw, masks, lora_alpha=16, blc_weight=wl['w']))
        else:
            layers.append(lambda x, masks, W0=W0, A=A, Bw=Bw: mo.vt_reference_form(
                x, W0, A[0], A[1], Bw, masks[0], masks[1], masks[2], lora_alpha=16, attn_weight=wl['w']))
    return 'port', layers


def _call_reference(layer, x, masks, avt, kind):
    if kind == 'port':
        return layer(x, masks)
    return layer(x, masks) if avt else layer(x, masks[0], masks[1], masks[2])


def cpu_reference_leg(wl: dict, seconds_budget: float = 25.0, steps: int = 5, warmup: int = 2,
                      dtypes=('f32', 'bf16')):
    """Times the reference's own CPU implementation of the path on the host cores (all of them).
    Sample: the 7 MokaLinear projections (frozen GEMM included, autograd backward) of ONE decoder layer on ONE
    sequence of S tokens, `warmup` untimed repetitions then the median of `steps` (>= 5 unless the time budget
    runs out), in fp32 and in bf16 (the dtype the reference trains in).  tokens/s = S / (num_layers * t_layer):
    an UPPER BOUND on what a whole CPU step could reach (attention, norms, lm_head, optimizer left out)."""
    import torch
    from moka_b200.layouts import VARIANT_AVT, VARIANT_VT, layout_masks
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    S = wl['S']
    avt = wl['variant'] == 'avt'
    hidden, inter, layers = _MODEL_DIMS[wl['model']]
    shapes = [(hidden, hidden)] * 4 + [(hidden, inter)] * 2 + [(inter, hidden)]
    masks = layout_masks(VARIANT_AVT if avt else VARIANT_VT, 1, S, wl['layout'])
    per = {}
    kind = 'port'
    for tag in dtypes:
        dt = torch.float32 if tag == 'f32' else torch.bfloat16
        kind, mods = _reference_layers(wl, shapes, dt)
        gen = torch.Generator().manual_seed(0)
        xs = [torch.randn(1, S, d_in, generator=gen).to(dt) for d_in, _ in shapes]

        def one_layer():
            t0 = time.perf_counter()
            for l, x0 in zip(mods, xs):
                x = x0.clone().requires_grad_(True)
                y = _call_reference(l, x, masks, avt, kind)
                y.backward(torch.ones_like(y))
            return time.perf_counter() - t0

        t_start = time.perf_counter()
        for _ in range(warmup):
            one_layer()
        times = []
        for i in range(max(1, steps)):
            times.append(one_layer())
            if i >= 4 and time.perf_counter() - t_start > seconds_budget / len(dtypes):
                break
        per[tag] = dict(t_layer_s=statistics.median(times), reps=len(times), warmup=warmup)
        del mods, xs
    best = min(per, key=lambda k: per[k]['t_layer_s'])
    t_layer = per[best]['t_layer_s']
    value = S / (layers * t_layer)
    what = ('the UNMODIFIED reference layers (oracle/_ref: ' +
            ('peft_hyper.tuners.lora.Linear' if avt else 'peft.tuners.lora.layer.Linear') + ')'
            if kind == 'reference' else "the oracle's op-by-op port (oracle/_ref absent)")
    sample = (f'{what}: 7 MokaLinear fwd+bwd (frozen GEMM included) of one decoder layer, 1 sequence of {S} tokens, '
              f"lora_dropout {wl['dropout']}, " +
              ', '.join(f"{k}: median {v['t_layer_s']:.3f} s/layer over {v['reps']} reps after {v['warmup']} warm-ups"
                        for k, v in per.items()) +
              f'; value = the faster dtype ({best}); tokens/s = S / ({layers} layers x t_layer), an upper bound on a '
              f'CPU step (attention, norms, lm_head, optimizer excluded)')
    return dict(value=value, unit='tokens/s', cores=threads, kind=kind, dtype=best, sample=sample,
                per_dtype={k: dict(tokens_per_s=S / (layers * v['t_layer_s']), **v) for k, v in per.items()}), t_layer


def gpu_comparators(wl, d, device, iters=5, warmup=2):
    """BASELINE.md section 4 comparators on the SAME GPU at the step's most frequent projection shape: the
    unmodified reference layer run eagerly on CUDA (its own Python loops and host syncs), the plain-PyTorch
    split-GEMM formulation (cuBLAS, gather-free) and this repo's MokaLinear module -- layer fwd+bwd, microseconds."""
    import torch
    import tools.microbench as mb
    from moka_b200.layouts import VARIANT_AVT, VARIANT_VT, layout_masks
    from moka_b200.layer import MokaLinearAVT, MokaLinearVT
    avt = wl['variant'] == 'avt'
    variant = VARIANT_AVT if avt else VARIANT_VT
    B, S, r, M = wl['B'], wl['S'], wl['r'], wl['M']
    bf = torch.bfloat16
    out = dict(shape=f'T={B * S} {d}->{d} r={r} bf16, layer fwd+bwd incl. the frozen GEMM')
    masks = [m.to(device) for m in layout_masks(variant, B, S, wl['layout'])]
    x0 = torch.randn(B, S, d, device=device).to(bf)
    dy = torch.randn(B, S, d, device=device).to(bf)

    def time_it(fn):
        for _ in range(warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / iters * 1e3, 1)

    try:
        kind, (ref,) = _reference_layers(wl, [(d, d)], bf, device=device)

        def run_ref():
            x = x0.clone().requires_grad_(True)
            _call_reference(ref, x, masks, avt, kind).backward(dy)
        out['eager_reference_cuda_us'] = time_it(run_ref)
        out['eager_reference_kind'] = kind
        del ref
    except Exception as exc:
        out['eager_reference_cuda_us'] = None
        out['eager_reference_error'] = str(exc)[:160]
    if avt:
        ours = MokaLinearAVT(d, d, rank=r, lora_alpha=16, lora_nums=M, blc_weight=wl['w'], lora_dropout=wl['dropout'],
                             loramethod='train', bias=False).to(device, bf)
        torch.nn.init.normal_(ours.lora_B0.weight, std=0.02)
        args = ([masks],)
    else:
        base = torch.nn.Linear(d, d, bias=False)
        ours = MokaLinearVT(base, 'image', r=r, lora_alpha=16, lora_dropout=wl['dropout'], attn_weight=wl['w'])
        ours.update_layer('text', r, lora_alpha=16, lora_dropout=wl['dropout'])
        ours.set_adapter(['image', 'text'])
        ours = ours.to(device, bf)
        torch.nn.init.normal_(ours.lora_B['text'].weight, std=0.02)
        args = (masks,)
    ours.train()

    def run_ours():
        x = x0.clone().requires_grad_(True)
        ours(x, *args[0]).backward(dy)
    out['moka_b200_us'] = time_it(run_ours)
    try:
        out['torch_split_gemm_adapter_only_us'] = round(mb.torch_split_gemm_us(variant, B, S, d, d, r, bf, iters, warmup), 1)
    except Exception as exc:
        out['torch_split_gemm_adapter_only_us'] = None
        out['torch_split_gemm_error'] = str(exc)[:160]
    return out


# --------------------------------------------------------------------------- main arm
def build_model(wl, device, ckpt):
    import torch
    from moka_b200.config import MokaConfig
    from moka_b200.inject import PeftMixedModel, get_peft_model
    from moka_b200.llama import LlamaConfig, LlamaForCausalLM
    cfg = {'7b': LlamaConfig.llama2_7b, '13b': LlamaConfig.llama2_13b, 'tiny': LlamaConfig.tiny}[wl['model']]()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    if wl['variant'] == 'avt':
        pm = get_peft_model(model, MokaConfig(variant='avt', r=wl['r'], num_modalities=wl['M'], lora_alpha=16,
                                              blc_weight=wl['w'], lora_dropout=wl['dropout']))
    else:
        c = MokaConfig(variant='vt', r=wl['r'], lora_alpha=16, attn_weight=wl['w'], lora_dropout=wl['dropout'])
        pm = PeftMixedModel(model, c, 'image')
        pm.add_adapter('text', c)
        pm.set_adapter(['image', 'text'])
    # the reference initialises B to zero; a fine-tune in progress has B != 0 -- use N(0, 0.02^2)
    g = torch.Generator(device=device).manual_seed(1234)
    with torch.no_grad():
        for n, p in pm.named_parameters():
            if 'lora_B' in n:
                p.normal_(mean=0.0, std=0.02, generator=g)
    model.gradient_checkpointing_enable(ckpt)
    pm.train()
    return pm, cfg


IMAGE_TOKEN_ID = 31999            # stand-in for the tokenizer's <image> id (VisualText/train/train.py:204)


def host_batch(wl, cfg, rank):
    """Synthetic batch in PINNED host memory, as the data pipeline holds it BEFORE any mask tensor exists
    (SURVEY 8f N1): token ids, labels and - AVT - the per-sample segment table (unified_arch.py:139-240); VT needs
    nothing else (its masks are a function of ids and labels, train.py:206-231)."""
    import torch
    from moka_b200.layouts import layout_ids_labels, layout_segments
    B, S = wl['B'], wl['S']
    gen = torch.Generator().manual_seed(1234 + rank)
    if wl['variant'] == 'avt':
        ids = torch.randint(3, cfg.vocab_size, (B, S), generator=gen)
        labels = ids.clone()
        labels[:, :int(S * 1680 / 2048)] = -100             # loss on the answer span only
        out = dict(input_ids=ids.pin_memory(), labels=labels.pin_memory(),
                   segments=layout_segments(B, S, wl['layout']).pin_memory())
    else:
        ids, labels = layout_ids_labels(B, S, wl['layout'], IMAGE_TOKEN_ID, cfg.vocab_size, seed=1234 + rank)
        out = dict(input_ids=ids.pin_memory(), labels=labels.pin_memory())
    return out


def to_device(hb, wl, device):
    """Host -> device copies of one step's inputs and the device-side token metadata (no host synchronisation:
    the status word is read once, after the timed region)."""
    from moka_b200.metadata import build_token_meta_avt_from_segments, build_token_meta_vt_from_ids
    ids = hb['input_ids'].to(device, non_blocking=True)
    labels = hb['labels'].to(device, non_blocking=True)
    if wl['variant'] == 'avt':
        seg = hb['segments'].to(device, non_blocking=True)
        meta = build_token_meta_avt_from_segments(seg, wl['S'], max_keys=128)
    else:
        meta = build_token_meta_vt_from_ids(ids, labels, IMAGE_TOKEN_ID, max_keys=128)
        ids = torch_where_image(ids)
    return ids, labels, dict(token_meta=meta)


def torch_where_image(ids):
    import torch
    return torch.where(ids == IMAGE_TOKEN_ID, torch.zeros_like(ids), ids)     # train.py:213


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='moka', choices=['moka', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--micro-batch', type=int, default=0)
    ap.add_argument('--zero-stage', type=int, default=1)
    ap.add_argument('--dropout', type=float, default=0.05,
                    help='lora_dropout; 0.05 is what both reference launch scripts train with')
    ap.add_argument('--grad-checkpointing', action='store_true')
    ap.add_argument('--flags', type=int, default=0, help='MokaDesc.flags (1 = force warp-FMA kernels)')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    wl['dropout'] = args.dropout
    if args.micro_batch:
        wl['B'] = args.micro_batch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(args.warmup, 3) if args.impl == 'moka' else args.warmup

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return 0
        base, t_layer = cpu_reference_leg(wl, seconds_budget=150.0, steps=max(args.steps, 5), warmup=max(args.warmup, 2))
        line = dict(metric=metric_name(wl), value=base['value'], unit='tokens/s',
                    impl='reference', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=t_layer * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype=base['dtype'], data='synthetic',
                    config=dict(workload=args.workload, lora_dropout=args.dropout, sample=base['sample']),
                    cpu_baseline=base,
                    e2e=dict(value=base['value'], unit='tokens/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ this repo's arm (GPU)
    import torch
    import torch.distributed as dist
    from moka_b200 import _cabi
    from moka_b200 import layer as moka_layer
    from moka_b200.dp import AdamWConfig, AdapterDataParallel
    _cabi.load()                                   # no CUDA library -> fail loudly, there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (B200)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    pm, cfg = build_model(wl, device, args.grad_checkpointing)
    if args.flags:
        for m in pm.moka_layers():
            m.kernel_flags = args.flags
    eng = AdapterDataParallel(pm, AdamWConfig(lr=1e-4), zero_stage=args.zero_stage, num_buckets=4)
    hb = host_batch(wl, cfg, rank)
    tokens_per_step = wl['B'] * wl['S'] * world
    h2d = sum(t.numel() * t.element_size() for t in hb.values())

    def step(dev_batch=None):
        ids, labels, kw = dev_batch if dev_batch is not None else to_device(hb, wl, device)
        loss, _ = pm(input_ids=ids, labels=labels, **kw)
        loss.backward()
        eng.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, resident, collect_events):
        dev_batch = to_device(hb, wl, device) if resident else None
        barrier()
        moka_layer.TIMING['events'].clear()
        moka_layer.TIMING['enabled'] = collect_events
        n0 = _cabi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n):
            loss = step(dev_batch)
            if not resident:
                last = float(loss.detach())                 # device -> host read of the step's result
        e1.record()
        barrier()
        moka_layer.TIMING['enabled'] = False
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _cabi.launch_count() - n0, last

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches, _ = timed(args.steps, resident=True, collect_events=True)
    events = list(moka_layer.TIMING['events'])
    ms_e2e, _, last_loss = timed(args.steps, resident=False, collect_events=False)
    clocks = sampler.stop() if rank == 0 else None
    to_device(hb, wl, device)[2]['token_meta'].raise_if_error()      # the deferred metadata check, outside the timed region

    # roofline of the fused MokaLinear fwd+bwd (algorithmic bytes of SURVEY 8d / DESIGN.md)
    alg_bytes, dur_ms, per_kind = 0.0, 0.0, {'fwd': [0.0, 0.0], 'bwd': [0.0, 0.0]}
    n_delta = 0
    by_shape = {}
    for kind, T, d_in, d_out, e, ev, delta in events:
        if delta:
            # MOKA_FLAG_DELTA_OUT: the residual read moved into the frozen GEMM's epilogue (beta = 1); the calls
            # timed here only write y / dx, so only those bytes are credited to them
            b = e * T * ((d_in + d_out) if kind == 'fwd' else (2 * d_in + d_out))
            n_delta += 1
        else:
            b = e * T * ((d_in + 2 * d_out) if kind == 'fwd' else (3 * d_in + d_out))
        t = ev[0].elapsed_time(ev[1])
        alg_bytes += b; dur_ms += t
        per_kind[kind][0] += b; per_kind[kind][1] += t
        bs = by_shape.setdefault(f'{d_in}x{d_out}', dict(fwd_ms=0.0, bwd_ms=0.0, fwd_bytes=0.0, bwd_bytes=0.0, calls=0,
                                                          fwd_list=[], bwd_list=[]))
        bs[kind + '_ms'] += t; bs[kind + '_bytes'] += b; bs['calls'] += 1
        bs[kind + '_list'].append(t * 1e3)
    peak, peak_kind = peak_hbm()
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9 if dur_ms > 0 else 0.0

    if rank == 0:
        value = tokens_per_step * args.steps / (ms_dev * 1e-3)
        e2e_value = tokens_per_step * args.steps / (ms_e2e * 1e-3)
        uses_tc = kernel_path_name(wl, cfg.hidden_size, args.flags)
        line = dict(
            metric=metric_name(wl), value=value, unit='tokens/s', n_gpus=world,
            steps=args.steps, warmup=warmup, ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling='weak',
            vs_baseline=None, dtype='bf16', data='synthetic', impl='moka_b200',
            config=dict(workload=args.workload, micro_batch_per_gpu=wl['B'], seq_len=wl['S'], rank=wl['r'],
                        modalities=wl['M'], lora_dropout=wl['dropout'], grad_checkpointing=args.grad_checkpointing,
                        parallelism=f'dp{world}-zero{args.zero_stage}', kernel_path=uses_tc,
                        adapted_linears=len(pm.moka_layers()), trainable_params=eng.trainable_numel(),
                        gemm_fusion=('delta-out: adapter term written alone, frozen GEMM accumulates onto it (beta=1)'
                                     if n_delta else 'off: adapter added in place after the frozen GEMM'),
                        l2='per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush',
                        base='frozen random-init Llama-2 (cuBLAS GEMMs + SDPA via PyTorch)'),
            e2e=dict(value=e2e_value, unit='tokens/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                     ms_per_step=ms_e2e / args.steps, last_loss=last_loss),
            gpu_launches=int(launches),
            # traffic: dram__bytes_read + dram__bytes_write of one forward + backward chain at the step's most
            # frequent shape, from the ncu --set full capture recorded in profiles/r02_ncu_traffic.json
            # (tools/ncu_traffic.py); null when no capture matches this workload's shape
            roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                          **ncu_traffic(wl, cfg.hidden_size),
                          peak_source=peak_kind,
                          kernel='fused MokaLinear fwd+bwd (libmoka.so calls, CUDA events per call)',
                          calls=len(events), delta_out_calls=n_delta,
                          bytes_per_token=('e(3 d_in + 2 d_out): fwd d_in + d_out, bwd 2 d_in + d_out [delta-out calls]' if n_delta
                                           else 'e(4 d_in + 3 d_out) [SURVEY 8d]'),
                          fwd_gbs=per_kind['fwd'][0] / max(per_kind['fwd'][1], 1e-9) / 1e6,
                          bwd_gbs=per_kind['bwd'][0] / max(per_kind['bwd'][1], 1e-9) / 1e6,
                          share_of_step=dur_ms / ms_dev,
                          by_shape={k: dict(calls=v['calls'],
                                            fwd_us=round(v['fwd_ms'] * 2e3 / max(v['calls'], 1), 1),
                                            bwd_us=round(v['bwd_ms'] * 2e3 / max(v['calls'], 1), 1),
                                            fwd_us_p10_p50_p90=[round(float(x), 1) for x in _pct(v['fwd_list'])],
                                            bwd_us_p10_p50_p90=[round(float(x), 1) for x in _pct(v['bwd_list'])],
                                            frac=round((v['fwd_bytes'] + v['bwd_bytes']) / max(v['fwd_ms'] + v['bwd_ms'], 1e-9)
                                                       / 1e6 / peak, 4))
                                    for k, v in by_shape.items()}),
            clocks=clocks)
        if world == 1:
            # the same calls timed ALONE (back to back, rotating buffers, no GEMMs around them, SM clock not under
            # the step's power cap) at the step's most frequent shape: explains the gap to the in-step figure above
            try:
                import tools.microbench as mb
                from moka_b200.layouts import VARIANT_AVT, VARIANT_VT
                d = cfg.hidden_size
                torch.cuda.empty_cache()
                iso = mb.run_point(VARIANT_AVT if wl['M'] == 3 else VARIANT_VT, wl['B'], wl['S'], d, d, wl['r'],
                                   torch.bfloat16, args.flags, 10, 3, wl['dropout'], nbuf=2)
                line['roofline']['isolated'] = dict(
                    shape=f"T={wl['B'] * wl['S']} {d}->{d} r={wl['r']} lora_dropout={wl['dropout']}",
                    fwd_us=iso['fwd_us'], bwd_us=iso['bwd_us'], achieved=iso['both_gbs'], frac=iso['both_gbs'] / peak,
                    note='tools/microbench.py run_point: CUDA events over 10 back-to-back fwd+bwd pairs')
            except Exception as exc:                       # never let the explanatory leg break the bench line
                line['roofline']['isolated'] = dict(error=str(exc)[:200])
        if world == 1 and not args.skip_cpu_baseline:
            base, _ = cpu_reference_leg(wl, seconds_budget=30.0, steps=5, warmup=2)
            line['cpu_baseline'] = base
            try:
                torch.cuda.empty_cache()
                line['gpu_comparators'] = gpu_comparators(wl, cfg.hidden_size, device)
            except Exception as exc:                       # explanatory leg only
                line['gpu_comparators'] = dict(error=str(exc)[:200])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
