"""Per-CTA fixed costs of k_wgrad_tc inside a real training step (MOKA_DEBUG_TIMING stamps of the last backward call):
entry -> first operands / first tile done (prologue + pipeline fill) and last tile done -> exit (dA flush, store drain),
separately for the CTAs of the first wave (which also wait for the attention backward) and of the second wave.
usage: python tools/wgrad_overheads.py [micro_batch]"""
import ctypes, os, sys
os.environ['MOKA_DEBUG_TIMING'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from moka_b200 import _cabi
from moka_b200.dp import AdamWConfig, AdapterDataParallel

wl = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD]); wl['dropout'] = 0.05
if len(sys.argv) > 1:
    wl['B'] = int(sys.argv[1])
dev = torch.device('cuda', 0)
pm, cfg = bench.build_model(wl, dev, False)
eng = AdapterDataParallel(pm, AdamWConfig(lr=1e-4), zero_stage=1, num_buckets=4)
hb = bench.host_batch(wl, cfg, 0)
emb = [m for m in pm.modules() if isinstance(m, torch.nn.Embedding)][0]
for _ in range(3):
    ids, labels, kw = bench.to_device(hb, wl, dev)
    e_in = emb(ids).detach().requires_grad_(True)
    loss, _ = pm(inputs_embeds=e_in, labels=labels, **kw)
    loss.backward(); eng.step()
torch.cuda.synchronize()
lib = _cabi.load()
n = 148 * 4 * 128
buf = (ctypes.c_ulonglong * n)()
_cabi.check(lib.moka_debug_read(buf, n), 'debug_read')
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 128).astype(np.int64)
ok = (a[:, 112] > 0) & (a[:, 113] > 0)
ok &= a[:, 112] > a[ok, 112].max() - 2_000_000
w = a[ok]
t0 = w[:, 112].min()
late = w[:, 112] > t0 + 60_000                      # entered > 60 us after the first CTA: second wave
for name, sel in (('first wave', ~late), ('second wave', late)):
    c = w[sel]
    if not len(c):
        continue
    rel = lambda i: (c[:, i] - c[:, 1]).mean()
    dur = (c[:, 113] - c[:, 112]).mean() / 1e3
    print(f'{name}: {len(c)} CTAs, mean duration {dur:.1f} us (globaltimer)')
    print(f'  cycles after CTA entry: first operands {rel(40):.0f}, first x ready {rel(41):.0f}, first products issued {rel(42):.0f}, '
          f'tile 0 accumulator {rel(8):.0f}, tile 0 done {rel(11):.0f}, tile 7 done {rel(11 + 28):.0f}')
    print(f'  end: last epilogue done -> dA final {(c[:, 5] - c[:, 4]).mean():.0f}, dA flush {(c[:, 6] - c[:, 5]).mean():.0f}, '
          f'flush done -> exit {(c[:, 7] - c[:, 6]).mean():.0f}; entry -> exit {(c[:, 7] - c[:, 1]).mean():.0f} cycles')
