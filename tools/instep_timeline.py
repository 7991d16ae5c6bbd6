"""Per-kernel wall-clock (globaltimer) durations of the LAST moka_forward / moka_backward call of a real bench step
(frozen Llama GEMMs around them, SM clock under the step's power cap) - to compare with tools/timeline.py, which runs
the same chain alone.  MOKA_DEBUG_TIMING=1 is set here.   usage: python tools/instep_timeline.py [micro_batch]"""
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
from moka_b200 import layer as moka_layer
moka_layer.TIMING['enabled'] = True
moka_layer.TIMING['marks'] = os.environ.get('MOKA_MARKS', '0') == '1'
emb = None
for m_ in pm.modules():
    if isinstance(m_, torch.nn.Embedding):
        emb = m_
for _ in range(3):
    ids, labels, kw = bench.to_device(hb, wl, dev)
    # the embeddings require grad here, so that layer 0's projections compute dx like every other layer's do and the
    # LAST backward call of the step (whose stamps survive) is a typical one
    e_in = emb(ids).detach().requires_grad_(True)
    loss, _ = pm(inputs_embeds=e_in, labels=labels, **kw)
    loss.backward(); eng.step()
torch.cuda.synchronize()
lib = _cabi.load()
n = 148 * 4 * 128
buf = (ctypes.c_ulonglong * n)()
_cabi.check(lib.moka_debug_read(buf, n), 'debug_read')
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 128).astype(np.int64)
slots = [('down', 100, 101), ('attn_fwd', 102, 103), ('up', 104, 105), ('prep', 106, 107), ('dgrad', 108, 109),
         ('attn_bwd', 110, 111), ('wgrad', 112, 113)]
print('last calls of the step (forward: layer L-1 down_proj 11008->4096; backward: layer 0, a 4096->4096 projection)')
print(f'{"kernel":10s} {"CTAs":>5s} | first entry -> last exit (us) | median CTA (us) | last exit - previous kernel\'s last exit (us)')
prev_end = None
for name, se, sx in slots:
    ent, ext = a[:, se], a[:, sx]
    ok = (ent > 0) & (ext > 0)
    if not ok.any():
        continue
    newest = ent[ok].max()
    ok &= ent > newest - 2_000_000
    e0, x1 = ent[ok].min(), ext[ok].max()
    gap = '' if prev_end is None or name == 'prep' else f'{(x1 - prev_end) / 1e3:8.2f}   (entered {(e0 - prev_end) / 1e3:7.2f} us after it)'
    print('%-10s %5d | %10.2f | %8.2f | %s' % (name, ok.sum(), (x1 - e0) / 1e3, np.median((ext - ent)[ok]) / 1e3, gap))
    prev_end = x1

# wgrad per-tile stamps of the last call (cycles relative to each CTA's first stamp, mean over CTAs)
w = a[(a[:, 112] > 0) & (a[:, 112] > a[:, 112].max() - 2_000_000)]
rel = w - w[:, 1:2]
mm = lambda i: rel[:, i].mean()
print('wgrad (last call) tiles: workers t_full / d_full / epilogue / done | mma op_full / x_full+ready / t_empty   (cycles)')
for j in range(8):
    b = 8 + j * 4
    print(f'  tile {j}: {mm(b):8.0f} {mm(b+1):8.0f} {mm(b+2):8.0f} {mm(b+3):8.0f} | {mm(40+j*3):8.0f} {mm(41+j*3):8.0f} {mm(42+j*3):8.0f}')
ev = [e for e in moka_layer.TIMING['events'] if e[0] == 'bwd'][-1]
print('last bwd call by CUDA events: %.1f us' % (ev[5][0].elapsed_time(ev[5][1]) * 1e3))
m0, m1 = a[-1, 0], a[-1, 1]
if not moka_layer.TIMING['marks']:
    raise SystemExit(0)
rows = {}
for name, se, sx in slots:
    ent, ext = a[:, se], a[:, sx]
    ok = (ent > 0) & (ext > 0)
    ok &= ent > ent[ok].max() - 2_000_000
    rows[name] = (ent[ok].min(), ext[ok].max())
print('marker before the call -> prep last exit %.1f us; marker before -> marker after %.1f us; wgrad last exit -> marker after %.1f us'
      % ((rows['prep'][1] - m0) / 1e3, (m1 - m0) / 1e3, (m1 - rows['wgrad'][1]) / 1e3))
print('prep first entry relative to the marker before: %.1f us' % ((rows['prep'][0] - m0) / 1e3))
evs = [e for e in moka_layer.TIMING['events']]
print('last 16 calls of the step (kind, d_in -> d_out, us):')
print('  '.join('%s %dx%d %.0f' % (e[0], e[2], e[3], e[5][0].elapsed_time(e[5][1]) * 1e3) for e in evs[-16:]))
fw = [e for e in evs if e[0] == 'fwd']
print('first 9 forward calls:', '  '.join('%dx%d %.0f' % (e[2], e[3], e[5][0].elapsed_time(e[5][1]) * 1e3) for e in fw[-224:-215]))
