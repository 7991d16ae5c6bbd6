"""HBM copy bandwidth of a plain torch copy (the MEASURED_PEAKS.json method: b.copy_(a), read + write bytes) timed
(a) alone and (b) interleaved with large bf16 GEMMs, i.e. under the power cap / clocks of a training step.
usage: python tools/copy_under_load.py"""
import json, statistics, subprocess, sys, threading
import torch
dev = 'cuda:0'
n = 1 << 29                                    # 512 Mi bf16 = 1 GiB per buffer
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)
g1 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
g2 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)


def copy_gbs(load):
    res = []
    for _ in range(30):
        if load:
            for _ in range(4):
                g1 @ g2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a); e1.record()
        torch.cuda.synchronize()
        res.append(2 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return res[5:]


for _ in range(3):
    b.copy_(a)
alone = copy_gbs(False)
# warm the power state up with GEMMs first
for _ in range(200):
    g1 @ g2
torch.cuda.synchronize()
load = copy_gbs(True)
clk = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw', '--format=csv,noheader'], capture_output=True, text=True).stdout.strip()
print(json.dumps(dict(copy_alone_gbs=round(statistics.median(alone), 1), copy_alone_max=round(max(alone), 1),
                      copy_between_gemms_gbs=round(statistics.median(load), 1), copy_between_gemms_min=round(min(load), 1),
                      clocks_after=clk)))
