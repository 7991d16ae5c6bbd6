"""profiles/r01_microbench.md from the two sweep logs of tools/microbench.py (BASELINE config 5).

    python tools/sweep_to_md.py gpurun_out/sweep_p0.jsonl gpurun_out/sweep_p05.jsonl profiles/r01_microbench.md "session tag"
"""
import json
import sys


def rows(path):
    return [json.loads(l) for l in open(path) if l.startswith('{')]


def main():
    p0, p05, out, tag = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else '')
    rnd = sys.argv[5] if len(sys.argv) > 5 else 'r01'
    L = []
    L.append(f'# {rnd} ({tag}): MokaLinear microbench sweep (BASELINE config 5)\n')
    L.append('B200, `tools/microbench.py`, T = 8192 tokens (4 x 2048), square d -> d, bf16, adapter fwd+bwd through the C ABI, CUDA')
    L.append('events over 10 back-to-back iterations, inputs rotated over three buffer sets (> L2).  `frac` = algorithmic bytes')
    L.append('`2 (4 d_in + 3 d_out) T` / time / 6532.5 GB/s (measured copy bandwidth).  `torch` = the same layer as plain PyTorch on the')
    L.append('same GPU (cuBLAS split GEMMs over all rows, eager softmax attention, autograd backward; W0 GEMM excluded on both sides).')
    L.append('VT = 2 modalities / 64 keys per sample (layout C2), AVT = 3 modalities / 96 keys (layout C3).')
    L.append('`path`: tc = tcgen05 chain (round 2: every r % 4 == 0, r = 4 included); fma = warp-FMA kernels.\n')
    L.append('## lora_dropout = 0\n')
    L.append('| variant | d | r | path | fwd us | bwd us | fwd+bwd us | GB/s | frac | torch us | speed-up |')
    L.append('|---|---|---|---|---|---|---|---|---|---|---|')
    for r in rows(p0):
        t = r.get('torch_split_gemm_us')
        L.append(f"| {r['variant']} | {r['d_in']} | {r['r']} | {'tc' if r['tc'] else 'fma'} | {r['fwd_us']} | {r['bwd_us']} | "
                 f"{r['both_us']} | {r['both_gbs']} | {r['frac']} | {t if t is not None else ''} | "
                 f"{round(t / r['both_us'], 2) if t else ''} |")
    L.append('\n## lora_dropout = 0.05 (what the reference trains with)\n')
    L.append('| variant | d | r | path | fwd us | bwd us | fwd+bwd us | GB/s | frac |')
    L.append('|---|---|---|---|---|---|---|---|---|')
    for r in rows(p05):
        L.append(f"| {r['variant']} | {r['d_in']} | {r['r']} | {'tc' if r['tc'] else 'fma'} | {r['fwd_us']} | {r['bwd_us']} | "
                 f"{r['both_us']} | {r['both_gbs']} | {r['frac']} |")
    L.append('')
    open(out, 'w').write('\n'.join(L))


if __name__ == '__main__':
    main()
