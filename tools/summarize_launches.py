"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches.md "title"
"""
import collections
import csv
import sys


def main():
    src, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else 'launch list')
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i
            break
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 2:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        n = r[ki].split('(')[0][:70]
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    moka = sum(t for n, (c, t) in agg.items() if 'moka::' in n)
    lines = [f'# {title}', '',
             f'source: `{src}` (`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are '
             'cold-cache and serialised: compare SHARES, not absolutes)', '',
             f'launches: {sum(v[0] for v in agg.values())}, total kernel time {tot / 1e6:.2f} ms, '
             f'libmoka.so kernels: {moka / 1e6:.2f} ms = {moka / tot * 100:.1f} % of the window', '',
             '| share | total ms | launches | avg us | kernel |', '|---|---|---|---|---|']
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        lines.append(f'| {t / tot * 100:.2f} % | {t / 1e6:.3f} | {c} | {t / c / 1e3:.1f} | `{n}` |')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[:12]))


if __name__ == '__main__':
    main()
