"""Print the top stall sites of a `ncu --page source --csv` dump."""
import csv
import sys


def main(path, topn=22):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    si, src = idx['# Samples'], idx['Source']
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        try:
            n = int(r[si])
        except Exception:
            continue
        data.append((n, r))
    tot = sum(n for n, _ in data)
    agg = {h: 0 for h in stall_cols}
    for n, r in data:
        for h in stall_cols:
            agg[h] += int(r[idx[h]] or 0)
    print('total samples', tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]})
    for n, r in sorted(data, key=lambda x: -x[0])[:topn]:
        st = {h: int(r[idx[h]] or 0) for h in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f'{n / tot * 100:5.1f}% {r[src][:72]:72s} {top}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
